/* hv_b200_ops.h -- operator-level C ABI of libhv_b200.so (sm_100a only).
 *
 * Each entry point is one hot-path operator of the CamAnimate denoising forward, on channels-last fp16 device
 * tensors (activations are (N, H, W, C) with N = batch*frames; "tokens x C" matrices are the same memory).  They
 * replace, one to one, the library calls the reference reaches through torch/diffusers:
 *
 *   hv_op_gemm            nn.Linear / 1x1 InflatedConv3d           src/models/resnet.py:9-15,211, transformer_3d.py:64,93,
 *                                                                  diffusers Attention.to_q/k/v/out, FeedForward (GEGLU)
 *   hv_op_conv3x3         3x3 InflatedConv3d (stride 1 / 2)        src/models/resnet.py:163,192,104; pose_adaptor.py:123,223
 *   hv_op_groupnorm       InflatedGroupNorm / nn.GroupNorm (+SiLU) src/models/resnet.py:18-26,218-235; transformer_3d.py:58,124
 *   hv_op_layernorm       nn.LayerNorm (+ temporal PE add)         src/models/attention.py:389-427; motion_module.py:244,256,273-277
 *   hv_op_attention       F.scaled_dot_product_attention over the spatial tokens of a frame, optional reference-bank
 *                         keys (ReferenceAttentionControl read hook)  src/models/attention.py:405; mutual_self_attention.py:147-186
 *   hv_op_temporal_attention  SDPA over the frame axis per pixel   src/models/motion_module.py:351-388
 *
 * All functions launch asynchronously on `stream`, never synchronise, return 0 on success or a negative hv_status,
 * and do not throw.  Pointers are caller-owned device memory, 16-byte aligned, row strides multiples of 8 elements.
 */
#ifndef HV_B200_OPS_H
#define HV_B200_OPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hv_stream_t; /* cudaStream_t */

enum hv_status {
  HV_OK = 0,
  HV_ERR_INVALID = -1,   /* bad argument (shape / alignment / unsupported size) */
  HV_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
  HV_ERR_TMA = -3,       /* tensor-map encoding failed */
  HV_ERR_MISSING = -4,   /* a required weight was never set */
  HV_ERR_STATE = -5      /* call order violation */
};

enum hv_act { HV_ACT_NONE = 0, HV_ACT_RELU = 1, HV_ACT_SILU = 2 };

/* Fused epilogue of hv_op_gemm / hv_op_conv3x3, applied in this order in fp32 with ONE fp16 rounding at the store (the
 * reference's eager fp16 modules round after every step; the fp32 oracle is the parity anchor):  v = acc + bias;
 * v += rowvec[row / rows_per_group]; v = act(v); v += residual[row].  With geglu != 0 the packed weight holds [128 hidden | 128 gate] rows per 256-row block and the
 * output is hidden * gelu_erf(gate) with N/2 columns. */
typedef struct hv_epilogue {
  const void* bias;      /* fp16 [N] or NULL */
  const void* rowvec;    /* fp16 [groups][rowvec_ld] or NULL (time-embedding projection per batch item) */
  int32_t rowvec_ld;
  int32_t rows_per_group;
  const void* residual;  /* fp16 [rows][ldr] or NULL */
  int32_t ldr;
  int32_t act;           /* hv_act */
  int32_t geglu;
  int32_t n_valid;       /* columns of the output actually written; 0 = all */
} hv_epilogue;

/* out[M, N] = A[M, K] * W[N, K]^T (+epilogue).  A may be split column-wise over two buffers (torch.cat of a skip
 * connection is never materialised): columns [0,K1) come from A (row stride lda), [K1,K) from A2 (row stride lda2);
 * K1 = 0 / A2 = NULL means a single source.  K1 must be a multiple of 64. */
int hv_op_gemm(const void* A, int64_t lda, const void* A2, int64_t lda2, int64_t K1, const void* W, void* out, int64_t ldc,
               int64_t M, int64_t N, int64_t K, const hv_epilogue* ep, hv_stream_t stream);

/* out[M][n*out_stride + j] = sum_k A[M][k] * X[n][j][k], n < batch, j < rows: the "swapped" GEMM that yields V^T with every
 * frame's token segment starting on a 16-byte boundary (out_stride % 8 == 0), as TMA readers of V^T require. */
int hv_op_gemm_batched_b(const void* A, int64_t lda, const void* X, int64_t ldx, void* out, int64_t ldc, int64_t M, int64_t batch,
                         int64_t rows, int64_t out_stride, int64_t K, const void* rowbias /* fp16 [M] added to row m, or NULL */,
                         hv_stream_t stream);

/* The small-channel 3x3 convolutions of the PoseGuider front (src/models/pose_guider.py:25-49) at their true channel counts (HBM-bound
 * layers at up to full image resolution; mma.sync implicit GEMM over a shared-memory halo tile): (Cin, Cout, stride) in
 * {(16,16,1), (16,32,2), (32,32,1), (32,96,2)}, padding 1, + bias, act.  X (NF, H, W, Cin) channels-last; Wp = hv_pack_conv3x3 with
 * Cin_pad = Cin, Cout_pad = Cout ([Cout][9 * Cin]); out (NF, Ho, Wo, ldo >= Cout), columns >= Cout are left untouched.
 * hv_op_pose_conv_in: the PoseGuider's conv_in (3 -> 16) read straight from the (B, 3, F, H, W) image, W in the reference layout (16,3,3,3),
 * out channels-last (B*F, H, W, 16). */
int hv_op_conv3x3_small(const void* X, const void* Wp, const void* bias, void* out, int64_t ldo, int64_t NF, int64_t H, int64_t W, int64_t Cin,
                        int64_t Cout, int32_t stride, int32_t act, hv_stream_t stream);
int hv_op_pose_conv_in(const void* X, const void* W, const void* bias, void* out, int64_t B, int64_t F, int64_t H, int64_t W_, int32_t act,
                       hv_stream_t stream);

/* Upsample3D (src/models/resnet.py:29-88): nearest 2x upsample followed by the 3x3 convolution, WITHOUT materialising the upsampled
 * tensor: every output parity (2y+py, 2x+px) is a 2x2 convolution of the source whose weights are sums of the 3x3 taps that land on
 * the same source pixel (4/9 of the multiply-adds; the summed weights are rounded to fp16 once, at packing).
 * X (NF, H, W, Cin) -> out (NF, 2H, 2W, Cout) given as [rows][ldc]; Wp = hv_pack_upconv2x2(W (Cout, Cin, 3, 3)) : [4][Cout][4 * Cin].
 * Epilogue: bias only.  Cin % 64 == 0; Cout a multiple of the 128/160/256-column tile the GEMM picks (320, 640, 1280 are). */
int hv_op_upconv2x2(const void* X, const void* Wp, void* out, int64_t ldc, int64_t NF, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                    const hv_epilogue* ep, hv_stream_t stream);
int hv_pack_upconv2x2(const void* W, void* out, int64_t Cout, int64_t Cin, hv_stream_t stream);

/* 3x3 convolution, padding 1, stride 1 or 2, over channels-last X (NF, H, W, Cin) -> out (NF, Ho, Wo, Cout) given as
 * a [rows][ldc] matrix.  Wp is the packed weight [Cout][9 * Cin] with k = (ky*3 + kx) * Cin + c (see hv_pack_conv3x3).
 * Cin must be a multiple of 64; stride 2 needs even H and W. */
int hv_op_conv3x3(const void* X, const void* Wp, void* out, int64_t ldc, int64_t NF, int64_t H, int64_t W, int64_t Cin,
                  int64_t Cout, int32_t stride, const hv_epilogue* ep, hv_stream_t stream);

/* Small-channel direct 3x3 convolution (any Cin/Cout, stride 1/2, padding 1) for conv_in and the PoseGuider stack.
 * W is the reference layout (Cout, Cin, 3, 3) fp16; act = hv_act. */
int hv_op_conv3x3_direct(const void* X, const void* W, const void* bias, void* out, int64_t NF, int64_t H, int64_t Wd,
                         int64_t Cin, int64_t Cout, int32_t stride, int32_t act, const void* add, hv_stream_t stream);

/* GroupNorm over (H*W, C/groups) per frame, fp32 statistics; optional SiLU; reads the channel concatenation of X
 * (C1 channels) and X2 (C2 channels, may be NULL/0) and writes one (NF, HW, C1+C2) tensor. stats: fp32 scratch
 * of hv_groupnorm_scratch_floats(...) floats (per-slab partial sums; the reduction is atomics-free and deterministic). */
size_t hv_groupnorm_scratch_floats(int64_t C, int64_t NF, int64_t HW, int32_t groups);
int hv_op_groupnorm(const void* X, int64_t C1, const void* X2, int64_t C2, const void* gamma, const void* beta, void* out,
                    int64_t NF, int64_t HW, int32_t groups, float eps, int32_t silu, float* stats, hv_stream_t stream);

/* LayerNorm over C per row.  Optional: `pre_add` fp16 [rows / rows_per_group][C] is added to x first and the sum is
 * written to `x_out` (the cross-attention residual); `pe` fp16 [F][C] is added to the normalised row
 * (row -> frame (row / hw) % F) as the motion module does before q/k/v. */
int hv_op_layernorm(const void* X, const void* gamma, const void* beta, void* out, int64_t rows, int64_t C, float eps,
                    const void* pre_add, int64_t rows_per_group, void* x_out, const void* pe, int64_t hw, int64_t F,
                    hv_stream_t stream);

/* Spatial multi-head attention.  Q: [NF*L][ldq] (head h at columns h*dpad, dpad = d rounded up to 16, pad columns zero), K
 * likewise.  Vt: [heads*dv][ldvt], dv = (d+1) rounded up to 16: rows h*dv..h*dv+d-1 = V of head h transposed, row h*dv+d = ONES (its
 * P*V column is the softmax denominator), remaining rows zero; holding
 * V transposed: frame n's tokens occupy columns [n*vt_stride, n*vt_stride + L), vt_stride % 8 == 0.  Optional bank keys/values per batch item (frame n uses bank n / F):
 * Kb [B][Lb][ldkb], Vbt [heads*dv][B*vbt_stride] (same row layout).  Frames with n < nf_nobank attend to their own L keys only (the
 * unconditional CFG half).  out: [NF*L][heads*d]. */
int hv_op_attention(const void* Q, const void* K, const void* Vt, void* out, int64_t NF, int64_t L, int32_t heads,
                    int32_t d, int64_t ldq, int64_t ldk, int64_t ldvt, int64_t ldo, const void* Kb, const void* Vbt,
                    int64_t Lb, int64_t ldkb, int64_t ldvbt, int64_t F, int64_t nf_nobank, int64_t vt_stride, int64_t vbt_stride,
                    hv_stream_t stream);

/* Attention over the frame axis: tokens are rows (b, f, p) of QKV [B*F*HW][3*C] (q | k | v), one problem per
 * (b, p, head) with F <= 32 keys. out [B*F*HW][C]. */
int hv_op_temporal_attention(const void* QKV, void* out, int64_t B, int64_t F, int64_t HW, int32_t heads, int32_t d,
                             hv_stream_t stream);

/* Layout / glue kernels. */
int hv_op_ncfhw_to_nhwc(const void* X, void* out, int64_t B, int64_t C, int64_t F, int64_t H, int64_t W, int32_t src_fp32,
                        hv_stream_t stream);
int hv_op_nhwc_to_ncfhw(const void* X, int64_t ldx, void* out, int64_t B, int64_t C, int64_t F, int64_t H, int64_t W,
                        hv_stream_t stream);
int hv_op_upsample2x(const void* X, void* out, int64_t NF, int64_t H, int64_t W, int64_t C, hv_stream_t stream);
int hv_op_add(const void* A, const void* B, void* out, int64_t n, hv_stream_t stream);
int hv_op_pixel_unshuffle(const void* X, void* out, int64_t B, int64_t C, int64_t F, int64_t H, int64_t W, int32_t r,
                          hv_stream_t stream);
/* Plucker-embedding producer on the device (Camera + ray_condition, src/dataset/dance_image_h_v_camera.py:17-130; called from
 * scripts/pose2vid.py:52-84) fused with CameraPoseEncoder's PixelUnshuffle(r) (pose_adaptor.py:222,233-236):
 * K: DEVICE fp32 [NF][4] = (fx, fy, cx, cy) in pixels; c2w: DEVICE fp32 [NF][16] row-major camera-to-world (relative poses);
 * out: channels-last fp16 (NF, H/r, W/r, 6*r*r) with channel = c*r*r + dy*r + dx, i.e. pixel_unshuffle(plucker (NF, 6, H, W)). */
int hv_op_plucker_unshuffle(const float* K, const float* c2w, void* out, int64_t NF, int64_t H, int64_t W, int32_t r, hv_stream_t stream);
/* y[M][N] = act_in(x[M][K]) * W[N][K]^T + bias, tiny M (time embedding, cross-attention collapse); fp32 math,
 * fp16 rounding of the result. act_in: hv_act applied to x first. */
int hv_op_small_linear(const void* X, const void* W, const void* bias, void* out, int64_t M, int64_t N, int64_t K,
                       int32_t act_in, hv_stream_t stream);
/* Timesteps(320, flip_sin_to_cos=True, shift=0): out[b] = [cos(t w_i) | sin(t w_i)] rounded to fp16. */
int hv_op_timestep_embedding(int64_t timestep, void* out, int64_t B, int64_t dim, hv_stream_t stream);

/* ---- per-timestep glue of Pose2VideoPipeline.__call__ on the device (src/pipelines/pipeline_pose2vid_long.py:516-563) ----
 * hv_op_window_gather: out[(r*Bl + b), c, i] = latents[b, c, frame_idx[i]] for r < repeat -- `latents[:, :, c].repeat(2, ...)` (:516-523).
 *   latents (Bl, C, Ftot, HW) fp16; frame_idx: DEVICE int32[Fw]; out (repeat*Bl, C, Fw, HW) fp16.
 * hv_op_cfg_ddim_step: for every latent element, mean over the windows that contain its frame of the window predictions
 *   (`noise_pred[:, :, c] += pred; counter[:, :, c] += 1; noise_pred / counter`, :550-556), classifier-free guidance
 *   `uncond + s (text - uncond)` (:557-559; skipped when pred_cond == NULL) and the DDIM update with eta = 0 (:561-563;
 *   diffusers DDIMScheduler.step): x0, eps from the model output (prediction_type 0 = v_prediction, 1 = epsilon),
 *   x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps, all in fp32, latents updated in place (one fp16 rounding).
 *   pred_uncond / pred_cond: HOST arrays of n_windows device pointers, each (Bl, C, Fw, HW) fp16 (the two halves of a
 *   CFG-doubled UNet output are pred and pred + Bl*C*Fw*HW); inv: DEVICE int32 [Ftot][K], entry = window*Fw + position of
 *   a window slot holding that frame, or -1; coef: DEVICE float [steps][4] = sqrt(a_t), sqrt(1-a_t), sqrt(a_prev),
 *   sqrt(1-a_prev); step_index: DEVICE int32 (row of coef used; NULL = row 0), read when the kernel runs so that a captured
 *   CUDA graph of one step replays for every timestep.
 * hv_op_advance_index: *index += 1 on the device (end of a graph-captured step). */
int hv_op_window_gather(const void* latents, const int32_t* frame_idx, void* out, int64_t Bl, int64_t C, int64_t Ftot, int64_t Fw, int64_t HW,
                        int32_t repeat, hv_stream_t stream);
int hv_op_cfg_ddim_step(const void* const* pred_uncond, const void* const* pred_cond, int32_t n_windows, const int32_t* inv, int32_t K,
                        const float* coef, const int32_t* step_index, void* latents, int64_t Bl, int64_t C, int64_t Ftot, int64_t Fw, int64_t HW,
                        float guidance_scale, int32_t prediction_type, hv_stream_t stream);
int hv_op_advance_index(int32_t* index, hv_stream_t stream);

/* Weight packing (device to device).  conv: fp16 (Cout, Cin, 3, 3) -> [Cout_pad][9*Cin_pad] (zero padded); geglu: rows of a [8C][K] matrix
 * (and its bias) interleaved in 128-row hidden/gate blocks; heads: [heads*d][K] -> [heads*dpad][K] with zero rows. */
int hv_pack_conv3x3(const void* W, void* out, int64_t Cout, int64_t Cin, int64_t Cout_pad, int64_t Cin_pad, hv_stream_t stream);
int hv_pack_geglu(const void* W, void* out, int64_t rows, int64_t K, hv_stream_t stream);
int hv_pack_heads(const void* W, void* out, int32_t heads, int32_t d, int32_t dpad, int64_t K, hv_stream_t stream);

/* Debug cross-checks (slow CUDA-core kernels, used by tests to localise a fault on the device). */
int hv_dbg_gemm(const void* A, int64_t lda, const void* W, float* out, int64_t M, int64_t N, int64_t K, hv_stream_t stream);

const char* hv_ops_last_error(void);
int hv_num_sms(void);

#ifdef __cplusplus
}
#endif
#endif
