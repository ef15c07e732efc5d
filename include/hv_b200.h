/* hv_b200.h -- network-level C ABI of libhv_b200.so: the drop-in boundary for the CamAnimate denoising path.
 *
 * The reference (zhenzhiwang/HumanVid) is pure Python and has no FFI; these entry points are what a binding of its
 * three hot-path modules would call (see INTEGRATION.md for the ctypes stub):
 *
 *   HV_KIND_UNET3D        UNet3DConditionModel.forward   src/models/unet_3d.py:397-577
 *   HV_KIND_POSE_GUIDER   PoseGuider.forward             src/models/pose_guider.py:51-61
 *   HV_KIND_CAMERA_ENCODER CameraPoseEncoder.forward     src/cameractrl/pose_adaptor.py:232-248
 *   hv_set_ref_bank / hv_clear_ref_banks                 ReferenceAttentionControl.update / .clear
 *                                                        src/models/mutual_self_attention.py:302-363
 *   HV_KIND_UNET2D_REF    UNet2DConditionModel.forward in "write" mode (the reference / "writer" UNet, once per clip)
 *                                                        src/models/unet_2d_condition.py:872-1308 (post-process removed
 *                                                        :1295-1299), write hook mutual_self_attention.py:137-146
 *
 * Contract: all tensor arguments are caller-owned CUDA device pointers (torch allocations), contiguous, fp16 unless
 * stated, in the reference's own layouts ((B,C,F,H,W) etc.).  The library owns only its packed weights, reference
 * banks and (unless the caller passes one) its workspace.  Every launch goes on the stream passed in and returns
 * without synchronising.  Functions return 0 or a negative hv_status and never throw; hv_last_error(handle) gives the
 * message.  A handle is bound to the device current at hv_create and is not thread-safe (the reference is
 * single-threaded Python under torch.no_grad).
 */
#ifndef HV_B200_H
#define HV_B200_H

#include "hv_b200_ops.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hv_model* hv_handle;

enum hv_kind { HV_KIND_UNET3D = 0, HV_KIND_POSE_GUIDER = 1, HV_KIND_CAMERA_ENCODER = 2, HV_KIND_UNET2D_REF = 3 };
enum hv_dtype { HV_F16 = 0, HV_F32 = 1 };
enum hv_forward_flags {
  HV_FLAG_CFG = 1,         /* batch = [uncond half ; cond half]: the first half ignores the reference banks
                              (mutual_self_attention.py:166-186) */
  HV_FLAG_UNCOND_ONLY = 2, /* the batch is an unconditional CFG half only (one unit of a (window x CFG-half) split over
                              GPUs): no item reads the banks */
  HV_FLAG_COND_ONLY = 4    /* the batch is a conditional CFG half only: its B items read the LAST B items of the banks
                              (banks are written for [uncond ; cond], pipeline_pose2vid_long.py:470-480) */
};

typedef struct hv_config {
  int32_t kind; /* hv_kind */
  /* UNet3DConditionModel (unet_3d.py:34-83 + configs/inference/inference_v2.yaml) */
  int32_t in_channels;           /* 4 */
  int32_t out_channels;          /* 4 */
  int32_t block_out_channels[4]; /* 320 640 1280 1280 */
  int32_t heads;                 /* attention_head_dim = 8 heads */
  int32_t cross_attention_dim;   /* 768 */
  int32_t norm_groups;           /* 32 */
  int32_t use_motion_module;     /* 1 (config 1 / pose2img: 0) */
  int32_t motion_max_len;        /* temporal_position_encoding_max_len = 32 */
  /* PoseGuider (pose_guider.py:17-49) */
  int32_t pg_cond_channels;      /* 3 */
  int32_t pg_block_channels[4];  /* 16 32 96 256 */
  int32_t pg_out_channels;       /* 320 */
  /* CameraPoseEncoder (pose_adaptor.py:162-230, pose_encoder_kwargs) */
  int32_t cam_downscale;         /* 8 */
  int32_t cam_cin;               /* 384 */
  int32_t cam_channels;          /* 320 */
  int32_t cam_nums_rb;           /* 2 */
  int32_t cam_heads;             /* 8 */
  int32_t cam_max_len;           /* 24 */
} hv_config;

int hv_create(const hv_config* cfg, hv_handle* out);
void hv_destroy(hv_handle h);
const char* hv_last_error(hv_handle h);

/* Copies (and converts to fp16) one state_dict entry under the reference's key name, e.g.
 * "down_blocks.0.resnets.1.conv1.weight".  Unknown keys are accepted and ignored (returns 0). */
int hv_set_weight(hv_handle h, const char* key, const void* dev_ptr, const int64_t* shape, int32_t ndim, int32_t dtype,
                  hv_stream_t stream);
/* Packs all weights into their kernel layouts; fails with HV_ERR_MISSING naming the first absent key. */
int hv_finalize(hv_handle h, hv_stream_t stream);

/* Reference-attention K/V banks in reader order (all TemporalBasicTransformerBlocks of the UNet in DFS order
 * down_blocks, up_blocks, mid_block, stable-sorted by descending width).  bank: (B_ref, L, C) fp16. */
int hv_num_ref_blocks(hv_handle h);
int hv_ref_block_dim(hv_handle h, int32_t block_idx);
int hv_set_ref_bank(hv_handle h, int32_t block_idx, const void* dev_ptr, int64_t B_ref, int64_t L, int64_t C, hv_stream_t stream);
int hv_clear_ref_banks(hv_handle h);

/* Bytes of scratch one forward needs at this shape (h, w = spatial size of the tensor entering the network). */
size_t hv_workspace_bytes(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width);
/* Grows the handle's private workspace to what this shape needs (cudaMalloc here, never inside a forward).  A forward
 * called with workspace == NULL uses it and fails with HV_ERR_STATE if it is too small. */
int hv_reserve_workspace(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width);
/* Device-resident timestep for CUDA-graph replay of a denoising loop (pipeline_pose2vid_long.py:454-563): when set, every
 * UNet forward of this handle embeds dev_table[*dev_index] (read on the device when the kernel runs) and ignores its
 * `timestep` argument.  Both NULL restores the host argument. */
int hv_set_timestep_source(hv_handle h, const int64_t* dev_table, const int32_t* dev_index);

/* sample (B,4,F,h,w), encoder_hidden_states (B,1,cross_attention_dim), pose_cond_fea (B,320,F,h,w) or NULL,
 * out (B,4,F,h,w); all fp16.  workspace may be NULL: the handle's reserved workspace (hv_reserve_workspace) is used. */
int hv_unet3d_forward(hv_handle h, const void* sample, int64_t timestep, const void* encoder_hidden_states, const void* pose_cond_fea,
                      void* out, int32_t B, int32_t F, int32_t height, int32_t width, uint32_t flags, void* workspace,
                      size_t ws_bytes, hv_stream_t stream);
/* Reference ("writer") UNet, kind HV_KIND_UNET2D_REF (hv_config as for the UNet3D with use_motion_module = 0; no
 * conv_norm_out / conv_out weights).  sample (B,4,h,w), encoder_hidden_states (B,1,cross_attention_dim);
 * bank_out[i] (i in reader order, see hv_num_ref_blocks / hv_ref_block_dim) receives block i's LayerNorm-1 output
 * (B, h_i*w_i, C_i) fp16 -- exactly what hv_set_ref_bank of the denoising UNet takes; hidden_out (B,C0,h,w) or NULL is
 * the forward's return value (the last up block's output). */
int hv_unet2d_reference_forward(hv_handle h, const void* sample, int64_t timestep, const void* encoder_hidden_states, void* hidden_out,
                                void* const* bank_out, int32_t n_banks, int32_t B, int32_t height, int32_t width, void* workspace,
                                size_t ws_bytes, hv_stream_t stream);
/* conditioning (B,3,F,H,W) -> out (B,320,F,H/8,W/8) */
int hv_pose_guider_forward(hv_handle h, const void* conditioning, void* out, int32_t B, int32_t F, int32_t H, int32_t W,
                           void* workspace, size_t ws_bytes, hv_stream_t stream);
/* plucker (B,6,F,H,W) -> out (B*F,320,H/8,W/8) (the single feature map of the one-level encoder) */
int hv_camera_encoder_forward(hv_handle h, const void* plucker, void* out, int32_t B, int32_t F, int32_t H, int32_t W,
                              void* workspace, size_t ws_bytes, hv_stream_t stream);

/* The same encoder fed by the cameras themselves instead of the 6-channel Plucker image: intrinsics DEVICE fp32
 * (B, F, 4) = (fx, fy, cx, cy) in pixels, c2w DEVICE fp32 (B, F, 4, 4) relative camera-to-world poses -- what
 * scripts/pose2vid.py:66-80 hands to ray_condition.  The embedding is generated inside the PixelUnshuffle producer. */
int hv_camera_encoder_forward_rays(hv_handle h, const float* intrinsics, const float* c2w, void* out, int32_t B, int32_t F, int32_t H,
                                   int32_t W, void* workspace, size_t ws_bytes, hv_stream_t stream);

/* Debug taps (per-layer error ladder, profiles/): the activation leaving the i-th block of the UNet forward (conv_in, every
 * resnet / transformer / motion module / down- / up-sampler, in execution order) is copied to dst[i] as channels-last fp16
 * (NF, H, W, C).  hv_debug_tap_count plans the forward at this shape and returns the number of taps; hv_debug_tap_info
 * names tap i (the reference's module path) and gives dims4 = {NF, H, W, C}; hv_debug_set_taps(h, NULL, 0) disables. */
int hv_debug_tap_count(hv_handle h, int32_t B, int32_t F, int32_t height, int32_t width);
int hv_debug_tap_info(hv_handle h, int32_t i, char* name, int32_t name_cap, int64_t* dims4);
int hv_debug_set_taps(hv_handle h, void* const* dst, int32_t n);

/* Kernel launches issued by the last forward on this handle (bench.py's gpu_launches). */
int64_t hv_last_launch_count(hv_handle h);

/* Optional per-operator-category device timing of the next forwards (CUDA events around every launch on the caller's
 * stream).  Categories: 0 tcgen05 GEMM (linear / 1x1), 1 tcgen05 implicit-GEMM 3x3 conv, 2 spatial attention,
 * 3 temporal attention, 4 GroupNorm/LayerNorm, 5 small linears.  hv_get_profile synchronises on the last event and
 * returns summed milliseconds, algorithmic FLOPs (2*MAC, unpadded shapes) and launch counts of the LAST forward. */
int hv_set_profiling(hv_handle h, int32_t enable);
int hv_get_profile(hv_handle h, double* ms, double* flops, int64_t* count, int32_t ncat);
/* One CSV line per timed launch of the last profiled forward: idx,cat,label,M,N,K,ms,tflops. */
int hv_dump_profile(hv_handle h, const char* path);

#ifdef __cplusplus
}
#endif
#endif
