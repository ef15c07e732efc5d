// Internal launcher declarations shared by the operator C ABI (capi_ops.cu) and the model runtime (model.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace hv {

cudaError_t launch_groupnorm(const __half* x1, int C1, const __half* x2, int C2, const __half* gamma, const __half* beta,
                             __half* out, int NF, int HW, int groups, float eps, int silu, float* stats, int num_sms,
                             cudaStream_t stream);
size_t groupnorm_scratch_floats(int C, int NF, int HW, int groups, int num_sms);
int groupnorm_num_launches(int C, int NF, int HW, int num_sms);
cudaError_t launch_layernorm(const __half* x, const __half* gamma, const __half* beta, __half* out, long long rows, int C, float eps,
                             const __half* pre_add, long long rows_per_group, __half* x_out, const __half* pe, int hw, int F,
                             cudaStream_t stream);
cudaError_t launch_temporal_attention(const __half* qkv, __half* out, int B, int F, int HW, int heads, int d, cudaStream_t stream);

struct AttnArgs {
  const __half* q; const __half* k; const __half* vt; __half* out;
  int NF, L, heads, d, dpad;
  long long ldq, ldk, ldvt, ldo;
  const __half* kb; const __half* vbt;   // bank keys [B][Lb][ldkb], bank values transposed [heads*d][ldvbt]
  int Lb; long long ldkb, ldvbt;
  int F, nf_nobank;
  long long vt_stride, vbt_stride;       // columns between consecutive frames / bank items in V^T (multiples of 8)
};
cudaError_t launch_attention(const AttnArgs& a, int num_sms, cudaStream_t stream);

cudaError_t launch_ncfhw_to_nhwc(const void* x, __half* out, int B, int C, int F, int H, int W, int src_fp32, cudaStream_t s, int ldo = 0);
cudaError_t launch_nhwc_to_ncfhw(const __half* x, int ldx, __half* out, int B, int C, int F, int H, int W, cudaStream_t s);
cudaError_t launch_upsample2x(const __half* x, __half* out, long long NF, int H, int W, int C, int num_sms, cudaStream_t s);
cudaError_t launch_add(const __half* a, const __half* b, __half* out, long long n, int num_sms, cudaStream_t s);
cudaError_t launch_pixel_unshuffle(const __half* x, __half* out, int B, int C, int F, int H, int W, int r, int num_sms, cudaStream_t s);
cudaError_t launch_plucker_unshuffle(const float* K, const float* c2w, __half* out, long long NF, int H, int W, int r, int num_sms, cudaStream_t s);
cudaError_t launch_small_linear(const __half* x, const __half* w, const __half* bias, __half* out, int M, int N, int K, int act_in,
                                cudaStream_t s);
cudaError_t launch_timestep_embedding(long long timestep, const long long* table, const int* index, __half* out, int B, int dim, cudaStream_t s);
cudaError_t launch_conv3x3_direct(const __half* x, const __half* w, const __half* bias, __half* out, long long NF, int H, int W, int Cin,
                                  int Cout, int stride, int act, const __half* add, int num_sms, cudaStream_t s);
cudaError_t launch_pack_conv3x3(const __half* w, __half* out, int Cout, int Cin, int Cout_pad, int Cin_pad, int num_sms, cudaStream_t s);
cudaError_t launch_conv3x3_direct_padded(const __half* x, const __half* w, const __half* bias, __half* out, long long NF, int H, int W,
                                         int Cin, int Cout, int ldo, int act, int num_sms, cudaStream_t s);
cudaError_t launch_pack_upconv2x2(const __half* w, __half* out, int Cout, int Cin, int num_sms, cudaStream_t s);
cudaError_t launch_pack_geglu(const __half* w, __half* out, int rows, int K, int num_sms, cudaStream_t s);
cudaError_t launch_pack_heads(const __half* w, __half* out, int heads, int d, int dpad, int K, int num_sms, cudaStream_t s);
// ---- small-channel convolutions of the PoseGuider front (smallconv.cu)
bool smallconv_supported(int cin, int cout, int stride);
cudaError_t launch_smallconv(const __half* x, const __half* wp, const __half* bias, __half* out, int NF, int H, int W, int cin, int cout, int stride,
                             int ldo, int act, cudaStream_t s);
cudaError_t launch_pg_conv_in(const __half* x, const __half* w, const __half* bias, __half* out, int B, int F, int H, int W, int act, int num_sms,
                              cudaStream_t s);

// ---- per-timestep glue of the denoising loop (step.cu)
constexpr int kMaxWindows = 32;
struct StepPreds {                       // window w's UNet prediction, (Bl, C, Fw, HW) fp16 each; cond unused without CFG
  const __half* uncond[kMaxWindows];
  const __half* cond[kMaxWindows];
};
cudaError_t launch_window_gather(const __half* latents, const int* idx, __half* out, int Bl, int C, int Ftot, int Fw, int HW, int R, int num_sms,
                                 cudaStream_t s);
cudaError_t launch_cfg_ddim_step(const StepPreds& preds, const int* inv, int K, const float* coef, const int* step_idx, __half* latents, int Bl, int C,
                                 int Ftot, int Fw, int HW, float guidance, int cfg, int epsilon, int num_sms, cudaStream_t s);
cudaError_t launch_advance_index(int* idx, cudaStream_t s);

cudaError_t launch_dbg_gemm(const __half* a, long long lda, const __half* w, float* out, int M, int N, int K, cudaStream_t s);

}  // namespace hv
