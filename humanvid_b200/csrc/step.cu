// Per-timestep glue of the denoising loop on the device: what Pose2VideoPipeline.__call__ does around the
// UNet forward with a dozen small eager ops per step (src/pipelines/pipeline_pose2vid_long.py:516-563):
//
//   window gather   latent_model_input = latents[:, :, window].repeat(2 if CFG)                          (:516-523)
//   step            noise_pred[:, :, window] += pred ; counter[:, :, window] += 1   for every window    (:550-552)
//                   noise_pred / counter ; uncond + s * (text - uncond)                                  (:555-559)
//                   DDIM update (eta = 0, v-prediction or epsilon): x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps  (:561-563)
//
// One elementwise kernel does accumulate -> /counter -> CFG -> DDIM in fp32 with a single fp16 rounding at the latent store,
// reading the DDIM coefficients of the CURRENT step from a device table indexed by a device-resident step counter, so a
// captured CUDA graph of one step replays for every timestep (hv_op_advance_index bumps the counter at the end of the step).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace hv {

namespace {

constexpr int kThreads = 256;

// out[(r*Bl + b), c, i, p] = latents[b, c, idx[i], p]   for r < R  (R = 2: the CFG-doubled batch)
__global__ void window_gather_kernel(const __half* __restrict__ latents, const int* __restrict__ idx, __half* __restrict__ out, int Bl, int C,
                                     int Ftot, int Fw, int HW, int R) {
  const long long total = static_cast<long long>(Bl) * C * Fw * HW;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(e % HW);
    long long t = e / HW;
    const int i = static_cast<int>(t % Fw);
    t /= Fw;   // t = b*C + c
    const __half v = latents[(t * Ftot + idx[i]) * HW + p];
    for (int r = 0; r < R; ++r) out[e + r * total] = v;
  }
}

__global__ void cfg_ddim_step_kernel(StepPreds preds, const int* __restrict__ inv, int K, const float* __restrict__ coef,
                                     const int* __restrict__ step_idx, __half* __restrict__ latents, int Bl, int C, int Ftot, int Fw, int HW,
                                     float guidance, int cfg, int epsilon) {
  const float4 cf = *reinterpret_cast<const float4*>(coef + 4 * (step_idx ? *step_idx : 0));   // sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)
  const long long total = static_cast<long long>(Bl) * C * Ftot * HW;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(e % HW);
    long long t = e / HW;
    const int f = static_cast<int>(t % Ftot);
    t /= Ftot;   // t = b*C + c
    float un = 0.f, tx = 0.f;
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
      const int j = inv[f * K + k];
      if (j < 0) continue;
      const int w = j / Fw, i = j - w * Fw;
      const long long off = (t * Fw + i) * HW + p;
      un += __half2float(preds.uncond[w][off]);
      if (cfg) tx += __half2float(preds.cond[w][off]);
      ++cnt;
    }
    const float inv_cnt = cnt > 0 ? 1.f / static_cast<float>(cnt) : 0.f;
    un *= inv_cnt;
    tx *= inv_cnt;
    const float m = cfg ? un + guidance * (tx - un) : un;
    const float x = __half2float(latents[e]);
    float x0, eps;
    if (!epsilon) {   // v-prediction
      x0 = cf.x * x - cf.y * m;
      eps = cf.x * m + cf.y * x;
    } else {
      eps = m;
      x0 = (x - cf.y * eps) / cf.x;
    }
    latents[e] = __float2half_rn(cf.z * x0 + cf.w * eps);
  }
}

__global__ void advance_index_kernel(int* idx) { *idx += 1; }

inline unsigned blocks_for(long long n, int num_sms) {
  long long b = (n + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(num_sms) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

cudaError_t launch_window_gather(const __half* latents, const int* idx, __half* out, int Bl, int C, int Ftot, int Fw, int HW, int R, int num_sms,
                                 cudaStream_t s) {
  const long long total = static_cast<long long>(Bl) * C * Fw * HW;
  window_gather_kernel<<<blocks_for(total, num_sms), kThreads, 0, s>>>(latents, idx, out, Bl, C, Ftot, Fw, HW, R);
  return cudaGetLastError();
}

cudaError_t launch_cfg_ddim_step(const StepPreds& preds, const int* inv, int K, const float* coef, const int* step_idx, __half* latents, int Bl, int C,
                                 int Ftot, int Fw, int HW, float guidance, int cfg, int epsilon, int num_sms, cudaStream_t s) {
  const long long total = static_cast<long long>(Bl) * C * Ftot * HW;
  cfg_ddim_step_kernel<<<blocks_for(total, num_sms), kThreads, 0, s>>>(preds, inv, K, coef, step_idx, latents, Bl, C, Ftot, Fw, HW, guidance, cfg, epsilon);
  return cudaGetLastError();
}

cudaError_t launch_advance_index(int* idx, cudaStream_t s) {
  advance_index_kernel<<<1, 1, 0, s>>>(idx);
  return cudaGetLastError();
}

}  // namespace hv
