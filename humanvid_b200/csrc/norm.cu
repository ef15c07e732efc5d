// GroupNorm (+SiLU) and LayerNorm (+cross-attention residual, +temporal positional encoding) on channels-last fp16.
// Both are HBM-bound passes: fp32 statistics, fp16 rounding exactly where the reference's eager fp16 modules round
// (after the norm, after SiLU, after the PE add).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hv {

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_h2(v[0], v[1]);
  u.y = pack_h2(v[2], v[3]);
  u.z = pack_h2(v[4], v[5]);
  u.w = pack_h2(v[6], v[7]);
  return u;
}

// ---------------------------------------------------------------- GroupNorm statistics
// grid (slabs, NF); thread = (row lane, 8-channel vector).  stats[n][g] = {sum, sumsq} accumulated with atomics.
__global__ void gn_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int groups,
                                int rows_per_block, float* __restrict__ stats) {
  extern __shared__ float sacc[];  // [groups][2]
  const int C = C1 + C2, vecs = C / 8, cpg = C / groups;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int rl = threadIdx.x / vecs, v = threadIdx.x % vecs, rpi = blockDim.x / vecs;
  if (rl < rpi) {
    const int c0 = v * 8;
    const __half* src;
    int ldc, cc;
    if (c0 < C1) { src = x1; ldc = C1; cc = c0; } else { src = x2; ldc = C2; cc = c0 - C1; }
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const int row_end = min(HW, (blockIdx.x + 1) * rows_per_block);
    for (int r = blockIdx.x * rows_per_block + rl; r < row_end; r += rpi) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<long long>(n) * HW + r) * ldc + cc));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i] += f[2 * i] + f[2 * i + 1];
        q[i] += f[2 * i] * f[2 * i] + f[2 * i + 1] * f[2 * i + 1];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // channels-per-group is even, so a half2 never straddles two groups
      const int g = (c0 + 2 * i) / cpg;
      atomicAdd(&sacc[2 * g], s[i]);
      atomicAdd(&sacc[2 * g + 1], q[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sacc[i]);
}

__global__ void gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, __half* __restrict__ out,
                                long long total_vecs, int HW, int groups, float eps, int silu, const float* __restrict__ stats) {
  const int C = C1 + C2, vecs = C / 8, cpg = C / groups;
  const float inv_cnt = 1.f / (static_cast<float>(HW) * cpg);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total_vecs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vecs;
    const int c0 = static_cast<int>(i % vecs) * 8;
    const int n = static_cast<int>(row / HW);
    const __half* src = c0 < C1 ? x1 + row * C1 + c0 : x2 + row * C2 + (c0 - C1);
    float f[8], gm[8], bt[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + c0)), gm);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta + c0)), bt);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = (c0 + 2 * k) / cpg;
      const float sum = stats[(static_cast<long long>(n) * groups + g) * 2], sq = stats[(static_cast<long long>(n) * groups + g) * 2 + 1];
      const float mean = sum * inv_cnt;
      const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float y = r16((f[2 * k + j] - mean) * rstd * gm[2 * k + j] + bt[2 * k + j]);
        if (silu) y = silu_f(y);
        f[2 * k + j] = y;
      }
    }
    *reinterpret_cast<uint4*>(out + row * C + c0) = pack8(f);
  }
}

// ---------------------------------------------------------------- LayerNorm: one warp per row
template <int MAXV>  // vectors (8 halves) per lane
__global__ void layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                 __half* __restrict__ out, long long rows, int C, float eps, const __half* __restrict__ pre_add,
                                 long long rows_per_group, __half* __restrict__ x_out, const __half* __restrict__ pe, int hw, int F) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int vecs = C / 8;
  float v[MAXV][8];
  float sum = 0.f;
  const __half* add = pre_add ? pre_add + (row / rows_per_group) * C : nullptr;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < vecs) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C + vi * 8)), v[k]);
      if (add) {
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(add + vi * 8)), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = r16(v[k][j] + a[j]);
        if (x_out) *reinterpret_cast<uint4*>(x_out + row * C + vi * 8) = pack8(v[k]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const __half* pe_row = pe ? pe + static_cast<long long>((row / hw) % F) * C : nullptr;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < vecs) {
      float gm[8], bt[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + vi * 8)), gm);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta + vi * 8)), bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = (v[k][j] - mean) * rstd * gm[j] + bt[j];
      if (pe_row) {
        float pv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(pe_row + vi * 8)), pv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = r16(v[k][j]) + pv[j];
      }
      *reinterpret_cast<uint4*>(out + row * C + vi * 8) = pack8(v[k]);
    }
  }
}

}  // namespace

cudaError_t launch_groupnorm(const __half* x1, int C1, const __half* x2, int C2, const __half* gamma, const __half* beta,
                             __half* out, int NF, int HW, int groups, float eps, int silu, float* stats, int num_sms,
                             cudaStream_t stream) {
  const int C = C1 + C2;
  if (C % 8 || C1 % 8 || C % groups || ((C / groups) & 1)) return cudaErrorInvalidValue;
  const int vecs = C / 8;
  if (vecs > 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(float) * 2 * NF * groups, stream);
  if (e != cudaSuccess) return e;
  int threads = vecs * (vecs >= 256 ? 1 : 256 / vecs);
  if (threads > 1024) threads = vecs;
  // enough blocks to fill the machine ~4x, at least 32 rows per block
  int slabs = (4 * num_sms + NF - 1) / NF;
  int rows_per_block = (HW + slabs - 1) / slabs;
  if (rows_per_block < 32) rows_per_block = 32;
  slabs = (HW + rows_per_block - 1) / rows_per_block;
  gn_stats_kernel<<<dim3(slabs, NF), threads, groups * 2 * sizeof(float), stream>>>(x1, C1, x2, C2, HW, groups, rows_per_block, stats);
  const long long total = static_cast<long long>(NF) * HW * vecs;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms * 16LL) blocks = num_sms * 16LL;
  gn_apply_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x1, C1, x2, C2, gamma, beta, out, total, HW, groups, eps, silu, stats);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const __half* x, const __half* gamma, const __half* beta, __half* out, long long rows, int C, float eps,
                             const __half* pre_add, long long rows_per_group, __half* x_out, const __half* pe, int hw, int F,
                             cudaStream_t stream) {
  if (C % 8) return cudaErrorInvalidValue;
  const int vecs = C / 8;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  if (rows_per_group <= 0) rows_per_group = 1;
  if (hw <= 0) hw = 1;
  if (F <= 0) F = 1;
  if (vecs <= 32)
    layernorm_kernel<1><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else if (vecs <= 64)
    layernorm_kernel<2><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else if (vecs <= 96)
    layernorm_kernel<3><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else if (vecs <= 160)
    layernorm_kernel<5><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace hv
