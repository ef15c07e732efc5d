// GroupNorm (+SiLU) and LayerNorm (+cross-attention residual, +temporal positional encoding) on channels-last fp16.
// Both are HBM-bound passes: fp32 statistics and fp32 arithmetic through norm -> SiLU / PE add, one fp16 rounding at the store.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tuning.h"

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

// ---------------------------------------------------------------- GroupNorm statistics (deterministic: no atomics)
// grid (slabs, NF); thread = (row lane, 8-channel vector).  Each block writes its per-group {sum, sumsq} to
// partial[n][slab][g]; gn_apply_kernel adds the slabs in a fixed order -> {mean, rstd}.
__global__ void gn_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW, int groups,
                                int rows_per_block, float* __restrict__ partial) {
  extern __shared__ float sred[];  // [blockDim.x][4][2]
  const int C = C1 + C2, vecs = C / 8, cpg = C / groups;
  const int n = blockIdx.y;
  const int rl = threadIdx.x / vecs, v = threadIdx.x % vecs, rpi = blockDim.x / vecs;
  const int c0 = v * 8;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  {
    const __half* src;
    int ldc, cc;
    if (c0 < C1) { src = x1; ldc = C1; cc = c0; } else { src = x2; ldc = C2; cc = c0 - C1; }
    const int row_end = min(HW, (blockIdx.x + 1) * rows_per_block);
    const __half* base = src + static_cast<long long>(n) * HW * ldc + cc;
    int r = blockIdx.x * rows_per_block + rl;
    for (; r + 3 * rpi < row_end; r += 4 * rpi) {  // four independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r + k * rpi) * ldc));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[i] += f[2 * i] + f[2 * i + 1];
          q[i] += f[2 * i] * f[2 * i] + f[2 * i + 1] * f[2 * i + 1];
        }
      }
    }
    for (; r < row_end; r += rpi) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * ldc));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i] += f[2 * i] + f[2 * i + 1];
        q[i] += f[2 * i] * f[2 * i] + f[2 * i + 1] * f[2 * i + 1];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sred[(threadIdx.x * 4 + i) * 2] = s[i];
    sred[(threadIdx.x * 4 + i) * 2 + 1] = q[i];
  }
  __syncthreads();
  // thread g sums, in a fixed order, every (row lane, half2) slot that belongs to group g (channels-per-group is even,
  // so a half2 never straddles two groups; the slots of a group are the contiguous half2 range [g*cpg/2, (g+1)*cpg/2))
  if (threadIdx.x < groups) {
    const int g = threadIdx.x, h0 = g * (cpg / 2), h1 = h0 + cpg / 2;
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < rpi; ++r)
      for (int hh = h0; hh < h1; ++hh) {
        const int slot = (r * vecs + hh / 4) * 4 + (hh & 3);
        ss += sred[slot * 2];
        qq += sred[slot * 2 + 1];
      }
    float* dst = partial + ((static_cast<long long>(n) * gridDim.x + blockIdx.x) * groups + g) * 2;
    dst[0] = ss;
    dst[1] = qq;
  }
}

// grid (slabs, NF), thread = (row lane, 8-channel vector) like the statistics kernel: the thread's channels are fixed, so
// gamma/beta/mean/rstd fold into one per-channel scale and shift kept in registers, and the row loop has no integer
// division and one FFMA + SiLU per element.
__global__ void gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, __half* __restrict__ out,
                                int HW, int groups, int rows_per_block, int silu, const float* __restrict__ partial, int slabs, float inv_cnt,
                                float eps) {
  const int C = C1 + C2, vecs = C / 8, cpg = C / groups;
  const int n = blockIdx.y;
  const int rl = threadIdx.x / vecs, v = threadIdx.x % vecs, rpi = blockDim.x / vecs;
  const int c0 = v * 8;
  // {mean, rstd} of this frame's groups: four lanes per group add the per-slab partial sums (each lane its slabs in order, then a fixed
  // two-step tree -> deterministic), full warps only so that the shuffles are well defined whatever the block size
  extern __shared__ float2 sstat[];   // [groups]
  {
    const int full = (blockDim.x >> 5) << 5;
    if (static_cast<int>(threadIdx.x) < full) {
      for (int g0 = 0; g0 < groups; g0 += full / 4) {
        const int g = g0 + static_cast<int>(threadIdx.x) / 4, l = threadIdx.x & 3;
        float ss = 0.f, qq = 0.f;
        if (g < groups)
          for (int sl = l; sl < slabs; sl += 4) {
            const float2 pp = __ldg(reinterpret_cast<const float2*>(partial + ((static_cast<long long>(n) * slabs + sl) * groups + g) * 2));
            ss += pp.x;
            qq += pp.y;
          }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        qq += __shfl_xor_sync(0xffffffffu, qq, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        qq += __shfl_xor_sync(0xffffffffu, qq, 2);
        if (g < groups && l == 0) {
          const float mean = ss * inv_cnt;
          sstat[g] = make_float2(mean, rsqrtf(fmaxf(qq * inv_cnt - mean * mean, 0.f) + eps));
        }
      }
    }
    __syncthreads();
  }
  float sc[8], sh[8];
  {
    float gm[8], bt[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + c0)), gm);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta + c0)), bt);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 st = sstat[(c0 + 2 * k) / cpg];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[2 * k + j] = st.y * gm[2 * k + j];
        sh[2 * k + j] = fmaf(-st.x, sc[2 * k + j], bt[2 * k + j]);
      }
    }
  }
  const __half* src;
  int ldc;
  if (c0 < C1) { src = x1 + static_cast<long long>(n) * HW * C1 + c0; ldc = C1; } else { src = x2 + static_cast<long long>(n) * HW * C2 + (c0 - C1); ldc = C2; }
  __half* dst = out + static_cast<long long>(n) * HW * C + c0;
  const int row_end = min(HW, (blockIdx.x + 1) * rows_per_block);
  auto emit = [&](const uint4& u, int r) {
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(f[j], sc[j], sh[j]);
      if (silu) y = __fdividef(y, 1.0f + __expf(-y));
      f[j] = y;
    }
    *reinterpret_cast<uint4*>(dst + static_cast<long long>(r) * C) = pack8(f);
  };
  int r = blockIdx.x * rows_per_block + rl;
  for (; r + 3 * rpi < row_end; r += 4 * rpi) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r + k * rpi) * ldc));
#pragma unroll
    for (int k = 0; k < 4; ++k) emit(u[k], r + k * rpi);
  }
  for (; r < row_end; r += rpi) emit(__ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * ldc)), r);
}

// ---------------------------------------------------------------- LayerNorm: one warp per row
// LPR lanes cooperate on one row (32/LPR rows per warp), each lane owning up to 5 vectors of 8 halves: C = 320 -> 8 lanes
// x 5 vectors, 640 -> 16 x 5, 1280 -> 32 x 5, so every lane is busy and five 16-byte loads per lane are in flight.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                 __half* __restrict__ out, long long rows, int C, float eps, const __half* __restrict__ pre_add,
                                 long long rows_per_group, __half* __restrict__ x_out, const __half* __restrict__ pe, int hw, int F) {
  constexpr int MAXV = 5;
  constexpr int RPW = 32 / LPR;   // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsel = lane / LPR;
  const int vecs = C / 8;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  // grid-stride over row groups: a few resident blocks per SM stream the whole tensor (10 368 one-shot blocks at level 0 spent a third of
  // the kernel in block turnover: 4.1 TB/s against the 6.5 TB/s a copy reaches)
  for (long long warp_global = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); warp_global * RPW < rows; warp_global += warps_total) {
  const long long row = warp_global * RPW + rsel;
  const bool row_ok = row < rows;
  float v[MAXV][8];
  float sum = 0.f;
  const __half* add = (pre_add && row_ok) ? pre_add + (row / rows_per_group) * C : nullptr;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = sub + LPR * k;
    if (row_ok && vi < vecs) unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C + vi * 8)), v[k]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = sub + LPR * k;
    if (row_ok && vi < vecs) {
      if (add) {
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(add + vi * 8)), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] += a[j];
        if (x_out) *reinterpret_cast<uint4*>(x_out + row * C + vi * 8) = pack8(v[k]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = sub + LPR * k;
    if (vi < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const __half* pe_row = (pe && row_ok) ? pe + static_cast<long long>((row / hw) % F) * C : nullptr;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = sub + LPR * k;
    if (row_ok && vi < vecs) {
      float gm[8], bt[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + vi * 8)), gm);
      unpack8(__ldg(reinterpret_cast<const uint4*>(beta + vi * 8)), bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = (v[k][j] - mean) * rstd * gm[j] + bt[j];
      if (pe_row) {
        float pv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(pe_row + vi * 8)), pv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] += pv[j];
      }
      *reinterpret_cast<uint4*>(out + row * C + vi * 8) = pack8(v[k]);
    }
  }
  }
}

}  // namespace

// scratch floats needed by launch_groupnorm: final {mean, rstd} + per-slab partials
// The split of a frame's rows into slabs (= the order its statistics are summed in) depends on HW only, never on the number of frames in
// the batch: a forward of one CFG half (a unit of the multi-GPU split) is then bitwise equal to that half of a CFG batch.
// (Measured and rejected: normalising ~56 MB chunks of frames at a time so that the apply pass re-reads from L2 what the statistics pass
// just pulled in -- the smaller launches lose far more bandwidth than the L2 hits return: 135 us -> 250 us at level 0.)
static void gn_plan(int C, int NF, int HW, int num_sms, int* threads, int* slabs, int* rows_per_block) {
  (void)NF;
  (void)num_sms;
  const int vecs = C / 8;
  int t = vecs * (vecs >= 256 ? 1 : 256 / vecs);
  if (t > 1024) t = vecs;
  int sl = 16;
  int rpb = (HW + sl - 1) / sl;
  if (rpb < 32) rpb = 32;
  sl = (HW + rpb - 1) / rpb;
  *threads = t;
  *slabs = sl;
  *rows_per_block = rpb;
}
size_t groupnorm_scratch_floats(int C, int NF, int HW, int groups, int num_sms) {
  int t, sl, rpb;
  gn_plan(C, NF, HW, num_sms, &t, &sl, &rpb);
  return static_cast<size_t>(2) * NF * groups * (1 + sl);
}

int groupnorm_num_launches(int, int, int, int) { return 2; }

cudaError_t launch_groupnorm(const __half* x1, int C1, const __half* x2, int C2, const __half* gamma, const __half* beta,
                             __half* out, int NF, int HW, int groups, float eps, int silu, float* stats, int num_sms,
                             cudaStream_t stream) {
  const int C = C1 + C2;
  if (C % 8 || C1 % 8 || C % groups || ((C / groups) & 1) || groups > 32 * 8) return cudaErrorInvalidValue;
  const int vecs = C / 8;
  if (vecs > 1024) return cudaErrorInvalidValue;
  int threads, slabs, rows_per_block;
  gn_plan(C, NF, HW, num_sms, &threads, &slabs, &rows_per_block);
  if (threads < groups || threads < 32) return cudaErrorInvalidValue;
  float* partial = stats + static_cast<size_t>(2) * NF * groups;
  gn_stats_kernel<<<dim3(slabs, NF), threads, threads * 8 * sizeof(float), stream>>>(x1, C1, x2, C2, HW, groups, rows_per_block, partial);
  // (the {mean, rstd} finalize is folded into the apply kernel: one launch less per GroupNorm, 83 per forward)
  gn_apply_kernel<<<dim3(slabs, NF), threads, groups * sizeof(float2), stream>>>(x1, C1, x2, C2, gamma, beta, out, HW, groups, rows_per_block, silu, partial,
                                                                                 slabs, 1.f / (static_cast<float>(HW) * (C / groups)), eps);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const __half* x, const __half* gamma, const __half* beta, __half* out, long long rows, int C, float eps,
                             const __half* pre_add, long long rows_per_group, __half* x_out, const __half* pe, int hw, int F,
                             cudaStream_t stream) {
  if (C % 8) return cudaErrorInvalidValue;
  const int vecs = C / 8;
  if (vecs > 160) return cudaErrorInvalidValue;
  if (rows_per_group <= 0) rows_per_group = 1;
  if (hw <= 0) hw = 1;
  if (F <= 0) F = 1;
  const int lpr = vecs <= 40 ? 8 : (vecs <= 80 ? 16 : 32);
  const int wpb = 8;
  const long long warps = (rows + (32 / lpr) - 1) / (32 / lpr);
  static const int ln_waves = static_cast<int>(tune_env("HV_LN_BLOCKS_PER_SM", 8));
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  long long blocks = (warps + wpb - 1) / wpb;
  if (ln_waves > 0 && blocks > static_cast<long long>(sms) * ln_waves) blocks = static_cast<long long>(sms) * ln_waves;
  const unsigned grid = static_cast<unsigned>(blocks);
  if (lpr == 8)
    layernorm_kernel<8><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else if (lpr == 16)
    layernorm_kernel<16><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  else
    layernorm_kernel<32><<<grid, wpb * 32, 0, stream>>>(x, gamma, beta, out, rows, C, eps, pre_add, rows_per_group, x_out, pe, hw, F);
  return cudaGetLastError();
}

}  // namespace hv
