// Microbenchmark: cost of issuing small tcgen05.mma instructions from one thread (SS and TS operand modes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I humanvid_b200/csrc -o tools/umma_issue tools/umma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace hv;

template <int N, bool TS>
__global__ void k(int count, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint64_t bd = umma_desc_k_sw128(aB) + 2 * (i & 3);
      if (TS) umma_f16_ts(tm, tm + 256 + (i & 7) * 8, bd, idesc, i != 0);
      else umma_f16_ss(tm, umma_desc_k_sw128(aA) + 2 * (i & 3), bd, idesc, i != 0);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// bursts of `burst` MMAs separated by `gap` idle clocks (GAPMODE 0: spin only; 1: commit + wait for the burst, then spin)
template <int N, int GAPMODE>
__global__ void kb(int burst, int nburst, int gap, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 16384);
    long long issue = 0, total = 0;
    uint32_t ph = 0;
    for (int b = 0; b < nburst; ++b) {
      const long long t0 = clock64();
      for (int i = 0; i < burst; ++i)
        umma_f16_ss(tm, umma_desc_k_sw128(aA) + 2 * (i & 3), umma_desc_k_sw128(aB) + 2 * (i & 3), idesc, (b | i) != 0);
      const long long t1 = clock64();
      issue += t1 - t0;
      if (GAPMODE == 1) {
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
        total += clock64() - t0;
      }
      const long long tg = clock64();
      while (clock64() - tg < gap) {}
    }
    if (GAPMODE == 0) { umma_commit(&bar); mbar_wait(&bar, 0); }
    out[blockIdx.x * 2] = issue;
    out[blockIdx.x * 2 + 1] = total;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int GAPMODE>
void runb(int burst, int nburst, int gap) {
  long long* d;
  const int grid = 148;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaFuncSetAttribute(kb<N, GAPMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kb<N, GAPMODE><<<grid, 128, 100 * 1024>>>(burst, nburst, gap, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[2048];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0, b = 0;
  for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  printf("N=%3d bursts of %d, gap %4d clk, %s: issue %.1f clk/burst (%.1f /MMA), issue+commit+wait %.1f clk/burst\n", N, burst, gap,
         GAPMODE ? "commit+wait each burst" : "no wait", a / grid / nburst, a / grid / nburst / burst, b / grid / nburst);
  cudaFree(d);
}

template <int N, bool TS>
void run(int count, int ctas_per_sm) {
  long long* d;
  const int grid = 148 * ctas_per_sm;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, TS><<<grid, 128, 100 * 1024>>>(count, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[2048];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0, b = 0;
  for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  printf("M=128 N=%3d K=16 %s x%d, %d CTA/SM: issue %.1f clk/MMA, issue+drain %.1f clk/MMA (floor %d)\n", N, TS ? "TS" : "SS", count, ctas_per_sm,
         a / grid / count, b / grid / count, 128 * N / 256);
  cudaFree(d);
}

int main() {
  run<64, false>(8, 1); run<64, false>(64, 1); run<64, false>(64, 2);
  run<48, false>(64, 1); run<48, true>(64, 1); run<48, true>(8, 1);
  run<128, false>(64, 1); run<256, false>(64, 1); run<256, false>(4, 1);
  runb<64, 0>(8, 16, 0); runb<64, 0>(8, 16, 200); runb<64, 0>(8, 16, 1000); runb<64, 0>(3, 16, 500); runb<64, 0>(1, 16, 500);
  runb<64, 1>(8, 16, 0); runb<64, 1>(8, 16, 500); runb<64, 1>(3, 16, 500); runb<64, 1>(1, 16, 500); runb<128, 1>(3, 16, 500);
  return 0;
}
