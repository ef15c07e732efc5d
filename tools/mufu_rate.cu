// MUFU throughput on sm_100a: ex2.approx.ftz.f32 vs ex2.approx.f16x2 (which SASS shows as TWO MUFU.EX2.F16 per PTX op) vs the
// degree-3 polynomial on the FMA pipe.  One CTA per SM, W warps, each lane runs N independent-chain exponentials.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_rate tools/mufu_rate.cu && tools/mufu_rate
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float x[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); x[i] = fmaf(x[i], -0.5f, -0.25f); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, clk, iters);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      const double ops = double(iters) * 8 * warps * 32 * (mode == 1 ? 2 : 1);   // results produced per SM
      printf("%s warps=%2d: %.1f results/clk/SM (%.0f clk)\n", mode == 0 ? "ex2.f32      " : mode == 1 ? "ex2.f16x2    " : "ex2.f32+ffma ", warps, ops / c, c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
