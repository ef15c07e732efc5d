// Microbenchmark: TMA box throughput per SM as a function of the box row width (request-rate vs byte-rate bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I humanvid_b200/csrc -o tools/tma_rate tools/tma_rate.cu humanvid_b200/csrc/tma.cpp -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "ptx.cuh"
using namespace hv;

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  static PFN_cuTensorMapEncodeTiled_v12000 f = nullptr;
  if (!f) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    f = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  return f;
}

// mode 0: loads (global -> smem, mbarrier), mode 1: stores (smem -> global, bulk group)
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap map, int iters, int box_rows, int row_bytes, int rows_total, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
    const int box_bytes = box_rows * row_bytes;
    const int rows_per_cta = rows_total / gridDim.x;
    const int r0 = blockIdx.x * rows_per_cta;
    const long long t0 = clock64();
    uint32_t ph[4] = {0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
      const int s = i & 3;
      const int row = r0 + (i * box_rows) % (rows_per_cta - box_rows);
      if (MODE == 0) {
        if (i >= 4) { mbar_wait(&bar[s], ph[s]); ph[s] ^= 1; }
        mbar_arrive_expect_tx(&bar[s], box_bytes);
        tma_load_2d(smem + s * 32768, &map, &bar[s], 0, row);
      } else {
        tma_store_2d(&map, smem + s * 32768, 0, row);
        tma_store_commit();
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
    }
    if (MODE == 0) {
      for (int s = 0; s < 4 && s < iters; ++s) { mbar_wait(&bar[s], ph[s]); }
    } else {
      tma_store_wait_all();
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

template <int MODE>
void run(int row_bytes, int box_rows, CUtensorMapSwizzle swz, const char* name) {
  const int rows_total = 148 * 4096;
  const size_t ld = 1024;  // bytes between rows in global memory (rows are strided like a wide activation matrix)
  void* d;
  cudaMalloc(&d, size_t(rows_total) * ld);
  cudaMemset(d, 0, size_t(rows_total) * ld);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, (cuuint64_t)rows_total};
  cuuint64_t strides[1] = {ld};
  cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%s rows of %d B: encode failed %d\n", name, row_bytes, (int)r); cudaFree(d); return; }
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int iters = 2000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  k<MODE><<<148, 32, 140 * 1024>>>(m, iters, box_rows, row_bytes, rows_total, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double rows = double(iters) * box_rows;
  printf("%-6s %s box %3d rows x %3d B: %.2f clk/row, %.1f B/clk/SM\n", MODE ? "store" : "load", name, box_rows, row_bytes, avg / rows,
         rows * row_bytes / avg);
  cudaFree(d);
  cudaFree(out);
}

int main() {
  run<0>(128, 128, CU_TENSOR_MAP_SWIZZLE_128B, "sw128");
  run<0>(128, 32, CU_TENSOR_MAP_SWIZZLE_128B, "sw128");
  run<0>(64, 32, CU_TENSOR_MAP_SWIZZLE_64B, "sw64 ");
  run<0>(80, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<0>(160, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<0>(256, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<0>(320, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<0>(512, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<1>(128, 32, CU_TENSOR_MAP_SWIZZLE_128B, "sw128");
  run<1>(64, 32, CU_TENSOR_MAP_SWIZZLE_64B, "sw64 ");
  run<1>(80, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<1>(160, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<1>(320, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  run<1>(512, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "none ");
  return 0;
}
