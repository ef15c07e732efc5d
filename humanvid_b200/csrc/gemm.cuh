// Host-visible parameter block + launcher of the tcgen05 GEMM / implicit-GEMM convolution kernel.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hv {

enum AMode : int { A_LINEAR = 0, A_CONV3X3 = 1, A_CONV3X3_S2 = 2, A_UPCONV2X2 = 3 };
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

struct GemmEpilogue {
  __half* out = nullptr;           // [rows][ldc] fp16
  int ldc = 0;
  int n_valid = 0;                 // columns actually written (<= N)
  const __half* bias = nullptr;    // [N] (packed order)
  const __half* rowvec = nullptr;  // per-row-group vector added after bias: rowvec[(row / rows_per_group) * rowvec_ld + n]
  int rowvec_ld = 0;
  int rows_per_group = 1;
  const __half* residual = nullptr;  // [rows][ldr]
  int ldr = 0;
  int act = ACT_NONE;
  int geglu = 0;                   // BLOCK_N=256 tile = [128 hidden | 128 gate] -> 128 outputs
  const __half* rowbias = nullptr; // [M]: added to every column of output row m (the "ones" rows of V^T)
  float* gn_stats = nullptr;       // reserved (fused GroupNorm statistics)
  int tma_io = 0;                  // output tile leaves (and the residual tile arrives) through per-warp shared-memory boxes + TMA
};

struct GemmProblem {
  int M = 0, N = 0;                // logical rows / packed columns of D
  int num_k_blocks = 0;            // ceil(K / 64)
  int bst_stages = 0;              // set by launch_gemm: A-ring depth of the B-stationary variant
  int pf_dist = 0;                 // set by launch_gemm: the producer prefetches A boxes this many k-blocks ahead into L2
  int cluster = 0;                 // 2-CTA clusters along N: each CTA loads half of the shared A tile (64-row box) and multicasts it
  int a_mode = A_LINEAR;
  int k_split = 0;                 // A_LINEAR: k-blocks taken from tensor map a0 (rest from a1); 0 = all from a0
  int cin_blocks = 0;              // conv: 64-channel blocks per tap
  int cin = 0;                     // conv s2: channels of the source (parity offset in the folded 2C axis)
  int H = 0, W = 0, NF = 0;        // conv: OUTPUT height / width / frame count
  int bn = 1, bh = 1, bw = 128;    // conv: tile box (bn*bh*bw == 128 or 256 rows per CTA tile)
  int tiles_n = 0, tiles_y = 0, tiles_x = 0;
  int b_par_rows = 0;              // A_UPCONV2X2: rows of the packed weight per output parity (B is [4 parities][b_par_rows][4 * Cin])
  // batched B operand (V^T = Wv * X^T per frame): B is a 3-D map (K, b_rows, b_batch); output columns of batch n
  // start at n * b_out_stride (a multiple of 8) so every frame's token segment is 16-byte aligned for TMA readers.
  int b_batch = 0, b_rows = 0, b_out_stride = 0;
};

// C[M, N] = A[M, K] * B[N, K]^T  (fp16 in, fp32 accumulate in TMEM, fused epilogue, fp16 out).
// a0/a1: TMA maps of the A operand (see make_* helpers in tma.h), b: TMA map of the packed weights [N][K].
// block_n must be 128, 160 or 256 (256 required for geglu); 160 = two exact tiles for the 320-wide layers.
// io_out / io_res (optional, linear A only, m_sub == 1): maps from make_map_2d_io with box 32 rows x gemm_io_box_cols(); when given,
// the epilogue stages each warp's 32-row slice in shared memory and moves it with TMA (full-line stores instead of one
// 16-byte fragment per thread and row).
cudaError_t launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmProblem& p,
                        const GemmEpilogue& e, int block_n, int num_sms, cudaStream_t stream, int m_sub = 1,
                        const CUtensorMap* io_out = nullptr, const CUtensorMap* io_res = nullptr);
inline int gemm_io_box_cols(int block_n, bool geglu) { return geglu ? block_n / 8 : block_n / 4; }
// Experimental (HV_GEMM_CLUSTER=1, default off, not yet validated on hardware): pairs of CTAs with the same m-block share the A
// tile through TMA multicast; the caller must then build the A map(s) with 64-row boxes.  True when this launch would use it.
bool gemm_wants_cluster(int64_t M, int64_t N, int block_n, int m_sub, bool batched_b);

// Picks the (bn, bh, bw) output-tile box with the least padding for an NF x H x W output.
// (rows = 128 or 256 output pixels per CTA tile; every box dimension is a power of two <= 256)
void choose_conv_box(int NF, int H, int W, int* bn, int* bh, int* bw, int rows = 128);

}  // namespace hv
