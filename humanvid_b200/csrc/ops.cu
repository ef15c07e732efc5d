#include "ops.h"

#include <stdarg.h>

#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

#include "gemm.cuh"
#include "kernels.h"
#include "tma.h"
#include "tuning.h"

namespace hv {

namespace {
thread_local char g_msg[512] = "";
int g_sms = 0;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_msg, sizeof g_msg, fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_msg; }
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return HV_ERR_CUDA;
}

int device_sms() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    if (prop.major != 10) {
      set_error("libhv_b200 needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
      return 0;
    }
    g_sms = prop.multiProcessorCount;
  }
  return g_sms;
}

static void fill_epilogue(GemmEpilogue& e, const hv_epilogue* ep, __half* out, int64_t ldc, int64_t N) {
  e.out = out;
  e.ldc = static_cast<int>(ldc);
  e.n_valid = static_cast<int>(N);
  if (ep) {
    e.bias = static_cast<const __half*>(ep->bias);
    e.rowvec = static_cast<const __half*>(ep->rowvec);
    e.rowvec_ld = ep->rowvec_ld;
    e.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
    e.residual = static_cast<const __half*>(ep->residual);
    e.ldr = ep->ldr;
    e.act = ep->act;
    e.geglu = ep->geglu;
    if (ep->geglu) e.n_valid = static_cast<int>(N / 2);
    if (ep->n_valid > 0) e.n_valid = ep->n_valid;
  }
}

// Tile width: the candidate that pads N the least (N=320 -> 2 x 160, 640 -> 4 x 160, 384 -> 3 x 128, 1280 -> 5 x 256);
// ties go to the wider tile (fewer A re-reads, better UMMA smem ratio) unless that leaves SMs idle.
static int pick_block_n(int64_t N, int64_t m_tiles, bool geglu, int sms) {
  if (geglu) return 256;
  {
    const int v = static_cast<int>(tune_env("HV_GEMM_BN", 0));  // tuning build only
    if (v == 128 || v == 160 || v == 256) return v;
  }
  const int cand[3] = {256, 160, 128};
  int best = 128;
  int64_t best_pad = -1;
  for (int bn : cand) {
    const int64_t padded = (N + bn - 1) / bn * bn;
    if (best_pad < 0 || padded < best_pad) { best_pad = padded; best = bn; }
  }
  if (best != 128 && m_tiles * ((N + best - 1) / best) < sms && m_tiles * ((N + 127) / 128) > m_tiles * ((N + best - 1) / best)) best = 128;
  return best;
}

// 256-row CTA tiles (two UMMA sub-tiles sharing one B stage) when K is large enough to amortise the then
// single-buffered accumulator's epilogue and there is enough work to fill the machine.
static int pick_m_sub(int64_t rows, int64_t N, int bn, int64_t K, int sms) {
  static const int64_t min_k = tune_env("HV_GEMM_MT2_MINK", 2816);  // measured: 1280 loses 24 % on the level-0 FF2 (exposed single-buffered epilogue)
  if (K < min_k) return 1;  // 3x3 convs (K >= 2880) and the widest linears only: below that the exposed epilogue costs more than the L2 traffic saved
  const int64_t tiles2 = ((rows + 255) / 256) * ((N + bn - 1) / bn);
  return tiles2 >= sms ? 2 : 1;
}

int op_gemm(const __half* A, int64_t lda, const __half* A2, int64_t lda2, int64_t K1, const __half* W, __half* out, int64_t ldc,
            int64_t M, int64_t N, int64_t K, const hv_epilogue* ep, cudaStream_t stream) {
  const int sms = device_sms();
  if (!sms) return HV_ERR_CUDA;
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) || (lda % 8) || (ldc % 8) || (N % 8)) {
    set_error("hv_op_gemm: bad shape M=%lld N=%lld K=%lld lda=%lld ldc=%lld", (long long)M, (long long)N, (long long)K, (long long)lda,
              (long long)ldc);
    return HV_ERR_INVALID;
  }
  const bool split = A2 != nullptr && K1 > 0;
  if (split && ((K1 % 64) || (lda2 % 8) || K1 >= K)) {
    set_error("hv_op_gemm: split K1=%lld must be a multiple of 64 and < K", (long long)K1);
    return HV_ERR_INVALID;
  }
  const bool geglu = ep && ep->geglu;
  if (geglu && (N % 256)) {
    set_error("hv_op_gemm: geglu needs N %% 256 == 0 (N=%lld)", (long long)N);
    return HV_ERR_INVALID;
  }
  CUtensorMap ma0, ma1, mb;
  const int64_t Ka = split ? K1 : K;
  const int64_t m_tiles = (M + 127) / 128;
  const int bn = pick_block_n(N, m_tiles, geglu, sms);
  const int m_sub = pick_m_sub(M, N, bn, K, sms);
  const bool cluster = gemm_wants_cluster(M, N, bn, m_sub, false);
  const int a_box_rows = cluster ? 64 : 128 * m_sub;   // clustered launch: each CTA fetches (and multicasts) half of the A tile
  if (!make_map_2d(&ma0, A, M, Ka, lda, a_box_rows)) { set_error("hv_op_gemm A map: %s", tma_last_error()); return HV_ERR_TMA; }
  ma1 = ma0;
  if (split && !make_map_2d(&ma1, A2, M, K - K1, lda2, a_box_rows)) { set_error("hv_op_gemm A2 map: %s", tma_last_error()); return HV_ERR_TMA; }
  if (!make_map_2d(&mb, W, N, K, K, bn)) { set_error("hv_op_gemm W map: %s", tma_last_error()); return HV_ERR_TMA; }
  GemmProblem p;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.num_k_blocks = static_cast<int>((K + 63) / 64);
  p.a_mode = A_LINEAR;
  p.k_split = split ? static_cast<int>(K1 / 64) : 0;
  p.cluster = cluster ? 1 : 0;
  GemmEpilogue e;
  fill_epilogue(e, ep, out, ldc, N);
  // 128-row tiles: the output (and residual) slices move by TMA through per-warp shared-memory boxes
  CUtensorMap mo, mr;
  const CUtensorMap *pmo = nullptr, *pmr = nullptr;
  static const int tma_io_env = static_cast<int>(tune_env("HV_GEMM_TMA_IO", 1));
  if (m_sub == 1 && tma_io_env) {
    const int bc = gemm_io_box_cols(bn, geglu);
    if (!make_map_2d_io(&mo, out, M, e.n_valid, ldc, 32, bc)) { set_error("hv_op_gemm out map: %s", tma_last_error()); return HV_ERR_TMA; }
    pmo = &mo;
    if (e.residual != nullptr && !geglu) {
      if (e.ldr % 8) { set_error("hv_op_gemm: residual ld %d must be a multiple of 8", e.ldr); return HV_ERR_INVALID; }
      if (!make_map_2d_io(&mr, e.residual, M, e.n_valid, e.ldr, 32, bc)) { set_error("hv_op_gemm residual map: %s", tma_last_error()); return HV_ERR_TMA; }
      pmr = &mr;
    }
  }
  cudaError_t err = launch_gemm(ma0, ma1, mb, p, e, bn, sms, stream, m_sub, pmo, pmr);
  if (err != cudaSuccess) return cuda_fail(err, "hv_op_gemm launch");
  return HV_OK;
}

int op_gemm_batched_b(const __half* A, int64_t lda, const __half* X, int64_t ldx, __half* out, int64_t ldc, int64_t M, int64_t batch,
                      int64_t rows, int64_t out_stride, int64_t K, const __half* rowbias, cudaStream_t stream) {
  const int sms = device_sms();
  if (!sms) return HV_ERR_CUDA;
  if (M <= 0 || batch <= 0 || rows <= 0 || (K % 8) || (lda % 8) || (ldx % 8) || (ldc % 8) || (out_stride % 8) || out_stride < rows) {
    set_error("hv_op_gemm_batched_b: bad shape M=%lld batch=%lld rows=%lld stride=%lld K=%lld", (long long)M, (long long)batch, (long long)rows,
              (long long)out_stride, (long long)K);
    return HV_ERR_INVALID;
  }
  CUtensorMap ma, mb;
  if (!make_map_2d(&ma, A, M, K, lda, 128)) { set_error("gemm_batched_b A map: %s", tma_last_error()); return HV_ERR_TMA; }
  const int64_t m_tiles = (M + 127) / 128;
  const int bn = (rows > 128 && m_tiles * batch * ((rows + 255) / 256) >= sms / 2) ? 256 : 128;
  if (!make_map_3d(&mb, X, batch, rows, K, ldx, bn)) { set_error("gemm_batched_b X map: %s", tma_last_error()); return HV_ERR_TMA; }
  GemmProblem p;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(rows);
  p.num_k_blocks = static_cast<int>((K + 63) / 64);
  p.a_mode = A_LINEAR;
  p.b_batch = static_cast<int>(batch);
  p.b_rows = static_cast<int>(rows);
  p.b_out_stride = static_cast<int>(out_stride);
  GemmEpilogue e;
  e.out = out;
  e.ldc = static_cast<int>(ldc);
  e.n_valid = static_cast<int>(rows);
  e.rowbias = rowbias;
  // the output tile leaves through the per-warp shared-memory boxes + TMA (full lines; per-thread 16-byte stores along V^T rows ran this
  // GEMM at 1.8 TB/s): a 3-D map clips every frame's ragged last tile at its own segment end
  CUtensorMap mo;
  const CUtensorMap* pmo = nullptr;
  if (tune_env("HV_GEMM_TMA_IO", 1)) {
    // (extent = rows rounded up to 8: the pad columns of a segment get finite values -- the attention multiplies them by p = 0)
    if (!make_map_3d_io(&mo, out, std::min<int64_t>(out_stride, (rows + 7) / 8 * 8), M, batch, ldc, out_stride, 32, gemm_io_box_cols(bn, false))) { set_error("gemm_batched_b out map: %s", tma_last_error()); return HV_ERR_TMA; }
    pmo = &mo;
  }
  cudaError_t err = launch_gemm(ma, ma, mb, p, e, bn, sms, stream, 1, pmo, nullptr);
  if (err != cudaSuccess) return cuda_fail(err, "hv_op_gemm_batched_b launch");
  return HV_OK;
}

int op_conv3x3(const __half* X, const __half* Wp, __half* out, int64_t ldc, int64_t NF, int64_t H, int64_t W, int64_t Cin,
               int64_t Cout, int stride, const hv_epilogue* ep, cudaStream_t stream) {
  const int sms = device_sms();
  if (!sms) return HV_ERR_CUDA;
  if ((Cin % 64) || (Cout % 8) || (ldc % 8) || (stride != 1 && stride != 2) || NF <= 0 || H <= 0 || W <= 0) {
    set_error("hv_op_conv3x3: bad shape NF=%lld H=%lld W=%lld Cin=%lld Cout=%lld stride=%d", (long long)NF, (long long)H, (long long)W,
              (long long)Cin, (long long)Cout, stride);
    return HV_ERR_INVALID;
  }
  const int64_t Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  GemmProblem p;
  p.a_mode = stride == 1 ? A_CONV3X3 : A_CONV3X3_S2;
  p.N = static_cast<int>(Cout);
  p.cin_blocks = static_cast<int>(Cin / 64);
  p.cin = static_cast<int>(Cin);
  p.num_k_blocks = 9 * p.cin_blocks;
  p.H = static_cast<int>(Ho);
  p.W = static_cast<int>(Wo);
  p.NF = static_cast<int>(NF);
  const int bn_guess = pick_block_n(Cout, (static_cast<int64_t>(NF) * Ho * Wo + 127) / 128, false, sms);
  const int m_sub = pick_m_sub(static_cast<int64_t>(NF) * Ho * Wo, Cout, bn_guess, 9 * Cin, sms);
  choose_conv_box(p.NF, p.H, p.W, &p.bn, &p.bh, &p.bw, 128 * m_sub);
  p.tiles_n = (p.NF + p.bn - 1) / p.bn;
  p.tiles_y = (p.H + p.bh - 1) / p.bh;
  p.tiles_x = (p.W + p.bw - 1) / p.bw;
  p.M = p.NF * p.H * p.W;
  CUtensorMap ma, mb;
  bool ok = stride == 1 ? make_map_nhwc(&ma, X, NF, H, W, Cin, p.bn, p.bh, p.bw) : make_map_nhwc_s2(&ma, X, NF, H, W, Cin, p.bn, p.bh, p.bw);
  if (!ok) { set_error("hv_op_conv3x3 X map: %s", tma_last_error()); return HV_ERR_TMA; }
  const int bn = bn_guess;
  if (!make_map_2d(&mb, Wp, Cout, 9 * Cin, 9 * Cin, bn)) { set_error("hv_op_conv3x3 W map: %s", tma_last_error()); return HV_ERR_TMA; }
  GemmEpilogue e;
  fill_epilogue(e, ep, out, ldc, Cout);
  cudaError_t err = launch_gemm(ma, ma, mb, p, e, bn, sms, stream, m_sub);
  if (err != cudaSuccess) return cuda_fail(err, "hv_op_conv3x3 launch");
  return HV_OK;
}

// Nearest-2x upsample + 3x3 conv (Upsample3D, resnet.py:68-71 + :49) as four 2x2 convolutions of the SOURCE tensor, one per output
// parity: the 4x larger upsampled tensor is never written or read and the conv does 4/9 of the multiply-adds.
int op_upconv2x2(const __half* X, const __half* Wp, __half* out, int64_t ldc, int64_t NF, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                 const hv_epilogue* ep, cudaStream_t stream) {
  const int sms = device_sms();
  if (!sms) return HV_ERR_CUDA;
  if ((Cin % 64) || (Cout % 8) || (ldc % 8) || NF <= 0 || H <= 0 || W <= 0 || (ep && (ep->residual || ep->rowvec || ep->geglu))) {
    set_error("hv_op_upconv2x2: bad shape NF=%lld H=%lld W=%lld Cin=%lld Cout=%lld (or unsupported epilogue)", (long long)NF, (long long)H, (long long)W,
              (long long)Cin, (long long)Cout);
    return HV_ERR_INVALID;
  }
  GemmProblem p;
  p.a_mode = A_UPCONV2X2;
  p.N = static_cast<int>(Cout);
  p.cin_blocks = static_cast<int>(Cin / 64);
  p.cin = static_cast<int>(Cin);
  p.num_k_blocks = 4 * p.cin_blocks;
  p.H = static_cast<int>(H);     // tile space = SOURCE pixels; the epilogue scatters to (2y + py, 2x + px)
  p.W = static_cast<int>(W);
  p.NF = static_cast<int>(NF);
  p.b_par_rows = static_cast<int>(Cout);
  const int64_t rows = NF * H * W;
  const int bn = pick_block_n(Cout, (4 * rows + 127) / 128, false, sms);
  if (Cout % bn) { set_error("hv_op_upconv2x2: Cout=%lld must be a multiple of the %d-column tile", (long long)Cout, bn); return HV_ERR_INVALID; }
  const int m_sub = pick_m_sub(4 * rows, Cout, bn, 4 * Cin, sms);
  choose_conv_box(p.NF, p.H, p.W, &p.bn, &p.bh, &p.bw, 128 * m_sub);
  p.tiles_n = (p.NF + p.bn - 1) / p.bn;
  p.tiles_y = (p.H + p.bh - 1) / p.bh;
  p.tiles_x = (p.W + p.bw - 1) / p.bw;
  p.M = p.NF * p.H * p.W;
  CUtensorMap ma, mb;
  if (!make_map_nhwc(&ma, X, NF, H, W, Cin, p.bn, p.bh, p.bw)) { set_error("hv_op_upconv2x2 X map: %s", tma_last_error()); return HV_ERR_TMA; }
  if (!make_map_2d(&mb, Wp, 4 * Cout, 4 * Cin, 4 * Cin, bn)) { set_error("hv_op_upconv2x2 W map: %s", tma_last_error()); return HV_ERR_TMA; }
  GemmEpilogue e;
  fill_epilogue(e, ep, out, ldc, Cout);
  cudaError_t err = launch_gemm(ma, ma, mb, p, e, bn, sms, stream, m_sub);
  if (err != cudaSuccess) return cuda_fail(err, "hv_op_upconv2x2 launch");
  return HV_OK;
}

}  // namespace hv
