// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N] = A[M, K] * B[N, K]^T, fp16 operands, fp32 accumulation in TMEM, fused epilogue, fp16 output.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  : cp.async.bulk.tensor loads of the A tile (128 rows x 64 k) and the B tile
//                               (BLOCK_N rows x 64 k) into a ring of 128B-swizzled shared-memory stages
//   warp 1      MMA issuer    : one thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16), accumulators
//                               double-buffered in TMEM (2 x BLOCK_N columns); owns TMEM alloc/dealloc
//   warps 2..17 epilogue      : tcgen05.ld of the accumulator (one output row per thread, four warps per TMEM lane
//                               quadrant splitting the columns), bias / time-embedding row vector / activation /
//                               GEGLU / residual in fp32, one fp16 rounding at the store
// The A operand is addressed through TMA only, so a 3x3 convolution over a channels-last (N,H,W,C) tensor is the
// same kernel: k-block kb = (tap, 64-channel block) and the tile of 128 output pixels is a (bn x bh x bw) box whose
// input window is fetched with the box shifted by the tap offset; TMA's out-of-bounds zero fill is the padding.
// Stride-2 convolutions view the same memory as (N, H/2, 2, W/2, 2*C) so each tap is again a dense box.
#include <cstdlib>

#include "gemm.cuh"
#include "ptx.cuh"
#include "tuning.h"

namespace hv {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB per 128-row sub-tile
constexpr int kStagePerWarp = 256;  // per epilogue warp: its <= 64 bias values as fp32 (LDS broadcast instead of an LDG + convert per use)
constexpr int kMaxStages = 8;
constexpr int kMaxSmem = 227 * 1024;
constexpr int GEMM_THREADS = 576;  // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue

// MT = number of 128-row UMMA sub-tiles per CTA tile.  MT = 2 (256 x BLOCK_N tile, both accumulators fed by the same B
// stage) halves the L2 -> shared-memory traffic per FLOP, which is what bounds the 128-row kernel (~58 B/clk/SM of TMA
// fill sustains only ~60 % of the tensor pipe at 48 KB per 128x256x64 k-block); it is used when K is large enough that
// the then single-buffered accumulator's epilogue is a small fraction of the tile (3x3 convs, wide linears).
template <int BLOCK_N, int MT>
struct Cfg {
  static constexpr int kATileBytes = MT * A_TILE_BYTES;
  static constexpr int kBTileBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kAccStages = (2 * MT * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int kAccCols = MT * BLOCK_N;                         // TMEM columns per accumulator stage
  static constexpr int kTmemCols = kAccStages * kAccCols <= 256 ? 256 : 512;
  // per epilogue warp: a 32-row x BLOCK_N/4-column fp16 box through which its slice of the residual arrives and its slice
  // of the output leaves by TMA (128-row tiles only; the 256-row conv tiles are compute-bound and keep direct stores)
  static constexpr int kIoPerWarp = MT == 1 ? 32 * (BLOCK_N / 4) * 2 : 0;
  static constexpr int kIoBytes = 16 * kIoPerWarp;
  static constexpr int kBarBytes = (2 * kMaxStages + 6 + 16) * 8 + 16;  // full[8] empty[8] tfull[2] tempty[2] bpanel[2] res[16] + TMEM slot
  static constexpr int kFixed = kIoBytes + kBarBytes + 16 * kStagePerWarp + 1024;  // + alignment slack
  static constexpr int kStagesFit = (227 * 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  // B-stationary mode: the CTA's whole BLOCK_N x K weight panel stays in shared memory, the ring holds A tiles only
  static constexpr int kBstFixed = kFixed;
};

struct TileCoord {
  int m0;          // linear: first row.   conv: unused
  int n0, y0, x0;  // conv: first frame / output row / output col of the box
  int py, px;      // A_UPCONV2X2: output parity of this tile (output pixel (2y + py, 2x + px) for source pixel (y, x))
};

__device__ __forceinline__ TileCoord tile_coord(const GemmProblem& p, int m_blk, int rows_per_tile) {
  TileCoord t;
  t.py = t.px = 0;
  if (p.a_mode == A_LINEAR) {
    t.m0 = m_blk * rows_per_tile;
    t.n0 = t.y0 = t.x0 = 0;
  } else {
    if (p.a_mode == A_UPCONV2X2) {   // the four parities of one source box are consecutive tiles (the box stays in L2 for all four)
      t.py = (m_blk >> 1) & 1;
      t.px = m_blk & 1;
      m_blk >>= 2;
    }
    int xb = m_blk % p.tiles_x;
    int r = m_blk / p.tiles_x;
    int yb = r % p.tiles_y;
    int nb = r / p.tiles_y;
    t.m0 = 0;
    t.n0 = nb * p.bn;
    t.y0 = yb * p.bh;
    t.x0 = xb * p.bw;
  }
  return t;
}

__device__ __forceinline__ void store8(__half* dst, const float* v) {
  uint4 u;
  u.x = pack_h2(v[0], v[1]);
  u.y = pack_h2(v[2], v[3]);
  u.z = pack_h2(v[4], v[5]);
  u.w = pack_h2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void ldsf8(const float* src, float* v) {  // shared memory, 16-byte aligned
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* src, float* v) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// BST ("B-stationary"): every tile a CTA visits has the same n_blk (the host makes gridDim.x a multiple of n_tiles), so
// the CTA loads its BLOCK_N x K weight panel into shared memory once and streams only A afterwards.  For the K <= 640
// linears of the two high-resolution levels this removes the per-tile re-fetch of B from L2 (as many bytes as A at
// BLOCK_N = 128..256), and L2 -> SM fill is what bounds those GEMMs.
template <int BLOCK_N, int MT, bool BST>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
            const __grid_constant__ CUtensorMap map_res, const GemmProblem p, const GemmEpilogue e) {
  using C = Cfg<BLOCK_N, MT>;
  constexpr int TILE_M = MT * BLOCK_M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nstages = BST ? p.bst_stages : C::kStages;
  constexpr int kRingStageBytes = BST ? C::kATileBytes : C::kStageBytes;
  uint8_t* ring = smem + (BST ? p.num_k_blocks * C::kBTileBytes : 0);
  uint8_t* io_base = ring + nstages * kRingStageBytes;   // 1024-byte aligned (every ring stage is a multiple of 1 KB)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(io_base + C::kIoBytes);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_tfull = bar_empty + kMaxStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_bpanel = bar_tempty + 2;
  uint64_t* bar_res = bar_bpanel + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_res + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = p.a_mode == A_LINEAR ? (p.M + TILE_M - 1) / TILE_M : p.tiles_n * p.tiles_y * p.tiles_x * (p.a_mode == A_UPCONV2X2 ? 4 : 1);
  const int n_per_batch = p.b_batch ? (p.b_rows + BLOCK_N - 1) / BLOCK_N : 0;
  const int n_tiles = p.b_batch ? p.b_batch * n_per_batch : (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int nkb = p.num_k_blocks;

  // 2-CTA cluster along N (p.cluster): the partner CTAs work on tiles t, t+1 of the same m-block (n_tiles and gridDim.x are
  // even), each loads 64 of the A tile's 128 rows and multicasts them into both CTAs' rings -> half the TMA requests per CTA
  // for A.  A ring slot may be refilled only when BOTH CTAs have consumed it: the MMA warps commit to both empty barriers.
  const uint32_t crank = p.cluster ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], p.cluster ? 2 : 1);
    }
    mbar_init(bar_bpanel, 1);
    for (int s = 0; s < 16; ++s) mbar_init(&bar_res[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 16);
    }  // (only stage 0 is used when the accumulator is single-buffered)
    fence_mbar_init();
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    if (e.tma_io) {
      tma_prefetch_desc(&map_out);
      tma_prefetch_desc(&map_res);
    }
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync();   // the partner's barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (BST && static_cast<int>(blockIdx.x) < num_tiles) {
        mbar_arrive_expect_tx(bar_bpanel, nkb * C::kBTileBytes);
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d(smem + kb * C::kBTileBytes, &map_b, bar_bpanel, kb * BLOCK_K, (blockIdx.x % n_tiles) * BLOCK_N);
      }
      // Optional L2 prefetch of the A stream, p.pf_dist k-blocks ahead of the loads (HV_GEMM_PF, default off).  Measured
      // NEGATIVE: +17 % time on the K = 320 linears, +35 % on the convs -- these kernels are bound by the request rate of the
      // SM's TMA unit (~0.45 box rows per clock, 128 B each = the 58 B/clk fill rate), and every prefetch is another request.
      int pf_tile = blockIdx.x, pf_kb = 0;
      auto pf_step = [&]() {
        if (pf_tile >= num_tiles) return;
        const TileCoord pc = tile_coord(p, pf_tile / n_tiles, TILE_M);
        if (p.a_mode == A_LINEAR) {
          if (p.k_split == 0 || pf_kb < p.k_split) tma_prefetch_2d(&map_a0, pf_kb * BLOCK_K, pc.m0);   // the box spans all TILE_M rows
          else tma_prefetch_2d(&map_a1, (pf_kb - p.k_split) * BLOCK_K, pc.m0);
        } else if (p.a_mode == A_CONV3X3) {
          const int tap = pf_kb / p.cin_blocks, cb = pf_kb - tap * p.cin_blocks;
          const int dy = tap / 3, dx = tap - dy * 3;
          tma_prefetch_4d(&map_a0, cb * BLOCK_K, pc.x0 + dx - 1, pc.y0 + dy - 1, pc.n0);
        }
        if (++pf_kb == nkb) { pf_kb = 0; pf_tile += gridDim.x; }
      };
      const bool do_pf = p.pf_dist > 0 && p.a_mode != A_CONV3X3_S2 && p.a_mode != A_UPCONV2X2;
      if (do_pf)
        for (int i = 0; i < p.pf_dist; ++i) pf_step();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const TileCoord tc = tile_coord(p, m_blk, TILE_M);
        for (int kb = 0; kb < nkb; ++kb) {
          if (do_pf) pf_step();
          mbar_wait(&bar_empty[stage], phase ^ 1);
          uint8_t* sa = ring + stage * kRingStageBytes;
          uint8_t* sb = sa + C::kATileBytes;
          mbar_arrive_expect_tx(&bar_full[stage], kRingStageBytes);
          if (p.a_mode == A_LINEAR && p.cluster) {
            uint8_t* half = sa + crank * (A_TILE_BYTES / 2);   // rows 64*crank .. +63 of the tile (SW128 pattern repeats every 8 rows)
            if (p.k_split == 0 || kb < p.k_split)
              tma_load_2d_mc(half, &map_a0, &bar_full[stage], kb * BLOCK_K, tc.m0 + crank * (BLOCK_M / 2), 0x3);
            else
              tma_load_2d_mc(half, &map_a1, &bar_full[stage], (kb - p.k_split) * BLOCK_K, tc.m0 + crank * (BLOCK_M / 2), 0x3);
          } else if (p.a_mode == A_LINEAR) {
            if (p.k_split == 0 || kb < p.k_split)
              tma_load_2d(sa, &map_a0, &bar_full[stage], kb * BLOCK_K, tc.m0);
            else
              tma_load_2d(sa, &map_a1, &bar_full[stage], (kb - p.k_split) * BLOCK_K, tc.m0);
          } else {
            const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            const int dy = tap / 3, dx = tap - dy * 3;
            if (p.a_mode == A_CONV3X3) {
              tma_load_4d(sa, &map_a0, &bar_full[stage], cb * BLOCK_K, tc.x0 + dx - 1, tc.y0 + dy - 1, tc.n0);
            } else if (p.a_mode == A_UPCONV2X2) {
              // nearest-2x upsample folded into the conv: output (2y+py, 2x+px) reads the 2x2 source pixels (y + a + py - 1, x + b + px - 1)
              // with the 3x3 taps that land on each pre-summed in the packed weight (tap = 2a + b)
              const int a = tap >> 1, b = tap & 1;
              tma_load_4d(sa, &map_a0, &bar_full[stage], cb * BLOCK_K, tc.x0 + b + tc.px - 1, tc.y0 + a + tc.py - 1, tc.n0);
            } else {  // stride 2: memory viewed as (N, H/2, 2, W/2, 2C); input row 2*yo+dy-1, col 2*xo+dx-1
              const int ph = dy == 1 ? 0 : 1, pw = dx == 1 ? 0 : 1;
              tma_load_5d(sa, &map_a0, &bar_full[stage], pw * p.cin + cb * BLOCK_K, tc.x0 - (dx == 0 ? 1 : 0), ph,
                          tc.y0 - (dy == 0 ? 1 : 0), tc.n0);
            }
          }
          if (!BST) {
            if (p.b_batch)
              tma_load_3d(sb, &map_b, &bar_full[stage], kb * BLOCK_K, (n_blk % n_per_batch) * BLOCK_N, n_blk / n_per_batch);
            else
              tma_load_2d(sb, &map_b, &bar_full[stage], kb * BLOCK_K, n_blk * BLOCK_N + (tc.py * 2 + tc.px) * p.b_par_rows);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: all 32 lanes run the loop with
    // warp-uniform operands, one elected lane issues (umma_*_w) -- keeps the descriptors in uniform registers
    {
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (BST && static_cast<int>(blockIdx.x) < num_tiles) mbar_wait(bar_bpanel, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * kRingStageBytes);
          const uint64_t bdesc = umma_desc_k_sw128(BST ? smem_u32(smem + kb * C::kBTileBytes) : sa + C::kATileBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = umma_desc_k_sw128(sa + mt * A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_f16_ss_w(d_tmem + mt * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (p.cluster) umma_commit_mc_w(&bar_empty[stage], 0x3);
          else umma_commit_w(&bar_empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&bar_tfull[acc]);
        if (++acc == C::kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps, 1 output row / thread)
    // The tensor pipe retires a 128x256x64 k-block in 512 clocks; for the K = 320..1280 linears of this network the
    // epilogue (fp16 rounding chain, GEGLU, residual) needs more issue slots per tile than the mainloop needs clocks, so
    // it is spread over 16 warps: four per TMEM lane quadrant, each owning a quarter of the tile's columns, in 16-column
    // chunks with the TMEM load of the next chunk and the residual of the next chunk in flight behind the math.
    const int q = warp & 3;             // TMEM lane quadrant this warp may read
    const int cs = (warp - 2) >> 2;     // column slice of the tile handled by this warp (0..3)
    const int r = q * 32 + lane;
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot + 4) + (warp - 2) * kStagePerWarp);
    int acc = 0;
    uint32_t acc_phase = 0, res_phase = 0;
    constexpr int SL = BLOCK_N / 4;     // columns per warp slice: 64 / 40 / 32
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const TileCoord tc = tile_coord(p, m_blk, TILE_M);
      // this warp's slice of the bias vector -> its private shared-memory strip (issued before the wait on the accumulator)
      if (e.bias != nullptr) {
        // lane l holds strip entries 2l, 2l+1: plain = this warp's SL columns; GEGLU = 32 hidden then their 32 gates
        int bcol;
        bool act;
        if (!e.geglu) { bcol = n_blk * BLOCK_N + cs * SL + lane * 2; act = lane * 2 < SL; }
        else { bcol = n_blk * BLOCK_N + cs * (BLOCK_N / 8) + (lane & 15) * 2 + (lane >> 4) * (BLOCK_N / 2); act = true; }
        float2 bv = make_float2(0.f, 0.f);
        if (act && p.b_batch == 0 && bcol < p.N) bv = __half22float2(*reinterpret_cast<const __half2*>(e.bias + bcol));
        __syncwarp();
        if (act) *reinterpret_cast<float2*>(sbias + lane * 2) = bv;
        __syncwarp();
      }
      const bool ts = MT == 1 && e.tma_io != 0;
      int ts_col0 = 0, ts_row0 = 0, ts_batch = 0;
      uint8_t* io = io_base + (warp - 2) * C::kIoPerWarp;
      if (ts) {
        // ---- 128-row linear tiles: this warp's 32 x SL slice goes through its shared-memory box; the residual slice is
        // fetched into the box by TMA while the MMAs run, updated in place (each thread touches only its own row) and
        // the finished box leaves with one TMA store: full 128-byte lines instead of 16 bytes per thread and row.
        constexpr int PITCH = SL * 2;                    // bytes per staged row: 128 / 80 / 64
        const bool has_res = e.residual != nullptr && !e.geglu;
        ts_row0 = tc.m0 + q * 32;
        ts_col0 = e.geglu ? n_blk * (BLOCK_N / 2) + cs * (BLOCK_N / 8) : n_blk * BLOCK_N + cs * SL;
        if (p.b_batch) {                                 // batched B (V^T): column inside the batch item's segment + the item (3-D output map)
          ts_col0 = (n_blk % n_per_batch) * BLOCK_N + cs * SL;
          ts_batch = n_blk / n_per_batch;
        }
        if (lane == 0) {
          tma_store_wait_read();                         // the previous tile's store has drained the box
          if (has_res) {
            mbar_arrive_expect_tx(&bar_res[warp - 2], 32 * PITCH);
            tma_load_2d(io, &map_res, &bar_res[warp - 2], ts_col0, ts_row0);
          }
        }
        __syncwarp();
        mbar_wait(&bar_tfull[acc], acc_phase);
        tc_fence_after();
        const long long row = tc.m0 + r;
        const bool row_ok = row < p.M;
        const __half* __restrict__ rv = (e.rowvec != nullptr && row_ok) ? e.rowvec + (row / e.rows_per_group) * e.rowvec_ld : nullptr;
        const bool has_bias = e.bias != nullptr;
        const float rowbias = (e.rowbias != nullptr && row_ok) ? __half2float(e.rowbias[row]) : 0.f;
        const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols;
        if (!e.geglu) {
          constexpr int NC = (SL + 15) / 16;
          const int colbase = ts_col0;
          const int swz = PITCH == 128 ? (lane & 7) : PITCH == 64 ? ((lane >> 1) & 3) : 0;
          uint8_t* myrow = io + lane * PITCH;
          const bool lean = e.rowvec == nullptr && e.act == ACT_NONE && e.rowbias == nullptr;
          uint32_t raw[2][16];
          auto fetch = [&](int c, int buf) {
            if (c * 16 + 16 <= SL) tmem_ld16(t_acc + cs * SL + c * 16, raw[buf]);
            else tmem_ld8(t_acc + cs * SL + c * 16, raw[buf]);
          };
          fetch(0, 0);
          if (has_res) {
            mbar_wait(&bar_res[warp - 2], res_phase);
            res_phase ^= 1;
          }
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            tmem_ld_wait();
            if (c + 1 < NC) fetch(c + 1, (c + 1) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = colbase + c * 16 + g * 8;
              if (c * 16 + g * 8 < SL && col < e.n_valid) {
                // fp32 all the way: accumulator + bias (+ time-embedding row vector) -> activation -> + residual, ONE rounding
                // at the store (the fp32 oracle is the parity anchor; the reference's eager fp16 path rounds after every op)
                float v[8], t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[c & 1][g * 8 + i]);
                if (has_bias) {
                  ldsf8(sbias + c * 16 + g * 8, t);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += t[i];
                }
                if (!lean) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += rowbias;
                  if (rv != nullptr) {
                    load8(rv + col, t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += t[i];
                  }
                  if (e.act == ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                  } else if (e.act == ACT_SILU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
                  }
                }
                uint4* sp4 = reinterpret_cast<uint4*>(myrow + (((2 * c + g) ^ swz) << 4));
                if (has_res) {
                  const uint4 rr = *sp4;
                  add_h2(v[0], v[1], rr.x);
                  add_h2(v[2], v[3], rr.y);
                  add_h2(v[4], v[5], rr.z);
                  add_h2(v[6], v[7], rr.w);
                }
                *sp4 = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
              }
            }
          }
        } else {
          constexpr int NC = (BLOCK_N / 8) / 16;
          const int hcol0 = cs * (BLOCK_N / 8);
          const int swz = (lane >> 1) & 3;               // 64-byte rows, 64B swizzle
          uint8_t* myrow = io + lane * 64;
          uint32_t hraw[2][16], graw[2][16];
          auto fetch = [&](int c, int buf) {
            tmem_ld16(t_acc + hcol0 + c * 16, hraw[buf]);
            tmem_ld16(t_acc + BLOCK_N / 2 + hcol0 + c * 16, graw[buf]);
          };
          fetch(0, 0);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            tmem_ld_wait();
            if (c + 1 < NC) fetch(c + 1, (c + 1) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float bh[8], bg[8];
              if (has_bias) {
                ldsf8(sbias + c * 16 + g * 8, bh);
                ldsf8(sbias + 32 + c * 16 + g * 8, bg);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) bh[i] = bg[i] = 0.f;
              }
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {   // hidden * gelu(gate) in fp32 (packed f32x2 math), one rounding at the store
                const f32x2 hh = add2(pk2u(hraw[c & 1][g * 8 + 2 * i], hraw[c & 1][g * 8 + 2 * i + 1]), pk2(bh[2 * i], bh[2 * i + 1]));
                const f32x2 gg = add2(pk2u(graw[c & 1][g * 8 + 2 * i], graw[c & 1][g * 8 + 2 * i + 1]), pk2(bg[2 * i], bg[2 * i + 1]));
                o[i] = geglu2(hh, gg);
              }
              *reinterpret_cast<uint4*>(myrow + (((2 * c + g) ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
        fence_proxy_async_smem();
      } else {
      if (e.residual != nullptr && !e.geglu) {
        // pull this thread's residual row segments towards L2 while the tile's MMAs are still running
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int rt = mt * BLOCK_M + r;
          long long prow;
          bool pok;
          if (p.a_mode == A_LINEAR) {
            prow = tc.m0 + rt;
            pok = prow < p.M;
          } else {
            const int ix = rt % p.bw, t2 = rt / p.bw;
            const int iy = t2 % p.bh, in = t2 / p.bh;
            const int n = tc.n0 + in, y = tc.y0 + iy, x = tc.x0 + ix;
            pok = n < p.NF && y < p.H && x < p.W;
            prow = (static_cast<long long>(n) * p.H + y) * p.W + x;
          }
          const int pcol = n_blk * BLOCK_N + cs * SL;
          if (pok && pcol < e.n_valid) {
            const __half* pa = e.residual + prow * e.ldr + pcol;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
            if (SL * 2 > 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 32));   // slices wider than 64 B may straddle another line
          }
        }
      }
      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int rt = mt * BLOCK_M + r;   // row inside the CTA tile
        long long row;
        bool row_ok;
        if (p.a_mode == A_LINEAR) {
          row = tc.m0 + rt;
          row_ok = row < p.M;
        } else {
          const int ix = rt % p.bw, t2 = rt / p.bw;
          const int iy = t2 % p.bh, in = t2 / p.bh;
          const int n = tc.n0 + in, y = tc.y0 + iy, x = tc.x0 + ix;
          row_ok = n < p.NF && y < p.H && x < p.W;
          row = p.a_mode == A_UPCONV2X2 ? (static_cast<long long>(n) * 2 * p.H + 2 * y + tc.py) * (2 * p.W) + 2 * x + tc.px
                                        : (static_cast<long long>(n) * p.H + y) * p.W + x;
        }
        const __half* __restrict__ rv = (e.rowvec != nullptr && row_ok) ? e.rowvec + (row / e.rows_per_group) * e.rowvec_ld : nullptr;
        const __half* __restrict__ res = (e.residual != nullptr && row_ok) ? e.residual + row * e.ldr : nullptr;
        const bool has_bias = e.bias != nullptr;
        const float rowbias = (e.rowbias != nullptr && row_ok) ? __half2float(e.rowbias[row]) : 0.f;
        __half* __restrict__ out = e.out + row * e.ldc;
        const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C::kAccCols + mt * BLOCK_N;

        if (!e.geglu) {
          constexpr int NC = (SL + 15) / 16;     // 16-column chunks (the last one is 8 wide when SL = 40)
          const int tile_col0 = p.b_batch ? (n_blk / n_per_batch) * p.b_out_stride + (n_blk % n_per_batch) * BLOCK_N : n_blk * BLOCK_N;
          const int col_lim = p.b_batch ? (n_blk / n_per_batch) * p.b_out_stride + ((p.b_rows + 7) & ~7) : e.n_valid;
          const int colbase = tile_col0 + cs * SL;
          uint32_t raw[2][16];
          uint4 rres[2][2];
          auto fetch = [&](int c, int buf) {
            if (c * 16 + 16 <= SL) tmem_ld16(t_acc + cs * SL + c * 16, raw[buf]);
            else tmem_ld8(t_acc + cs * SL + c * 16, raw[buf]);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = colbase + c * 16 + g * 8;
              rres[buf][g] = (res != nullptr && c * 16 + g * 8 < SL && col < col_lim) ? __ldg(reinterpret_cast<const uint4*>(res + col)) : make_uint4(0, 0, 0, 0);
            }
          };
          fetch(0, 0);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            tmem_ld_wait();
            if (c + 1 < NC) fetch(c + 1, (c + 1) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = colbase + c * 16 + g * 8;
              if (c * 16 + g * 8 < SL && col < col_lim) {
                float v[8], t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[c & 1][g * 8 + i]) + rowbias;
                if (has_bias) {
                  ldsf8(sbias + c * 16 + g * 8, t);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += t[i];
                }
                if (rv != nullptr) {
                  load8(rv + col, t);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += t[i];
                }
                if (e.act == ACT_RELU) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (e.act == ACT_SILU) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
                }
                if (res != nullptr) {
                  const uint4 rr = rres[c & 1][g];
                  add_h2(v[0], v[1], rr.x);
                  add_h2(v[2], v[3], rr.y);
                  add_h2(v[4], v[5], rr.z);
                  add_h2(v[6], v[7], rr.w);
                }
                if (row_ok) store8(out + col, v);
              }
            }
          }
        } else {
          // GEGLU (BLOCK_N == 256): tile columns [0,128) are "hidden", [128,256) the matching "gate" rows of the packed
          // weight; this warp owns hidden columns [cs*32, cs*32+32) and their gates.
          constexpr int NC = (BLOCK_N / 8) / 16;
          const int hcol0 = cs * (BLOCK_N / 8);
          uint32_t hraw[2][16], graw[2][16];
          auto fetch = [&](int c, int buf) {
            tmem_ld16(t_acc + hcol0 + c * 16, hraw[buf]);
            tmem_ld16(t_acc + BLOCK_N / 2 + hcol0 + c * 16, graw[buf]);
          };
          fetch(0, 0);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            tmem_ld_wait();
            if (c + 1 < NC) fetch(c + 1, (c + 1) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int ocol = n_blk * (BLOCK_N / 2) + hcol0 + c * 16 + g * 8;
              float bh[8], bg[8];
              if (has_bias) {
                ldsf8(sbias + c * 16 + g * 8, bh);
                ldsf8(sbias + 32 + c * 16 + g * 8, bg);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) bh[i] = bg[i] = 0.f;
              }
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {   // hidden * gelu(gate) in fp32 (packed f32x2 math), one rounding at the store
                const f32x2 hh = add2(pk2u(hraw[c & 1][g * 8 + 2 * i], hraw[c & 1][g * 8 + 2 * i + 1]), pk2(bh[2 * i], bh[2 * i + 1]));
                const f32x2 gg = add2(pk2u(graw[c & 1][g * 8 + 2 * i], graw[c & 1][g * 8 + 2 * i + 1]), pk2(bg[2 * i], bg[2 * i + 1]));
                o[i] = geglu2(hh, gg);
              }
              if (row_ok && ocol < e.n_valid) *reinterpret_cast<uint4*>(out + ocol) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }  // sub-tiles
      }  // direct-store path
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_tempty[acc]);
        if (ts) {
          if (p.b_batch) tma_store_3d(&map_out, io, ts_col0, ts_row0, ts_batch);
          else tma_store_2d(&map_out, io, ts_col0, ts_row0);
          tma_store_commit();
        }
      }
      if (++acc == C::kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (MT == 1 && e.tma_io && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync();   // no CTA leaves while its partner may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

}  // namespace

void choose_conv_box(int NF, int H, int W, int* bn, int* bh, int* bw, int rows) {
  long long best = -1;
  int b_n = 1, b_h = 1, b_w = rows;
  for (int w = rows; w >= 1; w >>= 1) {
    for (int h = rows / w; h >= 1; h >>= 1) {
      int n = rows / (w * h);
      auto up = [](int a, int b) { return static_cast<long long>((a + b - 1) / b) * b; };
      long long padded = up(NF, n) * up(H, h) * up(W, w);
      if (best < 0 || padded < best) {  // ties keep the widest box along W (longest contiguous TMA rows)
        best = padded;
        b_n = n;
        b_h = h;
        b_w = w;
      }
    }
  }
  *bn = b_n;
  *bh = b_h;
  *bw = b_w;
}

template <int BN, int MT>
static cudaError_t launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& mo, const CUtensorMap& mr,
                            const GemmProblem& p, const GemmEpilogue& e, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_kernel<BN, MT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, MT>::kSmemBytes);
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  if (p.cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid & ~1, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<BN, MT>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_kernel<BN, MT, false>, a0, a1, b, mo, mr, p, e);
  }
  gemm_kernel<BN, MT, false><<<grid, GEMM_THREADS, Cfg<BN, MT>::kSmemBytes, stream>>>(a0, a1, b, mo, mr, p, e);
  return cudaGetLastError();
}

// B-stationary launch; returns cudaErrorNotSupported when the shape does not qualify (caller falls back to the ring).
template <int BN>
static cudaError_t launch_bst(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& mo, const CUtensorMap& mr,
                              GemmProblem p, const GemmEpilogue& e, int m_tiles, int n_tiles, int num_sms, cudaStream_t stream) {
  using C = Cfg<BN, 1>;
  const int panel = p.num_k_blocks * C::kBTileBytes;
  int stages = (kMaxSmem - C::kBstFixed - panel) / C::kATileBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  const int per_col = num_sms / n_tiles;  // CTAs sharing one n_blk
  if (stages < 3 || per_col < 1 || m_tiles < 2 * per_col) return cudaErrorNotSupported;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_kernel<BN, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  p.bst_stages = stages;
  const int smem = panel + stages * C::kATileBytes + C::kBstFixed;
  if (p.cluster) {   // n_tiles is even, so the grid (a multiple of n_tiles) is too
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(per_col * n_tiles, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_kernel<BN, 1, true>, a0, a1, b, mo, mr, p, e);
  }
  gemm_kernel<BN, 1, true><<<per_col * n_tiles, GEMM_THREADS, smem, stream>>>(a0, a1, b, mo, mr, p, e);
  return cudaGetLastError();
}

bool gemm_wants_cluster(int64_t M, int64_t N, int block_n, int m_sub, bool batched_b) {
  static const int env = static_cast<int>(tune_env("HV_GEMM_CLUSTER", 0));
  if (!env || m_sub != 1 || batched_b) return false;
  const int64_t n_tiles = (N + block_n - 1) / block_n, m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  return (n_tiles % 2) == 0 && m_tiles * n_tiles >= 2 * 148;   // pairs exist and the persistent grid is full
}

cudaError_t launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmProblem& p_in,
                        const GemmEpilogue& e_in, int block_n, int num_sms, cudaStream_t stream, int m_sub, const CUtensorMap* io_out,
                        const CUtensorMap* io_res) {
  static const int pf_env = static_cast<int>(tune_env("HV_GEMM_PF", 0));  // experiment, off: the extra TMA requests cost more than the latency they hide
  GemmProblem p = p_in;
  p.pf_dist = pf_env;
  if (p.cluster && !(m_sub == 1 && p.a_mode == A_LINEAR && !p.b_batch && gemm_wants_cluster(p.M, p.N, block_n, m_sub, false)))
    return cudaErrorInvalidValue;   // the caller built 64-row A boxes for a launch that cannot use them
  GemmEpilogue e = e_in;
  e.tma_io = (io_out != nullptr && m_sub == 1 && p.a_mode == A_LINEAR && (e.residual == nullptr || e.geglu || io_res != nullptr)) ? 1 : 0;
  const CUtensorMap& mo = e.tma_io ? *io_out : a0;
  const CUtensorMap& mr = (e.tma_io && io_res) ? *io_res : mo;
  const int tile_m = BLOCK_M * m_sub;
  const int m_tiles = p.a_mode == A_LINEAR ? (p.M + tile_m - 1) / tile_m : p.tiles_n * p.tiles_y * p.tiles_x * (p.a_mode == A_UPCONV2X2 ? 4 : 1);
  const int n_tiles = p.b_batch ? p.b_batch * ((p.b_rows + block_n - 1) / block_n) : (p.N + block_n - 1) / block_n;
  const int tiles = m_tiles * n_tiles;
  if (tiles <= 0 || p.num_k_blocks <= 0) return cudaErrorInvalidValue;
  if (p.b_batch && (e.bias || e.rowvec || e.residual || e.geglu || (p.b_out_stride % 8))) return cudaErrorInvalidValue;
  if (e.geglu && block_n != 256) return cudaErrorInvalidValue;
  int grid = tiles < num_sms ? tiles : num_sms;
  if (p.cluster) grid &= ~1;
  static const int bst_env = static_cast<int>(tune_env("HV_GEMM_BST", 1));
  if (bst_env && m_sub == 1 && p.a_mode == A_LINEAR && !p.b_batch && n_tiles <= num_sms) {
    cudaError_t r = cudaErrorNotSupported;
    if (block_n == 256) r = launch_bst<256>(a0, a1, b, mo, mr, p, e, m_tiles, n_tiles, num_sms, stream);
    if (block_n == 160) r = launch_bst<160>(a0, a1, b, mo, mr, p, e, m_tiles, n_tiles, num_sms, stream);
    if (block_n == 128) r = launch_bst<128>(a0, a1, b, mo, mr, p, e, m_tiles, n_tiles, num_sms, stream);
    if (r != cudaErrorNotSupported) return r;
  }
  if (m_sub == 1) {
    if (block_n == 256) return launch_t<256, 1>(a0, a1, b, mo, mr, p, e, grid, stream);
    if (block_n == 160) return launch_t<160, 1>(a0, a1, b, mo, mr, p, e, grid, stream);
    if (block_n == 128) return launch_t<128, 1>(a0, a1, b, mo, mr, p, e, grid, stream);
  } else if (m_sub == 2) {
    if (block_n == 256) return launch_t<256, 2>(a0, a1, b, mo, mr, p, e, grid, stream);
    if (block_n == 160) return launch_t<160, 2>(a0, a1, b, mo, mr, p, e, grid, stream);
    if (block_n == 128) return launch_t<128, 2>(a0, a1, b, mo, mr, p, e, grid, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace hv
