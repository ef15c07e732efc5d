// Spatial multi-head self-attention over the tokens of one frame (reference: diffusers Attention + AttnProcessor2_0
// reached from src/models/attention.py:405-408, and the ReferenceAttentionControl read hook
// src/models/mutual_self_attention.py:147-186 that appends the reference image's tokens as extra keys/values).
//
// Flash-style, tcgen05: one CTA per (128-query tile, head, frame).
//   warp 0      TMA producer : Q tile once, then K / V^T tiles of 128 keys through a shared-memory ring
//   warp 1      MMA issuer   : S = Q K^T (UMMA 128x128xdpad) into TMEM, then PV = P V (UMMA 128 x dpad x 128) into TMEM
//   warps 2..5  softmax      : one query row per thread: double-buffered tcgen05.ld of S, one sweep per key tile
//                              (the max is only checked for growth), p = 2^(s*scale - m), P written to shared memory in the 128B-swizzled K-major layout the
//                              next UMMA reads.  O stays in TMEM across key tiles (UMMA accumulate); V^T carries a row of
//                              ones per head so O's extra column IS the softmax denominator; O is rescaled in TMEM
//                              (tcgen05.ld/st) only when a row max grows by more than 2^8 (lazy rescale).
// Keys come from two segments: the frame's own L tokens and, for frames of the conditional CFG half, the Lb tokens of
// the batch item's reference bank -- the reference instead materialises cat([x, bank.repeat(F)]) per frame and
// recomputes the unconditional half (mutual_self_attention.py:158-186).
// V is consumed transposed (V^T[channel][token], produced directly by a GEMM with swapped operands) so every UMMA
// operand in this file is K-major and shares one shared-memory descriptor format with gemm.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tma.h"
#include "tuning.h"

namespace hv {

namespace {

constexpr int QT = 128;   // queries per CTA
constexpr int KT = 128;   // keys per tile
constexpr float kRescaleThreshold = 8.0f;  // log2 units: O is rescaled only when the row max grows by more than 2^8

// PT ("P in TMEM"): the probabilities never touch shared memory -- the softmax threads store them (packed fp16) into 64
// TMEM columns and the P V MMA takes its A operand from there.  That frees 32 KB of shared memory (a third K/V stage), the
// 32 KB of st.shared + the proxy fence per key tile, and the 32 KB the P V MMAs used to read back.  Needs 128 (S) + 64 (P)
// + dv (O) <= 256 TMEM columns to keep two CTAs per SM, i.e. head dims up to 47.
template <int D, bool PT = false>
struct ACfg {
  static constexpr int kDpad = (D + 15) / 16 * 16;             // Q/K head stride (zero padded)
  static constexpr int kDv = (D + 1 + 15) / 16 * 16;           // V^T rows per head: d values, one row of ones, zero pad
  static constexpr int kKC = (kDpad + 63) / 64;               // 64-column chunks of Q / K rows
  static constexpr int kQBytes = kKC * QT * 128;
  static constexpr int kKBytes = kKC * KT * 128;
  static constexpr int kVChunk = kDv * 128;                    // [dv rows][64 keys]
  static constexpr int kVBytes = 2 * kVChunk;
  static constexpr int kPBytes = PT ? 0 : 2 * QT * 128;        // two 64-key chunks
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kMisc = 256 + 2048 + 1024;               // barriers, row-max exchange, alignment
  static constexpr int kStages = (PT && 2 * (kQBytes + 3 * kStageBytes + kMisc) <= 227 * 1024) ? 3
                                 : (kQBytes + kPBytes + 2 * kStageBytes + 2048 <= 227 * 1024) ? 2 : 1;
  static constexpr int kSmemBytes = kQBytes + kPBytes + kStages * kStageBytes + kMisc;
  static constexpr int kTmemCols = (128 + (PT ? 64 : 0) + kDv) <= 256 ? 256 : 512;
  static constexpr int kMinBlocks = (2 * kSmemBytes <= 227 * 1024 && kTmemCols == 256) ? 2 : 1;
};

struct AttnKernelArgs {
  __half* out;
  long long ldo;
  int L, Lb, F, nf_nobank, heads;
  int vt_stride, vbt_stride;
  float scale_log2;
  long long* dbg;   // PE == 31 only: per-CTA accumulated clock64() spans of the pipeline phases (16 slots per CTA)
};

// ---------------------------------------------------------------------------------------------------------------------
// Variant with EIGHT softmax warps (two threads per query row, one per 64-key half of the tile).  ncu on the 4-warp
// kernel showed the XU (ex2) pipe only ~50 % busy and ~0.3 IPC per scheduler: with one softmax warp per scheduler and
// CTA there is too little to overlap the fixed-latency chains (FFMA -> MUFU -> F2FP -> STS).  Here each scheduler holds
// two softmax warps per CTA (four with two CTAs per SM); the two threads of a row agree on the row max through shared
// memory and a 256-thread named barrier once per key tile.
constexpr int ATT8_THREADS = 320;

// PE > 0: every PE-th pair of exponentials of a full 32-key chunk is evaluated on the FMA pipe (exp2_poly3) instead of MUFU.
template <int D, int PE, bool PT>
__global__ void __launch_bounds__(ATT8_THREADS, ACfg<D, PT>::kMinBlocks)
attn_kernel8(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
             const __grid_constant__ CUtensorMap map_vt, const __grid_constant__ CUtensorMap map_kb,
             const __grid_constant__ CUtensorMap map_vbt, const AttnKernelArgs a) {
  using C = ACfg<D, PT>;
  auto WAIT = [](uint64_t* bar, uint32_t parity) {
    if constexpr (PE == 21) mbar_wait_nohint(bar, parity);
    else if constexpr (PE == 22) mbar_wait_poll(bar, parity);
    else mbar_wait(bar, parity);
  };
  constexpr int DP = C::kDpad, DV = C::kDv;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + C::kQBytes;
  uint8_t* sKV = sP + C::kPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + C::kStages * C::kStageBytes);
  uint64_t* bar_q = bars;
  uint64_t* bar_p = bars + 1;
  uint64_t* bar_pv = bars + 2;
  uint64_t* bar_s = bars + 3;        // [2]
  uint64_t* bar_sfree = bars + 5;    // [2]
  // K and V tiles travel through separate rings: a K slot is free again as soon as its S = Q K^T has been computed (one
  // key tile before the P V of the same tile), so K(j+2) is in flight a full tile earlier than a combined stage allows
  uint64_t* bar_k_full = bars + 7;
  uint64_t* bar_k_empty = bars + 7 + C::kStages;
  uint64_t* bar_v_full = bars + 7 + 2 * C::kStages;
  uint64_t* bar_v_empty = bars + 7 + 3 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7 + 4 * C::kStages);
  float* smax = reinterpret_cast<float*>(bars + 32);   // [2 (tile parity)][2 (half)][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int n = blockIdx.z;
  const bool use_bank = a.Lb > 0 && n >= a.nf_nobank;
  const int Ts = (a.L + KT - 1) / KT;
  const int T = Ts + (use_bank ? (a.Lb + KT - 1) / KT : 0);
  const int bidx = n / a.F;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_pv, 1);
    for (int hh = 0; hh < 2; ++hh) {
      mbar_init(&bar_s[hh], 1);
      mbar_init(&bar_sfree[hh], 128);
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&bar_k_full[s], 1);
      mbar_init(&bar_k_empty[s], 1);
      mbar_init(&bar_v_full[s], 1);
      mbar_init(&bar_v_empty[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_p = tmem_base + 128;                  // PT: 64 columns of packed fp16 probabilities (128 keys)
  const uint32_t tmem_o = tmem_base + (PT ? 192 : 128);

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, C::kQBytes);
#pragma unroll
      for (int kc = 0; kc < C::kKC; ++kc) tma_load_2d(sQ + kc * QT * 128, &map_q, bar_q, h * DP + kc * 64, n * a.L + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        uint8_t* sK = sKV + stage * C::kStageBytes;
        uint8_t* sV = sK + C::kKBytes;
        const bool self = j < Ts;
        const CUtensorMap* mk = self ? &map_k : &map_kb;
        const CUtensorMap* mv = self ? &map_vt : &map_vbt;
        const int tok = self ? n * a.L + j * KT : bidx * a.Lb + (j - Ts) * KT;
        const int vcol = self ? n * a.vt_stride + j * KT : bidx * a.vbt_stride + (j - Ts) * KT;
        mbar_wait(&bar_k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bar_k_full[stage], C::kKBytes);
#pragma unroll
        for (int kc = 0; kc < C::kKC; ++kc) tma_load_2d(sK + kc * KT * 128, mk, &bar_k_full[stage], h * DP + kc * 64, tok);
        mbar_wait(&bar_v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bar_v_full[stage], C::kVBytes);
        tma_load_2d(sV, mv, &bar_v_full[stage], vcol, h * DV);
        tma_load_2d(sV + C::kVChunk, mv, &bar_v_full[stage], vcol + 64, h * DV);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {   // all 32 lanes, warp-uniform operands, one elected lane issues (see umma_*_w)
      constexpr uint32_t idesc_pv = umma_idesc_f16(QT, DV);
      constexpr uint32_t idesc_sh = umma_idesc_f16(QT, KT / 2);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_s_half = [&](int stage, int hh) {
        const uint32_t aK = smem_u32(sKV + stage * C::kStageBytes) + hh * (KT / 2) * 128;
#pragma unroll
        for (int ks = 0; ks < DP / 16; ++ks) {
          const uint64_t ad = umma_desc_k_sw128(aQ + (ks / 4) * QT * 128) + 2 * (ks % 4);
          const uint64_t bd = umma_desc_k_sw128(aK + (ks / 4) * KT * 128) + 2 * (ks % 4);
          umma_f16_ss_w(tmem_s + hh * (KT / 2), ad, bd, idesc_sh, ks != 0 ? 1u : 0u);
        }
        umma_commit_w(&bar_s[hh]);
      };
      auto issue_pv = [&](int stage, int j) {
        const uint32_t aV = smem_u32(sKV + stage * C::kStageBytes + C::kKBytes);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t bd = umma_desc_k_sw128(aV + (kk / 4) * C::kVChunk) + 2 * (kk % 4);
          if constexpr (PT) {
            umma_f16_ts_w(tmem_o, tmem_p + kk * 8, bd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          } else {
            const uint64_t ad = umma_desc_k_sw128(aP + (kk / 4) * QT * 128) + 2 * (kk % 4);
            umma_f16_ss_w(tmem_o, ad, bd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit_w(bar_pv);
        umma_commit_w(&bar_v_empty[stage]);
      };
      int stage = 0;
      uint32_t phase = 0;
      constexpr bool kProf = PE == 31;
      long long pt[6] = {0, 0, 0, 0, 0, 0}, pc = 0;
      auto tick = [&](int slot) {
        if constexpr (kProf) {
          const long long now = clock64();
          pt[slot] += now - pc;
          pc = now;
        }
      };
      WAIT(bar_q, 0);
      WAIT(&bar_k_full[0], 0);
      tc_fence_after();
      issue_s_half(0, 0);
      issue_s_half(0, 1);
      umma_commit_w(&bar_k_empty[0]);
      if constexpr (kProf) pc = clock64();
      for (int j = 0; j < T; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == C::kStages) { nstage = 0; nphase ^= 1; }
        if constexpr (C::kStages >= 2) {
          if (j + 1 < T) {
            WAIT(&bar_k_full[nstage], nphase);
            tick(0);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              WAIT(&bar_sfree[hh], j & 1);
              tc_fence_after();
              tick(1 + hh);
              issue_s_half(nstage, hh);
              tick(3);
            }
            umma_commit_w(&bar_k_empty[nstage]);
          }
          WAIT(bar_p, j & 1);
          WAIT(&bar_v_full[stage], phase);
          tc_fence_after();
          tick(4);
          issue_pv(stage, j);
          tick(5);
        } else {
          WAIT(bar_p, j & 1);
          WAIT(&bar_v_full[stage], phase);
          tc_fence_after();
          issue_pv(stage, j);
          if (j + 1 < T) {
            WAIT(&bar_k_full[nstage], nphase);
            tc_fence_after();
            issue_s_half(nstage, 0);
            issue_s_half(nstage, 1);
            umma_commit_w(&bar_k_empty[nstage]);
          }
        }
        stage = nstage;
        phase = nphase;
      }
      if constexpr (kProf) if (lane == 0) {
        long long* d = a.dbg + (static_cast<long long>(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 24 + 12;
        for (int i = 0; i < 6; ++i) d[i] = pt[i];
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: two threads per query row
    const int qd = warp & 3;
    const int hf = (warp - 2) >> 2;               // which 64-key half of every tile this thread owns
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t my_s = tmem_s + lane_off + hf * 64;
    float m_used = -INFINITY;
    uint8_t* prow = sP + hf * (QT * 128) + r * 128;   // P chunk hf, row r
    const int sw = r & 7;
    const float sc = a.scale_log2;

    constexpr bool kProf = PE == 31;
    long long pt[7] = {0, 0, 0, 0, 0, 0, 0}, pc = 0;
    auto tick = [&](int slot) {
      if constexpr (kProf) {
        const long long now = clock64();
        pt[slot] += now - pc;
        pc = now;
      }
    };
    if constexpr (kProf) pc = clock64();
    for (int j = 0; j < T; ++j) {
      const bool self = j < Ts;
      const int kv_valid = (self ? min(KT, a.L - j * KT) : min(KT, a.Lb - (j - Ts) * KT)) - hf * 64;   // valid keys in my half (may be <= 0)
      WAIT(&bar_s[hf], j & 1);
      tc_fence_after();
      tick(0);
      // pass 1: my half's row max
      float mx0 = -INFINITY, mx1 = -INFINITY;
      constexpr bool kDbgNoPass1 = PE == 11 || PE == 12;   // timing experiments only (results wrong)
      constexpr bool kDbgNoSync = PE == 12;
      constexpr bool kDbgNoExp = PE == 13;
      if (kDbgNoPass1) mx0 = 0.f;
#pragma unroll
      for (int c = 0; c < (kDbgNoPass1 ? 0 : 2); ++c) {
        uint32_t raw[32];
        tmem_ld32(my_s + c * 32, raw);
        tmem_ld_wait();
        if ((c + 1) * 32 <= kv_valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(raw[i]));
            mx1 = fmaxf(mx1, __uint_as_float(raw[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kv_valid) mx0 = fmaxf(mx0, __uint_as_float(raw[i]));
        }
      }
      tick(1);
      float* sm = smax + (j & 1) * 256;
      sm[hf * 128 + r] = fmaxf(mx0, mx1) * sc;
      if (!kDbgNoSync) asm volatile("bar.sync 1, 256;" ::: "memory");
      tick(2);
      const float rowmax = fmaxf(sm[r], sm[128 + r]);
      const bool grow = rowmax > m_used + kRescaleThreshold;   // identical in both threads of the row
      const float m_new = grow ? rowmax : m_used;
      if (j > 0) {
        WAIT(bar_pv, (j - 1) & 1);   // previous P V done: P smem and O are ours again
        tc_fence_after();
        if (hf == 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_used - m_new) : 1.0f;
#pragma unroll
          for (int c = 0; c < DV / 16; ++c) {
            uint32_t t16[16];
            tmem_ld16(tmem_o + lane_off + c * 16, t16);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = __float_as_uint(__uint_as_float(t16[i]) * alpha);
            tmem_st16(tmem_o + lane_off + c * 16, t16);
          }
          tmem_st_wait();
        }
      }
      m_used = m_new;
      tick(3);
      // pass 2: p = 2^(s*scale - m) -> fp16 -> my 64-key chunk of P
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(my_s + c * 32, raw);
        tmem_ld_wait();
        if (c == 1) {   // all my reads of this S half are done
          tc_fence_before();
          mbar_arrive(&bar_sfree[hf]);
        }
        uint32_t pk[16];
        if ((c + 1) * 32 <= kv_valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0, a1;   // s * scale - m for two keys in one FFMA2 (bit-identical to two scalar FMAs)
            upk2(fma2(pk2u(raw[2 * i], raw[2 * i + 1]), pk2(sc, sc), pk2(-m_used, -m_used)), a0, a1);
            if (kDbgNoExp) pk[i] = pack_h2(a0, a1);
            else if (PE > 0 && PE < 10 && (i % (PE > 0 ? PE : 1)) == PE - 1) {   // this pair's exponentials on the FMA pipe (packed fp32x2)
              float e0, e1;
              exp2_poly3_x2(a0, a1, e0, e1);
              pk[i] = pack_h2(e0, e1);
            }
            else pk[i] = pack_h2(fast_exp2(a0), fast_exp2(a1));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = c * 32 + 2 * i;
            const float p0 = c0 < kv_valid ? fast_exp2(fmaf(__uint_as_float(raw[2 * i]), sc, -m_used)) : 0.f;
            const float p1 = c0 + 1 < kv_valid ? fast_exp2(fmaf(__uint_as_float(raw[2 * i + 1]), sc, -m_used)) : 0.f;
            pk[i] = pack_h2(p0, p1);
          }
        }
        if constexpr (PT) {
          tmem_st16(tmem_p + lane_off + hf * 32 + c * 16, pk);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(prow + (((c * 4 + u) ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      tick(4);
      if constexpr (PT) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_p);
      tick(5);
    }
    if constexpr (kProf) {
      if (threadIdx.x == 64 || threadIdx.x == 64 + 128) {   // warp 2 (half 0) and warp 6 (half 1), lane 0
        long long* d = a.dbg + (static_cast<long long>(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 24 + (hf ? 6 : 0);
        for (int i = 0; i < 6; ++i) d[i] = pt[i];
      }
    }
    // epilogue: O / denominator; the two threads of a row write alternate 8-column groups
    WAIT(bar_pv, (T - 1) & 1);
    tc_fence_after();
    float o[DV];
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t raw16[16];
      tmem_ld16(tmem_o + lane_off + c * 16, raw16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[c * 16 + i] = __uint_as_float(raw16[i]);
    }
    tc_fence_before();
    if (q0 + r < a.L) {
      const float inv = 1.f / o[D];
      __half* dst = a.out + (static_cast<long long>(n) * a.L + q0 + r) * a.ldo + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        if ((c & 1) == hf) {
          uint4 u;
          u.x = pack_h2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
          u.y = pack_h2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
          u.z = pack_h2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
          u.w = pack_h2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + c * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// "Ping-pong" variant for head dims < 48: ONE CTA per SM works on TWO 128-query tiles (groups A and B) of the same
// (frame, head) against one shared K/V stream.  What the per-phase clocks of attn_kernel8 showed: a tcgen05.mma of this size
// costs ~50 clocks to issue and ~300 to drain whatever its N, so S = Q K^T (6 MMAs) and P V (8 MMAs) put ~1100 clocks of
// pure latency between the end of one tile's exponentials and the start of the next tile's; two independent CTAs per SM
// fall into lockstep (both in the MUFU phase, then both waiting) instead of filling each other's gaps.  Here the two
// groups are forced out of phase: a token (two named barriers) lets only one group at a time into its MUFU phase, so one
// group's max / S / P V latency always hides behind the other group's exponentials, and the K/V tiles are fetched once
// for 256 queries.  TMEM (512 columns): per group 128 S + 64 P (packed fp16, A operand of the P V MMA) + dv O.
constexpr int ATTPP_THREADS = 576;   // warp 0 TMA, warp 1 MMA, warps 2-9 softmax A, warps 10-17 softmax B

template <int D>
struct PPCfg {
  static constexpr int kDpad = (D + 15) / 16 * 16;
  static constexpr int kDv = (D + 1 + 15) / 16 * 16;
  static constexpr int kKC = (kDpad + 63) / 64;
  static constexpr int kQBytes = kKC * QT * 128;                // one 128-query tile
  static constexpr int kKBytes = kKC * KT * 128;
  static constexpr int kVChunk = kDv * 128;
  static constexpr int kVBytes = 2 * kVChunk;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kMisc = kBarBytes + 4096 + 1024;          // barriers, row-max exchange (2 groups), alignment
  static constexpr int kStagesFit = (227 * 1024 - 2 * kQBytes - kMisc) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = 2 * kQBytes + kStages * kStageBytes + kMisc;
  static constexpr bool kFits = 192 + kDv <= 256 && kStages >= 2;
};

// MW = 2: one MMA-issuing warp PER GROUP (warp 1 -> group A, warp 18 -> group B).  With a single issuer the loop is in order
// (S_A, P V_A, S_B, P V_B): group B's next S = Q K^T cannot be issued before group A's probabilities arrive, which locks the two groups
// into the same phase (both in the MUFU phase, then both waiting -- ncu: XU pipe 69 % busy, B waits ~1.7x longer for S than A).
template <int D, int PE, bool TOKEN, int MW = 1>
__global__ void __launch_bounds__(ATTPP_THREADS + 32 * (MW - 1), 1)   // 18 (19) warps -> 5 on a scheduler -> 16K / (5 * 32) = 102 -> 96 registers
attn_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_vt, const __grid_constant__ CUtensorMap map_kb,
               const __grid_constant__ CUtensorMap map_vbt, const AttnKernelArgs a) {
  using C = PPCfg<D>;
  constexpr int DP = C::kDpad, DV = C::kDv, NS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 groups][kQBytes]
  uint8_t* sKV = sQ + 2 * C::kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NS * C::kStageBytes);
  uint64_t* bar_q = bars;
  uint64_t* bar_s = bars + 1;        // [2]  S of group g is in TMEM
  uint64_t* bar_sfree = bars + 3;    // [2]  all 256 threads of group g have read S
  uint64_t* bar_p = bars + 5;        // [2]  P of group g is in TMEM
  uint64_t* bar_pv = bars + 7;       // [2]  P V of group g has been accumulated
  uint64_t* bar_k_full = bars + 9;
  uint64_t* bar_k_empty = bars + 9 + NS;
  uint64_t* bar_v_full = bars + 9 + 2 * NS;
  uint64_t* bar_v_empty = bars + 9 + 3 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9 + 4 * NS);
  float* smax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);   // [2 groups][2 parity][2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * QT;
  const int h = blockIdx.y;
  const int n = blockIdx.z;
  const bool use_bank = a.Lb > 0 && n >= a.nf_nobank;
  const int Ts = (a.L + KT - 1) / KT;
  const int T = Ts + (use_bank ? (a.Lb + KT - 1) / KT : 0);
  const int bidx = n / a.F;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar_s[g], 1);
      mbar_init(&bar_sfree[g], 256);
      mbar_init(&bar_p[g], 256);
      mbar_init(&bar_pv[g], 1);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&bar_k_full[s], 1);
      mbar_init(&bar_k_empty[s], MW);   // every MMA issuer releases the K / V slot once ITS MMAs on it have completed
      mbar_init(&bar_v_full[s], 1);
      mbar_init(&bar_v_empty[s], MW);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // group g: S at +256g (128 columns), P at +256g+128 (64), O at +256g+192 (DV)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, 2 * C::kQBytes);
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int kc = 0; kc < C::kKC; ++kc)
          tma_load_2d(sQ + g * C::kQBytes + kc * QT * 128, &map_q, bar_q, h * DP + kc * 64, n * a.L + q0 + g * QT);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        uint8_t* sK = sKV + stage * C::kStageBytes;
        uint8_t* sV = sK + C::kKBytes;
        const bool self = j < Ts;
        const CUtensorMap* mk = self ? &map_k : &map_kb;
        const CUtensorMap* mv = self ? &map_vt : &map_vbt;
        const int tok = self ? n * a.L + j * KT : bidx * a.Lb + (j - Ts) * KT;
        const int vcol = self ? n * a.vt_stride + j * KT : bidx * a.vbt_stride + (j - Ts) * KT;
        mbar_wait(&bar_k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bar_k_full[stage], C::kKBytes);
#pragma unroll
        for (int kc = 0; kc < C::kKC; ++kc) tma_load_2d(sK + kc * KT * 128, mk, &bar_k_full[stage], h * DP + kc * 64, tok);
        mbar_wait(&bar_v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bar_v_full[stage], C::kVBytes);
        tma_load_2d(sV, mv, &bar_v_full[stage], vcol, h * DV);
        tma_load_2d(sV + C::kVChunk, mv, &bar_v_full[stage], vcol + 64, h * DV);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 18) {
    {   // all 32 lanes run the issue loop with warp-uniform operands; one elected lane issues (see umma_*_w)
      constexpr uint32_t idesc_pv = umma_idesc_f16(QT, DV);
      constexpr uint32_t idesc_s = umma_idesc_f16(QT, KT);
      auto issue_s = [&](int g, int stage) {
        const uint32_t aQ = smem_u32(sQ + g * C::kQBytes);
        const uint32_t aK = smem_u32(sKV + stage * C::kStageBytes);
#pragma unroll
        for (int ks = 0; ks < DP / 16; ++ks) {
          const uint64_t ad = umma_desc_k_sw128(aQ + (ks / 4) * QT * 128) + 2 * (ks % 4);
          const uint64_t bd = umma_desc_k_sw128(aK + (ks / 4) * KT * 128) + 2 * (ks % 4);
          umma_f16_ss_w(tmem_base + g * 256, ad, bd, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit_w(&bar_s[g]);
      };
      auto issue_pv = [&](int g, int stage, int j) {
        const uint32_t aV = smem_u32(sKV + stage * C::kStageBytes + C::kKBytes);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t bd = umma_desc_k_sw128(aV + (kk / 4) * C::kVChunk) + 2 * (kk % 4);
          umma_f16_ts_w(tmem_base + g * 256 + 192, tmem_base + g * 256 + 128 + kk * 8, bd, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit_w(&bar_pv[g]);
      };
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k_full[0], 0);
      tc_fence_after();
      if (MW == 2) {
        // ---- one issuer per group: S(j+1) goes out as soon as the group has read S(j); P V(j) as soon as its P(j) is in TMEM
        const int g = warp == 1 ? 0 : 1;
        issue_s(g, 0);
        umma_commit_w(&bar_k_empty[0]);
        for (int j = 0; j < T; ++j) {
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == NS) { nstage = 0; nphase ^= 1; }
          if (j + 1 < T) {
            mbar_wait(&bar_k_full[nstage], nphase);
            mbar_wait(&bar_sfree[g], j & 1);
            tc_fence_after();
            issue_s(g, nstage);
            umma_commit_w(&bar_k_empty[nstage]);
          }
          mbar_wait(&bar_p[g], j & 1);
          mbar_wait(&bar_v_full[stage], phase);
          tc_fence_after();
          issue_pv(g, stage, j);
          umma_commit_w(&bar_v_empty[stage]);
          stage = nstage;
          phase = nphase;
        }
      } else {
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit_w(&bar_k_empty[0]);
      for (int j = 0; j < T; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == NS) { nstage = 0; nphase ^= 1; }
        const bool more = j + 1 < T;
        if (more) mbar_wait(&bar_k_full[nstage], nphase);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (more) {
            mbar_wait(&bar_sfree[g], j & 1);
            tc_fence_after();
            issue_s(g, nstage);
            if (g == 1) umma_commit_w(&bar_k_empty[nstage]);
          }
          mbar_wait(&bar_p[g], j & 1);
          if (g == 0) mbar_wait(&bar_v_full[stage], phase);
          tc_fence_after();
          issue_pv(g, stage, j);
          if (g == 1) umma_commit_w(&bar_v_empty[stage]);
        }
        stage = nstage;
        phase = nphase;
      }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: group g, two threads per query row
    const int g = (warp - 2) >> 3;
    const int w = (warp - 2) & 7;
    const int qd = warp & 3;                        // TMEM lane quadrant of this warp
    const int hf = w >> 2;                          // which 64-key half of every tile this thread owns
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 256, tmem_p = tmem_s + 128, tmem_o = tmem_s + 192;
    const uint32_t my_s = tmem_s + lane_off + hf * 64;
    float* gmax = smax + g * 512;
    float m_used = -INFINITY;
    const float sc = a.scale_log2;
    // token: barrier 3 = "A may run its MUFU phase", barrier 4 = "B may"; B hands A the first turn
    if (TOKEN && g == 1) asm volatile("bar.arrive 3, 512;" ::: "memory");
    constexpr bool kProf = PE == 31;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
    auto tick = [&](int slot) {
      if constexpr (kProf) {
        const long long now = clock64();
        pt[slot] += now - pc;
        pc = now;
      }
    };
    if constexpr (kProf) pc = clock64();

    for (int j = 0; j < T; ++j) {
      const bool self = j < Ts;
      const int kv_valid = (self ? min(KT, a.L - j * KT) : min(KT, a.Lb - (j - Ts) * KT)) - hf * 64;
      mbar_wait(&bar_s[g], j & 1);
      tc_fence_after();
      tick(0);
      // pass 1: my half's row max.  The second 32-key chunk stays in registers for pass 2; the first is re-read from
      // TMEM, the load being issued here so that its latency hides behind the row-max exchange.
      float mx0 = -INFINITY, mx1 = -INFINITY;
      uint32_t raw0[32], raw1[32];
      tmem_ld32(my_s, raw0);
      tmem_ld32(my_s + 32, raw1);
      tmem_ld_wait();
      if (64 <= kv_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(raw0[i]), __uint_as_float(raw1[i])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(raw0[i + 1]), __uint_as_float(raw1[i + 1])));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < kv_valid) mx0 = fmaxf(mx0, __uint_as_float(raw0[i]));
          if (32 + i < kv_valid) mx1 = fmaxf(mx1, __uint_as_float(raw1[i]));
        }
      }
      tick(1);
      tmem_ld32(my_s, raw0);     // (re-)fetch chunk 0 for pass 2; completes during the exchange below
      // (measured alternatives, both slower on the same B200: computing the whole row's max in both threads of a row instead of this exchange
      // + 256-thread barrier costs 64 more TMEM columns and 32 more FMNMX3 per thread-tile: +3 %; re-ordering the warps into four softmax
      // warpgroups + a producer warpgroup and moving registers to the softmax warps with setmaxnreg (104 / 64) removes the ~50 bytes of
      // spills of the 96-register build but is another +0.7 %)
      float* sm = gmax + (j & 1) * 256;
      sts_f32(sm + hf * 128 + r, fmaxf(mx0, mx1) * sc);
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
      const float rowmax = fmaxf(lds_f32(sm + r), lds_f32(sm + 128 + r));
      const bool grow = rowmax > m_used + kRescaleThreshold;   // identical in both threads of the row
      const float m_new = grow ? rowmax : m_used;
      tick(2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_sfree[g]);   // all my reads of S are done: the MMA warp may overwrite it with the next tile's scores
      if (j > 0) {
        mbar_wait(&bar_pv[g], (j - 1) & 1);   // previous P V done: P and O are ours again
        tc_fence_after();
        if (hf == 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_used - m_new) : 1.0f;
#pragma unroll
          for (int c = 0; c < DV / 16; ++c) {
            uint32_t t16[16];
            tmem_ld16(tmem_o + lane_off + c * 16, t16);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = __float_as_uint(__uint_as_float(t16[i]) * alpha);
            tmem_st16(tmem_o + lane_off + c * 16, t16);
          }
          tmem_st_wait();
        }
      }
      m_used = m_new;
      tick(3);
      if (TOKEN) {
        if (g == 0) asm volatile("bar.sync 3, 512;" ::: "memory");
        else asm volatile("bar.sync 4, 512;" ::: "memory");
      }
      tick(4);
      // pass 2: p = 2^(s*scale - m) -> packed fp16 -> my 32 columns of P
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t (&raw)[32] = c == 0 ? raw0 : raw1;
        uint32_t pk[16];
        if ((c + 1) * 32 <= kv_valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a0, a1;   // s * scale - m for two keys in one FFMA2 (bit-identical to two scalar FMAs)
            upk2(fma2(pk2u(raw[2 * i], raw[2 * i + 1]), pk2(sc, sc), pk2(-m_used, -m_used)), a0, a1);
            if (PE > 0 && PE < 10 && (i % (PE > 0 ? PE : 1)) == PE - 1) {   // this pair's exponentials on the FMA pipe (packed fp32x2)
              float e0, e1;
              exp2_poly3_x2(a0, a1, e0, e1);
              pk[i] = pack_h2(e0, e1);
            }
            else pk[i] = pack_h2(fast_exp2(a0), fast_exp2(a1));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = c * 32 + 2 * i;
            const float p0 = c0 < kv_valid ? fast_exp2(fmaf(__uint_as_float(raw[2 * i]), sc, -m_used)) : 0.f;
            const float p1 = c0 + 1 < kv_valid ? fast_exp2(fmaf(__uint_as_float(raw[2 * i + 1]), sc, -m_used)) : 0.f;
            pk[i] = pack_h2(p0, p1);
          }
        }
        tmem_st16(tmem_p + lane_off + hf * 32 + c * 16, pk);
      }
      tick(5);
      if (TOKEN) {   // hand the MUFU phase to the other group (B's last turn has no taker)
        if (g == 0) asm volatile("bar.arrive 4, 512;" ::: "memory");
        else if (j + 1 < T) asm volatile("bar.arrive 3, 512;" ::: "memory");
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bar_p[g]);
      tick(6);
    }
    if constexpr (kProf) {
      if (lane == 0 && (w == 0 || w == 4)) {
        long long* d = a.dbg + (static_cast<long long>(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * 32 + g * 16 + (hf ? 8 : 0);
        for (int i = 0; i < 7; ++i) d[i] = pt[i];
      }
    }
    // epilogue: O / denominator; the two threads of a row write alternate 8-column groups
    mbar_wait(&bar_pv[g], (T - 1) & 1);
    tc_fence_after();
    float o[DV];
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t raw16[16];
      tmem_ld16(tmem_o + lane_off + c * 16, raw16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[c * 16 + i] = __uint_as_float(raw16[i]);
    }
    tc_fence_before();
    const int qrow = q0 + g * QT + r;
    if (qrow < a.L) {
      const float inv = 1.f / o[D];
      __half* dst = a.out + (static_cast<long long>(n) * a.L + qrow) * a.ldo + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        if ((c & 1) == hf) {
          uint4 u;
          u.x = pack_h2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
          u.y = pack_h2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
          u.z = pack_h2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
          u.w = pack_h2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + c * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int PE, bool TOKEN, int MW = 1>
cudaError_t launch_attn_pp(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mvt, const CUtensorMap& mkb, const CUtensorMap& mvbt,
                           const AttnKernelArgs& ka, int L, int heads, int NF, cudaStream_t stream) {
  using C = PPCfg<D>;
  if constexpr (!C::kFits) {
    return cudaErrorInvalidValue;
  } else {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(attn_pp_kernel<D, PE, TOKEN, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
      if (e != cudaSuccess) return e;
      attr = true;
    }
    dim3 grid((L + 2 * QT - 1) / (2 * QT), heads, NF);
    if constexpr (PE == 31) {   // phase-timing experiment: blocking, prints per-key-tile averages
      AttnKernelArgs kd = ka;
      const long long nct = static_cast<long long>(grid.x) * grid.y * grid.z;
      if (cudaMalloc(&kd.dbg, nct * 32 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
      cudaMemsetAsync(kd.dbg, 0, nct * 32 * sizeof(long long), stream);
      attn_pp_kernel<D, PE, TOKEN, MW><<<grid, ATTPP_THREADS + 32 * (MW - 1), C::kSmemBytes, stream>>>(mq, mk, mvt, mkb, mvbt, kd);
      cudaStreamSynchronize(stream);
      long long* hb = static_cast<long long*>(malloc(nct * 32 * sizeof(long long)));
      cudaMemcpy(hb, kd.dbg, nct * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
      double acc[32] = {0};
      for (long long c = 0; c < nct; ++c)
        for (int i = 0; i < 32; ++i) acc[i] += static_cast<double>(hb[c * 32 + i]);
      const double tiles = static_cast<double>(nct) * ((L + KT - 1) / KT);
      const char* sn[7] = {"wait S", "pass1 (ld+max)", "max xchg+bar", "ld wait+PV wait", "wait token", "pass2", "handoff+st+arrive"};
      fprintf(stderr, "[attn_pp phase clocks per key tile, D=%d L=%d token=%d stages=%d]\n", D, L, (int)TOKEN, C::kStages);
      for (int i = 0; i < 7; ++i)
        fprintf(stderr, "  %-18s A.h0 %7.1f  A.h1 %7.1f  B.h0 %7.1f  B.h1 %7.1f\n", sn[i], acc[i] / tiles, acc[8 + i] / tiles, acc[16 + i] / tiles, acc[24 + i] / tiles);
      free(hb);
      cudaFree(kd.dbg);
      return cudaGetLastError();
    }
    attn_pp_kernel<D, PE, TOKEN, MW><<<grid, ATTPP_THREADS + 32 * (MW - 1), C::kSmemBytes, stream>>>(mq, mk, mvt, mkb, mvbt, ka);
    return cudaGetLastError();
  }
}

template <int D, int PE, bool PT>
cudaError_t launch_attn8(dim3 grid, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mvt, const CUtensorMap& mkb,
                         const CUtensorMap& mvbt, AttnKernelArgs ka, int L, cudaStream_t stream) {
  using C = ACfg<D, PT>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel8<D, PE, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_kernel8<D, PE, PT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  if constexpr (PE == 31) {   // phase-timing experiment: blocking, prints the per-key-tile averages
    const long long nct = static_cast<long long>(grid.x) * grid.y * grid.z;
    if (cudaMalloc(&ka.dbg, nct * 24 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
    cudaMemsetAsync(ka.dbg, 0, nct * 24 * sizeof(long long), stream);
    attn_kernel8<D, PE, PT><<<grid, ATT8_THREADS, C::kSmemBytes, stream>>>(mq, mk, mvt, mkb, mvbt, ka);
    cudaStreamSynchronize(stream);
    long long* h = static_cast<long long*>(malloc(nct * 24 * sizeof(long long)));
    cudaMemcpy(h, ka.dbg, nct * 24 * sizeof(long long), cudaMemcpyDeviceToHost);
    double acc[24] = {0};
    for (long long c = 0; c < nct; ++c)
      for (int i = 0; i < 24; ++i) acc[i] += static_cast<double>(h[c * 24 + i]);
    const double tiles = static_cast<double>(nct) * ((L + KT - 1) / KT);
    const char* sn[6] = {"wait S", "pass1", "max xchg+bar", "wait PV(+rescale)", "pass2", "fence+arrive"};
    const char* mn[6] = {"wait k_full", "wait sfree0", "wait sfree1", "issue S", "wait P,V", "issue PV"};
    fprintf(stderr, "[attn8 phase clocks per key tile, D=%d L=%d PT=%d stages=%d]\n", D, L, (int)PT, C::kStages);
    for (int i = 0; i < 6; ++i) fprintf(stderr, "  softmax h0 %-18s %8.1f   h1 %8.1f\n", sn[i], acc[i] / tiles, acc[6 + i] / tiles);
    for (int i = 0; i < 6; ++i) fprintf(stderr, "  mma        %-18s %8.1f\n", mn[i], acc[12 + i] / tiles);
    free(h);
    cudaFree(ka.dbg);
    return cudaGetLastError();
  } else {
    attn_kernel8<D, PE, PT><<<grid, ATT8_THREADS, C::kSmemBytes, stream>>>(mq, mk, mvt, mkb, mvbt, ka);
    return cudaGetLastError();
  }
}

template <int D>
cudaError_t launch_attn_t(const AttnArgs& a, cudaStream_t stream) {
  using C = ACfg<D>;
  CUtensorMap mq, mk, mvt, mkb, mvbt;
  const long long tokens = static_cast<long long>(a.NF) * a.L;
  if (!make_map_2d(&mq, a.q, tokens, a.ldq, a.ldq, QT)) return cudaErrorInvalidValue;
  if (!make_map_2d(&mk, a.k, tokens, a.ldk, a.ldk, KT)) return cudaErrorInvalidValue;
  if (!make_map_2d(&mvt, a.vt, static_cast<long long>(a.heads) * C::kDv, static_cast<long long>(a.NF) * a.vt_stride, a.ldvt, C::kDv)) return cudaErrorInvalidValue;
  mkb = mk;
  mvbt = mvt;
  const int B = a.NF / (a.F > 0 ? a.F : 1);
  if (a.Lb > 0) {
    if (!make_map_2d(&mkb, a.kb, static_cast<long long>(B) * a.Lb, a.ldkb, a.ldkb, KT)) return cudaErrorInvalidValue;
    if (!make_map_2d(&mvbt, a.vbt, static_cast<long long>(a.heads) * C::kDv, static_cast<long long>(B) * a.vbt_stride, a.ldvbt, C::kDv))
      return cudaErrorInvalidValue;
  }
  AttnKernelArgs ka;
  ka.out = a.out;
  ka.ldo = a.ldo;
  ka.L = a.L;
  ka.Lb = a.Lb;
  ka.F = a.F > 0 ? a.F : 1;
  ka.nf_nobank = a.nf_nobank;
  ka.heads = a.heads;
  ka.vt_stride = static_cast<int>(a.vt_stride);
  ka.vbt_stride = static_cast<int>(a.vbt_stride);
  ka.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(D));
  ka.dbg = nullptr;
  dim3 grid((a.L + QT - 1) / QT, a.heads, a.NF);
  constexpr bool kCanPT = 128 + 64 + C::kDv <= 256;
#ifdef HV_TUNING
  // A/B experiments (build with -DHV_TUNING, see humanvid_b200/build.py): HV_ATTN_POLY = 0 / 2 / 3 / 4 (every n-th exponential pair on the
  // FMA pipe) or 31 (blocking run that prints per-phase clocks); HV_ATTN_MW = 1 / 2 MMA-issuing warps; HV_ATTN_PP = 0 (one query tile per
  // CTA) / 1 (token hand-off between the groups) / 2 (free-running groups).  The release build has exactly one path per head dim.
  static const int mode = [] { const char* v = getenv("HV_ATTN_POLY"); return v ? atoi(v) : 4; }();
  static const int mw = [] { const char* v = getenv("HV_ATTN_MW"); return v ? atoi(v) : 2; }();
  static const int pp = [] { const char* v = getenv("HV_ATTN_PP"); return v ? atoi(v) : 2; }();
  if (PPCfg<D>::kFits && pp != 0 && a.L > QT) {
#define HV_PP(PE_, TOK_, MW_) return launch_attn_pp<D, PE_, TOK_, MW_>(mq, mk, mvt, mkb, mvbt, ka, a.L, a.heads, a.NF, stream)
#define HV_PP_MODE(TOK_, MW_) \
    switch (mode) { case 0: HV_PP(0, TOK_, MW_); case 2: HV_PP(2, TOK_, MW_); case 3: HV_PP(3, TOK_, MW_); case 31: HV_PP(31, TOK_, MW_); default: HV_PP(4, TOK_, MW_); }
    if (pp == 1 && mw == 2) { HV_PP_MODE(true, 2) }
    if (pp == 1) { HV_PP_MODE(true, 1) }
    if (mw == 2) { HV_PP_MODE(false, 2) }
    HV_PP_MODE(false, 1)
#undef HV_PP_MODE
#undef HV_PP
  }
  if (mode == 31) return launch_attn8<D, 31, kCanPT>(grid, mq, mk, mvt, mkb, mvbt, ka, a.L, stream);
#else
  // d < 48 (level 0, d = 40): two query tiles per CTA, one MMA issuer per tile, every 4th exponential pair on the FMA pipe
  if (PPCfg<D>::kFits && a.L > QT) return launch_attn_pp<D, 4, false, 2>(mq, mk, mvt, mkb, mvbt, ka, a.L, a.heads, a.NF, stream);
#endif
  return launch_attn8<D, 0, kCanPT>(grid, mq, mk, mvt, mkb, mvbt, ka, a.L, stream);
}

}  // namespace

cudaError_t launch_attention(const AttnArgs& a, int /*num_sms*/, cudaStream_t stream) {
  switch (a.d) {
    case 8: return launch_attn_t<8>(a, stream);
    case 16: return launch_attn_t<16>(a, stream);
    case 32: return launch_attn_t<32>(a, stream);
    case 40: return launch_attn_t<40>(a, stream);
    case 80: return launch_attn_t<80>(a, stream);
    case 160: return launch_attn_t<160>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace hv
