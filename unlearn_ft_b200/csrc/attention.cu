// Fused flash-style attention for sm_100a, head_dim 64, no mask (reference: F.scaled_dot_product_attention at
// pdm/models/unet/blocks.py:275-277, called by HeadGatedAttnProcessor2 for self (Lq = Lk in {4096,1024,256,64}) and
// cross (Lk = 77) attention).
//
// One CTA = one (sample, head, 128-query block); 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0 (one lane) : TMA producer  - Q once, then K/V blocks of 128 keys through a 2-stage smem ring
//   warp 1 (one lane) : MMA issuer    - S = Q K^T (tcgen05.mma M128 N128 K64 -> TMEM), O_blk = P V (M128 N64 K128)
//   warps 2..5        : softmax       - each thread owns one query row (= one TMEM lane): online softmax in the
//                                       log2 domain, P written as bf16 into a swizzle-128B K-major smem tile that the
//                                       second MMA consumes, running output kept in registers (64 fp32)
// Scores never touch HBM: traffic per launch = Q + K + V + O (+ LSE).
#include "common.cuh"
#include "../../include/b200pdm.h"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace b200 {
extern std::atomic<uint64_t> g_launches;
void set_err(const char* fmt, const char* a);
int make_map_public(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box);
int make_map_f32_3d(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld,
                    uint32_t box_rows);

constexpr int kQ = 128, kKV = 128, kD = 64;
// Share of the forward exponentials moved from MUFU.EX2 to the FMA pipe (exp2_fma): elements with (i & mask) == mask; 7 -> 12.5 %,
// 3 -> 25 %, 1 -> 50 %, 64 -> none.  While the softmax warps still waited ~1500 cycles per key block for their next S the share
// made no difference (round-2 experiment F: 25 % / 50 % slower); with S issued early (s_read barrier) the exponential phase of
// the two in-phase query tiles IS the MUFU rate (2100 of 2048 cycles per key block), and a small share pays: L = 4096, 5 heads,
// batch 16, same box: none 525.8 us, 25 % 512.6 us, 12.5 % 508.4 us.  (The polynomial costs 9 issue slots per element.)
#ifndef B200PDM_EXP_FMA_MASK
#define B200PDM_EXP_FMA_MASK 7
#endif
constexpr int kExpFmaMask = B200PDM_EXP_FMA_MASK;
constexpr int kTileBytes = 128 * 128;  // [128 rows][64 bf16] swizzle-128B tile

struct AttnFwdParams {
  int B, H, Lq, Lk, nkv;
  float scale_log2;  // softmax scale * log2(e)
  bf16* out;
  int64_t ldo;
  float* lse;  // [B, H, Lq] log2-domain log-sum-exp (may be null)
  int causal;  // 1: query i attends keys <= i (CLIP text encoder); masked per row in the ragged path
#ifdef B200PDM_DIAG
  long long* dbg;  // optional phase timers of CTA 0 / warp 2 (`make diag`)
#else
  static constexpr long long* dbg = nullptr;
#endif
};

// ----------------------------------------------------------------------------------------------------------------
// Forward: one CTA = one (sample, head, 256-query block) = two 128-row query tiles; one CTA per SM.
//   warp 0      : TMA producer  - both Q tiles once, then K/V blocks of 128 keys through a 3-stage smem ring
//   warp 1      : MMA issuer    - per key block j and query tile t:  S_t(j+1) = Q_t K_{j+1}^T as soon as the softmax warps
//                                 hold S_t(j) in registers (s_read), O_t += P_t(j) V_j when P_t(j) has been written (p_full);
//                                 P_t is the PV MMA's A operand straight from tensor memory (tcgen05.mma, TS form)
//   warps 2..17 : four softmax warpgroups: query tile t = wg >> 1, key half = wg & 1; a thread owns one query row (= TMEM
//                 lane) and 64 of the block's 128 keys.  The scores are read out of TMEM once and kept in registers,
//                 exponentiated against a LAZILY updated running maximum (the O accumulator in TMEM is rescaled, and the
//                 block redone from registers, only when a row maximum grew by more than 2^8, which after the first blocks
//                 is rare), and written back to tensor memory as bf16 P.  12.5 % of the exponentials run on the FMA pipe.
// O leaves TMEM once, at the end.  TMEM: S_t at 128 t, O_t at 256 + 64 t, P_t at 384 + 64 t.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kFwd2Threads = 64 + 4 * 128;   // TMA warp, MMA warp, four softmax warpgroups
constexpr int kKVStages = 3;
constexpr float kRescaleThreshold = 8.f;   // log2 domain
constexpr int kPCol = 384;                 // TMEM: S_t at 128 t, O_t at 256 + 64 t, P_t (bf16 pairs) at 384 + 64 t

__global__ void __launch_bounds__(kFwd2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* sKV = smem + 2 * kTileBytes;                 // kKVStages x (K tile, V tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kKVStages * 2 * kTileBytes);   // (P lives in tensor memory: no smem tile)
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                 // [3]
  uint64_t* kv_empty = kv_full + kKVStages;     // [3]
  uint64_t* s_full = kv_empty + kKVStages;      // [2]
  uint64_t* p_full = s_full + 2;                // [2]
  uint64_t* pv_done = p_full + 2;               // [2]
  uint64_t* s_read = pv_done + 2;               // [2] the softmax warps hold S_t(j) in registers: S_t(j + 1) may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_read + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kQ;
  const int n_qt = min(2, (p.Lq - q0 + kQ - 1) / kQ);   // query tiles with at least one valid row

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) {
      printf("b200pdm attention: dynamic smem not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_read[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;   // S_t: columns [128 t, 128 t + 128);  O_t: columns [256 + 64 t, 256 + 64 t + 64)
  pdl_wait();   // barrier init / TMEM allocation above overlapped the predecessor's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, n_qt * kTileBytes);
      for (int t = 0; t < n_qt; ++t) tma_load_4d(sQ + t * kTileBytes, &tm_q, q_full, 0, q0 + t * kQ, h, b);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < p.nkv; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[stage], 2 * kTileBytes);
        tma_load_4d(sKV + stage * 2 * kTileBytes, &tm_k, &kv_full[stage], 0, j * kKV, h, b);
        tma_load_4d(sKV + stage * 2 * kTileBytes + kTileBytes, &tm_v, &kv_full[stage], 0, j * kKV, h, b);
      }
      if (++stage == kKVStages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_bf16(kKV, 0, 0);  // S: N = 128 keys, A/B K-major
    const uint32_t idesc_o = make_idesc_bf16(kD, 0, 1);   // PV: N = 64, B (= V) MN-major
    const uint64_t dk = make_smem_desc_sw128(0, 16, 1024);       // K-major operand template
    const uint64_t dv = make_smem_desc_sw128(0, 8192, 1024);     // MN-major (V) template
    const uint32_t q_lo = (smem_u32(sQ) & 0x3FFFF) >> 4, kv_lo = (smem_u32(sKV) & 0x3FFFF) >> 4;
    auto issue_s = [&](int t, int stage) {   // S_t = Q_t K^T
      const uint32_t a = q_lo + t * (kTileBytes >> 4), bq = kv_lo + stage * (2 * kTileBytes >> 4);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) umma_bf16(tmem + t * 128, dk | (a + k * 2), dk | (bq + k * 2), idesc_s, k > 0);
      umma_commit(&s_full[t]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one())
      for (int t = 0; t < n_qt; ++t) issue_s(t, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < p.nkv; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kKVStages) nstage = 0, nphase ^= 1;
      if (j + 1 < p.nkv) {
        mbar_wait(&kv_full[nstage], nphase);
        tc_fence_after();
        // S_t(j + 1) goes out as soon as the softmax warps have pulled S_t(j) into registers -- it runs on the tensor core
        // while they exponentiate block j.  (Issued after PV_t(j), as before the scores were read in a single pass, it left
        // every softmax warp waiting ~1500 cycles per key block for its next S.)
        for (int t = 0; t < n_qt; ++t) {
          mbar_wait(&s_read[t], j & 1);
          tc_fence_after();
          if (elect_one()) issue_s(t, nstage);
        }
      }
      for (int t = 0; t < n_qt; ++t) {
        const bool mprof = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
        const long long tm0 = mprof ? clock64() : 0;
        mbar_wait(&p_full[t], j & 1);   // P_t(j) is in tensor memory, S_t(j) has been read out of it
        tc_fence_after();
        if (mprof) p.dbg[4 + t * 8] += clock64() - tm0;
        if (elect_one()) {
          const uint32_t bv = kv_lo + (stage * 2 * kTileBytes + kTileBytes) / 16;
#pragma unroll
          for (int k = 0; k < kKV / 16; ++k)   // A = P_t straight from tensor memory: 8 columns (16 bf16 keys) per K step
            umma_bf16_ts(tmem + 256 + t * 64, tmem + kPCol + t * 64 + k * 8, dv | (bv + k * 128), idesc_o,
                         (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&pv_done[t]);
        }
      }
      if (elect_one()) umma_commit(&kv_empty[stage]);   // both tiles' PV(j) (and S(j), long before) have read this stage
      stage = nstage, phase = nphase;
    }
  } else if (((warp - 2) >> 3) < n_qt) {
    // ===================== softmax warpgroups =====================
    // FOUR warpgroups: query tile t = wg >> 1, key half = wg & 1.  A thread owns one query row (= TMEM lane) and 64 of the 128
    // keys of a block.  It reads its 64 scores out of tensor memory ONCE (TMEM reads run at 64 B/clk/SM: the two passes of the
    // round-1 kernel -- maximum, then exponentials -- cost 4096 cycles per pair of 128x128 tiles, twice the MUFU bound and
    // exactly what that kernel measured), keeps them in registers, exponentiates at once against the STALE running maximum
    // while the second half of the read is still in flight, and writes the bf16 probabilities back into tensor memory, where the
    // PV MMA takes them as its A operand (no shared-memory round trip for P).  The two threads of a row agree on the block
    // maximum afterwards through a 4-byte shared-memory exchange and a 64-thread named barrier; only if a row's maximum grew
    // by more than 2^8 -- rare after the first blocks -- is the accumulator rescaled and the block redone from registers.
    const int wg = (warp - 2) >> 2;
    const int t = wg >> 1, half = wg & 1;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_s = tmem + t * 128 + half * 64 + lane_base;          // this thread's 64 score columns
    const uint32_t tmem_o = tmem + 256 + t * 64 + half * 32 + lane_base;     // the 32 O columns it rescales / writes
    const uint32_t tmem_p = tmem + kPCol + t * 64 + half * 32 + lane_base;   // its 64 probabilities = 32 packed columns
    float* xch = reinterpret_cast<float*>(bars + 32);                        // [parity 2][tile 2][quarter 4][half 2][32 lanes]
    const int bar_id = 1 + t * 4 + qd;                                       // the two warps that share these 32 rows
    auto xslot = [&](int par, int hh) { return xch + (((par * 2 + t) * 4 + qd) * 2 + hh) * 32 + lane; };
    float m = -INFINITY, l = 0.f;
    // One 32-key chunk: p = exp2(s * c - mref) -> 16 packed bf16 columns of this row's P; accumulates the row sum (lsum) and
    // the running maximum of the raw scores (mx).
    auto exp_chunk = [&](const uint32_t(&vv)[32], int c, int vh, float mref, float& lsum, float& mx) {
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-mref, -mref);
      float2 ls2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float l4 = 0.f, mxa = mx, mxb = mx;
      uint32_t pk[16];
      if (vh == 64) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {   // packed pairs: one FFMA2 + one FADD2 per two scores
          const float s0 = __uint_as_float(vv[2 * i]), s1 = __uint_as_float(vv[2 * i + 1]);
          mxa = fmaxf(mxa, s0), mxb = fmaxf(mxb, s1);
          const float2 xs = ffma2(make_float2(s0, s1), sc2, nm2);
          // (kExpFmaMask: share of the exponentials evaluated on the FMA pipe instead of MUFU.EX2, see the top of the file)
          const float2 e = make_float2(((2 * i) & kExpFmaMask) == kExpFmaMask ? exp2_fma(xs.x) : exp2f(xs.x),
                                       ((2 * i + 1) & kExpFmaMask) == kExpFmaMask ? exp2_fma(xs.y) : exp2f(xs.y));
          ls2[i & 1] = fadd2(ls2[i & 1], e);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e.x, e.y);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const bool ok = c * 32 + 2 * i + u < vh;
            const float sv = __uint_as_float(vv[2 * i + u]);
            if (ok) mxa = fmaxf(mxa, sv);
            e2[u] = ok ? exp2f(fmaf(sv, p.scale_log2, -mref)) : 0.f;
            l4 += e2[u];
          }
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e2[0], e2[1]);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
      }
      tmem_st_32x16(tmem_p + c * 16, pk);
      lsum += l4 + ((ls2[0].x + ls2[0].y) + (ls2[1].x + ls2[1].y));
      mx = fmaxf(mxa, mxb);
    };
    auto max_chunk = [&](const uint32_t(&vv)[32], int c, int vh, float& mx) {
      float m4[4] = {mx, mx, mx, mx};
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (vh == 64 || c * 32 + i < vh) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(vv[i]));
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    };
    // Optional MUFU turn-taking between the two query tiles (-DB200PDM_ATTN_PINGPONG; named barriers 9 / 10 pass a token between
    // the tiles' 256 threads each).  Phase timers (make diag, B200PDM_ATTN_DBG) show the tiles running IN PHASE: both
    // exponentiate at the MUFU rate (2300-2700 cycles per key block) and then both wait ~1100 cycles for their PV / next-S MMAs.
    // Taking turns does not help: with only one tile's two warps per scheduler active, the in-order MUFU -> FADD2 / F2FP
    // dependence (ptxas places consumers one pair behind their producers) makes a tile's exponentials latency-bound at ~1750
    // cycles, and 2 x 1750 equals the in-phase period.  Measured equal (568 us), so it stays off.
#ifdef B200PDM_ATTN_PINGPONG
    const bool pingpong = n_qt == 2;
#else
    const bool pingpong = false;
#endif
    auto token_wait = [&]() {
      if (pingpong) asm volatile("bar.sync %0, 512;" ::"r"(9 + t) : "memory");
    };
    auto token_pass = [&]() {
      if (pingpong) asm volatile("bar.arrive %0, 512;" ::"r"(9 + (t ^ 1)) : "memory");
    };
    if (pingpong && t == 1) asm volatile("bar.arrive %0, 512;" ::"r"(9) : "memory");   // tile 0 goes first
    for (int j = 0; j < p.nkv; ++j) {
      int valid = min(kKV, p.Lk - j * kKV);  // keys in this block
      if (p.causal) valid = min(valid, q0 + t * kQ + r - j * kKV + 1);   // ... that this row may attend (can be <= 0)
      const int vh = min(64, max(0, valid - half * 64));                 // ... of them in this thread's half
      const bool prof = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && qd == 2 && half == 0;
      const long long tp0 = prof ? clock64() : 0;
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const long long tp1 = prof ? clock64() : 0;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_s, v0);
      tmem_ld_32x32(tmem_s + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_read[t]);           // the 64 scores live in registers from here on
      // PV_t(j - 1) must have consumed the previous P_t (and updated O_t) before this block writes P_t: its MMA is issued AFTER
      // S_t(j) now, so this is a real -- short -- wait (a completed phase returns at once, also for j = 0)
      mbar_wait(&pv_done[t], (j + 1) & 1);
      tc_fence_after();
      const long long tp2 = prof ? clock64() : 0;
      float mx = -INFINITY, lsum = 0.f;
      bool redo = false;
      if (j == 0) {
        // first block: no reference maximum yet -- maximum first, then the exponentials
        max_chunk(v0, 0, vh, mx);
        tmem_ld_wait();
        max_chunk(v1, 1, vh, mx);
        *xslot(0, half) = mx * p.scale_log2;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        m = fmaxf(mx * p.scale_log2, *xslot(0, half ^ 1));
        if (valid <= 0) m = 0.f;                        // (causal: nothing to attend in this block -- all its p are 0)
        redo = true;
      } else {
        token_wait();
        exp_chunk(v0, 0, vh, m, lsum, mx);
        tmem_ld_wait();
        exp_chunk(v1, 1, vh, m, lsum, mx);
        token_pass();
        *xslot(j & 1, half) = mx * p.scale_log2;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        float m_tile = fmaxf(mx * p.scale_log2, *xslot(j & 1, half ^ 1));
        if (valid <= 0) m_tile = m;
        const bool grow = m_tile - m > kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {              // (both warps of these rows see the same values: same branch)
          const float alpha = grow ? exp2f(m - m_tile) : 1.f;
          if (grow) m = m_tile;
          l *= alpha;
          tc_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {      // 16 columns at a time: the 64 scores stay in registers meanwhile
            uint32_t o[16];
            tmem_ld_32x16(tmem_o + hh * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(tmem_o + hh * 16, o);
          }
          redo = true;
        }
      }
      if (redo) {
        if (j == 0) token_wait();
        lsum = 0.f;
        float dummy = -INFINITY;
        exp_chunk(v0, 0, vh, m, lsum, dummy);
        exp_chunk(v1, 1, vh, m, lsum, dummy);
        if (j == 0) token_pass();
      }
      l += lsum;
      const long long tp3 = prof ? clock64() : 0;
      tmem_st_wait();       // P_t (and a rescaled O_t) are in tensor memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (prof) {
        const long long tp4 = clock64();
        long long* d = p.dbg + t * 8;
        d[0] += tp1 - tp0, d[1] += tp2 - tp1, d[2] += tp3 - tp2, d[3] += tp4 - tp3;
      }
    }
    mbar_wait(&pv_done[t], (p.nkv - 1) & 1);
    tc_fence_after();
    // row sum over both halves
    const int par = p.nkv & 1;                      // a slot the last block's exchange did not use
    *xslot(par, half) = l;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float l_tot = l + *xslot(par, half ^ 1);
    const int q = q0 + t * kQ + r;
    const float inv = 1.f / l_tot;
    bf16* dst = p.out + (static_cast<int64_t>(b) * p.Lq + q) * p.ldo + h * kD + half * 32;
    {
      uint32_t o[32];
      tmem_ld_32x32(tmem_o, o);
      tmem_ld_wait();
      if (q < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = __uint_as_float(o[g * 8 + i]) * inv;
          *reinterpret_cast<bf16x8*>(dst + g * 8) = pack8(t8);
        }
      }
    }
    if (half == 0 && q < p.Lq && p.lse) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Lq + q] = m + log2f(l_tot);
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int make_qkv_map(CUtensorMap* map, const void* ptr, int64_t ld, int B, int H, int L) {
  // dims (d, token, head, batch)
  uint64_t dims[4] = {64, (uint64_t)L, (uint64_t)H, (uint64_t)B};
  uint64_t str[4] = {1, (uint64_t)ld, 64, (uint64_t)L * (uint64_t)ld};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_map_public(map, ptr, 4, dims, str, box);
}

}  // namespace b200

using namespace b200;

extern "C" int b200pdm_attention_fwd_ex(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                        void* out, int64_t ldo, float* lse, int batch, int heads, int lq, int lk,
                                        float scale, int causal, b200pdm_stream_t stream_);
extern "C" int b200pdm_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                     void* out, int64_t ldo, float* lse, int batch, int heads, int lq, int lk,
                                     float scale, b200pdm_stream_t stream_) {
  return b200pdm_attention_fwd_ex(q, ldq, k, ldk, v, ldv, out, ldo, lse, batch, heads, lq, lk, scale, 0, stream_);
}
extern "C" int b200pdm_attention_fwd_ex(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                        void* out, int64_t ldo, float* lse, int batch, int heads, int lq, int lk,
                                        float scale, int causal, b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !out || batch <= 0 || heads <= 0 || lq <= 0 || lk <= 0) return B200PDM_ERR_ARG;
  if (ldo % 8 || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_err("attention_fwd: output pitch/base must be 16-byte aligned", "");
    return B200PDM_ERR_ARG;
  }
  CUtensorMap mq, mk, mv;
  int rc = make_qkv_map(&mq, q, ldq, batch, heads, lq);
  if (rc) return rc;
  rc = make_qkv_map(&mk, k, ldk, batch, heads, lk);
  if (rc) return rc;
  rc = make_qkv_map(&mv, v, ldv, batch, heads, lk);
  if (rc) return rc;
  AttnFwdParams p;
  p.B = batch, p.H = heads, p.Lq = lq, p.Lk = lk, p.nkv = (lk + kKV - 1) / kKV;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(out), p.ldo = ldo, p.lse = lse;
  p.causal = causal ? 1 : 0;
#ifdef B200PDM_DIAG
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) dbg_on = getenv("B200PDM_ATTN_DBG") ? 1 : 0;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(long long), stream);
  }
  p.dbg = dbg_on ? dbg_buf : nullptr;
#endif
  cudaError_t e;
  const size_t smem = (2 + 2 * kKVStages) * kTileBytes + 256 + 4096;   // Q and K/V tiles, barriers, row-maximum exchange
  static bool attr_set = false;   // once, not per launch (graph capture)
  if (!attr_set) {
    e = cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_err("attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return B200PDM_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((lq + 2 * kQ - 1) / (2 * kQ), heads, batch);
  launch_pdl<true>(attn_fwd2_kernel, grid, kFwd2Threads, smem, stream, mq, mk, mv, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err("attention_fwd launch: %s", cudaGetErrorString(e));
    return B200PDM_ERR_CUDA;
  }
  g_launches++;
#ifdef B200PDM_DIAG
  if (dbg_on) {
    long long hbuf[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[attn dbg] Lq=%d Lk=%d nkv=%d | tile0: wait_s=%lld tmem_ld=%lld exp+max+xchg=%lld st_wait+arrive=%lld mma_wait_p=%lld | "
            "tile1: wait_s=%lld tmem_ld=%lld exp+max+xchg=%lld st_wait+arrive=%lld mma_wait_p=%lld (cycles of CTA 0, summed over the key blocks)\n",
            lq, lk, p.nkv, hbuf[0], hbuf[1], hbuf[2], hbuf[3], hbuf[4], hbuf[8], hbuf[9], hbuf[10], hbuf[11], hbuf[12]);
  }
#endif
  return B200PDM_OK;
}

// ================================================================================================================
// Backward.  One CTA = one (sample, head, 128-key block); loops over 128-query blocks i:
//   S = Q_i K^T, dP = dO_i V^T                      (tensor core -> TMEM)
//   P = exp2(S*c - LSE), dS = P o (dP - D) * scale   (softmax threads; bf16 tiles to smem)
//   dV += P^T dO_i, dK += dS^T Q_i                   (accumulate in TMEM across i)
//   dQ_i = dS K                                      (TMEM -> fp32 vector atomics into dq_acc)
// The same swizzle-128B smem tiles are consumed K-major or MN-major by changing only the UMMA descriptors, so K, Q,
// dO, P and dS are each staged once.  D = rowsum(dO o O) is precomputed by attn_delta_kernel.
// ================================================================================================================
namespace b200 {

struct AttnBwdParams {
  int B, H, Lq, Lk, nq;
  float scale, scale_log2;
  const float* lse;    // [B, H, Lq] (log2 domain)
#ifdef B200PDM_DIAG
  long long* dbg;      // optional phase timers (`make diag`)
#else
  static constexpr long long* dbg = nullptr;
#endif
  const float* delta;  // [B, H, Lq]
  float* dq_acc;       // fp32 [B*Lq, ld_dq], head h at columns [64h, 64h+64)
  int64_t ld_dq;
  bf16* dk;
  int64_t ld_dk;
  bf16* dv;
  int64_t ld_dv;
};

__global__ void attn_delta_kernel(const bf16* __restrict__ dO, int64_t lddo, const bf16* __restrict__ O, int64_t ldo,
                                  float* __restrict__ delta, int B, int H, int Lq) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);  // (b*Lq + q)*H + h
  if (row >= (int64_t)B * Lq * H) return;
  const int h = (int)(row % H);
  const int64_t tok = row / H;
  const int b = (int)(tok / Lq), q = (int)(tok % Lq);
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dO + tok * lddo + h * 64 + 2 * lane));
  float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(O + tok * ldo + h * 64 + 2 * lane));
  float s = warp_sum(a.x * c.x + a.y * c.y);
  if (lane == 0) delta[((int64_t)b * H + h) * Lq + q] = s;
}

__device__ __forceinline__ void st_tile_row32(uint32_t row_base, int c, int sw, const float (&f)[32]) {
  // 32 consecutive columns [32c, 32c+32) of this thread's row into the two-subtile K-major swizzle-128B layout
  const uint32_t base = row_base + (c >> 1) * kTileBytes;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    __nv_bfloat162 a0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
    __nv_bfloat162 a1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
    __nv_bfloat162 a2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
    __nv_bfloat162 a3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
    const int chunk = ((c & 1) * 4 + g) ^ sw;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + chunk * 16),
                 "r"(*reinterpret_cast<uint32_t*>(&a0)), "r"(*reinterpret_cast<uint32_t*>(&a1)),
                 "r"(*reinterpret_cast<uint32_t*>(&a2)), "r"(*reinterpret_cast<uint32_t*>(&a3))
                 : "memory");
  }
}

constexpr int kBwdThreads = 448;   // TMA warp, MMA warp, 8 softmax warps (two per TMEM lane quarter, 64 key columns each), 4 dQ-drain warps

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdParams p) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytes;
  uint8_t* sQ = smem + 2 * kTileBytes;    // [2]
  uint8_t* sdO = smem + 4 * kTileBytes;   // [2]
  uint8_t* sP = smem + 6 * kTileBytes;    // two sub-tiles
  uint8_t* sdS = smem + 8 * kTileBytes;   // two sub-tiles
  uint8_t* sDQ = smem + 10 * kTileBytes;  // fp32 dQ tile as two [128 rows x 32 cols] swizzle-128B sub-tiles (bulk reduce source)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * kTileBytes);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_empty = bars + 8;
  uint64_t* acc_full = bars + 9;
  uint64_t* sdp_read = bars + 10;  // the softmax warps hold S_i and dP_i in registers: the TMEM columns may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kb * kKV;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(acc_full, 1);
    mbar_init(sdp_read, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 128, t_dv = tmem + 256, t_dk = tmem + 320, t_dq = tmem + 384;
  pdl_wait();   // barrier init / TMEM allocation above overlapped the predecessor's tail

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * kTileBytes);
      tma_load_4d(sK, &tm_k, kv_full, 0, k0, h, b);
      tma_load_4d(sV, &tm_v, kv_full, 0, k0, h, b);
      for (int i = 0; i < p.nq; ++i) {
        const int s = i & 1;
        mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[s], 2 * kTileBytes);
        tma_load_4d(sQ + s * kTileBytes, &tm_q, &qdo_full[s], 0, i * kQ, h, b);
        tma_load_4d(sdO + s * kTileBytes, &tm_do, &qdo_full[s], 0, i * kQ, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id_kk128 = make_idesc_bf16(128, 0, 0);  // S, dP : A K-major, B K-major, N = 128
      const uint32_t id_mm64 = make_idesc_bf16(64, 1, 1);    // dV, dK: A MN-major (P^T / dS^T), B MN-major, N = 64
      const uint32_t id_km64 = make_idesc_bf16(64, 0, 1);    // dQ    : A K-major (dS), B MN-major (K), N = 64
      mbar_wait(kv_full, 0);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      // S_j = Q_j K^T and dP_j = dO_j V^T of query tile j.  Tile j + 1's pair is issued as soon as the softmax warps have
      // pulled S_j / dP_j out of tensor memory (sdp_read), i.e. it runs on the tensor core WHILE they exponentiate tile j;
      // round 1 issued it after dV_j / dK_j, which left the tensor core idle for the whole softmax phase.
      auto issue_s_dp = [&](int j) {
        const int sj = j & 1;
        mbar_wait(&qdo_full[sj], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ + sj * kTileBytes), da = smem_u32(sdO + sj * kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S = Q K^T
          umma_bf16(t_s, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                    id_kk128, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dP = dO V^T
          umma_bf16(t_dp, make_smem_desc_sw128(da + k * 32, 16, 1024),
                    make_smem_desc_sw128(v_addr + k * 32, 16, 1024), id_kk128, k > 0);
        umma_commit(s_full);
      };
      issue_s_dp(0);
      for (int i = 0; i < p.nq; ++i) {
        const int s = i & 1;
        const uint32_t q_addr = smem_u32(sQ + s * kTileBytes), do_addr = smem_u32(sdO + s * kTileBytes);
        if (i + 1 < p.nq) {
          mbar_wait(sdp_read, i & 1);
          tc_fence_after();
          issue_s_dp(i + 1);
        }
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV += P^T dO   (A: M' = keys over the two sub-tiles (LBO), K' = query rows)
          umma_bf16(t_dv, make_smem_desc_sw128(p_addr + k * 2048, kTileBytes, 1024),
                    make_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), id_mm64, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK += dS^T Q
          umma_bf16(t_dk, make_smem_desc_sw128(ds_addr + k * 2048, kTileBytes, 1024),
                    make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), id_mm64, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&qdo_empty[s]);   // Q_i / dO_i may be refilled
        if (i > 0) {   // dQ_{i-1} is drained one tile late (after the softmax of tile i), so nobody ever waits for a dQ MMA
          mbar_wait(dq_empty, (i - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ_i = dS K  (A K-major over the two sub-tiles, B = K tile read MN-major)
          umma_bf16(t_dq, make_smem_desc_sw128(ds_addr + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), id_km64, k > 0);
        umma_commit(dq_full);   // ... and, MMAs completing in order, P / dS of tile i may be overwritten
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 10) {
    // ===================== dQ drain warps =====================
    // dQ tile j: TMEM -> fp32 smem tile -> ONE bulk tensor reduce-add per 32-column half (instead of 2048 vector atomics), by
    // four warps of their own (one per TMEM lane quarter): with the softmax warps draining dQ themselves, the ~1000 cycles per
    // query tile sat on the critical path of the kernel (softmax -> drain -> next softmax).
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const int sw = r & 7;
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    const bool issuer = threadIdx.x == 320;
    for (int j = 0; j < p.nq; ++j) {
      mbar_wait(dq_full, j & 1);
      tc_fence_after();
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile left smem
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v0[32];
        tmem_ld_32x32(t_dq + lane_base + hh * 32, v0);
        tmem_ld_wait();
        const uint32_t row0 = smem_u32(sDQ) + hh * kTileBytes + r * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) sts128(row0 + ((g ^ sw) << 4), v0[4 * g], v0[4 * g + 1], v0[4 * g + 2], v0[4 * g + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);   // the TMEM accumulator may be overwritten by the next dQ
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tm_dq)),
                       "r"(smem_u32(sDQ + c * kTileBytes)), "r"(h * kD + c * 32), "r"(j * kQ), "r"(b)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all dQ reduces have been performed
  } else {
    // ===================== softmax warps =====================
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;   // which 64 of the 128 key columns this warp handles (and which of dV / dK it writes)
    const int r = qd * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t p_row = smem_u32(sP) + r * 128, ds_row = smem_u32(sdS) + r * 128;
    const int sw = r & 7;
    const int valid_k = min(kKV, p.Lk - k0);
    const int64_t bh = (int64_t)b * p.H + h;
    for (int i = 0; i < p.nq; ++i) {
      const int q = i * kQ + r;
      const bool q_ok = q < p.Lq;
      const float lse = q_ok ? p.lse[bh * p.Lq + q] : 0.f;
      const float dlt = q_ok ? p.delta[bh * p.Lq + q] : 0.f;
      const bool prof = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64;
      const long long tp0 = prof ? clock64() : 0;
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      const long long tp1 = prof ? clock64() : 0;
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nl2 = make_float2(-lse, -lse);
      const float2 nd2 = make_float2(-dlt, -dlt), ss2 = make_float2(p.scale, p.scale);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {   // one 32-column chunk at a time (the drain warps share the register file now)
        const int c = half * 2 + cc;
        uint32_t s_[32], d_[32];
        tmem_ld_32x32(t_s + lane_base + c * 32, s_);
        tmem_ld_32x32(t_dp + lane_base + c * 32, d_);
        tmem_ld_wait();
        if (cc == 1) {   // S / dP of this tile live in registers now: the MMA warp may issue the next tile's pair
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_read);
        }
        float pf[32], dsf[32];
        if (q_ok && valid_k == kKV) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {   // packed pairs: FFMA2, FADD2, 2 x FMUL2 per two elements
            const float2 xs = ffma2(make_float2(__uint_as_float(s_[2 * j]), __uint_as_float(s_[2 * j + 1])), sc2, nl2);
            const float2 pe = make_float2(exp2f(xs.x), exp2f(xs.y));
            const float2 dd = fmul2(fadd2(make_float2(__uint_as_float(d_[2 * j]), __uint_as_float(d_[2 * j + 1])), nd2), ss2);
            const float2 ds = fmul2(pe, dd);
            pf[2 * j] = pe.x, pf[2 * j + 1] = pe.y;
            dsf[2 * j] = ds.x, dsf[2 * j + 1] = ds.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pe = exp2f(fmaf(__uint_as_float(s_[j]), p.scale_log2, -lse));
            pe = (q_ok && (c * 32 + j < valid_k)) ? pe : 0.f;
            pf[j] = pe;
            dsf[j] = pe * (__uint_as_float(d_[j]) - dlt) * p.scale;
          }
        }
        if (cc == 0 && i > 0) {   // dQ_{i-1} (hence dV_{i-1}, dK_{i-1}) has read P / dS of the previous tile
          mbar_wait(dq_full, (i - 1) & 1);
          tc_fence_after();
        }
        st_tile_row32(p_row, c, sw, pf);
        st_tile_row32(ds_row, c, sw, dsf);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (prof) {
        const long long tp2 = clock64();
        p.dbg[0] += tp1 - tp0, p.dbg[1] += tp2 - tp1;
      }
    }
    // accumulated dV / dK for key row k0 + r
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int key = k0 + r;
    {
      const int which = half;   // first warp group writes dV, second dK
      bf16* dst = (which == 0 ? p.dv : p.dk) + ((int64_t)b * p.Lk + key) * (which == 0 ? p.ld_dv : p.ld_dk) + h * kD;
      const uint32_t t = (which == 0 ? t_dv : t_dk) + lane_base;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t + c * 32, v);
        tmem_ld_wait();
        if (key < p.Lk) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t8[j] = __uint_as_float(v[g * 8 + j]);
            *reinterpret_cast<bf16x8*>(dst + c * 32 + g * 8) = pack8(t8);
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__global__ void cast2d_f32_to_bf16_kernel(const float* __restrict__ x, int64_t ldx, bf16* __restrict__ y, int64_t ldy,
                                          int64_t rows, int cols8) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = rows * cols8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols8;
    int c0 = (int)(i - r * cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(x + r * ldx + c0);
    const float4 c = *reinterpret_cast<const float4*>(x + r * ldx + c0 + 4);
    float f[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    *reinterpret_cast<bf16x8*>(y + r * ldy + c0) = pack8(f);
  }
}

}  // namespace b200

// q, k, v, out, dout: bf16 matrices (head h = columns [64h, 64h+64)); lse from the forward; dq/dk/dv: bf16 outputs.
// workspace: fp32 [B*H*Lq (delta) + B*Lq*H*64 (dq accumulator)], caller provided.
extern "C" int b200pdm_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                     const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                     void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                     float* workspace, int batch, int heads, int lq, int lk, float scale,
                                     b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !out || !dout || !lse || !dq || !dk || !dv || !workspace) return B200PDM_ERR_ARG;
  if (lddq % 8 || lddk % 8 || lddv % 8) return B200PDM_ERR_ARG;
  float* delta = workspace;
  float* dq_acc = workspace + (((int64_t)batch * heads * lq + 3) / 4) * 4;
  const int64_t ld_acc = (int64_t)heads * 64;
  const int64_t rows = (int64_t)batch * lq * heads;
  launch_pdl(attn_delta_kernel, (unsigned)((rows + 7) / 8), 256, 0, stream, reinterpret_cast<const bf16*>(dout), lddo,
                                                                   reinterpret_cast<const bf16*>(out), ldo, delta,
                                                                   batch, heads, lq);
  if (cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)batch * lq * ld_acc, stream) != cudaSuccess)
    return B200PDM_ERR_CUDA;
  CUtensorMap mq, mk, mv, mdo;
  int rc = make_qkv_map(&mq, q, ldq, batch, heads, lq);
  if (rc) return rc;
  rc = make_qkv_map(&mk, k, ldk, batch, heads, lk);
  if (rc) return rc;
  rc = make_qkv_map(&mv, v, ldv, batch, heads, lk);
  if (rc) return rc;
  rc = make_qkv_map(&mdo, dout, lddo, batch, heads, lq);
  if (rc) return rc;
  AttnBwdParams p;
  p.B = batch, p.H = heads, p.Lq = lq, p.Lk = lk, p.nq = (lq + kQ - 1) / kQ;
  p.scale = scale, p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse, p.delta = delta, p.dq_acc = dq_acc, p.ld_dq = ld_acc;
#ifdef B200PDM_DIAG
  static long long* bdbg = nullptr;
  static int bdbg_on = -1;
  if (bdbg_on < 0) bdbg_on = getenv("B200PDM_ATTN_DBG") ? 1 : 0;
  if (bdbg_on) {
    if (!bdbg) cudaMalloc(&bdbg, 16 * sizeof(long long));
    cudaMemsetAsync(bdbg, 0, 16 * sizeof(long long), stream);
  }
  p.dbg = bdbg_on ? bdbg : nullptr;
#endif
  p.dk = reinterpret_cast<bf16*>(dk), p.ld_dk = lddk, p.dv = reinterpret_cast<bf16*>(dv), p.ld_dv = lddv;
  CUtensorMap mdq;
  rc = make_map_f32_3d(&mdq, dq_acc, (uint64_t)ld_acc, (uint64_t)lq, (uint64_t)batch, (uint64_t)ld_acc, kQ);
  if (rc) return rc;
  const size_t smem = 12 * kTileBytes + 256;
  cudaError_t e;
  static bool attr_set = false;   // once, not per launch (graph capture)
  if (!attr_set) {
    e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_err("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return B200PDM_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((lk + kKV - 1) / kKV, heads, batch);
  launch_pdl<true>(attn_bwd_kernel, grid, kBwdThreads, smem, stream, mq, mk, mv, mdo, mdq, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err("attention_bwd launch: %s", cudaGetErrorString(e));
    return B200PDM_ERR_CUDA;
  }
  const int64_t nrow = (int64_t)batch * lq;
  const int cols8 = heads * 8;
  int64_t blocks = (nrow * cols8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(cast2d_f32_to_bf16_kernel, (int)blocks, 256, 0, stream, dq_acc, ld_acc, reinterpret_cast<bf16*>(dq), lddq, nrow,
                                                            cols8);
  e = cudaGetLastError();
  if (e != cudaSuccess) return B200PDM_ERR_CUDA;
  g_launches += 3;
#ifdef B200PDM_DIAG
  if (bdbg_on) {
    long long hbuf[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hbuf, bdbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[attn bwd dbg] Lq=%d Lk=%d nq=%d | wait_s=%lld softmax=%lld wait_dq=%lld drain_dq=%lld\n", lq, lk, p.nq, hbuf[0],
            hbuf[1], hbuf[2], hbuf[3]);
  }
#endif
  return B200PDM_OK;
}
