// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution core for sm_100a.
//
//   warp 0             : TMA producer   - cp.async.bulk.tensor tiles of A and B into a swizzle-128B smem ring (3-8 stages)
//   warp 1             : MMA issuer     - tcgen05.mma (M = 128, or 256 across a cta_group::2 pair; N = block_n; K = 16) into a
//                                         double-buffered TMEM accumulator (tall tiles: both buffers, two 128-row sub-tiles)
//   warps 2..9         : epilogue       - two warpgroups: tcgen05.ld accumulator rows (one row per lane), alpha / bias /
//                                         time-embedding row bias / residual / GEGLU, 256-bit row stores; no shared memory
//   (both single-thread roles run warp-uniform loops and issue through elect.sync)
//
// The A/B tiles are addressed by a small "operand program" evaluated by the producer, which is what turns the same kernel into:
// linear fwd/dgrad/wgrad (K-major or MN-major 2D operands, batched), 3x3/1x1 convolution fprop (NHWC activations fetched
// with 4D TMA boxes, out-of-bounds = zero padding; stride-1 3x3: one box of R + 2 image rows per (channel block, kw) serves
// the three kh taps -- GemmDev::kh3), convolution dgrad (tap flip + weight matrix read MN-major) and convolution wgrad (both
// operands MN-major, K = output pixels).  A host-side cost model (plan_gemm) picks block_n, pair mode, tall tiles and split-K
// (per-split fp32 slabs in the caller's workspace + an ordered finalize pass: deterministic).
//
// Reference call sites replaced: see include/b200pdm.h (linear/conv entries).
#include "common.cuh"
#include "../../include/b200pdm.h"

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <string>
#include <stdio.h>
#include <string.h>

namespace b200 {

std::atomic<uint64_t> g_launches{0};
static char g_err[512] = "";
void set_err(const char* fmt, const char* a = "") { snprintf(g_err, sizeof(g_err), fmt, a); }
const char* get_err() { return g_err; }

// ------------------------------------------------------------------------------------------------
// Driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// rank-R bf16 tensor map, 128B swizzle, zero OOB fill. dims/strides innermost first; strides in ELEMENTS
// (strides[0] is implicit 1 and ignored).
static int make_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box, const uint32_t* estr) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled entry point not available");
    return B200PDM_ERR_DRIVER;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = estr ? estr[i] : 1;
  }
  for (int i = 1; i < rank; ++i) {
    gstr[i - 1] = strides_el[i] * 2;
    if (gstr[i - 1] % 16 != 0) {
      set_err("tensor map stride not a multiple of 16 bytes");
      return B200PDM_ERR_ARG;
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    set_err("tensor map base not 16-byte aligned");
    return B200PDM_ERR_ARG;
  }
  static int promo = -1;
  if (promo < 0) {
    const char* e = getenv("B200PDM_L2PROMO");
    promo = e ? atoi(e) : 256;
  }
  const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu %llu box %u %u %u %u)",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0], bx[1],
             rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return B200PDM_ERR_DRIVER;
  }
  return B200PDM_OK;
}

// fp32 [batch][rows][cols] view (row pitch ld, batch pitch rows*ld elements), [box_rows x 32 col] boxes, 128B swizzle:
// target of the attention backward's bulk tensor reduce-add of dQ tiles.
int make_map_f32_3d(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld,
                    uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return B200PDM_ERR_DRIVER;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {ld * 4, rows * ld * 4};
  cuuint32_t bx[3] = {32, box_rows, 1}, es[3] = {1, 1, 1};
  if ((gstr[0] % 16) || (reinterpret_cast<uintptr_t>(ptr) & 15)) return B200PDM_ERR_ARG;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled (fp32 reduce map) failed");
    return B200PDM_ERR_DRIVER;
  }
  return B200PDM_OK;
}

int make_map_public(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_el,
                    const uint32_t* box) {
  return make_map(map, ptr, rank, dims, strides_el, box, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Device-side parameters
// ------------------------------------------------------------------------------------------------
// n / d for 0 <= n < 2^31 without the ~200-cycle integer-division sequence: the single producer / epilogue threads sit on
// the critical path of every k-block, so their address arithmetic must stay a handful of instructions.
struct FastDiv {
  uint32_t mul, shr;
};
static FastDiv make_fastdiv(int64_t d64) {
  uint32_t d = d64 < 1 ? 1u : static_cast<uint32_t>(d64);
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  uint64_t m = ((1ull << 32) * ((1ull << s) - d)) / d + 1;
  FastDiv f;
  f.mul = static_cast<uint32_t>(m), f.shr = s;
  return f;
}
__device__ __forceinline__ int fdiv(int n, FastDiv d) {
  return static_cast<int>((__umulhi(static_cast<uint32_t>(n), d.mul) + static_cast<uint32_t>(n)) >> d.shr);
}

struct OpDev {
  int mode;
  int Z1;
  int cblks;   // 64-wide channel blocks per tap (conv modes)
  int taps;
  int Wo, HoWo;  // output grid (pixel decomposition)
  int stride;
  int flip;
  int pad;       // leading zero padding of the convolution window (1 = "same" 3x3; 0 = diffusers Downsample2D(padding=0))
  FastDiv fd_Wo, fd_HoWo;
};

struct GemmDev {
  OpDev a, b;
  int M;             // valid output rows (per batch)
  int n_per_group;   // valid output cols per N-group
  int n_groups;      // 1, or taps for conv wgrad
  int64_t out_group_stride;
  int64_t out_split_stride;   // elements between the per-split partial-sum slabs (deterministic split-K), else 0
  int tiles_m, tiles_n_per_group, Z, splits;
  int kblocks, kb_per_split;
  int k_taps;        // conv fprop/dgrad: K runs over (channel block, tap), taps innermost; else 1
  int block_n, stages;
  int pair;          // 1: cta_group::2 -- one M=256 MMA per CTA pair, each CTA stages its 128 A rows and HALF of B
  int cluster;       // CTAs per cluster (1, 2, 4): B tile loaded once per cluster and multicast
  int m_sub;         // 128-row sub-tiles per CTA and tile (2 = "tall" tile: two accumulators share every B stage)
  int tiles_m_super; // ceil(tiles_m / (cluster * m_sub))
  FastDiv fd_tiles_per_split, fd_slab, fd_tiles_n, fd_tiles_n_per_group, fd_Z1, fd_taps, fd_rows_per_group;
  uint32_t idesc;
  // epilogue
  void* out;
  int out_fp32;
  int64_t ldo, obs1, obs2;
  const float* bias;
  const float* rowbias;
  int64_t ld_rowbias;
  int rows_per_group;
  const bf16* residual;
  int64_t ldr, rbs1, rbs2;
  float alpha;
  int accumulate;
#ifdef B200PDM_DIAG   // `make diag` -> libb200pdm_diag.so (tools/diag_gemm_*.py); the product library carries none of this
  long long* dbg;  // optional per-role wait-cycle counters
  int dbg_mode;    // B200PDM_GEMM_DBGMODE: 1 = quarter of the MMAs, 2 = no A loads, 4 = no B loads, 8/128 = epilogue cuts
#else
  static constexpr long long* dbg = nullptr;
  static constexpr int dbg_mode = 0;
#endif
  int epi_groups;  // epilogue warpgroups: group g takes every epi_groups-th 32-column chunk of a tile
  // fused GEGLU epilogue (B = [2F, K] projection weight): a tile's accumulator holds the VALUE columns of bh = block_n / 2
  // hidden units in [0, bh) and their GATE columns in [bh, 2 bh); out[:, f] = value * gelu_erf(gate)
  int geglu_F;     // F (0 = off)
  bf16* aux;       // optional [M, 2F] pre-activations (value | gate) saved for the backward pass
  int64_t ld_aux;
  // Row-shared 3x3 taps ("kh3", stride-1 "same" convolutions whose CTA rows are R whole image rows): K runs over
  // (channel block, kw, kh), kh innermost.  One TMA box of R + 2 image rows (x shifted by kw - 1, zero-filled outside the
  // image) is staged per (channel block, kw) and serves the three kh taps: tap kh reads the box from image row kh on
  // (W * 128 bytes = whole swizzle atoms, so only the descriptor start address moves).  A-operand bytes through
  // L2 -> shared memory drop to (R + 2) / (3 R) of one box per tap; the B tiles keep their own, finer ring.
  int kh3;           // 0 = off
  int a_slots;       // A ring depth (kh3)
  int a_slot_bytes;  // (R + 2) * W * 128
  int kh_row_bytes;  // W * 128: one image row of a 64-channel block
  int rank_stride;   // first 128-row block of CTA `rank` in a super tile: m_super * cluster * m_sub + rank * rank_stride
  int sub_stride;    // ... and of its sub-tile s: + s * sub_stride   (normal: 1 / cluster; kh3: m_sub / 1 = contiguous rows)
};

// Where a tile sits: decoded once per tile by each role.
struct TileCoord {
  int split, z1, z2, m_tile, grp, nt;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int t, int tiles_per_split, int tiles_n, int cluster,
                                                 int rank) {
  TileCoord c;
  c.split = p.splits > 1 ? fdiv(t, p.fd_tiles_per_split) : 0;
  int r = t - c.split * tiles_per_split;
  int z = 0;
  if (p.Z > 1) {
    z = fdiv(r, p.fd_slab);
    r -= z * (p.tiles_m_super * tiles_n);
  }
  c.z2 = p.Z > 1 ? fdiv(z, p.fd_Z1) : 0;
  c.z1 = z - c.z2 * p.a.Z1;
  const int m_super = fdiv(r, p.fd_tiles_n);
  const int n_tile = r - m_super * tiles_n;
  c.m_tile = m_super * cluster * p.m_sub + rank * p.rank_stride;   // sub-tile s of the CTA: + s * p.sub_stride
  c.grp = p.n_groups > 1 ? fdiv(n_tile, p.fd_tiles_n_per_group) : 0;
  c.nt = n_tile - c.grp * p.tiles_n_per_group;
  return c;
}

// (kh, kw) of tap 0..8 (row-major 3x3), optionally flipped (dgrad); centre for 1x1.
__device__ __forceinline__ void tap_offsets(int taps, int tap, int flip, int* kh, int* kw) {
  int h = 1, w = 1;
  if (taps == 9) {
    h = (tap * 11) >> 5;
    w = tap - 3 * h;
    if (flip) h = 2 - h, w = 2 - w;
  }
  *kh = h, *kw = w;
}

#ifndef B200PDM_DIAG
#define DBG_WAIT(slot, stmt) \
  do {                       \
    stmt;                    \
  } while (0)
#else
#define DBG_WAIT(slot, stmt)                                              \
  do {                                                                    \
    if (p.dbg && blockIdx.x == 0) {                                       \
      long long t0__ = clock64();                                         \
      stmt;                                                               \
      if ((threadIdx.x & 31) == 0 || (slot) == 6)                         \
        atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (slot)),  \
                  static_cast<unsigned long long>(clock64() - t0__));     \
    } else {                                                              \
      stmt;                                                               \
    }                                                                     \
  } while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStageABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kAtomBytes = 64 * 64 * 2;              // one [64 k][64 mn] MN-major atom = 8 KiB
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
#ifndef B200PDM_MAX_EPI_GROUPS
#define B200PDM_MAX_EPI_GROUPS 2
#endif
constexpr int kMaxEpiGroups = B200PDM_MAX_EPI_GROUPS;   // epilogue warpgroups (4 warps each)
constexpr int kMaxThreads = 64 + 128 * kMaxEpiGroups;

template <bool PAIR>
__device__ __forceinline__ void ld4(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  if (PAIR)
    tma_load_4d_2sm(dst, map, bar, c0, c1, c2, c3);
  else
    tma_load_4d(dst, map, bar, c0, c1, c2, c3);
}
template <bool PAIR>
__device__ __forceinline__ void ld3(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  if (PAIR)
    tma_load_3d_2sm(dst, map, bar, c0, c1, c2);
  else
    tma_load_3d(dst, map, bar, c0, c1, c2);
}

// Per-tile operand cursors: everything that does not change along K is computed once per tile.
struct ACursor {
  int c1, c2, c3;   // K2D/MN2D: (row0, z1, z2); CONV_ACT: (w0*stride - 1, h0*stride - 1, n0)
};
__device__ __forceinline__ ACursor make_a_cursor(const OpDev& op, const TileCoord& tc, int m_tile) {
  ACursor c;
  if (op.mode == B200PDM_OP_CONV_ACT) {
    const int pix0 = m_tile * kBlockM;
    const int n0 = fdiv(pix0, op.fd_HoWo);
    const int rem = pix0 - n0 * op.HoWo;
    const int h0 = fdiv(rem, op.fd_Wo);
    const int w0 = rem - h0 * op.Wo;
    c.c1 = w0 * op.stride - op.pad, c.c2 = h0 * op.stride - op.pad, c.c3 = n0;
  } else {
    c.c1 = m_tile * kBlockM, c.c2 = tc.z1, c.c3 = tc.z2;
  }
  return c;
}

template <bool PAIR>
__device__ __forceinline__ void load_a(const OpDev& op, const CUtensorMap* map, uint8_t* dst, uint64_t* bar,
                                       const ACursor& c, int kb, int cb, int tap) {
  switch (op.mode) {
    case B200PDM_OP_K2D:
      ld4<PAIR>(dst, map, bar, kb * kBlockK, c.c1, c.c2, c.c3);
      break;
    case B200PDM_OP_MN2D:
      ld4<PAIR>(dst, map, bar, c.c1, kb * kBlockK, c.c2, c.c3);
      ld4<PAIR>(dst + kAtomBytes, map, bar, c.c1 + 64, kb * kBlockK, c.c2, c.c3);
      break;
    case B200PDM_OP_CONV_ACT: {
      int kh, kw;
      tap_offsets(op.taps, tap, op.flip, &kh, &kw);
      ld4<PAIR>(dst, map, bar, cb * kBlockK, c.c1 + kw, c.c2 + kh, c.c3);
      break;
    }
    default:
      break;
  }
}

// B tile of n-tile `nt` (tap group `grp` for wgrad).  pair mode (cta_group::2): CTA `rank` stages only its half of the
// tile (rows / atoms [rank*half, (rank+1)*half)) at the start of its own stage buffer.  Otherwise, with cluster > 1, each
// CTA fetches its 1/cluster slice and multicasts it to every CTA of the cluster.
struct BCursor {
  int n0, kh, kw;
};
template <bool PAIR>
__device__ __forceinline__ void load_b(const OpDev& op, const CUtensorMap* map, uint8_t* dst, uint64_t* bar,
                                       const BCursor& c, const TileCoord& tc, int block_n, int kb, int cb, int tap,
                                       int cluster, int rank, int geglu_F = 0) {
  if (geglu_F) {   // K-major [2F, K] weight: value rows [n0, n0 + bh) and gate rows [F + n0, F + n0 + bh), bh = block_n / 2
    if (PAIR) {    // cta_group::2 splits B along N: CTA 0 supplies the value half, CTA 1 the gate half
      tma_load_4d_2sm(dst, map, bar, kb * kBlockK, c.n0 + rank * geglu_F, tc.z1, tc.z2);
    } else {
      tma_load_4d(dst, map, bar, kb * kBlockK, c.n0, tc.z1, tc.z2);
      tma_load_4d(dst + (block_n / 2) * 128, map, bar, kb * kBlockK, c.n0 + geglu_F, tc.z1, tc.z2);
    }
    return;
  }
  const int parts = PAIR ? 2 : cluster;
  const uint16_t mask = static_cast<uint16_t>((1u << cluster) - 1);
  const int rows = block_n / parts;            // K-major slice
  const int atoms = (block_n / 64) / parts;    // MN-major slice (host guarantees divisibility when parts > 1)
  const int j0 = rank * atoms;
  // destination of this CTA's slice: pair mode packs it at the start of the stage, multicast keeps tile order
  uint8_t* kdst = PAIR ? dst : dst + rank * rows * 128;
  switch (op.mode) {
    case B200PDM_OP_K2D:
      if (PAIR)
        tma_load_4d_2sm(kdst, map, bar, kb * kBlockK, c.n0 + rank * rows, tc.z1, tc.z2);
      else if (cluster == 1)
        tma_load_4d(dst, map, bar, kb * kBlockK, c.n0, tc.z1, tc.z2);
      else
        tma_load_4d_mc(kdst, map, bar, kb * kBlockK, c.n0 + rank * rows, tc.z1, tc.z2, mask);
      break;
    case B200PDM_OP_MN2D:
      for (int j = 0; j < atoms; ++j) {
        if (PAIR)
          tma_load_4d_2sm(dst + j * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), kb * kBlockK, tc.z1, tc.z2);
        else if (cluster == 1)
          tma_load_4d(dst + j * kAtomBytes, map, bar, c.n0 + 64 * j, kb * kBlockK, tc.z1, tc.z2);
        else
          tma_load_4d_mc(dst + (j0 + j) * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), kb * kBlockK, tc.z1, tc.z2, mask);
      }
      break;
    case B200PDM_OP_CONV_W:
      if (PAIR)
        tma_load_3d_2sm(kdst, map, bar, cb * kBlockK, tap, c.n0 + rank * rows);
      else if (cluster == 1)
        tma_load_3d(dst, map, bar, cb * kBlockK, tap, c.n0);
      else
        tma_load_3d_mc(kdst, map, bar, cb * kBlockK, tap, c.n0 + rank * rows, mask);
      break;
    case B200PDM_OP_CONV_WT:
      for (int j = 0; j < atoms; ++j) {
        if (PAIR)
          tma_load_3d_2sm(dst + j * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), tap, cb * kBlockK);
        else if (cluster == 1)
          tma_load_3d(dst + j * kAtomBytes, map, bar, c.n0 + 64 * j, tap, cb * kBlockK);
        else
          tma_load_3d_mc(dst + (j0 + j) * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), tap, cb * kBlockK, mask);
      }
      break;
    case B200PDM_OP_CONV_ACT_MN: {
      const int pix0 = kb * kBlockK;
      const int b0 = fdiv(pix0, op.fd_HoWo);
      const int rem = pix0 - b0 * op.HoWo;
      const int h0 = fdiv(rem, op.fd_Wo);
      const int w0 = rem - h0 * op.Wo;
      const int cw = w0 * op.stride + c.kw - op.pad, ch = h0 * op.stride + c.kh - op.pad;
      for (int j = 0; j < atoms; ++j) {
        if (PAIR)
          tma_load_4d_2sm(dst + j * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), cw, ch, b0);
        else if (cluster == 1)
          tma_load_4d(dst + j * kAtomBytes, map, bar, c.n0 + 64 * j, cw, ch, b0);
        else
          tma_load_4d_mc(dst + (j0 + j) * kAtomBytes, map, bar, c.n0 + 64 * (j0 + j), cw, ch, b0, mask);
      }
      break;
    }
    default:
      break;
  }
}

template <int A_MN, int B_MN, bool PAIR>
__global__ void __launch_bounds__(kMaxThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmDev p) {
  pdl_trigger();   // PDL: successors may start their prologues while this grid runs
  unsigned long long gt_entry = 0;
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_b_bytes = (PAIR ? p.block_n / 2 : p.block_n) * 128;
  const int stage_a_bytes = p.m_sub * kStageABytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (p.kh3 ? p.a_slots * p.a_slot_bytes : p.stages * stage_a_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + p.stages * stage_b_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint64_t* afull_bar = tempty_bar + 2;        // [4]  kh3: A ring
  uint64_t* aempty_bar = afull_bar + 4;        // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 4);
  // per epilogue warp: 1 KiB slice holding (bias + time-embedding row bias) of the tile's columns, read back as broadcasts
  uint8_t* sEpi = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(aempty_bar + 6) + 127) & ~uintptr_t(127));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.tiles_n_per_group * p.n_groups;
  // "super tiles": one per cluster = `cluster` consecutive m-tiles sharing the same B tile
  const int tiles_per_split = p.tiles_m_super * tiles_n * p.Z;
  const int total_tiles = tiles_per_split * p.splits;
  constexpr bool pair = PAIR;   // compile-time: kernels containing cta_group::2 instructions must be launched as pairs
  const int cluster = pair ? 2 : p.cluster;       // m-tiles per super tile == CTAs per cluster
  const int mcast = pair ? 1 : p.cluster;          // TMA multicast width (pair mode does not multicast)
  const int rank = cluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = blockIdx.x / cluster, tile_step = gridDim.x / cluster;
  const uint16_t cmask = static_cast<uint16_t>((1u << cluster) - 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], (pair && (p.dbg_mode & 6) == 6) ? 2 : 1);   // (diagnostic no-load mode: both CTAs arrive)
      mbar_init(&empty_bar[i], mcast);     // every CTA multicasting into this stage must have consumed it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (pair ? 8 : 4) * p.epi_groups);   // pair: the leader's MMA waits for both CTAs' epilogue warps
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], mcast);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (pair) {
      tmem_alloc_2sm(tmem_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cluster > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the predecessor's tail; no global access before this point
  const long long t_kernel0 = (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[8] = (long long)(gt - gt_entry);   // prologue, ns
  }

  if (warp == 0) {
    if (p.kh3) {
      // ===================== TMA producer, row-shared 3x3 taps (see GemmDev::kh3) =====================
      int stage = 0, slot = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord tc = decode_tile(p, t, tiles_per_split, tiles_n, cluster, rank);
        const ACursor ac = make_a_cursor(p.a, tc, tc.m_tile);   // (x0 - 1 = -1, first image row - 1, image)
        BCursor bc;
        bc.n0 = tc.nt * p.block_n;
        bc.kh = bc.kw = 0;
        const int g0 = (tc.split * p.kb_per_split) / 3;
        const int g1 = min(p.kblocks, (tc.split + 1) * p.kb_per_split) / 3;
        int cb = g0 / 3;
        int kw = g0 - 3 * cb;
        for (int g = g0; g < g1; ++g) {
          mbar_wait(&aempty_bar[slot], aphase ^ 1);
          if (elect_one()) {
            if (!pair || rank == 0) mbar_expect_tx(&afull_bar[slot], (pair ? 2 : 1) * p.a_slot_bytes);
            ld4<PAIR>(sA + slot * p.a_slot_bytes, &tma_a, &afull_bar[slot], cb * kBlockK, ac.c1 + kw, ac.c2, ac.c3);
          }
          for (int kh = 0; kh < 3; ++kh) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              if (!pair || rank == 0) mbar_expect_tx(&full_bar[stage], (pair ? 2 : 1) * stage_b_bytes);
              // the weight tap whose window offset is (kh, kw) on the activation side (dgrad walks the window mirrored)
              const int tap = p.a.flip ? (2 - kh) * 3 + (2 - kw) : kh * 3 + kw;
              load_b<PAIR>(p.b, &tma_b, sB + stage * stage_b_bytes, &full_bar[stage], bc, tc, p.block_n, 0, cb, tap, mcast, rank, 0);
            }
            if (++stage == p.stages) stage = 0, phase ^= 1;
          }
          if (++kw == 3) kw = 0, ++cb;
          if (++slot == p.a_slots) slot = 0, aphase ^= 1;
        }
      }
    } else {
      // ===================== TMA producer =====================
      // On the critical path of every k-block: no divisions, operand cursors hoisted per tile, and the whole warp runs the
      // loop so that every TMA operand is warp-uniform (one elected lane issues).
      const uint32_t tx_bytes = stage_a_bytes + stage_b_bytes;
      const uint32_t a_bytes = (p.dbg_mode & 2) ? 0 : stage_a_bytes, b_bytes = (p.dbg_mode & 4) ? 0 : stage_b_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord tc = decode_tile(p, t, tiles_per_split, tiles_n, cluster, rank);
        const ACursor ac0 = make_a_cursor(p.a, tc, tc.m_tile);
        const ACursor ac1 = make_a_cursor(p.a, tc, tc.m_tile + p.sub_stride);   // second sub-tile of a tall tile
        BCursor bc;
        bc.n0 = p.geglu_F ? tc.nt * (p.block_n / 2) : tc.nt * p.block_n;
        tap_offsets(p.b.taps, tc.grp, 0, &bc.kh, &bc.kw);
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        int cb = p.k_taps > 1 ? fdiv(kb0, p.fd_taps) : kb0;   // K = (channel block, tap), taps innermost
        int tap = kb0 - cb * p.k_taps;
        for (int kb = kb0; kb < kb1; ++kb) {
          DBG_WAIT(0, mbar_wait(&empty_bar[stage], phase ^ 1));
          uint8_t* a_dst = sA + stage * stage_a_bytes;
          uint8_t* b_dst = sB + stage * stage_b_bytes;
          if (elect_one()) {
            if (p.dbg_mode & 6) {   // diagnostics: drop one or both operand streams (results are garbage)
              if (!pair || rank == 0) mbar_expect_tx(&full_bar[stage], (pair ? 2 : 1) * (a_bytes + b_bytes));
              else if ((p.dbg_mode & 6) == 6) mbar_arrive_remote(&full_bar[stage], 0);   // keeps the peer in lock step
              if (a_bytes) {
                load_a<PAIR>(p.a, &tma_a, a_dst, &full_bar[stage], ac0, kb, cb, tap);
                if (p.m_sub == 2) load_a<PAIR>(p.a, &tma_a, a_dst + kStageABytes, &full_bar[stage], ac1, kb, cb, tap);
              }
              if (b_bytes) load_b<PAIR>(p.b, &tma_b, b_dst, &full_bar[stage], bc, tc, p.block_n, kb, cb, tap, mcast, rank, p.geglu_F);
            } else {
              // pair: both CTAs' loads complete_tx on the LEADER's full barrier, which expects the bytes of the whole pair
              if (!pair || rank == 0) mbar_expect_tx(&full_bar[stage], (pair ? 2 : 1) * tx_bytes);
              DBG_WAIT(6, load_a<PAIR>(p.a, &tma_a, a_dst, &full_bar[stage], ac0, kb, cb, tap);
                       if (p.m_sub == 2) load_a<PAIR>(p.a, &tma_a, a_dst + kStageABytes, &full_bar[stage], ac1, kb, cb, tap);
                       load_b<PAIR>(p.b, &tma_b, b_dst, &full_bar[stage], bc, tc, p.block_n, kb, cb, tap, mcast, rank, p.geglu_F));
            }
          }
          if (++tap == p.k_taps) tap = 0, ++cb;
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!(pair && rank != 0)) {
      // ===================== MMA issuer (pair mode: leader CTA only) =====================
      // Warp-uniform loop, one elected lane issues the MMAs and their commits (see elect_one()).
      const uint32_t a_lo0 = (smem_u32(sA) & 0x3FFFF) >> 4, b_lo0 = (smem_u32(sB) & 0x3FFFF) >> 4;
      const uint32_t a_kstep = A_MN ? (2048 >> 4) : (32 >> 4), b_kstep = B_MN ? (2048 >> 4) : (32 >> 4);
      const uint64_t a_hi = make_smem_desc_sw128(0, A_MN ? kAtomBytes : 16, 1024);
      const uint64_t b_hi = make_smem_desc_sw128(0, B_MN ? kAtomBytes : 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t unit = 0;   // accumulator units issued so far: unit u lives in TMEM slot u & 1, barrier phase (u >> 1) & 1
      if (p.kh3) {
        // row-shared 3x3 taps: one A box per (channel block, kw), three B stages (kh = 0, 1, 2) read it at row offsets
        int slot = 0;
        uint32_t aphase = 0;
        for (int t = first_tile; t < total_tiles; t += tile_step) {
          const int split = p.splits > 1 ? fdiv(t, p.fd_tiles_per_split) : 0;
          const int kb0 = split * p.kb_per_split;
          const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1; kb += 3) {
            mbar_wait(&afull_bar[slot], aphase);
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + stage * (stage_b_bytes >> 4);
              for (int sub = 0; sub < p.m_sub; ++sub) {
                const uint32_t u = unit + sub, acc = u & 1;
                if (kb == kb0 && kh == 0) {
                  mbar_wait(&tempty_bar[acc], ((u >> 1) & 1) ^ 1);
                  tc_fence_after();
                }
                const uint32_t tmem_d = tmem_base + acc * kAccStride;
                const uint32_t a_lo = a_lo0 + (slot * p.a_slot_bytes + kh * p.kh_row_bytes + sub * kStageABytes) / 16;
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint64_t adesc = a_hi | (a_lo + k * a_kstep), bdesc = b_hi | (b_lo + k * b_kstep);
                    const uint32_t accum = (kb > kb0 || kh > 0 || k > 0) ? 1u : 0u;
                    if (pair)
                      umma_bf16_2sm(tmem_d, adesc, bdesc, p.idesc, accum);
                    else
                      umma_bf16(tmem_d, adesc, bdesc, p.idesc, accum);
                  }
                  if (sub == p.m_sub - 1) {
                    if (pair) {
                      umma_commit_2sm_mc(&empty_bar[stage], 3);
                      if (kh == 2) umma_commit_2sm_mc(&aempty_bar[slot], 3);
                    } else if (cluster == 1) {
                      umma_commit(&empty_bar[stage]);
                      if (kh == 2) umma_commit(&aempty_bar[slot]);
                    } else {
                      umma_commit_mc(&empty_bar[stage], cmask);
                      if (kh == 2) umma_commit_mc(&aempty_bar[slot], cmask);
                    }
                  }
                  if (kb + 3 >= kb1 && kh == 2) {
                    if (pair)
                      umma_commit_2sm_mc(&tfull_bar[acc], 3);
                    else
                      umma_commit(&tfull_bar[acc]);
                  }
                }
              }
              if (++stage == p.stages) stage = 0, phase ^= 1;
            }
            if (++slot == p.a_slots) slot = 0, aphase ^= 1;
          }
          unit += p.m_sub;
        }
      } else
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const int split = p.splits > 1 ? fdiv(t, p.fd_tiles_per_split) : 0;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          DBG_WAIT(1, mbar_wait(&full_bar[stage], phase));
          tc_fence_after();
          const long long t_iss0 = (p.dbg && blockIdx.x == 0) ? clock64() : 0;
          const uint32_t b_lo = b_lo0 + stage * (stage_b_bytes >> 4);
          for (int sub = 0; sub < p.m_sub; ++sub) {
            const uint32_t u = unit + sub, acc = u & 1;
            if (kb == kb0) {   // the epilogue must have drained this accumulator (two units ago)
              DBG_WAIT(2, mbar_wait(&tempty_bar[acc], ((u >> 1) & 1) ^ 1));
              tc_fence_after();
            }
            const uint32_t tmem_d = tmem_base + acc * kAccStride;
            const uint32_t a_lo = a_lo0 + (stage * stage_a_bytes + sub * kStageABytes) / 16;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if ((p.dbg_mode & 1) && k > 0) break;   // diagnostics: a quarter of the MMA work
                const uint64_t adesc = a_hi | (a_lo + k * a_kstep), bdesc = b_hi | (b_lo + k * b_kstep);
                if (pair)
                  umma_bf16_2sm(tmem_d, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                else
                  umma_bf16(tmem_d, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
              if (sub == p.m_sub - 1) {
                if (pair)
                  umma_commit_2sm_mc(&empty_bar[stage], 3);  // frees the stage in both CTAs of the pair
                else if (cluster == 1)
                  umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                else
                  umma_commit_mc(&empty_bar[stage], cmask);  // ... in every CTA of the cluster (their TMAs write here too)
              }
              if (kb == kb1 - 1) {
                if (pair)
                  umma_commit_2sm_mc(&tfull_bar[acc], 3);  // accumulator halves complete in both CTAs
                else
                  umma_commit(&tfull_bar[acc]);  // accumulator complete
              }
            }
          }
          if (p.dbg && blockIdx.x == 0 && lane == 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 7), static_cast<unsigned long long>(clock64() - t_iss0));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        unit += p.m_sub;
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2 + 4*epi_groups) =====================
    // A warp may only read the TMEM lane quarter (warp id % 4); each warpgroup covers all four quarters = the 128 rows
    // of a sub-tile and takes every epi_groups-th 32-column chunk.  Each lane owns one output row: it pulls 32 fp32
    // columns out of TMEM, applies alpha / bias / time-embedding row bias / residual and writes its 64 bytes of bf16 as
    // two 256-bit stores (whole 32-byte sectors; the neighbouring chunk, handled by the other warpgroup at the same time,
    // completes the 128-byte line in L2).  Deliberately no shared memory: TMA fills + UMMA operand reads already use
    // ~all of the smem bandwidth, and every staged variant (TMA store, warp transpose) measured ~1000 cycles per chunk
    // waiting for smem behind them.
    {
      const int wg = (warp - 2) >> 2;
      const int q = warp & 3;
      const bool dbg_thread = p.dbg && blockIdx.x == 0 && threadIdx.x == 64;
      const int out_el = p.out_fp32 ? 4 : 2;
      auto aligned_to = [&](int bytes) {   // every chunk start of every row is `bytes`-aligned
        return ((reinterpret_cast<uintptr_t>(p.out) % bytes) == 0) && ((p.ldo * out_el) % bytes == 0) &&
               ((p.obs1 * out_el) % bytes == 0) && ((p.obs2 * out_el) % bytes == 0) &&
               ((p.out_group_stride * out_el) % bytes == 0) && ((p.out_split_stride * out_el) % bytes == 0) &&
               ((p.block_n * out_el) % bytes == 0);
      };
      const bool out_v32 = aligned_to(32), out_v16 = aligned_to(16);
      const bool res_v32 = p.residual && ((reinterpret_cast<uintptr_t>(p.residual) & 31) == 0) && (p.ldr % 16 == 0) &&
                           (p.rbs1 % 16 == 0) && (p.rbs2 % 16 == 0) && (p.block_n % 16 == 0);
      const bool rb_vec = p.rowbias && ((reinterpret_cast<uintptr_t>(p.rowbias) & 15) == 0) && (p.ld_rowbias % 4 == 0);
      // Column offsets shared by all 32 rows of the warp (bias, and the row bias when a warp never straddles two row
      // groups) are summed once per unit into the warp's smem slice; per-lane loads of one address are far slower.
      const bool rb_uniform = p.rowbias && (p.rows_per_group % 32 == 0);
      const bool use_cb = p.bias || rb_uniform;
      const uint32_t cb_s = smem_u32(sEpi + (warp - 2) * 1024);
      uint32_t unit = 0;   // accumulator units drained so far (slot = unit & 1, phase = (unit >> 1) & 1)
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord tc = decode_tile(p, t, tiles_per_split, tiles_n, cluster, rank);
        if (p.geglu_F) {
          // ---------------- fused GEGLU epilogue (see GemmDev::geglu_F) ----------------
          const int bh = p.block_n >> 1;
          const int f_base = tc.nt * bh;
          const int n_f = min(bh, p.geglu_F - f_base);           // valid hidden units of this tile
          for (int sub = 0; sub < p.m_sub; ++sub, ++unit) {
            const int acc = unit & 1;
            const uint32_t acc_phase = (unit >> 1) & 1;
            const int row = (tc.m_tile + sub * p.sub_stride) * kBlockM + q * 32 + lane;
            const bool row_ok = row < p.M;
            {   // bias of the tile's accumulator columns -> this warp's smem slice (value half, then gate half)
              float b8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = lane * 8 + j;
                const int cf = c < bh ? c : c - bh;
                b8[j] = (p.bias && c < p.block_n && cf < n_f) ? __ldg(p.bias + f_base + cf + (c < bh ? 0 : p.geglu_F)) : 0.f;
              }
              __syncwarp();
              sts128(cb_s + lane * 32, __float_as_uint(b8[0]), __float_as_uint(b8[1]), __float_as_uint(b8[2]), __float_as_uint(b8[3]));
              sts128(cb_s + lane * 32 + 16, __float_as_uint(b8[4]), __float_as_uint(b8[5]), __float_as_uint(b8[6]), __float_as_uint(b8[7]));
              __syncwarp();
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
            bf16* o_row = reinterpret_cast<bf16*>(p.out) + static_cast<int64_t>(row) * p.ldo + f_base;
            bf16* a_row = p.aux ? p.aux + static_cast<int64_t>(row) * p.ld_aux + f_base : nullptr;
            const bool o16 = ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (p.ldo % 8 == 0);
            const bool a16 = p.aux && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && (p.ld_aux % 8 == 0) && (p.geglu_F % 8 == 0);
            auto store32 = [&](bf16* dst, const float* f, int nv, bool vec16) {
              if (vec16) {   // whole 16- / 8-column groups as 256- / 128-bit stores (full 32-byte sectors where the row piece
                             // is sector-aligned: 128-bit row stores reach less than half the store rate), the rest element-wise
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                  pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                const bool v32 = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
                int done = 0;
                if (v32 && nv >= 16) {
                  const uint32_t lo[8] = {pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]};
                  stg256(dst, lo);
                  done = 16;
                  if (nv == 32) {
                    const uint32_t hi[8] = {pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]};
                    stg256(dst + 16, hi);
                    done = 32;
                  }
                }
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                  if (g4 * 8 >= done && g4 * 8 + 8 <= nv) {
                    *reinterpret_cast<uint4*>(dst + g4 * 8) = make_uint4(pk[4 * g4], pk[4 * g4 + 1], pk[4 * g4 + 2], pk[4 * g4 + 3]);
                    done = g4 * 8 + 8;
                  }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j >= done && j < nv) dst[j] = __float2bfloat16(f[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) dst[j] = __float2bfloat16(f[j]);
              }
            };
            for (int c0 = wg * 32; c0 < n_f; c0 += 32 * p.epi_groups) {
              const int nv = min(32, n_f - c0);
              uint32_t vh[32], vg[32];
              tmem_ld_32x32(taddr + c0, vh);
              tmem_ld_32x32(taddr + bh + c0, vg);
              tmem_ld_wait();
              float h[32], gt[32];
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) {
                const float4 bh4 = lds128f(cb_s + (c0 + g4 * 4) * 4), bg4 = lds128f(cb_s + (bh + c0 + g4 * 4) * 4);
                h[4 * g4] = __uint_as_float(vh[4 * g4]) + bh4.x, h[4 * g4 + 1] = __uint_as_float(vh[4 * g4 + 1]) + bh4.y;
                h[4 * g4 + 2] = __uint_as_float(vh[4 * g4 + 2]) + bh4.z, h[4 * g4 + 3] = __uint_as_float(vh[4 * g4 + 3]) + bh4.w;
                gt[4 * g4] = __uint_as_float(vg[4 * g4]) + bg4.x, gt[4 * g4 + 1] = __uint_as_float(vg[4 * g4 + 1]) + bg4.y;
                gt[4 * g4 + 2] = __uint_as_float(vg[4 * g4 + 2]) + bg4.z, gt[4 * g4 + 3] = __uint_as_float(vg[4 * g4 + 3]) + bg4.w;
              }
              if (!row_ok) continue;
              if (a_row) {   // pre-activations for the backward pass, rounded exactly as the unfused path stored them
                store32(a_row + c0, h, nv, a16);
                store32(a_row + p.geglu_F + c0, gt, nv, a16);
#pragma unroll
                for (int j = 0; j < 32; ++j) {   // ... and the product is formed from the ROUNDED values, like geglu_fwd did
                  h[j] = __bfloat162float(__float2bfloat16(h[j]));
                  gt[j] = __bfloat162float(__float2bfloat16(gt[j]));
                }
              }
              float y[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) y[j] = h[j] * gelu_erf_fast(gt[j]);
              store32(o_row + c0, y, nv, o16);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (pair && rank != 0)
                mbar_arrive_remote(&tempty_bar[acc], 0);
              else
                mbar_arrive(&tempty_bar[acc]);
            }
          }
          continue;
        }
        const int col_base = tc.nt * p.block_n;                // within the N group
        const int n_cols = min(p.block_n, p.n_per_group - col_base);   // valid columns of this tile
        for (int sub = 0; sub < p.m_sub; ++sub, ++unit) {
          const int acc = unit & 1;
          const uint32_t acc_phase = (unit >> 1) & 1;
          const int m_tile = tc.m_tile + sub * p.sub_stride;
          const int row = m_tile * kBlockM + q * 32 + lane;      // own row (TMEM lane)
          const bool row_ok = row < p.M;
          const int64_t out_off =
              tc.z1 * p.obs1 + tc.z2 * p.obs2 + tc.grp * p.out_group_stride + tc.split * p.out_split_stride +
              static_cast<int64_t>(row) * p.ldo + col_base;
          const bf16* res_row = (p.residual && row_ok)
                                    ? p.residual + tc.z1 * p.rbs1 + tc.z2 * p.rbs2 + static_cast<int64_t>(row) * p.ldr + col_base
                                    : nullptr;
          const float* rb_row = (p.rowbias && !rb_uniform && row_ok)
                                    ? p.rowbias + static_cast<int64_t>(fdiv(row, p.fd_rows_per_group)) * p.ld_rowbias + col_base
                                    : nullptr;
          // The residual operand of the NEXT unit is pulled into L2 now (one 128-byte line per prefetch, this lane's row):
          // its loads sit on the unit's critical path -- on the small-K projections (K <= 1280: epilogue-bound, the
          // accumulator is ready before the previous unit has drained) every chunk otherwise exposes an HBM round trip.
          if (p.residual && wg == 0) {
            int nt_t = t, nt_sub = sub + 1;
            if (nt_sub == p.m_sub) nt_t = t + tile_step, nt_sub = 0;
            if (nt_t < total_tiles) {
              const TileCoord nc = (nt_t == t) ? tc : decode_tile(p, nt_t, tiles_per_split, tiles_n, cluster, rank);
              const int nrow = (nc.m_tile + nt_sub * p.sub_stride) * kBlockM + q * 32 + lane;
              if (nrow < p.M) {
                const int ncol = nc.nt * p.block_n;
                const int ncols = min(p.block_n, p.n_per_group - ncol);
                const char* r0 = reinterpret_cast<const char*>(p.residual + nc.z1 * p.rbs1 + nc.z2 * p.rbs2 +
                                                               static_cast<int64_t>(nrow) * p.ldr + ncol);
                const char* r1 = r0 + ncols * 2;
                for (const char* a = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(r0) & ~uintptr_t(127)); a < r1; a += 128)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
              }
            }
          }
          if (use_cb) {
            const int row_w = m_tile * kBlockM + q * 32;   // first row of the warp
            const float* rbw = (rb_uniform && row_w < p.M)
                                   ? p.rowbias + static_cast<int64_t>(fdiv(row_w, p.fd_rows_per_group)) * p.ld_rowbias + col_base
                                   : nullptr;
            float b8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = lane * 8 + j;
              float x = 0.f;
              if (c < n_cols) {
                if (p.bias) x = __ldg(p.bias + col_base + c);
                if (rbw) x += __ldg(rbw + c);
              }
              b8[j] = x;
            }
            __syncwarp();   // previous unit's readers are done
            sts128(cb_s + lane * 32, __float_as_uint(b8[0]), __float_as_uint(b8[1]), __float_as_uint(b8[2]), __float_as_uint(b8[3]));
            sts128(cb_s + lane * 32 + 16, __float_as_uint(b8[4]), __float_as_uint(b8[5]), __float_as_uint(b8[6]), __float_as_uint(b8[7]));
            __syncwarp();
          }

          // ---- the residual does not depend on the accumulator: the loads of a chunk are issued one chunk ahead (those of the
          // first chunk before the wait for the accumulator), so their L2 latency hides behind the previous chunk's TMEM read,
          // arithmetic and stores.  (Partial chunks -- block_n % 32 == 16, pruned widths such as 170 = 5 * 32 + 10 -- keep the
          // vector accesses for their whole 16- / 8-column groups: element-wise tails cost more than the rest of the tile.)
          uint32_t resn[2][8];
          auto res_issue = [&](int c0) {
            const int nv = min(32, n_cols - c0);
            if (res_row && res_v32 && nv >= 16) {
              ldg256(res_row + c0, resn[0]);
              if (nv == 32) ldg256(res_row + c0 + 16, resn[1]);
            }
          };
          if (wg * 32 < n_cols && !(p.dbg_mode & 128)) res_issue(wg * 32);

          if (dbg_thread) {
            DBG_WAIT(3, mbar_wait(&tfull_bar[acc], acc_phase));
          } else {
            mbar_wait(&tfull_bar[acc], acc_phase);
          }
          tc_fence_after();
          const long long t_epi0 = dbg_thread ? clock64() : 0;
          const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);

          for (int c0 = wg * 32; c0 < n_cols; c0 += 32 * p.epi_groups) {
            const int nv = min(32, n_cols - c0);                  // valid columns in this chunk
            if (p.dbg_mode & 128) continue;
            uint32_t resv[2][8];
            const bool res_fast = res_row && res_v32 && nv >= 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) resv[0][j] = resn[0][j], resv[1][j] = resn[1][j];
            if (c0 + 32 * p.epi_groups < n_cols) res_issue(c0 + 32 * p.epi_groups);
            // ---- TMEM -> registers (lane == row).  (Issuing the next chunk's load early was measured: no gain.)
            uint32_t v[32];
            if (p.block_n - c0 >= 32) {
              tmem_ld_32x32(taddr + c0, v);
            } else {  // block_n % 32 == 16 tail
              tmem_ld_32x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
              for (int jj = 16; jj < 32; ++jj) v[jj] = 0;
            }
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
            if (use_cb) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = lds128f(cb_s + (c0 + g * 4) * 4);
                f[4 * g] += b4.x, f[4 * g + 1] += b4.y, f[4 * g + 2] += b4.z, f[4 * g + 3] += b4.w;
              }
            }
            if (rb_row) {
              if (rb_vec && nv == 32) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb_row + c0) + g);
                  f[4 * g] += b4.x, f[4 * g + 1] += b4.y, f[4 * g + 2] += b4.z, f[4 * g + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) f[j] += __ldg(rb_row + c0 + j);
              }
            }
            if (res_fast) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j < 8 || nv == 32) {
                  const uint32_t w = resv[j >> 3][j & 7];
                  f[2 * j] += __uint_as_float(w << 16);
                  f[2 * j + 1] += __uint_as_float(w & 0xffff0000u);
                }
              }
              if (nv < 32) {
#pragma unroll
                for (int j = 16; j < 32; ++j)
                  if (j < nv) f[j] += __bfloat162float(res_row[c0 + j]);
              }
            } else if (res_row) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) f[j] += __bfloat162float(res_row[c0 + j]);
            }
            if (!row_ok || (p.dbg_mode & 8)) continue;

            if (!p.out_fp32) {
              bf16* o = reinterpret_cast<bf16*>(p.out) + out_off + c0;
              if (nv == 32 && out_v16) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                  pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                if (out_v32) {
                  const uint32_t lo[8] = {pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]};
                  const uint32_t hi[8] = {pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]};
                  stg256(o, lo);
                  stg256(o + 16, hi);
                } else {
#pragma unroll
                  for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(o + g * 8) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
              } else if (out_v16) {   // partial chunk: whole 8-column groups as 128-/256-bit stores, the rest element-wise
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                  pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                int done = 0;
                if (out_v32 && nv >= 16) {
                  const uint32_t lo[8] = {pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]};
                  stg256(o, lo);
                  done = 16;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  if (8 * g >= done && 8 * g + 8 <= nv) {
                    *reinterpret_cast<uint4*>(o + g * 8) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    done = 8 * g + 8;
                  }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j >= done && j < nv) o[j] = __float2bfloat16(f[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) o[j] = __float2bfloat16(f[j]);
              }
            } else {
              float* o = reinterpret_cast<float*>(p.out) + out_off + c0;
              if (nv == 32 && out_v16) {
                if (p.accumulate) {
#pragma unroll
                  for (int g = 0; g < 8; ++g)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + g * 4), "f"(f[g * 4]),
                                 "f"(f[g * 4 + 1]), "f"(f[g * 4 + 2]), "f"(f[g * 4 + 3])
                                 : "memory");
                } else if (out_v32) {
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const uint32_t w8[8] = {__float_as_uint(f[8 * g]),     __float_as_uint(f[8 * g + 1]), __float_as_uint(f[8 * g + 2]),
                                            __float_as_uint(f[8 * g + 3]), __float_as_uint(f[8 * g + 4]), __float_as_uint(f[8 * g + 5]),
                                            __float_as_uint(f[8 * g + 6]), __float_as_uint(f[8 * g + 7])};
                    stg256(o + g * 8, w8);
                  }
                } else {
#pragma unroll
                  for (int g = 0; g < 8; ++g)
                    *reinterpret_cast<float4*>(o + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                }
              } else {
                int done = 0;
                if (out_v16) {
#pragma unroll
                  for (int g = 0; g < 8; ++g) {
                    if (4 * g + 4 <= nv) {
                      if (p.accumulate)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + g * 4), "f"(f[g * 4]),
                                     "f"(f[g * 4 + 1]), "f"(f[g * 4 + 2]), "f"(f[g * 4 + 3])
                                     : "memory");
                      else
                        *reinterpret_cast<float4*>(o + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                      done = 4 * g + 4;
                    }
                  }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j >= done && j < nv) {
                    if (p.accumulate)
                      atomicAdd(o + j, f[j]);
                    else
                      o[j] = f[j];
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (dbg_thread)
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 5), static_cast<unsigned long long>(clock64() - t_epi0));
          if (lane == 0) {
            if (pair && rank != 0)
              mbar_arrive_remote(&tempty_bar[acc], 0);   // tell the leader CTA's MMA thread
            else
              mbar_arrive(&tempty_bar[acc]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[4] += clock64() - t_kernel0;
  if (cluster > 1) cluster_sync_all();   // no CTA exits while a peer may still arrive on / multicast into its smem
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[9] = (long long)(gt - gt_entry);   // CTA 0 entry -> after the final cluster sync, ns
  }
  if (warp == 1) {
    tc_fence_after();
    if (pair)
      tmem_dealloc_2sm(tmem_base, kTmemCols);
    else
      tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Pixel box (bw, bh, bn) covering `pixels` consecutive output pixels (full rows), or failure.
static bool pixel_box(int pixels, int Ho, int Wo, int* bw, int* bh, int* bn) {
  if (Wo >= pixels) {
    if (Wo % pixels) return false;
    *bw = pixels, *bh = 1, *bn = 1;
    return true;
  }
  if (pixels % Wo) return false;
  int rows = pixels / Wo;
  if (Ho >= rows) {
    if (Ho % rows) return false;
    *bw = Wo, *bh = rows, *bn = 1;
    return true;
  }
  if (rows % Ho) return false;
  *bw = Wo, *bh = Ho, *bn = rows / Ho;
  return true;
}

// Row-shared 3x3 taps (GemmDev::kh3): image rows R covered by one CTA's 128 * m_sub consecutive output pixels, or 0 when the
// operand / tiling does not qualify (stride-1 "same" 3x3 window, whole image rows per CTA, no ragged last super tile).
static int kh3_rows(const b200pdm_operand& a, int64_t M, int m_sub, int cs) {
  static int env_off = -1;
  if (env_off < 0) env_off = getenv("B200PDM_NO_KH3") ? 1 : 0;
  if (env_off || a.mode != B200PDM_OP_CONV_ACT || a.taps != 9 || a.stride > 1 || a.no_pad) return 0;
  const int W = a.w_out, H = a.h_out, rows = kBlockM * m_sub;
  if (W < 8 || W > 128 || W % 8 || rows % W) return 0;
  const int R = rows / W;
  if (R < 1 || R > H || H % R || R + 2 > 256) return 0;
  if (M % ((int64_t)rows * cs)) return 0;
  return R;
}

static int pair_enabled() {
  static int env_pair = -1;
  if (env_pair < 0) env_pair = getenv("B200PDM_NO_PAIR") ? 0 : 1;
  return env_pair;
}

// Tile plan: block_n, tall tiles, split-K factor and pair mode chosen with a small cost model (cycles per CTA slot).
//   per k-block   max(MMA = 2*bn cycles per 128-row sub-tile, smem fill = bytes / ~40 B per cycle per SM  [measured
//                 L2->SM delivery limit; a tall tile shares each B stage between two sub-tiles])
//   per tile      k-blocks * that; the epilogue of a normal tile overlaps the next tile's main loop (second TMEM
//                 accumulator), a tall tile uses both accumulators, so its drain is exposed
//   total         waves over the 148 (or 74 pair) slots, + a finalize pass when a bf16 output is split along K
struct Plan {
  int bn = 0, splits = 1, pair = 0, m_sub = 1;
  double cost = 1e30;
};
static Plan plan_gemm(int64_t n, int n_groups, bool b_mn, int tiles_m, int Z, int kblocks, bool can_split,
                      bool split_needs_finalize, int fixed_bn, int fixed_splits = 0, int bn_step = 0,
                      const b200pdm_operand* conv_a = nullptr, int64_t M = 0) {
  static int env_msub = -1;
  if (env_msub < 0) {
    const char* e = getenv("B200PDM_MSUB");
    env_msub = e ? atoi(e) : 0;
  }
#ifdef B200PDM_DIAG   // plan sweeps (tools/sweep_gemm_plans.py): B200PDM_PLAN="bn,m_sub,pair,splits", 0 / -1 = planner's choice
  int f_bn = 0, f_msub = 0, f_pair = -1, f_splits = 0;
  if (const char* e = getenv("B200PDM_PLAN")) sscanf(e, "%d,%d,%d,%d", &f_bn, &f_msub, &f_pair, &f_splits);
  if (f_bn > 0 && fixed_bn <= 0) fixed_bn = f_bn;
  if (f_splits > 0 && fixed_splits <= 0 && (can_split || f_splits == 1)) fixed_splits = f_splits;
#endif
  const int g = bn_step > 0 ? bn_step : (b_mn ? 64 : 16);   // (fused GEGLU: halves of 32-column chunks -> steps of 64)
  const int n_pad = static_cast<int>((n + g - 1) / g * g);
  static const int split_cands[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
  Plan best;
  for (int bn = g; bn <= 256; bn += g) {
    if (fixed_bn > 0 && bn != fixed_bn) continue;
    if (fixed_bn <= 0 && bn > n_pad) break;
    if (fixed_bn <= 0 && bn < 64 && bn != n_pad) continue;   // tiny tiles only when N itself is tiny
    const int tiles_n = static_cast<int>((n + bn - 1) / bn) * n_groups;
    int pair = (pair_enabled() && tiles_m >= 2 && (b_mn ? ((bn / 64) % 2 == 0) : true)) ? 1 : 0;
#ifdef B200PDM_DIAG
    if (f_pair == 0) pair = 0;
    if (f_pair == 1 && !pair) continue;
#endif
    const int cs = pair ? 2 : 1;
    const int slots = 148 / cs;
    const double epi_unit = 400.0 * ((bn + 31) / 32) / 2.0;   // two epilogue warpgroups
    for (int m_sub = 1; m_sub <= 2; ++m_sub) {
      if (env_msub > 0 && m_sub != env_msub) continue;
#ifdef B200PDM_DIAG
      if (f_msub > 0 && m_sub != f_msub) continue;
#endif
      if (m_sub == 2 && tiles_m < 2 * cs) break;
      const long base_tiles = (long)((tiles_m + cs * m_sub - 1) / (cs * m_sub)) * tiles_n * Z;
      double a_bytes = 16384.0 * m_sub;
      if (conv_a) {   // row-shared 3x3 taps: one (R + 2)-row box per three k-blocks
        const int R = kh3_rows(*conv_a, M, m_sub, cs);
        if (R > 0) a_bytes = (R + 2) * conv_a->w_out * 128.0 / 3.0;
      }
      const double bytes = a_bytes + (pair ? bn / 2 : bn) * 128.0;
      const double kcyc = std::max(2.0 * bn * m_sub, bytes / 40.0) + 40.0;
      for (int s : split_cands) {
        if (fixed_splits > 0) {
          if (s != 1) break;
          s = fixed_splits;
        } else {
          if (s > 1 && !can_split) break;
          if (s > 1 && kblocks / s < 8) break;
        }
        const int kb = (kblocks + s - 1) / s;
        const long tiles = base_tiles * s;
        const long waves = (tiles + slots - 1) / slots;
        double tile_cyc = (m_sub == 1 ? std::max(kb * kcyc, epi_unit) : kb * kcyc + 4.0 * epi_unit) + 800.0;
        if (s > 1) tile_cyc += epi_unit * m_sub;   // atomics epilogue is slower and less overlapped
        double cost = waves * tile_cyc + (m_sub == 1 ? epi_unit : 0.0);
        if (s > 1 && split_needs_finalize) cost += 14000.0;
        // accumulating fp32 outputs (weight gradients): every tile of every split adds its block to the output with
        // red.global.add -- a pass over the output per split, shared by all SMs.  Fitted on the brute-force plan sweep of the
        // step's wgrad shapes (tools/sweep_gemm_plans.py, profiles/r2_gemm_plan_sweep.txt): regret 4.2 % -> 2.2 %.
        if (can_split && !split_needs_finalize) cost += 0.12 * (double)tiles * bn * m_sub * 128.0 / 148.0;
        if (cost < best.cost) best.bn = bn, best.splits = s, best.pair = pair, best.m_sub = m_sub, best.cost = cost;
      }
    }
  }
  return best;
}

static int build_operand_map(const b200pdm_operand& op, bool is_a, int block_n, CUtensorMap* map, OpDev* dev,
                             int64_t mn_extent, int64_t k_extent, int Z1, int Z2, int cluster = 1, int kh3_R = 0) {
  memset(dev, 0, sizeof(*dev));
  dev->mode = op.mode;
  dev->Z1 = Z1 > 0 ? Z1 : 1;
  dev->taps = op.taps > 0 ? op.taps : 1;
  dev->stride = op.stride > 0 ? op.stride : 1;
  dev->flip = op.flip;
  dev->pad = op.no_pad ? 0 : 1;
  const int rows = is_a ? kBlockM : block_n / cluster;   // K-major B: each CTA loads (and multicasts) its slice
  uint64_t dims[5], str[5];
  uint32_t box[5], es[5] = {1, 1, 1, 1, 1};
  switch (op.mode) {
    case B200PDM_OP_K2D: {
      dims[0] = k_extent, dims[1] = mn_extent, dims[2] = Z1, dims[3] = Z2;
      str[0] = 1, str[1] = op.ld, str[2] = (Z1 > 1 ? op.bs1 : op.ld), str[3] = (Z2 > 1 ? op.bs2 : op.ld);
      box[0] = 64, box[1] = rows, box[2] = 1, box[3] = 1;
      return make_map(map, op.ptr, 4, dims, str, box, es);
    }
    case B200PDM_OP_MN2D: {
      dims[0] = mn_extent, dims[1] = k_extent, dims[2] = Z1, dims[3] = Z2;
      str[0] = 1, str[1] = op.ld, str[2] = (Z1 > 1 ? op.bs1 : op.ld), str[3] = (Z2 > 1 ? op.bs2 : op.ld);
      box[0] = 64, box[1] = 64, box[2] = 1, box[3] = 1;
      return make_map(map, op.ptr, 4, dims, str, box, es);
    }
    case B200PDM_OP_CONV_ACT:
    case B200PDM_OP_CONV_ACT_MN: {
      const int pixels = (op.mode == B200PDM_OP_CONV_ACT) ? kBlockM : kBlockK;
      int bw, bh, bn;
      if (!pixel_box(pixels, op.h_out, op.w_out, &bw, &bh, &bn)) {
        set_err("conv: output grid does not tile into full-row pixel boxes");
        return B200PDM_ERR_UNSUPPORTED;
      }
      dev->Wo = op.w_out;
      dev->HoWo = op.h_out * op.w_out;
      dev->fd_Wo = make_fastdiv(dev->Wo);
      dev->fd_HoWo = make_fastdiv(dev->HoWo);
      dev->cblks = (op.channels + 63) / 64;
      dims[0] = op.channels, dims[1] = op.w_in, dims[2] = op.h_in, dims[3] = op.batch;
      str[0] = 1, str[1] = op.ld, str[2] = (uint64_t)op.w_in * op.ld, str[3] = (uint64_t)op.h_in * op.w_in * op.ld;
      box[0] = 64, box[1] = bw * dev->stride, box[2] = bh * dev->stride, box[3] = bn;
      es[1] = dev->stride, es[2] = dev->stride;
      if (kh3_R > 0) box[1] = op.w_out, box[2] = kh3_R + 2, box[3] = 1;   // R + 2 whole image rows (GemmDev::kh3)
      if (dev->taps == 1 && dev->stride == 1) {
        // 1x1: the (kw-1, kh-1) shift is zero because load_a/load_b force the centre tap
      }
      return make_map(map, op.ptr, 4, dims, str, box, es);
    }
    case B200PDM_OP_CONV_W: {
      dev->cblks = (op.channels + 63) / 64;
      dims[0] = op.channels, dims[1] = dev->taps, dims[2] = op.out_channels;
      str[0] = 1, str[1] = op.ld, str[2] = (uint64_t)dev->taps * op.ld;
      box[0] = 64, box[1] = 1, box[2] = rows;
      return make_map(map, op.ptr, 3, dims, str, box, es);
    }
    case B200PDM_OP_CONV_WT: {
      dev->cblks = (op.out_channels + 63) / 64;  // K runs over O
      dims[0] = op.channels, dims[1] = dev->taps, dims[2] = op.out_channels;
      str[0] = 1, str[1] = op.ld, str[2] = (uint64_t)dev->taps * op.ld;
      box[0] = 64, box[1] = 1, box[2] = 64;
      return make_map(map, op.ptr, 3, dims, str, box, es);
    }
  }
  set_err("unknown operand mode");
  return B200PDM_ERR_ARG;
}

// Split-K of a NON-accumulating output (forward / dgrad GEMMs of the 8x8 / 16x16 levels, where one wave would leave most SMs
// idle): split s writes its partial tile with plain stores into its own fp32 slab of the CALLER's workspace
// (b200pdm_*_workspace() bytes), and one finalize pass adds the slabs in split order, applies bias / time-embedding /
// residual and converts.  No atomics, no memset, nothing allocated here: bit-identical from run to run, and safe under
// CUDA-graph capture and with any number of streams (round 1 kept a library-owned, lazily re-allocated scratch per "lane").
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int64_t ldws, int splits, int64_t slab,
                                       void* __restrict__ out, int out_fp32,
                                       int64_t ldo, const float* __restrict__ bias, const float* __restrict__ rowbias,
                                       int64_t ld_rowbias, int rows_per_group, const bf16* __restrict__ residual,
                                       int64_t ldr, int64_t M, int N) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int nq = (N + 3) / 4;
  const int64_t total = M * nq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / nq;
    const int n0 = (int)(i - m * nq) * 4;
    float4 a = *reinterpret_cast<const float4*>(ws + m * ldws + n0);
    for (int sp = 1; sp < splits; ++sp) {   // fixed order: deterministic
      const float4 b = *reinterpret_cast<const float4*>(ws + sp * slab + m * ldws + n0);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    float f[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j;
      if (n < N) {
        if (bias) f[j] += bias[n];
        if (rowbias) f[j] += rowbias[(m / rows_per_group) * ld_rowbias + n];
        if (residual) f[j] += __bfloat162float(residual[m * ldr + n]);
        if (out_fp32)
          reinterpret_cast<float*>(out)[m * ldo + n] = f[j];
        else
          reinterpret_cast<bf16*>(out)[m * ldo + n] = __float2bfloat16(f[j]);
      }
    }
  }
}

static int launch_finalize(const b200pdm_gemm_desc* d, const float* ws, int64_t ldws, int splits, int64_t slab,
                           cudaStream_t stream) {
  const int64_t total = d->M * ((d->N + 3) / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(splitk_finalize_kernel, (int)blocks, 256, 0, stream, ws, ldws, splits, slab, d->out, d->out_fp32, d->ldo, d->bias, d->rowbias,
                                                         d->ld_rowbias, d->rows_per_group > 0 ? d->rows_per_group : 1,
                                                         reinterpret_cast<const bf16*>(d->residual), d->ldr, d->M,
                                                         (int)d->N);
  if (cudaGetLastError() != cudaSuccess) {
    set_err("split-K finalize launch failed");
    return B200PDM_ERR_CUDA;
  }
  g_launches++;
  return B200PDM_OK;
}

// Optional per-shape timing table (B200PDM_GEMM_TRACE=1): every launch is bracketed by events and synchronised, so it
// serialises the stream -- diagnostics only, never enabled in bench.py.
struct TraceRow {
  long count = 0;
  double ms = 0, flops = 0;
};
static std::map<std::string, TraceRow> g_trace;
static int g_trace_on = -1;
static int trace_on() {
  if (g_trace_on < 0) g_trace_on = getenv("B200PDM_GEMM_TRACE") ? 1 : 0;
  return g_trace_on;
}

// Shape of the tile problem a descriptor poses + the plan chosen for it (shared by the launch and the workspace query, so
// both see the same split decision).
struct Shaped {
  bool a_mn, b_mn, acc_out;
  int Z1, Z2, n_groups, n_per_group, tiles_m, Z, kblocks;
  Plan plan;
  int64_t ldws, slab;   // split-K slabs: row pitch / elements per slab (0 when the plan needs no workspace)
};
static int shape_and_plan(const b200pdm_gemm_desc* d, bool have_workspace, Shaped* s, bool slab_pass = false) {
  s->a_mn = d->a.mode == B200PDM_OP_MN2D;
  s->b_mn = d->b.mode == B200PDM_OP_MN2D || d->b.mode == B200PDM_OP_CONV_WT || d->b.mode == B200PDM_OP_CONV_ACT_MN;
  s->Z1 = d->Z1 > 0 ? d->Z1 : 1, s->Z2 = d->Z2 > 0 ? d->Z2 : 1;
  s->n_groups = 1, s->n_per_group = static_cast<int>(d->geglu ? 2 * d->N : d->N);   // GEGLU: value + gate columns
  if (d->b.mode == B200PDM_OP_CONV_ACT_MN) {   // conv wgrad: one N group per tap
    s->n_groups = d->b.taps > 0 ? d->b.taps : 1;
    s->n_per_group = d->b.channels;
  }
  s->tiles_m = cdiv(d->M, kBlockM);
  s->Z = s->Z1 * s->Z2;
  if (d->a.mode == B200PDM_OP_CONV_ACT) {
    const int taps = d->a.taps > 0 ? d->a.taps : 1;
    s->kblocks = taps * cdiv(d->a.channels, 64);
  } else if (d->b.mode == B200PDM_OP_CONV_ACT_MN) {
    s->kblocks = cdiv((int64_t)d->b.batch * d->b.h_out * d->b.w_out, 64);
  } else {
    s->kblocks = cdiv(d->K, 64);
  }
  if (s->kblocks <= 0 || d->M <= 0 || d->N <= 0) {
    set_err("gemm: empty problem");
    return B200PDM_ERR_ARG;
  }
  s->acc_out = d->out_fp32 && d->accumulate;
  const bool slabs_ok = have_workspace && !s->acc_out && s->Z == 1 && s->n_groups == 1 &&
                        (int64_t)d->M * d->N * 4 <= (64ll << 20);
  const b200pdm_operand* conv_a = (d->a.mode == B200PDM_OP_CONV_ACT && d->a.taps == 9) ? &d->a : nullptr;
  if (d->geglu)   // one tile = bh value + bh gate columns of the same hidden units; never split along K
    s->plan = plan_gemm(s->n_per_group, 1, false, s->tiles_m, s->Z, s->kblocks, false, false, d->block_n, 0, 64);
  else if (d->splits > 1 && (s->acc_out || slab_pass))   // split factor fixed by the caller (accumulating outputs) / the slab pass
    s->plan = plan_gemm(s->n_per_group, s->n_groups, s->b_mn, s->tiles_m, s->Z, s->kblocks, false, false, d->block_n, d->splits,
                        0, conv_a, d->M);
  else
    s->plan = plan_gemm(s->n_per_group, s->n_groups, s->b_mn, s->tiles_m, s->Z, s->kblocks, s->acc_out || slabs_ok,
                        !s->acc_out, d->block_n, 0, 0, conv_a, d->M);
  if (s->plan.bn <= 0) {
    set_err("gemm: no valid tile plan (bad block_n?)");
    return B200PDM_ERR_ARG;
  }
  {   // normalise the split factor to what the kernel will really run (ceil division can leave the last splits empty): the
      // slab count the finalize pass adds up must equal the number of slabs that get written
    int sp = s->plan.splits > s->kblocks ? s->kblocks : s->plan.splits;
    if (sp < 1) sp = 1;
    int per = cdiv(s->kblocks, sp);
    if (conv_a) per = (per + 2) / 3 * 3;   // 3x3 taps: a split owns whole (channel block, kw) groups of three k-blocks
    s->plan.splits = cdiv(s->kblocks, per);
  }
  s->ldws = s->slab = 0;
  if (s->plan.splits > 1 && !s->acc_out && !slab_pass) {
    s->ldws = (d->N + 7) / 8 * 8;
    s->slab = d->M * s->ldws;
  }
  return B200PDM_OK;
}

static size_t gemm_workspace_bytes(const b200pdm_gemm_desc* d) {
  Shaped s;
  if (!d || shape_and_plan(d, true, &s) != B200PDM_OK) return 0;
  return (size_t)s.slab * s.plan.splits * sizeof(float);
}

static int launch_gemm(const b200pdm_gemm_desc* d, void* workspace, size_t ws_bytes, cudaStream_t stream,
                       int64_t split_stride = 0);
static int launch_gemm(const b200pdm_gemm_desc* d, void* workspace, size_t ws_bytes, cudaStream_t stream,
                       int64_t split_stride) {
  if (!d || !d->a.ptr || !d->b.ptr || !d->out) {
    set_err("gemm: null pointer");
    return B200PDM_ERR_ARG;
  }
  if (d->geglu && (d->a.mode != B200PDM_OP_K2D || d->b.mode != B200PDM_OP_K2D || d->out_fp32 || d->rowbias || d->residual ||
                   d->accumulate || (d->Z1 > 1) || (d->Z2 > 1))) {
    set_err("gemm: the fused GEGLU epilogue takes plain K-major operands, a bf16 output and an optional bias only");
    return B200PDM_ERR_ARG;
  }
  Shaped sh;
  int rc0 = shape_and_plan(d, workspace != nullptr, &sh, split_stride != 0);
  if (rc0) return rc0;
  const bool a_mn = sh.a_mn, b_mn = sh.b_mn, acc_out = sh.acc_out;
  const int Z1 = sh.Z1, Z2 = sh.Z2, kblocks = sh.kblocks;
  const Plan plan = sh.plan;

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.n_groups = sh.n_groups;
  p.n_per_group = sh.n_per_group;
  p.out_group_stride = d->b.mode == B200PDM_OP_CONV_ACT_MN ? d->ldo / p.n_groups : 0;  // dW row = [taps][I_ld]
  p.out_split_stride = split_stride;
  p.M = static_cast<int>(d->M);
  p.tiles_m = sh.tiles_m;
  p.Z = sh.Z;
  p.kblocks = kblocks;

  if (plan.splits > 1 && !acc_out && split_stride == 0) {
    const size_t need = (size_t)sh.slab * plan.splits * sizeof(float);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 31)) {
      set_err("gemm: workspace smaller than b200pdm_*_workspace() asked for, or not 32-byte aligned");
      return B200PDM_ERR_ARG;
    }
    float* ws = reinterpret_cast<float*>(workspace);
    b200pdm_gemm_desc d2 = *d;
    d2.out = ws, d2.out_fp32 = 1, d2.ldo = sh.ldws, d2.accumulate = 0, d2.splits = plan.splits, d2.block_n = plan.bn;
    d2.bias = nullptr, d2.rowbias = nullptr, d2.residual = nullptr;
    int rc2 = launch_gemm(&d2, nullptr, 0, stream, sh.slab);
    if (rc2) return rc2;
    return launch_finalize(d, ws, sh.ldws, plan.splits, sh.slab, stream);
  }
  int block_n = plan.bn;
  p.block_n = block_n;
  p.tiles_n_per_group = cdiv(p.n_per_group, block_n);
  int splits = plan.splits;
  if (splits > kblocks) splits = kblocks;
  p.kb_per_split = cdiv(kblocks, splits);
  if (d->a.mode == B200PDM_OP_CONV_ACT && d->a.taps == 9) p.kb_per_split = (p.kb_per_split + 2) / 3 * 3;   // as in shape_and_plan
  splits = cdiv(kblocks, p.kb_per_split);
  p.splits = splits;

  // cluster size: share the B tile among `cluster` m-tiles (TMA multicast); measured no gain on B200 -> default 1
  static int env_cluster = -1;
  if (env_cluster < 0) {
    const char* e = getenv("B200PDM_CLUSTER");
    env_cluster = e ? atoi(e) : 1;
    if (env_cluster != 1 && env_cluster != 2 && env_cluster != 4) env_cluster = 1;
  }
  int cluster = env_cluster;
  while (cluster > 1) {
    const bool div_ok = b_mn ? ((block_n / 64) % cluster == 0) : ((block_n / cluster) % 8 == 0);
    if (div_ok && p.tiles_m >= cluster) break;
    cluster >>= 1;
  }
  // cta_group::2 pairs: halve the B bytes each SM has to receive per MMA (the L2->SM path is the limiter)
  p.pair = plan.pair;
  if (p.pair) cluster = 2;
  p.cluster = p.pair ? 1 : cluster;
  p.m_sub = plan.m_sub;
  p.tiles_m_super = cdiv(p.tiles_m, cluster * p.m_sub);

  const int kh3_R = kh3_rows(d->a, d->M, p.m_sub, cluster);
  p.rank_stride = kh3_R ? p.m_sub : 1;
  p.sub_stride = kh3_R ? 1 : cluster;
  CUtensorMap map_a, map_b;
  int rc = build_operand_map(d->a, true, block_n, &map_a, &p.a, d->M, d->K, Z1, Z2, 1, kh3_R);
  if (rc) return rc;
  // (GEGLU: the weight has 2F rows and is always fetched in half-tile boxes, one for the value and one for the gate rows)
  rc = build_operand_map(d->b, false, block_n, &map_b, &p.b, d->geglu ? 2 * d->N : d->N, d->K, Z1, Z2, d->geglu ? 2 : cluster);
  if (rc) return rc;

  {
    const int tiles_n = p.tiles_n_per_group * p.n_groups;
    p.k_taps = (d->a.mode == B200PDM_OP_CONV_ACT && p.a.taps > 1) ? p.a.taps : 1;
    if ((d->b.mode == B200PDM_OP_CONV_W || d->b.mode == B200PDM_OP_CONV_WT) && p.b.taps != p.k_taps) {
      set_err("gemm: operand tap counts disagree");
      return B200PDM_ERR_ARG;
    }
    p.fd_tiles_per_split = make_fastdiv((int64_t)p.tiles_m_super * tiles_n * p.Z);
    p.fd_slab = make_fastdiv((int64_t)p.tiles_m_super * tiles_n);
    p.fd_tiles_n = make_fastdiv(tiles_n);
    p.fd_tiles_n_per_group = make_fastdiv(p.tiles_n_per_group);
    p.fd_Z1 = make_fastdiv(p.a.Z1);
    p.fd_taps = make_fastdiv(p.k_taps);
    p.fd_rows_per_group = make_fastdiv(d->rows_per_group > 0 ? d->rows_per_group : 1);
  }

  {
    static int env_eg = -1;
    if (env_eg < 0) {
      const char* e = getenv("B200PDM_EPI_GROUPS");
      env_eg = e ? atoi(e) : 0;
    }
    p.epi_groups = kMaxEpiGroups;   // measured: never slower than one group, up to 1.4x faster on small-K shapes
    if (env_eg >= 1 && env_eg <= kMaxEpiGroups) p.epi_groups = env_eg;
  }
  int stage_bytes = p.m_sub * kStageABytes + (p.pair ? block_n / 2 : block_n) * 128;
  const int fixed_bytes = 1024 + (2 * 8 + 6 + 8) * 8 + 64 + 128 + 4 * p.epi_groups * 1024;
  int a_region = 0;
  if (kh3_R > 0) {   // A ring of 3 (else 2) boxes, the rest of shared memory for B stages (at least 3)
    const int a_slot = (kh3_R + 2) * d->a.w_out * 128, b_stage = (p.pair ? block_n / 2 : block_n) * 128;
    const int room = 227 * 1024 - fixed_bytes;
    int slots = 3;
    if ((room - slots * a_slot) / b_stage < 4) slots = 2;
    if ((room - slots * a_slot) / b_stage >= 3) {
      p.kh3 = 1, p.a_slots = slots, p.a_slot_bytes = a_slot, p.kh_row_bytes = d->a.w_out * 128;
      a_region = slots * a_slot, stage_bytes = b_stage;
    } else {   // does not fit: fall back to one box per tap (needs the standard tensor map and row mapping)
      p.rank_stride = 1, p.sub_stride = cluster;
      rc = build_operand_map(d->a, true, block_n, &map_a, &p.a, d->M, d->K, Z1, Z2);
      if (rc) return rc;
    }
  }
  int stages = (227 * 1024 - fixed_bytes - a_region) / stage_bytes;
  if (stages > 8) stages = 8;
  static int env_stages = -1;
  if (env_stages < 0) {
    const char* e = getenv("B200PDM_STAGES");
    env_stages = e ? atoi(e) : 0;
  }
  if (env_stages > 0 && stages > env_stages) stages = env_stages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.idesc = make_idesc_bf16(block_n, a_mn ? 1 : 0, b_mn ? 1 : 0, p.pair ? 256 : 128);

  p.out = d->out;
  p.out_fp32 = d->out_fp32;
  p.ldo = d->ldo;
  p.obs1 = d->obs1;
  p.obs2 = d->obs2;
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.ld_rowbias = d->ld_rowbias;
  p.rows_per_group = d->rows_per_group > 0 ? d->rows_per_group : 1;
  p.residual = reinterpret_cast<const bf16*>(d->residual);
  p.ldr = d->ldr;
  p.rbs1 = d->rbs1;
  p.rbs2 = d->rbs2;
  p.alpha = d->alpha;
  p.accumulate = d->accumulate;
  p.geglu_F = d->geglu ? static_cast<int>(d->N) : 0;
  p.aux = reinterpret_cast<bf16*>(d->aux);
  p.ld_aux = d->ld_aux;
#ifdef B200PDM_DIAG
  static long long* dbg_buf = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) dbg_on = getenv("B200PDM_GEMM_DBG") ? 1 : 0;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(long long), stream);
  }
  p.dbg = dbg_on ? dbg_buf : nullptr;
#endif
  const size_t smem = (size_t)fixed_bytes + (size_t)a_region + (size_t)stages * stage_bytes;
  const int threads = 64 + 128 * p.epi_groups;
#ifdef B200PDM_DIAG
  {
    const char* e = getenv("B200PDM_GEMM_DBGMODE");   // re-read every launch so a script can toggle it
    p.dbg_mode = e ? atoi(e) : 0;
  }
#endif

  const long total_tiles = (long)p.tiles_m_super * p.tiles_n_per_group * p.n_groups * p.Z * p.splits;  // per cluster
  const int max_clusters = num_sms() / cluster;
  int grid = static_cast<int>(total_tiles < max_clusters ? total_tiles : max_clusters) * cluster;
  {
    const char* e = getenv("B200PDM_GRID");   // diagnostics: cap the number of CTAs
    if (e && atoi(e) > 0 && atoi(e) < grid) grid = atoi(e) / cluster * cluster;
  }

  auto launch = [&](auto kern) -> int {
    static std::map<const void*, bool> attr_set;   // once per kernel instantiation (not per launch: graph capture)
    cudaError_t e = cudaSuccess;
    if (!attr_set[reinterpret_cast<const void*>(kern)]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        set_err("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return B200PDM_ERR_CUDA;
      }
      attr_set[reinterpret_cast<const void*>(kern)] = true;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (trace_on()) {
      cudaEventCreate(&t0), cudaEventCreate(&t1);
      cudaEventRecord(t0, stream);
    }
    {
      static int use_pdl = -1;
      if (use_pdl < 0) use_pdl = getenv("B200PDM_NO_PDL") ? 0 : 1;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      int na = 0;
      if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
        ++na;
      }
      if (use_pdl) {   // may become resident during the predecessor's tail; the kernel calls griddepcontrol.wait itself
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      cfg.attrs = attr, cfg.numAttrs = na;
      cudaLaunchKernelEx(&cfg, kern, map_a, map_b, p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_err("gemm launch: %s", cudaGetErrorString(e));
      return B200PDM_ERR_CUDA;
    }
    g_launches++;
    if (trace_on()) {
      cudaEventRecord(t1, stream);
      cudaEventSynchronize(t1);
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, t1);
      char key[256];
      const double kk = (double)kblocks * 64;
      snprintf(key, sizeof(key), "a%d b%d M=%lld N=%d(x%d) K=%.0f Z=%d bn=%d msub=%d split=%d tiles=%ld grid=%d stages=%d cl=%d kh3=%d",
               d->a.mode, d->b.mode, (long long)d->M, p.n_per_group, p.n_groups, kk, p.Z, block_n, p.m_sub, p.splits,
               total_tiles, grid, stages, cluster, p.kh3 ? p.a_slots : 0);
#ifdef B200PDM_DIAG
      if (p.dbg) {
        long long h[16];
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        int tiles_cta0 = (int)((total_tiles + grid / cluster - 1) / (grid / cluster));
        fprintf(stderr, "[gemm dbg] %s | cta0: total=%lld prod_wait_empty=%lld mma_wait_full=%lld mma_wait_tempty=%lld "
                "epi_wait_tfull=%lld epi_busy=%lld tma_issue=%lld mma_issue=%lld (~%d tiles, %d kblocks) | prologue=%lld ns, "
                "cta0 lifetime=%lld ns, kernel (events)=%.1f us\n", key, h[4], h[0], h[1], h[2], h[3], h[5], h[6], h[7], tiles_cta0,
                p.kb_per_split, h[8], h[9], ms * 1e3);
      }
#endif
      TraceRow& r = g_trace[key];
      r.count++, r.ms += ms;
      r.flops += 2.0 * (double)d->M * p.n_per_group * p.n_groups * kk * p.Z;
      cudaEventDestroy(t0), cudaEventDestroy(t1);
    }
    return B200PDM_OK;
  };
  if (p.pair) {
    if (!a_mn && !b_mn) return launch(gemm_kernel<0, 0, true>);
    if (!a_mn && b_mn) return launch(gemm_kernel<0, 1, true>);
    if (a_mn && !b_mn) return launch(gemm_kernel<1, 0, true>);
    return launch(gemm_kernel<1, 1, true>);
  }
  if (!a_mn && !b_mn) return launch(gemm_kernel<0, 0, false>);
  if (!a_mn && b_mn) return launch(gemm_kernel<0, 1, false>);
  if (a_mn && !b_mn) return launch(gemm_kernel<1, 0, false>);
  return launch(gemm_kernel<1, 1, false>);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200pdm_version(void) { return 100; }
const char* b200pdm_last_error(void) { return get_err(); }
uint64_t b200pdm_launch_count(void) { return g_launches.load(); }
// Host-only query of the tile planner (no device work): what launch_gemm would choose for a GEMM whose N per group is `n`,
// with `tiles_m` 128-row tiles, `Z` batch slabs and `kblocks` 64-wide K blocks.  out = {block_n, splits, pair, m_sub, stages,
// tiles (CTA-slot units incl. splits), slots (148 or 74 pairs)}.  Used by the CPU tests and tools/plan_report.py.
int b200pdm_gemm_plan(int64_t n, int n_groups, int b_mn, int tiles_m, int Z, int kblocks, int can_split,
                      int split_needs_finalize, int* out) {
  if (!out || n <= 0 || n_groups <= 0 || tiles_m <= 0 || Z <= 0 || kblocks <= 0) return B200PDM_ERR_ARG;
  const Plan plan = plan_gemm(n, n_groups, b_mn != 0, tiles_m, Z, kblocks, can_split != 0, split_needs_finalize != 0, 0);
  if (plan.bn <= 0) return B200PDM_ERR_ARG;
  const int epi_groups = kMaxEpiGroups;
  const int stage_bytes = plan.m_sub * kStageABytes + (plan.pair ? plan.bn / 2 : plan.bn) * 128;
  const int fixed_bytes = 1024 + (2 * 8 + 6) * 8 + 64 + 128 + 4 * epi_groups * 1024;
  int stages = (227 * 1024 - fixed_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const int cs = plan.pair ? 2 : 1;
  const long tiles_n = (long)((n + plan.bn - 1) / plan.bn) * n_groups;
  const long tiles = (long)((tiles_m + cs * plan.m_sub - 1) / (cs * plan.m_sub)) * tiles_n * Z * plan.splits;
  out[0] = plan.bn, out[1] = plan.splits, out[2] = plan.pair, out[3] = plan.m_sub, out[4] = stages;
  out[5] = (int)tiles, out[6] = 148 / cs;
  return B200PDM_OK;
}

int b200pdm_gemm_trace_enable(int on) {
  g_trace_on = on ? 1 : 0;
  if (on) g_trace.clear();
  return B200PDM_OK;
}
int b200pdm_gemm_trace_totals(double* ms, double* flops, int64_t* launches) {
  double m = 0, f = 0;
  int64_t n = 0;
  for (auto& kv : g_trace) m += kv.second.ms, f += kv.second.flops, n += kv.second.count;
  if (ms) *ms = m;
  if (flops) *flops = f;
  if (launches) *launches = n;
  return B200PDM_OK;
}
int b200pdm_gemm_trace_dump(const char* path) {
  FILE* f = fopen(path, "w");
  if (!f) return B200PDM_ERR_ARG;
  fprintf(f, "ms_total\tcount\tms_avg\ttflops\tshape\n");
  for (auto& kv : g_trace)
    fprintf(f, "%.4f\t%ld\t%.4f\t%.1f\t%s\n", kv.second.ms, kv.second.count, kv.second.ms / kv.second.count,
            kv.second.flops / (kv.second.ms * 1e-3) / 1e12, kv.first.c_str());
  fclose(f);
  g_trace.clear();
  return B200PDM_OK;
}

size_t b200pdm_gemm_workspace(const b200pdm_gemm_desc* desc) { return gemm_workspace_bytes(desc); }

int b200pdm_gemm(const b200pdm_gemm_desc* desc, void* workspace, size_t ws_bytes, b200pdm_stream_t stream) {
  return launch_gemm(desc, workspace, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Descriptor builders shared by each entry point and its *_workspace twin (pointers may be null in the query).
static void desc_linear_fwd(b200pdm_gemm_desc& d, const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                            const void* residual, int64_t ldr, void* out, int64_t ldo, int out_fp32, int64_t M, int64_t N,
                            int64_t K) {
  memset(&d, 0, sizeof(d));
  d.a.mode = B200PDM_OP_K2D, d.a.ptr = x, d.a.ld = ldx;
  d.b.mode = B200PDM_OP_K2D, d.b.ptr = w, d.b.ld = ldw;
  d.M = M, d.N = N, d.K = K, d.Z1 = 1, d.Z2 = 1;
  d.out = out, d.out_fp32 = out_fp32, d.ldo = ldo;
  d.bias = bias, d.residual = residual, d.ldr = ldr, d.alpha = 1.f;
}
static void desc_linear_dgrad(b200pdm_gemm_desc& d, const void* dy, int64_t lddy, const void* w, int64_t ldw,
                              const void* residual, int64_t ldr, void* dx, int64_t lddx, int64_t M, int64_t N, int64_t K) {
  // dx[M,K] = dy[M,N] . w[N,K]: reduction over N; B(n'=k, k'=n) = w[n][k] is MN-major.
  memset(&d, 0, sizeof(d));
  d.a.mode = B200PDM_OP_K2D, d.a.ptr = dy, d.a.ld = lddy;
  d.b.mode = B200PDM_OP_MN2D, d.b.ptr = w, d.b.ld = ldw;
  d.M = M, d.N = K, d.K = N, d.Z1 = 1, d.Z2 = 1;
  d.out = dx, d.out_fp32 = 0, d.ldo = lddx;
  d.residual = residual, d.ldr = ldr, d.alpha = 1.f;
}
static void desc_conv_fwd(b200pdm_gemm_desc& d, const void* x, int64_t ldx, const void* w, int64_t w_ild, const float* bias,
                          const float* rowbias, int64_t ld_rowbias, const void* residual, int64_t ldr, void* out, int64_t ldo,
                          int batch, int h_in, int w_in, int c_in, int c_out, int ksize, int stride, int no_pad = 0) {
  const int h_out = h_in / stride, w_out = w_in / stride;
  memset(&d, 0, sizeof(d));
  d.a.mode = B200PDM_OP_CONV_ACT, d.a.ptr = x, d.a.ld = ldx;
  d.a.batch = batch, d.a.h_in = h_in, d.a.w_in = w_in, d.a.channels = c_in;
  d.a.h_out = h_out, d.a.w_out = w_out, d.a.stride = stride, d.a.taps = ksize * ksize, d.a.no_pad = no_pad;
  d.b.mode = B200PDM_OP_CONV_W, d.b.ptr = w, d.b.ld = w_ild, d.b.channels = c_in, d.b.out_channels = c_out;
  d.b.taps = ksize * ksize;
  d.M = (int64_t)batch * h_out * w_out, d.N = c_out, d.K = (int64_t)ksize * ksize * c_in, d.Z1 = 1, d.Z2 = 1;
  d.out = out, d.out_fp32 = 0, d.ldo = ldo;
  d.bias = bias, d.rowbias = rowbias, d.ld_rowbias = ld_rowbias, d.rows_per_group = h_out * w_out;
  d.residual = residual, d.ldr = ldr, d.alpha = 1.f;
}
static void desc_conv_dgrad(b200pdm_gemm_desc& d, const void* dy, int64_t lddy, const void* w, int64_t w_ild,
                            const void* residual, int64_t ldr, void* dx, int64_t lddx, int batch, int h, int w_sp, int c_in,
                            int c_out, int ksize) {
  memset(&d, 0, sizeof(d));
  d.a.mode = B200PDM_OP_CONV_ACT, d.a.ptr = dy, d.a.ld = lddy;
  d.a.batch = batch, d.a.h_in = h, d.a.w_in = w_sp, d.a.channels = c_out;
  d.a.h_out = h, d.a.w_out = w_sp, d.a.stride = 1, d.a.taps = ksize * ksize, d.a.flip = 1;
  d.b.mode = B200PDM_OP_CONV_WT, d.b.ptr = w, d.b.ld = w_ild, d.b.channels = c_in, d.b.out_channels = c_out;
  d.b.taps = ksize * ksize;
  d.M = (int64_t)batch * h * w_sp, d.N = c_in, d.K = (int64_t)ksize * ksize * c_out, d.Z1 = 1, d.Z2 = 1;
  d.out = dx, d.out_fp32 = 0, d.ldo = lddx;
  d.residual = residual, d.ldr = ldr, d.alpha = 1.f;
}

extern "C" {

size_t b200pdm_linear_fwd_workspace(int64_t M, int64_t N, int64_t K, int out_fp32) {
  b200pdm_gemm_desc d;
  desc_linear_fwd(d, nullptr, K, nullptr, K, nullptr, nullptr, 0, nullptr, N, out_fp32, M, N, K);
  return gemm_workspace_bytes(&d);
}
int b200pdm_linear_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const void* residual,
                       int64_t ldr, void* out, int64_t ldo, int out_fp32, int64_t M, int64_t N, int64_t K, void* workspace,
                       size_t ws_bytes, b200pdm_stream_t stream) {
  b200pdm_gemm_desc d;
  desc_linear_fwd(d, x, ldx, w, ldw, bias, residual, ldr, out, ldo, out_fp32, M, N, K);
  return launch_gemm(&d, workspace, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

size_t b200pdm_linear_dgrad_workspace(int64_t M, int64_t N, int64_t K) {
  b200pdm_gemm_desc d;
  desc_linear_dgrad(d, nullptr, N, nullptr, K, nullptr, 0, nullptr, K, M, N, K);
  return gemm_workspace_bytes(&d);
}
int b200pdm_linear_dgrad(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* residual, int64_t ldr,
                         void* dx, int64_t lddx, int64_t M, int64_t N, int64_t K, void* workspace, size_t ws_bytes,
                         b200pdm_stream_t stream) {
  b200pdm_gemm_desc d;
  desc_linear_dgrad(d, dy, lddy, w, ldw, residual, ldr, dx, lddx, M, N, K);
  return launch_gemm(&d, workspace, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int b200pdm_linear_geglu_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* out, int64_t ldo,
                             void* pre, int64_t ldp, int64_t M, int64_t F, int64_t K, b200pdm_stream_t stream) {
  b200pdm_gemm_desc d;
  desc_linear_fwd(d, x, ldx, w, ldw, bias, nullptr, 0, out, ldo, 0, M, F, K);
  d.geglu = 1, d.aux = pre, d.ld_aux = ldp;
  return launch_gemm(&d, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

int b200pdm_linear_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw, int64_t M,
                         int64_t N, int64_t K, b200pdm_stream_t stream) {
  // dw[N,K] += dy[M,N]^T . x[M,K]: reduction over M; A(m'=n, k'=m) = dy[m][n] MN-major, B(n'=k, k'=m) = x[m][k] MN-major.
  b200pdm_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a.mode = B200PDM_OP_MN2D, d.a.ptr = dy, d.a.ld = lddy;
  d.b.mode = B200PDM_OP_MN2D, d.b.ptr = x, d.b.ld = ldx;
  d.M = N, d.N = K, d.K = M, d.Z1 = 1, d.Z2 = 1;
  d.out = dw, d.out_fp32 = 1, d.ldo = lddw, d.alpha = 1.f, d.accumulate = 1;
  return launch_gemm(&d, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

size_t b200pdm_conv_fwd_workspace(int batch, int h_in, int w_in, int c_in, int c_out, int ksize, int stride) {
  if ((ksize != 3 && ksize != 1) || (stride != 1 && stride != 2)) return 0;
  b200pdm_gemm_desc d;
  desc_conv_fwd(d, nullptr, c_in, nullptr, c_in, nullptr, nullptr, 0, nullptr, 0, nullptr, c_out, batch, h_in, w_in, c_in, c_out,
                ksize, stride);
  return gemm_workspace_bytes(&d);
}
int b200pdm_conv_fwd(const void* x, int64_t ldx, const void* w, int64_t w_ild, const float* bias, const float* rowbias,
                     int64_t ld_rowbias, const void* residual, int64_t ldr, void* out, int64_t ldo, int batch, int h_in,
                     int w_in, int c_in, int c_out, int ksize, int stride, void* workspace, size_t ws_bytes,
                     b200pdm_stream_t stream) {
  if ((ksize != 3 && ksize != 1) || (stride != 1 && stride != 2)) {
    set_err("conv_fwd: unsupported ksize/stride");
    return B200PDM_ERR_UNSUPPORTED;
  }
  b200pdm_gemm_desc d;
  desc_conv_fwd(d, x, ldx, w, w_ild, bias, rowbias, ld_rowbias, residual, ldr, out, ldo, batch, h_in, w_in, c_in, c_out, ksize,
                stride);
  return launch_gemm(&d, workspace, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int b200pdm_conv_fwd_nopad(const void* x, int64_t ldx, const void* w, int64_t w_ild, const float* bias, void* out, int64_t ldo,
                           int batch, int h_in, int w_in, int c_in, int c_out, int stride, b200pdm_stream_t stream) {
  if (stride != 1 && stride != 2) return B200PDM_ERR_UNSUPPORTED;
  b200pdm_gemm_desc d;
  desc_conv_fwd(d, x, ldx, w, w_ild, bias, nullptr, 0, nullptr, 0, out, ldo, batch, h_in, w_in, c_in, c_out, 3, stride, 1);
  return launch_gemm(&d, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

size_t b200pdm_conv_dgrad_workspace(int batch, int h, int w_sp, int c_in, int c_out, int ksize) {
  if (ksize != 3 && ksize != 1) return 0;
  b200pdm_gemm_desc d;
  desc_conv_dgrad(d, nullptr, c_out, nullptr, c_in, nullptr, 0, nullptr, c_in, batch, h, w_sp, c_in, c_out, ksize);
  return gemm_workspace_bytes(&d);
}
int b200pdm_conv_dgrad(const void* dy, int64_t lddy, const void* w, int64_t w_ild, const void* residual, int64_t ldr,
                       void* dx, int64_t lddx, int batch, int h, int w_sp, int c_in, int c_out, int ksize, void* workspace,
                       size_t ws_bytes, b200pdm_stream_t stream) {
  if (ksize != 3 && ksize != 1) {
    set_err("conv_dgrad: unsupported ksize");
    return B200PDM_ERR_UNSUPPORTED;
  }
  b200pdm_gemm_desc d;
  desc_conv_dgrad(d, dy, lddy, w, w_ild, residual, ldr, dx, lddx, batch, h, w_sp, c_in, c_out, ksize);
  return launch_gemm(&d, workspace, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int b200pdm_conv_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t w_ild, int batch,
                       int h_in, int w_in, int c_in, int c_out, int ksize, int stride, b200pdm_stream_t stream) {
  if ((ksize != 3 && ksize != 1) || (stride != 1 && stride != 2)) {
    set_err("conv_wgrad: unsupported ksize/stride");
    return B200PDM_ERR_UNSUPPORTED;
  }
  const int h_out = h_in / stride, w_out = w_in / stride;
  const int taps = ksize * ksize;
  b200pdm_gemm_desc d;
  memset(&d, 0, sizeof(d));
  const int64_t pixels = (int64_t)batch * h_out * w_out;
  d.a.mode = B200PDM_OP_MN2D, d.a.ptr = dy, d.a.ld = lddy;  // A(m'=co, k'=pixel)
  d.b.mode = B200PDM_OP_CONV_ACT_MN, d.b.ptr = x, d.b.ld = ldx;
  d.b.batch = batch, d.b.h_in = h_in, d.b.w_in = w_in, d.b.channels = c_in;
  d.b.h_out = h_out, d.b.w_out = w_out, d.b.stride = stride, d.b.taps = taps;
  d.M = c_out, d.N = (int64_t)taps * c_in, d.K = pixels, d.Z1 = 1, d.Z2 = 1;
  d.out = dw, d.out_fp32 = 1, d.ldo = (int64_t)taps * w_ild, d.alpha = 1.f, d.accumulate = 1;
  return launch_gemm(&d, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
