// GroupNorm(+SiLU) and LayerNorm, forward and backward, over channels-last bf16 activations.
// HBM-bound: 16-byte vector loads where the row pitch allows, fp32 statistics, warp-shuffle reductions.
//
// GroupNorm layout note: in NHWC a (sample, group) slab is hw rows of `cpg` contiguous channels, cpg in
// {10,20,30,40,60,80} for this U-Net -- not a multiple of 8 in general, so the statistics kernel
// walks whole 16-byte chunks of each pixel row and attributes elements to groups by channel index.
#include "common.cuh"
#include "../../include/b200pdm.h"

#include <atomic>

namespace b200 {
extern std::atomic<uint64_t> g_launches;
void set_err(const char* fmt, const char* a);

// ------------------------------------------------------------------------------------------------
// GroupNorm.  All heavy passes stream whole pixel rows with 16-byte vectors (8 channels per thread) and use
// per-(sample, channel) fp32 coefficient tables so that the inner loops are free of integer divisions and scalar
// parameter loads:
//   tab[0] = scale = gamma * rstd         tab[1] = shift = beta - mean * scale        (z = x*scale + shift)
//   tab[2] = R     = rstd                 tab[3] = MR    = mean * rstd                (xhat = x*R - MR)
// Each table is [B, ldc] with ldc = round8(C); pad lanes hold zeros.
// ------------------------------------------------------------------------------------------------
enum { GN_FWD_STATS = 0, GN_BWD_STATS = 1 };

// Column statistics per (sample, channel).  grid (B, ceil(cvec/32), slices), block (32, 8):
//   FWD: acc0 += x, acc1 += x^2        BWD: dz = dy * act'(z); acc0 += dz, acc1 += dz * xhat
template <int MODE>
__global__ void gn_colstats_kernel(const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ dy, int64_t lddy,
                                   const float* __restrict__ tab, float* __restrict__ out0, float* __restrict__ out1,
                                   int B, int hw, int C, int ldc, int silu, int slices) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float red[256][17];   // [row lane * VL + vector lane][16 partial sums]
  const int VL = blockDim.x, RL = blockDim.y;   // (8|16|32) x (256 / VL): see apply_geom()
  const int b = blockIdx.x;
  const int cv = blockIdx.y * VL + threadIdx.x;
  const int c0 = cv * 8;
  const int s = blockIdx.z;
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  if (c0 < C) {
    const int nv = min(8, C - c0);
    float sc[8], sh[8], rr[8], mr[8];
    if (MODE == GN_BWD_STATS) {
      const int64_t tb = (int64_t)b * ldc + c0;
      const int64_t plane = (int64_t)B * ldc;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = tab[tb + j], sh[j] = tab[plane + tb + j], rr[j] = tab[2 * plane + tb + j], mr[j] = tab[3 * plane + tb + j];
      }
    }
    const int64_t row0 = (int64_t)b * hw;
    // U rows per trip, every load issued before the first use (one row per trip left a single 16-byte load in flight per
    // thread).  Whole vectors are always loadable (pitches are multiples of 8 elements); lanes >= nv are masked to zero.
    constexpr int U = (MODE == GN_FWD_STATS) ? 8 : 4;
    for (int p = p0 + threadIdx.y; p < p1; p += RL * U) {
      bf16x8 xv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + u * RL;
        if (pp < p1) {
          xv[u] = *reinterpret_cast<const bf16x8*>(x + (row0 + pp) * ldx + c0);
          if (MODE == GN_BWD_STATS) dv[u] = *reinterpret_cast<const bf16x8*>(dy + (row0 + pp) * lddy + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p + u * RL >= p1) break;
        float xf[8];
        unpack8(xv[u], xf);
        if (nv < 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) xf[j] = j < nv ? xf[j] : 0.f;
        }
        if (MODE == GN_FWD_STATS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a0[j] += xf[j], a1[j] = fmaf(xf[j], xf[j], a1[j]);
        } else {
          float df[8];
          unpack8(dv[u], df);
          if (nv < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) df[j] = j < nv ? df[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float dz = df[j];
            if (silu) dz *= silu_grad_f(fmaf(xf[j], sc[j], sh[j]));
            a0[j] += dz;
            a1[j] = fmaf(dz, fmaf(xf[j], rr[j], -mr[j]), a1[j]);
          }
        }
      }
    }
  }
  const int tid = threadIdx.y * VL + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[tid][j] = a0[j], red[tid][8 + j] = a1[j];
  __syncthreads();
  // 256 threads reduce VL vectors x 16 values over the RL row lanes
  for (int k = tid; k < VL * 16; k += 256) {
    const int vx = k >> 4, jj = k & 15;
    float t = 0.f;
    for (int y = 0; y < RL; ++y) t += red[y * VL + vx][jj];
    const int c = (blockIdx.y * VL + vx) * 8 + (jj & 7);
    // one plain store per (slice, sample, channel): the finalize kernel adds the slices in order, so the statistics -- and with
    // them the whole forward pass, whose other kernels are reproducible already -- no longer depend on the order in which
    // blocks retire.  (Round 1 accumulated with atomicAdd: the 1e-7 reordering noise flips bf16 roundings downstream, and every
    // later rounding amplifies the difference -- sqrt(eps * ulp) per layer -- until two runs of the same step differ by the
    // full bf16 noise level, 1e-2 in features and gradients: profiles/r2_determinism_*.txt.)
    if (c < C) (jj < 8 ? out0 : out1)[(int64_t)s * 2 * B * ldc + (int64_t)b * ldc + c] = t;
  }
}

// fwd finalize: one thread per (b, c): group sums from planes 4/5 (sum x, sum x^2) -> the four coefficient tables.
// Slice sums of the channels [c_lo, c_hi) of sample b into shared memory, each channel by one thread in slice order (coalesced:
// consecutive threads read consecutive channels of a slice plane).
__device__ __forceinline__ void gn_gather_slices(const float* __restrict__ part, int slices, int64_t plane, int64_t base, int c_lo,
                                                 int c_hi, float* sh0, float* sh1) {
  for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    float t0 = 0.f, t1 = 0.f;
    const float* pp = part + base + c;
    int sl = 0;
    // eight slices' loads in flight before the first add (same order of additions: bit-identical).  The rolled loop issued one
    // pair of loads per L2 round trip -- with 11 ... 37 slices that chain WAS the kernel (7.6 us for 64 blocks of trivial work).
    for (; sl + 8 <= slices; sl += 8, pp += 16 * plane) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = pp[(int64_t)(2 * u) * plane], b[u] = pp[(int64_t)(2 * u + 1) * plane];
#pragma unroll
      for (int u = 0; u < 8; ++u) t0 += a[u], t1 += b[u];
    }
    for (; sl < slices; ++sl, pp += 2 * plane) t0 += pp[0], t1 += pp[plane];
    sh0[c - c_lo] = t0, sh1[c - c_lo] = t1;
  }
  __syncthreads();
}
constexpr int kGnGroupsPerBlock = 8;
constexpr int kGnMaxCpg = 160;   // channels per group (widest: 5120-channel maps never occur; 2560 / 32 = 80)

// fwd finalize: grid (B, ceil(groups / 8)); per-slice partial sums (sum x, sum x^2) -> the four coefficient tables.
__global__ void gn_fwd_finalize_kernel(float* __restrict__ tab, const float* __restrict__ part, int slices,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, int B, int C, int cpg,
                                       int ldc, float inv_n, float eps) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float sh0[kGnGroupsPerBlock * kGnMaxCpg], sh1[kGnGroupsPerBlock * kGnMaxCpg];
  const int b = blockIdx.x;
  const int c_lo = blockIdx.y * kGnGroupsPerBlock * cpg, c_hi = min(C, c_lo + kGnGroupsPerBlock * cpg);
  const int64_t plane = (int64_t)B * ldc, base = (int64_t)b * ldc;
  gn_gather_slices(part, slices, plane, base, c_lo, c_hi, sh0, sh1);
  for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    const int g0 = ((c - c_lo) / cpg) * cpg;
    float s = 0.f, q = 0.f;
    for (int k = 0; k < cpg; ++k) s += sh0[g0 + k], q += sh1[g0 + k];
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    tab[base + c] = sc;
    tab[plane + base + c] = beta[c] - mean * sc;
    tab[2 * plane + base + c] = rstd;
    tab[3 * plane + base + c] = mean * rstd;
  }
}

// y = act(x * scale + shift)
// y = [silu](x * scale[b, c] + shift[b, c]).  grid (B, channel-vector chunks, row slices), block (VL, 256 / VL): a thread
// owns 8 fixed channels of one sample -- its coefficients stay in registers -- and walks the rows of its slice (no index
// divisions, one 16-byte load and store per row).
__global__ void gn_apply_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ tab,
                                bf16* __restrict__ y, int64_t ldy, int B, int hw, int C, int ldc, int silu, int slices) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int b = blockIdx.x;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= C) return;
  const int nv = min(8, C - c0);
  const int64_t plane = (int64_t)B * ldc;
  const float* tp = tab + (int64_t)b * ldc + c0;   // ldc % 8 == 0 -> 32-byte aligned vector loads
  const float4 s0 = *reinterpret_cast<const float4*>(tp), s1 = *reinterpret_cast<const float4*>(tp + 4);
  const float4 h0 = *reinterpret_cast<const float4*>(tp + plane), h1 = *reinterpret_cast<const float4*>(tp + plane + 4);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  const int s = blockIdx.z;
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  const int64_t row0 = (int64_t)b * hw;
  constexpr int U = 4;   // rows per trip: all loads first, then the arithmetic and the stores
  const int RL = blockDim.y;
  for (int p = p0 + threadIdx.y; p < p1; p += RL * U) {
    bf16x8 xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * RL < p1) xv[u] = *reinterpret_cast<const bf16x8*>(x + (row0 + p + u * RL) * ldx + c0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * RL;
      if (pp >= p1) break;
      float f[8];
      unpack8(xv[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = fmaf(f[j], sc[j], sh[j]);
        f[j] = silu ? silu_f(v) : v;
      }
      bf16* yp = y + (row0 + pp) * ldy + c0;
      if (nv == 8) {
        *reinterpret_cast<bf16x8*>(yp) = pack8(f);
      } else {   // tail vector of a width that is not a multiple of 8: never write past column C
        for (int j = 0; j < nv; ++j) yp[j] = __float2bfloat16(f[j]);
      }
    }
  }
}

// bwd finalize: one thread per (b, c).  sum0 = sum dz, sum1 = sum dz*xhat per (b, c)  ->
//   (per-slice partial sums in `part`, added in slice order)  dgamma[c] += sum1, dbeta[c] += sum0 ;
//   ws[0] <- P = -rstd^2 * s2/n ,  ws[1] <- Q = -rstd*s1/n + mean*rstd^2*s2/n
__global__ void gn_bwd_finalize_kernel(float* __restrict__ ws, const float* __restrict__ part, int slices,
                                       const float* __restrict__ tab, const float* __restrict__ gamma,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C, int cpg, int ldc,
                                       float inv_n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float sh0[kGnGroupsPerBlock * kGnMaxCpg], sh1[kGnGroupsPerBlock * kGnMaxCpg];
  const int b = blockIdx.x;
  const int c_lo = blockIdx.y * kGnGroupsPerBlock * cpg, c_hi = min(C, c_lo + kGnGroupsPerBlock * cpg);
  const int64_t plane = (int64_t)B * ldc, base = (int64_t)b * ldc;
  gn_gather_slices(part, slices, plane, base, c_lo, c_hi, sh0, sh1);
  for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    const int g0 = ((c - c_lo) / cpg) * cpg;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      const float ga = gamma[c_lo + g0 + k];
      s1 += ga * sh0[g0 + k], s2 += ga * sh1[g0 + k];
    }
    atomicAdd(dbeta + c, sh0[c - c_lo]);
    atomicAdd(dgamma + c, sh1[c - c_lo]);
    const float rstd = tab[2 * plane + base + c], mr = tab[3 * plane + base + c];
    ws[base + c] = -rstd * rstd * s2 * inv_n;
    ws[plane + base + c] = -rstd * s1 * inv_n + mr * rstd * s2 * inv_n;
  }
}

// dx = scale * dz + x * P + Q (+ residual)
// dx = scale * dz + x * P + Q (+ residual), dz = dy * silu'(x * scale + shift).  Same thread layout as gn_apply_kernel.
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                                    const float* __restrict__ tab, const float* __restrict__ ws,
                                    const bf16* __restrict__ res, int64_t ldr, bf16* __restrict__ dx, int64_t lddx,
                                    int B, int hw, int C, int ldc, int silu, int slices) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int b = blockIdx.x;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= C) return;
  const int nv = min(8, C - c0);
  const int64_t plane = (int64_t)B * ldc;
  const float* tp = tab + (int64_t)b * ldc + c0;
  const float* wp = ws + (int64_t)b * ldc + c0;
  const float4 s0 = *reinterpret_cast<const float4*>(tp), s1 = *reinterpret_cast<const float4*>(tp + 4);
  const float4 h0 = *reinterpret_cast<const float4*>(tp + plane), h1 = *reinterpret_cast<const float4*>(tp + plane + 4);
  const float4 q0 = *reinterpret_cast<const float4*>(wp), q1 = *reinterpret_cast<const float4*>(wp + 4);
  const float4 r0 = *reinterpret_cast<const float4*>(wp + plane), r1 = *reinterpret_cast<const float4*>(wp + plane + 4);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  const float P[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  const float Q[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const int s = blockIdx.z;
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  const int64_t row0 = (int64_t)b * hw;
  constexpr int U = 4;   // rows per trip: all loads (x, dy, residual) first
  const int RL = blockDim.y;
  for (int p = p0 + threadIdx.y; p < p1; p += RL * U) {
    bf16x8 xv[U], dv[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + p + u * RL;
      if (p + u * RL < p1) {
        xv[u] = *reinterpret_cast<const bf16x8*>(x + row * ldx + c0);
        dv[u] = *reinterpret_cast<const bf16x8*>(dy + row * lddy + c0);
        if (res) rv[u] = *reinterpret_cast<const bf16x8*>(res + row * ldr + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * RL >= p1) break;
      const int64_t row = row0 + p + u * RL;
      float xf[8], df[8], o[8];
      unpack8(xv[u], xf);
      unpack8(dv[u], df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = df[j];
        if (silu) dz *= silu_grad_f(fmaf(xf[j], sc[j], sh[j]));
        o[j] = fmaf(sc[j], dz, fmaf(xf[j], P[j], Q[j]));
      }
      if (res) {  // fused gradient merge: dx += residual (e.g. the skip/shortcut branch's gradient)
        float rf[8];
        unpack8(rv[u], rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rf[j];
      }
      bf16* op = dx + row * lddx + c0;
      if (nv == 8) {
        *reinterpret_cast<bf16x8*>(op) = pack8(o);
      } else {
        for (int j = 0; j < nv; ++j) op[j] = __float2bfloat16(o[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C % 8 == 0 (C in {320, 640, 1280}); row cached in registers.
// ------------------------------------------------------------------------------------------------
template <int MAXV, int R>
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                              const float* __restrict__ beta, bf16* __restrict__ y, int64_t ldy,
                              float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int C, float eps) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  // One warp normalises R consecutive rows; all of their loads are issued before any arithmetic and the rows stay PACKED
  // (bf16) in registers between the passes: short rows (640 bytes at C = 320) need many warps x rows in flight per SM to
  // reach HBM bandwidth, so the register footprint decides the throughput.
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int nvec = C >> 3;
  bf16x8 raw[R][MAXV];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int64_t row = row0 + rr;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec && row < rows) raw[rr][k] = *reinterpret_cast<const bf16x8*>(x + row * ldx + vi * 8);
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int64_t row = row0 + rr;
    if (row >= rows) break;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      if (lane + k * 32 < nvec) {
        float v[8];
        unpack8(raw[rr][k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[j];
      }
    }
    sum = warp_sum(sum);
    const float m = sum / C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      if (lane + k * 32 < nvec) {
        float v[8];
        unpack8(raw[rr][k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - m;
          sq += d * d;
        }
      }
    }
    sq = warp_sum(sq);
    const float r = rsqrtf(sq / C + eps);
    if (lane == 0) {
      if (mean) mean[row] = m;
      if (rstd) rstd[row] = r;
    }
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vi = lane + k * 32;
      if (vi < nvec) {
        float v[8], o[8];
        unpack8(raw[rr][k], v);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[j] - m) * r * gg[j] + bb[j];
        *reinterpret_cast<bf16x8*>(y + row * ldy + vi * 8) = pack8(o);
      }
    }
  }
}

// LN backward: dx per row (one warp per row); dgamma/dbeta partials accumulated per block in smem then atomics.
template <int MAXV, int R>
__global__ void __launch_bounds__(256, (MAXV <= 2 ? 2 : 1))
ln_bwd_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
              const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
              const bf16* __restrict__ res, int64_t ldr, bf16* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
              float* __restrict__ dbeta, int64_t rows, int C, int rows_per_block, int vec_red) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  extern __shared__ float sh[];  // dgamma[C], dbeta[C]
  float* sdg = sh;
  float* sdb = sh + C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nvec = C >> 3;
  float adg[MAXV][8], adb[MAXV][8];
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) adg[k][j] = 0.f, adb[k][j] = 0.f;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(rows, row_begin + rows_per_block);
  // R rows per warp iteration, all loads first: more bytes in flight per SM (short rows are latency-bound otherwise)
  for (int64_t rowb = row_begin + warp * R; rowb < row_end; rowb += nwarps * R) {
    bf16x8 xr[R][MAXV], dr[R][MAXV], rr_[R][MAXV];
    float mm[R], rs[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int64_t row = rowb + q;
      const bool ok = row < row_end;
      mm[q] = ok ? mean[row] : 0.f;
      rs[q] = ok ? rstd[row] : 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (ok && vi < nvec) {
          xr[q][k] = *reinterpret_cast<const bf16x8*>(x + row * ldx + vi * 8);
          dr[q][k] = *reinterpret_cast<const bf16x8*>(dy + row * lddy + vi * 8);
          if (res) rr_[q][k] = *reinterpret_cast<const bf16x8*>(res + row * ldr + vi * 8);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int64_t row = rowb + q;
      if (row >= row_end) break;
      const float m = mm[q], r = rs[q];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nvec) {
          float xv[8], dv[8];
          unpack8(xr[q][k], xv);
          unpack8(dr[q][k], dv);
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - m) * r, gd = gg[j] * dv[j];
            s1 += gd;
            s2 += gd * xh;
            adg[k][j] += dv[j] * xh;
            adb[k][j] += dv[j];
          }
        }
      }
      s1 = warp_sum(s1) / C;
      s2 = warp_sum(s2) / C;
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int vi = lane + k * 32;
        if (vi < nvec) {
          float xv[8], dv[8], o[8];
          unpack8(xr[q][k], xv);
          unpack8(dr[q][k], dv);
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = r * (gg[j] * dv[j] - s1 - (xv[j] - m) * r * s2);
          if (res) {
            float rf[8];
            unpack8(rr_[q][k], rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += rf[j];
          }
          *reinterpret_cast<bf16x8*>(dx + row * lddx + vi * 8) = pack8(o);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sdg[vi * 8 + j], adg[k][j]);
        atomicAdd(&sdb[vi * 8 + j], adb[k][j]);
      }
    }
  }
  __syncthreads();
  // block partials -> global.  One 128-bit reduction per four channels (C % 8 == 0): with one resident wave of blocks this is
  // C / 2 vector atomics per block instead of the 2 C scalar ones per block over two waves that made the C = 1280 backward
  // atomics-bound (round 1: 758 k scalar atomics on 2560 addresses for a 31 MB problem, 0.18 of HBM bandwidth).
  if (vec_red) {
    for (int i = threadIdx.x * 4; i < C; i += blockDim.x * 4) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + i), "f"(sdg[i]), "f"(sdg[i + 1]), "f"(sdg[i + 2]),
                   "f"(sdg[i + 3])
                   : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + i), "f"(sdb[i]), "f"(sdb[i + 1]), "f"(sdb[i + 2]),
                   "f"(sdb[i + 3])
                   : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&dgamma[i], sdg[i]);
      atomicAdd(&dbeta[i], sdb[i]);
    }
  }
}

}  // namespace b200

using namespace b200;

template <typename Kern>
static int occupancy_of(Kern kern, int* cache) {
  if (*cache <= 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, 0) != cudaSuccess || n <= 0) n = 2;
    *cache = n;
  }
  return *cache;
}

extern "C" {

// stats: fp32 [4][batch][round8(C)]: coefficient tables scale, shift, rstd, mean*rstd (saved for the backward pass);
// scratch: b200pdm_groupnorm_scratch_floats() fp32 of per-slice partial sums (contents irrelevant on entry).
// Launch geometry of the GroupNorm apply passes: VL lanes across 8-channel vectors (the narrowest of 8 / 16 / 32 that wastes
// the fewest lanes on the last chunk), 256 / VL row lanes, and enough row slices for ~8 blocks per SM.
struct ApplyGeom {
  dim3 grid, block;
  int slices;
};
// `occ` = resident blocks per SM of the kernel being launched.  The row slices are chosen so that the grid is (just under) a
// whole number of waves of 148 * occ blocks: with the old fixed "~8 blocks per SM" rule the 2-blocks-per-SM backward kernels
// ran 4.05 waves -- a fifth, almost empty wave cost 19 % -- and the forward ones 1.6.
static ApplyGeom apply_geom(int batch, int hw, int cvec, int occ) {
  int best_vl = 32, best_waste = 1 << 30;
  for (int vl = 8; vl <= 32; vl *= 2) {
    const int waste = (cvec + vl - 1) / vl * vl - cvec;
    if (waste < best_waste || (waste == best_waste && vl > best_vl)) best_waste = waste, best_vl = vl;
  }
  const int chunks = (cvec + best_vl - 1) / best_vl;
  const int rl = 256 / best_vl;
  const int bps = batch * chunks;                     // blocks per row slice
  const int cap = 148 * (occ > 0 ? occ : 1);          // blocks per wave
  int max_slices = hw / (4 * rl);                     // at least 4 rows per thread
  if (max_slices < 1) max_slices = 1;
  int slices = 1;
  double best = -1.0;
  for (int w = 1; w <= 4; ++w) {
    int sl = cap * w / bps;
    if (sl < 1) continue;
    if (sl > max_slices) sl = max_slices;
    const double util = (double)sl * bps / ((double)cap * ((sl * bps + cap - 1) / cap));
    if (util > best + 0.03) best = util, slices = sl;   // prefer fewer, longer blocks unless a deeper grid fills clearly better
  }
  ApplyGeom g;
  g.grid = dim3(batch, chunks, slices), g.block = dim3(best_vl, rl), g.slices = slices;
  return g;
}
// Upper bound of the row slices apply_geom() can choose (it depends on the kernel's measured occupancy, at most 8 blocks of 256
// threads per SM): sizes the per-slice partial-sum scratch without a device query.
static int max_stat_slices(int batch, int hw, int cvec) {
  int best_vl = 32, best_waste = 1 << 30;
  for (int vl = 8; vl <= 32; vl *= 2) {
    const int waste = (cvec + vl - 1) / vl * vl - cvec;
    if (waste < best_waste || (waste == best_waste && vl > best_vl)) best_waste = waste, best_vl = vl;
  }
  const int chunks = (cvec + best_vl - 1) / best_vl, rl = 256 / best_vl;
  int max_slices = hw / (4 * rl);
  if (max_slices < 1) max_slices = 1;
  const int by_grid = 148 * 8 * 4 / (batch * chunks) + 1;
  return max_slices < by_grid ? max_slices : by_grid;
}
size_t b200pdm_groupnorm_scratch_floats(int batch, int hw, int C) {
  const int ldc = (C + 7) / 8 * 8;
  return (size_t)2 * max_stat_slices(batch, hw, (C + 7) / 8) * batch * ldc;
}
size_t b200pdm_groupnorm_bwd_workspace_floats(int batch, int hw, int C) {
  const int ldc = (C + 7) / 8 * 8;
  return (size_t)2 * batch * ldc + b200pdm_groupnorm_scratch_floats(batch, hw, C);
}
int b200pdm_groupnorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* stats, float* scratch, int batch, int hw, int C, int groups, float eps, int silu,
                          b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (groups <= 0 || C % groups || ldx % 8 || ldy % 8) {
    set_err("groupnorm: C %% groups != 0 or pitches not multiples of 8", "");
    return B200PDM_ERR_ARG;
  }
  const int cpg = C / groups;
  const int ldc = (C + 7) / 8 * 8;
  const int64_t plane = (int64_t)batch * ldc;
  if (!stats || !scratch) return B200PDM_ERR_ARG;
  const int cvec = (C + 7) / 8;
  static int occ_stats = 0, occ_apply = 0;
  const ApplyGeom sg = apply_geom(batch, hw, cvec, occupancy_of(gn_colstats_kernel<GN_FWD_STATS>, &occ_stats));
  if (sg.slices > max_stat_slices(batch, hw, cvec)) return B200PDM_ERR_ARG;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  launch_pdl(gn_colstats_kernel<GN_FWD_STATS>, sg.grid, sg.block, 0, stream, xb, ldx, nullptr, 0, nullptr, scratch,
             scratch + plane, batch, hw, C, ldc, 0, sg.slices);
  B200_CHECK_LAUNCH();
  if (cpg > kGnMaxCpg) return B200PDM_ERR_UNSUPPORTED;
  const dim3 fgrid(batch, (groups + kGnGroupsPerBlock - 1) / kGnGroupsPerBlock);
  launch_pdl(gn_fwd_finalize_kernel, fgrid, 256, 0, stream, stats, scratch, sg.slices, gamma, beta, batch, C, cpg, ldc,
             1.f / ((float)hw * cpg), eps);
  B200_CHECK_LAUNCH();
  const ApplyGeom ag = apply_geom(batch, hw, cvec, occupancy_of(gn_apply_kernel, &occ_apply));
  launch_pdl(gn_apply_kernel, ag.grid, ag.block, 0, stream, xb, ldx, stats, reinterpret_cast<bf16*>(y), ldy, batch, hw, C, ldc,
             silu, ag.slices);
  B200_CHECK_LAUNCH();
  g_launches += 3;
  return B200PDM_OK;
}

// workspace: fp32, b200pdm_groupnorm_bwd_workspace_floats() elements ([2][batch][round8(C)] coefficients + per-slice partial sums).
int b200pdm_groupnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* beta, const float* stats, const void* residual, int64_t ldr, void* dx,
                          int64_t lddx, float* dgamma, float* dbeta, float* workspace, int batch, int hw, int C,
                          int groups, int silu, b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (groups <= 0 || C % groups || ldx % 8 || lddy % 8 || lddx % 8 || (residual && ldr % 8)) return B200PDM_ERR_ARG;
  (void)beta;
  const int cpg = C / groups;
  const int ldc = (C + 7) / 8 * 8;
  const int64_t plane = (int64_t)batch * ldc;
  const int cvec = (C + 7) / 8;
  static int occ_stats = 0, occ_apply = 0;
  const ApplyGeom sg = apply_geom(batch, hw, cvec, occupancy_of(gn_colstats_kernel<GN_BWD_STATS>, &occ_stats));
  if (!workspace || sg.slices > max_stat_slices(batch, hw, cvec)) return B200PDM_ERR_ARG;
  float* part = workspace + 2 * plane;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  const bf16* dyb = reinterpret_cast<const bf16*>(dy);
  launch_pdl(gn_colstats_kernel<GN_BWD_STATS>, sg.grid, sg.block, 0, stream, xb, ldx, dyb, lddy, stats, part,
             part + plane, batch, hw, C, ldc, silu, sg.slices);
  B200_CHECK_LAUNCH();
  if (cpg > kGnMaxCpg) return B200PDM_ERR_UNSUPPORTED;
  const dim3 fgrid(batch, (groups + kGnGroupsPerBlock - 1) / kGnGroupsPerBlock);
  launch_pdl(gn_bwd_finalize_kernel, fgrid, 256, 0, stream, workspace, part, sg.slices, stats, gamma, dgamma, dbeta,
             batch, C, cpg, ldc, 1.f / ((float)hw * cpg));
  B200_CHECK_LAUNCH();
  const ApplyGeom ag = apply_geom(batch, hw, cvec, occupancy_of(gn_bwd_apply_kernel, &occ_apply));
  launch_pdl(gn_bwd_apply_kernel, ag.grid, ag.block, 0, stream, dyb, lddy, xb, ldx, stats, workspace,
             reinterpret_cast<const bf16*>(residual), ldr, reinterpret_cast<bf16*>(dx), lddx, batch, hw, C, ldc, silu, ag.slices);
  B200_CHECK_LAUNCH();
  g_launches += 3;
  return B200PDM_OK;
}

int b200pdm_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* mean, float* rstd, int64_t rows, int C, float eps, b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (C % 8 || C > 8 * 32 * 8 || ldx % 8 || ldy % 8) {
    set_err("layernorm: C must be a multiple of 8 and <= 2048", "");
    return B200PDM_ERR_UNSUPPORTED;
  }
  const int wpb = 8;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  bf16* yb = reinterpret_cast<bf16*>(y);
  auto nblocks = [&](int r) { return (int)((rows + (int64_t)wpb * r - 1) / ((int64_t)wpb * r)); };
  if (C <= 8 * 32 * 2)
    launch_pdl(ln_fwd_kernel<2, 4>, nblocks(4), wpb * 32, 0, stream, xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  else if (C <= 8 * 32 * 3)
    launch_pdl(ln_fwd_kernel<3, 2>, nblocks(2), wpb * 32, 0, stream, xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  else if (C <= 8 * 32 * 5)
    launch_pdl(ln_fwd_kernel<5, 1>, nblocks(1), wpb * 32, 0, stream, xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  else
    launch_pdl(ln_fwd_kernel<8, 1>, nblocks(1), wpb * 32, 0, stream, xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

int b200pdm_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* mean, const float* rstd, const void* residual, int64_t ldr, void* dx,
                          int64_t lddx, float* dgamma, float* dbeta, int64_t rows, int C, b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (C % 8 || C > 8 * 32 * 5 || ldx % 8 || lddy % 8 || lddx % 8) {
    set_err("layernorm_bwd: C must be a multiple of 8 and <= 1280", "");
    return B200PDM_ERR_UNSUPPORTED;
  }
  int blocks = 148 * (C <= 512 ? 2 : 1);       // one resident wave (launch bounds: 2 blocks / SM up to C = 512, else 1)
  int rows_per_block = (int)((rows + blocks - 1) / blocks);
  if (rows_per_block < 8) rows_per_block = 8;
  const int vec_red = ((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0;
  blocks = (int)((rows + rows_per_block - 1) / rows_per_block);
  const size_t sh = sizeof(float) * 2 * C;
  const bf16* dyb = reinterpret_cast<const bf16*>(dy);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  bf16* dxb = reinterpret_cast<bf16*>(dx);
  const bf16* rb = reinterpret_cast<const bf16*>(residual);
  if (residual && ldr % 8) return B200PDM_ERR_UNSUPPORTED;
  if (C <= 8 * 32 * 2)
    launch_pdl(ln_bwd_kernel<2, 2>, blocks, 256, sh, stream, dyb, lddy, xb, ldx, gamma, mean, rstd, rb, ldr, dxb, lddx, dgamma,
                                                    dbeta, rows, C, rows_per_block, vec_red);
  else if (C <= 8 * 32 * 3)
    launch_pdl(ln_bwd_kernel<3, 1>, blocks, 256, sh, stream, dyb, lddy, xb, ldx, gamma, mean, rstd, rb, ldr, dxb, lddx, dgamma,
                                                    dbeta, rows, C, rows_per_block, vec_red);
  else
    launch_pdl(ln_bwd_kernel<5, 1>, blocks, 256, sh, stream, dyb, lddy, xb, ldx, gamma, mean, rstd, rb, ldr, dxb, lddx, dgamma,
                                                    dbeta, rows, C, rows_per_block, vec_red);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

}  // extern "C"
