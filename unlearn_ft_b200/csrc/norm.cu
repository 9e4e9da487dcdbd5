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
// GroupNorm statistics: grid (B, groups, slices); each block reduces hw/slices pixels of one (b, g) and
// atomically accumulates (sum, sumsq) into ws[2*(b*G+g)].  A second tiny kernel finalises mean/rstd.
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ ws, int hw, int cpg,
                                int groups, int slices) {
  const int b = blockIdx.x, g = blockIdx.y, s = blockIdx.z;
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  const bf16* base = x + ((int64_t)b * hw) * ldx + g * cpg;
  float sum = 0.f, sq = 0.f;
  const int n = (p1 - p0) * cpg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int p = i / cpg, c = i - p * cpg;
    float v = __bfloat162float(base[(int64_t)(p0 + p) * ldx + c]);
    sum += v;
    sq += v * v;
  }
  __shared__ float red[32];
  sum = block_sum(sum, red);
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) {
    atomicAdd(&ws[2 * (b * groups + g)], sum);
    atomicAdd(&ws[2 * (b * groups + g) + 1], sq);
  }
}

// Vectorised statistics when cpg % 2 == 0 and ldx % 2 == 0: one thread handles bf16 pairs.
__global__ void gn_stats_kernel_v2(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ ws, int hw, int cpg,
                                   int groups, int slices) {
  const int b = blockIdx.x, g = blockIdx.y, s = blockIdx.z;
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  const bf16* base = x + ((int64_t)b * hw) * ldx + g * cpg;
  const int hp = cpg >> 1;
  float sum = 0.f, sq = 0.f;
  const int n = (p1 - p0) * hp;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int p = i / hp, c = i - p * hp;
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + (int64_t)(p0 + p) * ldx + 2 * c));
    sum += v.x + v.y;
    sq += v.x * v.x + v.y * v.y;
  }
  __shared__ float red[32];
  sum = block_sum(sum, red);
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) {
    atomicAdd(&ws[2 * (b * groups + g)], sum);
    atomicAdd(&ws[2 * (b * groups + g) + 1], sq);
  }
}

// in place: (sum, sumsq) -> (mean, rstd)
__global__ void gn_finalize_kernel(float* __restrict__ ws, int n, float inv_count, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = ws[2 * i] * inv_count;
  float var = ws[2 * i + 1] * inv_count - m * m;
  var = fmaxf(var, 0.f);
  ws[2 * i] = m;
  ws[2 * i + 1] = rsqrtf(var + eps);
}

// Apply: y = act(gamma * (x - mean) * rstd + beta); one thread per 8 channels when C % 8 == 0, scalar tail otherwise.
__global__ void gn_apply_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ stats,
                                bf16* __restrict__ y, int64_t ldy, int hw, int C, int cpg,
                                int groups, int silu, int64_t total_vec, int cvec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / cvec;
    int cv = (int)(i - row * cvec);
    int b = (int)(row / hw);
    int c0 = cv * 8;
    const bf16* xp = x + row * ldx + c0;
    bf16* yp = y + row * ldy + c0;
    float f[8];
    int nv = min(8, C - c0);
    if (nv == 8) {
      unpack8(*reinterpret_cast<const bf16x8*>(xp), f);
    } else {
      for (int j = 0; j < 8; ++j) f[j] = j < nv ? __bfloat162float(xp[j]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = c0 + j;
      if (j < nv) {
        int g = c / cpg;
        float m = __ldg(stats + 2 * (b * groups + g)), r = __ldg(stats + 2 * (b * groups + g) + 1);
        float v = (f[j] - m) * r * __ldg(gamma + c) + __ldg(beta + c);
        f[j] = silu ? silu_f(v) : v;
      }
    }
    if (nv == 8) {
      *reinterpret_cast<bf16x8*>(yp) = pack8(f);
    } else {
      for (int j = 0; j < nv; ++j) yp[j] = __float2bfloat16(f[j]);
    }
  }
}

// Backward pass 1: per (b, g) sums  s1 = sum(gamma*dz), s2 = sum(gamma*dz*xhat)  (atomics into ws), and per-channel
// dgamma += sum(dz*xhat), dbeta += sum(dz) where dz = dy * act'(z).
// grid (B, groups, slices), block = 256 threads; thread t owns channel (t % cpg_pad) lanes for coalescing-ish access.
__global__ void gn_bwd_reduce_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ stats,
                                     float* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int hw, int cpg, int groups, int silu, int slices) {
  extern __shared__ float sh[];  // [2 * cpg] per-channel partials + 32 reduction scratch
  float* ch_dg = sh;
  float* ch_db = sh + cpg;
  float* red = sh + 2 * cpg;
  const int b = blockIdx.x, g = blockIdx.y, s = blockIdx.z;
  for (int i = threadIdx.x; i < 2 * cpg; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int p0 = (int)((int64_t)hw * s / slices), p1 = (int)((int64_t)hw * (s + 1) / slices);
  const float m = stats[2 * (b * groups + g)], r = stats[2 * (b * groups + g) + 1];
  const int64_t row0 = (int64_t)b * hw;
  // thread -> (pixel lane, channel): channel = threadIdx.x % cpg for threads < floor(blockDim/cpg)*cpg
  const int ppb = blockDim.x / cpg;  // pixels processed per block iteration
  float s1 = 0.f, s2 = 0.f, dg = 0.f, db = 0.f;
  const int c = threadIdx.x % cpg;
  const int pl = threadIdx.x / cpg;
  if (pl < ppb) {
    const int ch = g * cpg + c;
    const float ga = gamma[ch], be = beta[ch];
    for (int p = p0 + pl; p < p1; p += ppb) {
      float xv = __bfloat162float(x[(row0 + p) * ldx + ch]);
      float dv = __bfloat162float(dy[(row0 + p) * lddy + ch]);
      float xh = (xv - m) * r;
      float dz = dv;
      if (silu) dz *= silu_grad_f(xh * ga + be);
      dg += dz * xh;
      db += dz;
      s1 += ga * dz;
      s2 += ga * dz * xh;
    }
    atomicAdd(&ch_dg[c], dg);
    atomicAdd(&ch_db[c], db);
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&ws[2 * (b * groups + g)], s1);
    atomicAdd(&ws[2 * (b * groups + g) + 1], s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    atomicAdd(&dgamma[g * cpg + i], ch_dg[i]);
    atomicAdd(&dbeta[g * cpg + i], ch_db[i]);
  }
}

// Backward pass 2: dx = rstd * (gamma*dz - s1/n - xhat * s2/n)
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ stats,
                                    const float* __restrict__ ws, const bf16* __restrict__ res, int64_t ldr,
                                    bf16* __restrict__ dx, int64_t lddx, int hw, int C, int cpg, int groups, int silu,
                                    float inv_n, int64_t total_vec, int cvec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / cvec;
    int cv = (int)(i - row * cvec);
    int b = (int)(row / hw);
    int c0 = cv * 8;
    int nv = min(8, C - c0);
    float xf[8], df[8], o[8];
    const bf16* xp = x + row * ldx + c0;
    const bf16* dp = dy + row * lddy + c0;
    if (nv == 8) {
      unpack8(*reinterpret_cast<const bf16x8*>(xp), xf);
      unpack8(*reinterpret_cast<const bf16x8*>(dp), df);
    } else {
      for (int j = 0; j < 8; ++j) {
        xf[j] = j < nv ? __bfloat162float(xp[j]) : 0.f;
        df[j] = j < nv ? __bfloat162float(dp[j]) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = 0.f;
      if (j < nv) {
        int c = c0 + j;
        int g = c / cpg;
        int sg = b * groups + g;
        float m = __ldg(stats + 2 * sg), r = __ldg(stats + 2 * sg + 1);
        float ga = __ldg(gamma + c);
        float xh = (xf[j] - m) * r;
        float dz = df[j];
        if (silu) dz *= silu_grad_f(xh * ga + __ldg(beta + c));
        float s1 = __ldg(ws + 2 * sg) * inv_n, s2 = __ldg(ws + 2 * sg + 1) * inv_n;
        o[j] = r * (ga * dz - s1 - xh * s2);
      }
    }
    if (res) {  // fused gradient merge: dx += residual (e.g. the skip/shortcut branch's gradient)
      const bf16* rp = res + row * ldr + c0;
      if (nv == 8) {
        float rf[8];
        unpack8(*reinterpret_cast<const bf16x8*>(rp), rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rf[j];
      } else {
        for (int j = 0; j < nv; ++j) o[j] += __bfloat162float(rp[j]);
      }
    }
    bf16* op = dx + row * lddx + c0;
    if (nv == 8) {
      *reinterpret_cast<bf16x8*>(op) = pack8(o);
    } else {
      for (int j = 0; j < nv; ++j) op[j] = __float2bfloat16(o[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C % 8 == 0 (C in {320, 640, 1280}); row cached in registers.
// ------------------------------------------------------------------------------------------------
template <int MAXV>  // max 16-byte vectors per lane: C <= MAXV*8*32
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                              const float* __restrict__ beta, bf16* __restrict__ y, int64_t ldy,
                              float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 3;
  float v[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int vi = lane + k * 32;
    if (vi < nvec) {
      unpack8(*reinterpret_cast<const bf16x8*>(x + row * ldx + vi * 8), v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    }
  }
  sum = warp_sum(sum);
  const float m = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[k][j] - m;
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
    int vi = lane + k * 32;
    if (vi < nvec) {
      float o[8];
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8), g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + vi * 8), b1 = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - m) * r * gg[j] + bb[j];
      *reinterpret_cast<bf16x8*>(y + row * ldy + vi * 8) = pack8(o);
    }
  }
}

// LN backward: dx per row (one warp per row); dgamma/dbeta partials accumulated per block in smem then atomics.
template <int MAXV>
__global__ void ln_bwd_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                              const float* __restrict__ gamma, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const bf16* __restrict__ res, int64_t ldr,
                              bf16* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, int64_t rows, int C, int rows_per_block) {
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
  for (int64_t row = row_begin + warp; row < row_end; row += nwarps) {
    const float m = mean[row], r = rstd[row];
    float xh[MAXV][8], gd[MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      int vi = lane + k * 32;
      if (vi < nvec) {
        float xv[8], dv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + row * ldx + vi * 8), xv);
        unpack8(*reinterpret_cast<const bf16x8*>(dy + row * lddy + vi * 8), dv);
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8), g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[k][j] = (xv[j] - m) * r;
          gd[k][j] = gg[j] * dv[j];
          s1 += gd[k][j];
          s2 += gd[k][j] * xh[k][j];
          adg[k][j] += dv[j] * xh[k][j];
          adb[k][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      int vi = lane + k * 32;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = r * (gd[k][j] - s1 - xh[k][j] * s2);
        if (res) {
          float rf[8];
          unpack8(*reinterpret_cast<const bf16x8*>(res + row * ldr + vi * 8), rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += rf[j];
        }
        *reinterpret_cast<bf16x8*>(dx + row * lddx + vi * 8) = pack8(o);
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
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], sdg[i]);
    atomicAdd(&dbeta[i], sdb[i]);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

// stats: fp32 [2 * batch * groups], interleaved (mean, rstd) per (sample, group).
int b200pdm_groupnorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* stats, int batch, int hw, int C, int groups, float eps, int silu,
                          b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (groups <= 0 || C % groups) {
    set_err("groupnorm: C %% groups != 0", "");
    return B200PDM_ERR_ARG;
  }
  const int cpg = C / groups;
  const int n = batch * groups;
  if (cudaMemsetAsync(stats, 0, sizeof(float) * 2 * n, stream) != cudaSuccess) return B200PDM_ERR_CUDA;
  // enough blocks to fill the machine: B*groups*slices >= ~4*148
  int slices = (4 * 148 + n - 1) / n;
  if (slices < 1) slices = 1;
  if (slices > hw / 8) slices = hw / 8 > 0 ? hw / 8 : 1;
  dim3 grid(batch, groups, slices);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  if ((cpg % 2 == 0) && (ldx % 2 == 0))
    gn_stats_kernel_v2<<<grid, 256, 0, stream>>>(xb, ldx, stats, hw, cpg, groups, slices);
  else
    gn_stats_kernel<<<grid, 256, 0, stream>>>(xb, ldx, stats, hw, cpg, groups, slices);
  B200_CHECK_LAUNCH();
  gn_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(stats, n, 1.f / ((float)hw * cpg), eps);
  B200_CHECK_LAUNCH();
  const int cvec = (C + 7) / 8;
  const int64_t total_vec = (int64_t)batch * hw * cvec;
  int blocks = (int)((total_vec + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_apply_kernel<<<blocks, 256, 0, stream>>>(xb, ldx, gamma, beta, stats, reinterpret_cast<bf16*>(y), ldy, hw, C, cpg,
                                             groups, silu, total_vec, cvec);
  B200_CHECK_LAUNCH();
  g_launches += 4;
  return B200PDM_OK;
}

// workspace: fp32 [2 * batch * groups].
int b200pdm_groupnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* beta, const float* stats, const void* residual, int64_t ldr, void* dx,
                          int64_t lddx, float* dgamma, float* dbeta, float* workspace, int batch, int hw, int C,
                          int groups, int silu, b200pdm_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (groups <= 0 || C % groups) return B200PDM_ERR_ARG;
  const int cpg = C / groups;
  const int n = batch * groups;
  if (cpg > 256) {
    set_err("groupnorm_bwd: channels per group > 256 unsupported", "");
    return B200PDM_ERR_UNSUPPORTED;
  }
  if (cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * n, stream) != cudaSuccess) return B200PDM_ERR_CUDA;
  int slices = (4 * 148 + n - 1) / n;
  if (slices < 1) slices = 1;
  if (slices > hw / 8) slices = hw / 8 > 0 ? hw / 8 : 1;
  dim3 grid(batch, groups, slices);
  const size_t sh = sizeof(float) * (2 * cpg + 32);
  gn_bwd_reduce_kernel<<<grid, 256, sh, stream>>>(reinterpret_cast<const bf16*>(dy), lddy,
                                                  reinterpret_cast<const bf16*>(x), ldx, gamma, beta, stats, workspace,
                                                  dgamma, dbeta, hw, cpg, groups, silu, slices);
  B200_CHECK_LAUNCH();
  const int cvec = (C + 7) / 8;
  const int64_t total_vec = (int64_t)batch * hw * cvec;
  int blocks = (int)((total_vec + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_bwd_apply_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), lddy,
                                                 reinterpret_cast<const bf16*>(x), ldx, gamma, beta, stats, workspace,
                                                 reinterpret_cast<const bf16*>(residual), ldr,
                                                 reinterpret_cast<bf16*>(dx), lddx, hw, C, cpg, groups, silu,
                                                 1.f / ((float)hw * cpg), total_vec, cvec);
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
  const int blocks = (int)((rows + wpb - 1) / wpb);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  bf16* yb = reinterpret_cast<bf16*>(y);
  if (C <= 8 * 32 * 2)
    ln_fwd_kernel<2><<<blocks, wpb * 32, 0, stream>>>(xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  else if (C <= 8 * 32 * 5)
    ln_fwd_kernel<5><<<blocks, wpb * 32, 0, stream>>>(xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
  else
    ln_fwd_kernel<8><<<blocks, wpb * 32, 0, stream>>>(xb, ldx, gamma, beta, yb, ldy, mean, rstd, rows, C, eps);
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
  int blocks = 148 * 2;
  int rows_per_block = (int)((rows + blocks - 1) / blocks);
  if (rows_per_block < 8) rows_per_block = 8;
  blocks = (int)((rows + rows_per_block - 1) / rows_per_block);
  const size_t sh = sizeof(float) * 2 * C;
  const bf16* dyb = reinterpret_cast<const bf16*>(dy);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  bf16* dxb = reinterpret_cast<bf16*>(dx);
  const bf16* rb = reinterpret_cast<const bf16*>(residual);
  if (residual && ldr % 8) return B200PDM_ERR_UNSUPPORTED;
  if (C <= 8 * 32 * 2)
    ln_bwd_kernel<2><<<blocks, 256, sh, stream>>>(dyb, lddy, xb, ldx, gamma, mean, rstd, rb, ldr, dxb, lddx, dgamma,
                                                 dbeta, rows, C, rows_per_block);
  else
    ln_bwd_kernel<5><<<blocks, 256, sh, stream>>>(dyb, lddy, xb, ldx, gamma, mean, rstd, rb, ldr, dxb, lddx, dgamma,
                                                 dbeta, rows, C, rows_per_block);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

}  // extern "C"
