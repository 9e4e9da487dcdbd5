// HBM-bound elementwise / small-reduction kernels: GEGLU, row softmax, column sums, add/copy with pitches,
// SiLU on the time embedding, nearest-2x upsample and adjoints, layout conversion at the model boundary,
// sinusoidal timestep embedding, forward-diffusion prep.  All use 16-byte bf16x8 accesses on the common path and
// grid-stride loops sized to a multiple of the SM count.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200pdm.h"

#include <atomic>

namespace b200 {
extern std::atomic<uint64_t> g_launches;
void set_err(const char* fmt, const char* a);

static inline int grid_for(int64_t work_items, int threads) {
  int64_t b = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ---------------------------------------------------------------- GEGLU
// proj [rows, 2F] : value half [0,F), gate half [F,2F).  out = value * gelu_erf(gate).
__global__ void geglu_fwd_kernel(const bf16* __restrict__ p, int64_t ldp, bf16* __restrict__ o, int64_t ldo,
                                 int64_t rows, int F, int fvec) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = rows * fvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / fvec;
    int c0 = (int)(i - r * fvec) * 8;
    float h[8], g[8], y[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p + r * ldp + c0), h);
    unpack8(*reinterpret_cast<const bf16x8*>(p + r * ldp + F + c0), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = h[j] * gelu_erf_f(g[j]);
    *reinterpret_cast<bf16x8*>(o + r * ldo + c0) = pack8(y);
  }
}
__global__ void geglu_bwd_kernel(const bf16* __restrict__ d, int64_t ldd, const bf16* __restrict__ p, int64_t ldp,
                                 bf16* __restrict__ dp, int64_t lddp, int64_t rows, int F, int fvec) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = rows * fvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / fvec;
    int c0 = (int)(i - r * fvec) * 8;
    float h[8], g[8], dy[8], dh[8], dg[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p + r * ldp + c0), h);
    unpack8(*reinterpret_cast<const bf16x8*>(p + r * ldp + F + c0), g);
    unpack8(*reinterpret_cast<const bf16x8*>(d + r * ldd + c0), dy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dh[j] = dy[j] * gelu_erf_f(g[j]);
      dg[j] = dy[j] * h[j] * gelu_erf_grad_f(g[j]);
    }
    *reinterpret_cast<bf16x8*>(dp + r * lddp + c0) = pack8(dh);
    *reinterpret_cast<bf16x8*>(dp + r * lddp + F + c0) = pack8(dg);
  }
}

// ---------------------------------------------------------------- step-front producers (frozen encoders)
// erf-GELU over a [rows, C] bf16 matrix (CLIP text MLP: transformers CLIPMLP, hidden_act "gelu").
__global__ void gelu_kernel(const bf16* __restrict__ x, int64_t ldx, bf16* __restrict__ y, int64_t ldy, int64_t rows, int cvec) {
  pdl_trigger();
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cvec;
    const int c0 = (int)(i - r * cvec) * 8;
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + r * ldx + c0), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf_f(v[j]);
    *reinterpret_cast<bf16x8*>(y + r * ldy + c0) = pack8(v);
  }
}
// CLIPTextEmbeddings: out[b*L + l, :] = token_embedding[ids[b, l], :] + position_embedding[l, :]   (fp32 tables, bf16 out)
__global__ void clip_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                  bf16* __restrict__ out, int64_t ldo, int64_t rows, int L, int C, int vocab) {
  pdl_trigger();
  const int cvec = C >> 2;
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cvec;
    const int c0 = (int)(i - r * cvec) * 4;
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4 a = *reinterpret_cast<const float4*>(tok + id * C + c0);
    const float4 b = *reinterpret_cast<const float4*>(pos + (r % L) * (int64_t)C + c0);
    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x + b.x, a.y + b.y), hi = __floats2bfloat162_rn(a.z + b.z, a.w + b.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo), pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + r * ldo + c0) = pk;
  }
}
// AutoencoderKL.encode(...).latent_dist.sample() * scaling_factor (diffusers DiagonalGaussianDistribution): moments = NHWC
// bf16 [B*hw, 2*Cz] (mean | logvar); latents NCHW fp32 = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps) * scale.
__global__ void vae_sample_kernel(const bf16* __restrict__ mom, int64_t ldm, const float* __restrict__ eps, float* __restrict__ z,
                                  float* __restrict__ mean_out, int B, int Cz, int hw, float scale) {
  pdl_trigger();
  const int64_t total = (int64_t)B * Cz * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const int c = (int)((i / hw) % Cz);
    const int b = (int)(i / ((int64_t)hw * Cz));
    const bf16* m = mom + ((int64_t)b * hw + p) * ldm;
    const float mu = __bfloat162float(m[c]);
    const float lv = fminf(fmaxf(__bfloat162float(m[Cz + c]), -30.f), 20.f);
    const float sd = __expf(0.5f * lv);
    z[i] = (mu + sd * (eps ? eps[i] : 0.f)) * scale;
    if (mean_out) mean_out[i] = mu;
  }
}

// ---------------------------------------------------------------- runtime gates of the un-pruned network
// Width gate (pdm/models/gates.py:15-28 VirtualGate / :56-62 LinearWidthGate): y[r, c] = x[r, c] * gate[b(r) % Bg][g(c)], with
// b(r) = r / rows_per_sample and g(c) = (c % period) / group_size (period: q|k|v or value|gate column blocks share one gate).
__global__ void gate_scale_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gate, int ldg,
                                  bf16* __restrict__ y, int64_t ldy, int64_t rows, int C, int rows_per_sample, int period,
                                  int group_size, int Bg) {
  pdl_trigger();
  const int cvec = (C + 7) >> 3;
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cvec;
    const int c0 = (int)(i - r * cvec) * 8;
    const float* g = gate + (int64_t)((r / rows_per_sample) % Bg) * ldg;
    if (c0 + 8 <= C) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + r * ldx + c0), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= g[((c0 + j) % period) / group_size];
      *reinterpret_cast<bf16x8*>(y + r * ldy + c0) = pack8(v);
    } else {
      for (int c = c0; c < C; ++c) y[r * ldy + c] = __float2bfloat16(__bfloat162float(x[r * ldx + c]) * g[(c % period) / group_size]);
    }
  }
}
// d gate[bg][g] += sum over the rows of the samples b = bg (mod Bg) and the columns of group g of dy[r, c] * x[r, c].
// grid (column chunks of 256, samples); block 256: thread = one column, loop over the sample's rows, then one atomic per column.
__global__ void gate_grad_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                                 float* __restrict__ dgate, int ldg, int C, int rows_per_sample, int period, int group_size,
                                 int Bg, int row_slices) {
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  const int per = (rows_per_sample + row_slices - 1) / row_slices;
  const int r0 = blockIdx.z * per, r1 = min(rows_per_sample, r0 + per);
  const int64_t base = (int64_t)b * rows_per_sample;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r)
    acc += __bfloat162float(dy[(base + r) * lddy + c]) * __bfloat162float(x[(base + r) * ldx + c]);
  atomicAdd(dgate + (int64_t)(b % Bg) * ldg + (c % period) / group_size, acc);
}
// Depth gate (pdm/models/gates.py:43-49): y = (1 - m_b) * inp + m_b * out, m = gate[b % Bg].
__global__ void depth_blend_kernel(const bf16* __restrict__ inp, int64_t ldi, const bf16* __restrict__ out, int64_t ldo,
                                   const float* __restrict__ gate, bf16* __restrict__ y, int64_t ldy, int64_t rows, int C,
                                   int rows_per_sample, int Bg) {
  pdl_trigger();
  const int cvec = C >> 3;
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cvec;
    const int c0 = (int)(i - r * cvec) * 8;
    const float m = gate[(r / rows_per_sample) % Bg];
    float a[8], b[8];
    unpack8(*reinterpret_cast<const bf16x8*>(inp + r * ldi + c0), a);
    unpack8(*reinterpret_cast<const bf16x8*>(out + r * ldo + c0), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (1.f - m) * a[j] + m * b[j];
    *reinterpret_cast<bf16x8*>(y + r * ldy + c0) = pack8(a);
  }
}
// backward: d_inp = (1 - m) dy, d_out = m dy, d m[b % Bg] += sum dy * (out - inp)
__global__ void depth_blend_bwd_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ inp, int64_t ldi,
                                       const bf16* __restrict__ out, int64_t ldo, const float* __restrict__ gate,
                                       bf16* __restrict__ d_inp, int64_t ldgi, bf16* __restrict__ d_out, int64_t ldgo,
                                       float* __restrict__ dgate, int C, int rows_per_sample, int Bg, int row_slices) {
  pdl_trigger();
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float m = gate[b % Bg];
  const int cvec = C >> 3;
  const int per = (rows_per_sample + row_slices - 1) / row_slices;
  const int r0 = blockIdx.x * per, r1 = min(rows_per_sample, r0 + per);
  const int64_t base = (int64_t)b * rows_per_sample;
  float acc = 0.f;
  for (int64_t i = (int64_t)r0 * cvec + threadIdx.x; i < (int64_t)r1 * cvec; i += blockDim.x) {
    const int64_t r = base + i / cvec;
    const int c0 = (int)(i % cvec) * 8;
    float g[8], a[8], o[8], gi[8], go[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dy + r * lddy + c0), g);
    unpack8(*reinterpret_cast<const bf16x8*>(inp + r * ldi + c0), a);
    unpack8(*reinterpret_cast<const bf16x8*>(out + r * ldo + c0), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc += g[j] * (o[j] - a[j]);
      gi[j] = (1.f - m) * g[j];
      go[j] = m * g[j];
    }
    *reinterpret_cast<bf16x8*>(d_inp + r * ldgi + c0) = pack8(gi);
    *reinterpret_cast<bf16x8*>(d_out + r * ldgo + c0) = pack8(go);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && dgate) atomicAdd(dgate + (b % Bg), acc);
}

// ---------------------------------------------------------------- row softmax (frozen encoders: head dims other than 64)
// one block per row; fp32 scores in, bf16 probabilities out. p = softmax(scale * s).
__global__ void softmax_fwd_kernel(const float* __restrict__ s, int64_t lds, bf16* __restrict__ p, int64_t ldp, int cols,
                                   float scale) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* sr = s + row * lds;
  bf16* pr = p + row * ldp;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, sr[c]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf((sr[c] - mx) * scale);
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) pr[c] = __float2bfloat16(__expf((sr[c] - mx) * scale) * inv);
}

// ---------------------------------------------------------------- column sums (bias gradients)
// grid (col_blocks, row_slices); block (32, 8): thread (tx,ty) sums column tx over its rows; smem reduce over ty.
// Scalar variant for views whose base or pitch is not 16-byte aligned (column slices at pruned, odd offsets).
__global__ void colsum_scalar_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t rows, int cols,
                                     int rows_per_slice) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t r1 = min(rows, r0 + rows_per_slice);
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += __bfloat162float(x[r * ldx + c]);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

__global__ void colsum_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t rows, int cols,
                              int rows_per_slice) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  // block (32, 8): lane -> 8 consecutive columns (one 16-byte load per row), threadIdx.y -> row lane; ldx is a multiple of 8
  // and rows are padded to it, so the last (partial) vector of a row may be read but is masked when written.
  __shared__ float sh[8][32][9];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t r1 = min(rows, r0 + rows_per_slice);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols) {
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + r * ldx + c0), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;   // 256 threads <-> 32 vectors x 8 columns
  const int vx = tid >> 3, jj = tid & 7;
  const int c = (blockIdx.x * 32 + vx) * 8 + jj;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sh[y][vx][jj];
    atomicAdd(out + c, t);
  }
}

// out[g, c] += sum over rows r of group g (r / rows_per_group == g); grid (col_blocks, groups * slices_per_group)
__global__ void colsum_grouped_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t ldo,
                                      int cols, int rows_per_group, int slices, int rows_per_slice) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y / slices, s = blockIdx.y - g * slices;
  const int64_t base = (int64_t)g * rows_per_group;
  const int r0 = s * rows_per_slice, r1 = min(rows_per_group, r0 + rows_per_slice);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += __bfloat162float(x[(base + r) * ldx + c]);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    atomicAdd(out + (int64_t)g * ldo + c, t);
  }
}
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}

// ---------------------------------------------------------------- add / copy with pitches
__global__ void add_kernel(const bf16* __restrict__ a, int64_t lda, const bf16* __restrict__ b, int64_t ldb,
                           bf16* __restrict__ o, int64_t ldo, int64_t rows, int C, int cvec) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cvec;
    int c0 = (int)(i - r * cvec) * 8;
    int nv = min(8, C - c0);
    if (nv == 8) {
      float x[8], y[8];
      unpack8(*reinterpret_cast<const bf16x8*>(a + r * lda + c0), x);
      unpack8(*reinterpret_cast<const bf16x8*>(b + r * ldb + c0), y);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += y[j];
      *reinterpret_cast<bf16x8*>(o + r * ldo + c0) = pack8(x);
    } else {
      for (int j = 0; j < nv; ++j)
        o[r * ldo + c0 + j] =
            __float2bfloat16(__bfloat162float(a[r * lda + c0 + j]) + __bfloat162float(b[r * ldb + c0 + j]));
    }
  }
}
__global__ void copy2d_kernel(const bf16* __restrict__ s, int64_t lds, bf16* __restrict__ d, int64_t ldd, int64_t rows,
                              int C, int cvec) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cvec;
    int c0 = (int)(i - r * cvec) * 8;
    int nv = min(8, C - c0);
    if (nv == 8) {
      *reinterpret_cast<bf16x8*>(d + r * ldd + c0) = *reinterpret_cast<const bf16x8*>(s + r * lds + c0);
    } else {
      for (int j = 0; j < nv; ++j) d[r * ldd + c0 + j] = s[r * lds + c0 + j];
    }
  }
}

// ---------------------------------------------------------------- SiLU on the (small) time embedding
__global__ void silu_f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(silu_f(x[i]));
}
__global__ void silu_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, bf16* __restrict__ dx,
                                int64_t n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = __float2bfloat16(__bfloat162float(dy[i]) * silu_grad_f(x[i]));
}

// ---------------------------------------------------------------- nearest 2x upsample, adjoint, zero insertion
// mode 0: y[b,2h+dy,2w+dx,:] = x[b,h,w,:]   (index over y)
// mode 1: dx[b,h,w,:] = sum_{dy,dx} dy[b,2h+dy,2w+dx,:]   (index over x)
// mode 2: y[b,2h,2w,:] = x ; other positions 0   (index over y)
template <int MODE>
__global__ void resample2x_kernel(const bf16* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd,
                                  int batch, int h, int w, int C, int cvec) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  // (h, w) is the LOW resolution grid
  const int H2 = 2 * h, W2 = 2 * w;
  const int64_t total = (MODE == 1 ? (int64_t)batch * h * w : (int64_t)batch * H2 * W2) * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / cvec;
    int c0 = (int)(i - pix * cvec) * 8;
    int nv = min(8, C - c0);
    if (MODE == 1) {
      int ww = (int)(pix % w);
      int64_t t = pix / w;
      int hh = (int)(t % h);
      int b = (int)(t / h);
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const bf16* sp = src + (((int64_t)b * H2 + 2 * hh + dy) * W2 + 2 * ww + dx) * lds + c0;
          if (nv == 8) {
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(sp), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          } else {
            for (int j = 0; j < nv; ++j) acc[j] += __bfloat162float(sp[j]);
          }
        }
      bf16* dp = dst + pix * ldd + c0;
      if (nv == 8)
        *reinterpret_cast<bf16x8*>(dp) = pack8(acc);
      else
        for (int j = 0; j < nv; ++j) dp[j] = __float2bfloat16(acc[j]);
    } else {
      int ww = (int)(pix % W2);
      int64_t t = pix / W2;
      int hh = (int)(t % H2);
      int b = (int)(t / H2);
      bf16* dp = dst + pix * ldd + c0;
      const bool zero = (MODE == 2) && ((hh | ww) & 1);
      const bf16* sp = src + (((int64_t)b * h + (hh >> 1)) * w + (ww >> 1)) * lds + c0;
      if (nv == 8) {
        bf16x8 v;
        if (zero) {
          v = zero8();
        } else {
          v = *reinterpret_cast<const bf16x8*>(sp);
        }
        *reinterpret_cast<bf16x8*>(dp) = v;
      } else {
        for (int j = 0; j < nv; ++j) dp[j] = zero ? __float2bfloat16(0.f) : sp[j];
      }
    }
  }
}

// ---------------------------------------------------------------- model-boundary layout conversion
// NCHW fp32 -> NHWC bf16 (tiny C: 4 latent channels) and back.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t ldy, int batch, int C,
                                    int hw) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = (int64_t)batch * hw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int p = (int)(t % hw);
    int b = (int)(t / hw);
    y[((int64_t)b * hw + p) * ldy + c] = __float2bfloat16(x[((int64_t)b * C + c) * hw + p]);
  }
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ y, int batch, int C,
                                    int hw) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = (int64_t)batch * hw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int p = (int)(i % hw);
    int64_t t = i / hw;
    int c = (int)(t % C);
    int b = (int)(t / C);
    y[i] = __bfloat162float(x[((int64_t)b * hw + p) * ldx + c]);
  }
}

// ---------------------------------------------------------------- timestep embedding (flip_sin_to_cos, shift 0)
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, bf16* __restrict__ out, int64_t ldo, int batch,
                                          int dim) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int half = dim / 2;
  const int total = batch * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int b = i / half, k = i - b * half;
    // diffusers get_timestep_embedding: exponent = -ln(max_period) * arange(half) / (half - shift), shift = 0
    float freq = expf(-9.210340371976184f * (float)k / (float)half);
    float arg = (float)t[b] * freq;
    float s, c;
    sincosf(arg, &s, &c);
    out[(int64_t)b * ldo + k] = __float2bfloat16(c);          // flip_sin_to_cos: [cos | sin]
    out[(int64_t)b * ldo + half + k] = __float2bfloat16(s);
  }
}

// ---------------------------------------------------------------- sampling: CFG combine + DDIM step (eta = 0, v-prediction)
// One launch per denoising step: guided = u + g (c - u); x0 = sqrt(a_t) x - sqrt(1 - a_t) v; eps = sqrt(a_t) v + sqrt(1 - a_t) x;
// x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps.  The step counter, the next U-Net timestep tensor and the duplicated latent
// batch live in device memory so that the identical launch sequence of one step can be replayed from a CUDA graph 50 times.
__global__ void cfg_ddim_step_kernel(const float* __restrict__ model_out, float* __restrict__ latents, float* __restrict__ latent_in,
                                     const float* __restrict__ alphas_cumprod, const int64_t* __restrict__ timesteps,
                                     int* __restrict__ step_idx, int64_t* __restrict__ t_dev, int n, int64_t chw, int num_steps,
                                     int train_T, float guidance, unsigned int* __restrict__ done_ctr) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int idx = *step_idx;
  const int64_t t = timesteps[idx];
  const int64_t t_prev = t - train_T / num_steps;
  const float a_t = alphas_cumprod[t];
  const float a_prev = t_prev >= 0 ? alphas_cumprod[t_prev] : alphas_cumprod[0];   // set_alpha_to_one = False
  const float sa = sqrtf(a_t), sb = sqrtf(1.f - a_t), pa = sqrtf(a_prev), pb = sqrtf(1.f - a_prev);
  const int64_t total = (int64_t)n * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = model_out[i], c = model_out[total + i];
    const float v = u + guidance * (c - u);
    const float x = latents[i];
    const float x0 = sa * x - sb * v;
    const float eps = sa * v + sb * x;
    const float xp = pa * x0 + pb * eps;
    latents[i] = xp;
    latent_in[i] = xp;            // uncond half
    latent_in[total + i] = xp;    // text half
  }
  // last block to finish advances the step state (nobody reads step_idx / t_dev any more in this launch)
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    const int next = idx + 1;
    const int64_t tn = timesteps[next < num_steps ? next : num_steps - 1];
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) t_dev[i] = tn;
    if (threadIdx.x == 0) {
      *step_idx = next;
      *done_ctr = 0;
    }
  }
}

// Same fused step for diffusers PNDMScheduler(skip_prk_steps=True) -- the scheduler of the SD-2.1 hub config that
// scripts/metrics/generate_fid_images.py:113 loads: linear multistep (PLMS) over the last four guided model outputs.
// `timesteps` holds the N + 1 entries of PNDMScheduler.set_timesteps (the second one repeated); `ets` is a ring of the last four
// model outputs [4][n*chw], `cur_sample` the sample kept by the warm-up step; state = {counter, done counter, ets count, ring head}.
__global__ void cfg_pndm_step_kernel(const float* __restrict__ model_out, float* __restrict__ latents, float* __restrict__ latent_in,
                                     const float* __restrict__ alphas_cumprod, const int64_t* __restrict__ timesteps,
                                     int* __restrict__ state, int64_t* __restrict__ t_dev, float* __restrict__ ets,
                                     float* __restrict__ cur_sample, int n, int64_t chw, int num_steps, int train_T, float guidance) {
  pdl_trigger();
  const int k = state[0], n_ets0 = state[2], head0 = state[3];
  const int ratio = train_T / num_steps;
  int64_t t = timesteps[k < num_steps + 1 ? k : num_steps], t_prev = t - ratio;
  if (k == 1) t_prev = t, t = t + ratio;                       // second evaluation of the first interval (step_plms, counter == 1)
  const bool push = k != 1;
  const int n_ets = push ? min(n_ets0 + 1, 4) : n_ets0;
  const int head = push ? (head0 + 1) & 3 : head0;             // ring slot of the newest entry
  const float a_t = alphas_cumprod[t];
  const float a_prev = t_prev >= 0 ? alphas_cumprod[t_prev] : alphas_cumprod[0];   // set_alpha_to_one = False
  const float b_t = 1.f - a_t, b_prev = 1.f - a_prev;
  const float sample_coeff = sqrtf(a_prev / a_t);
  const float denom = a_t * sqrtf(b_prev) + sqrtf(a_t * b_t * a_prev);
  const float sa = sqrtf(a_t), sb = sqrtf(b_t);
  const int64_t total = (int64_t)n * chw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = model_out[i], c = model_out[total + i];
    const float g = u + guidance * (c - u);
    if (push) ets[(int64_t)head * total + i] = g;
    const float e1 = push ? g : ets[(int64_t)head * total + i];
    float m;
    float x = latents[i];
    if (k == 0) {
      m = g;
      cur_sample[i] = x;
    } else if (k == 1) {
      m = 0.5f * (g + e1);
      x = cur_sample[i];
    } else {
      const float e2 = ets[(int64_t)((head + 3) & 3) * total + i];
      if (n_ets == 2) {
        m = 0.5f * (3.f * e1 - e2);
      } else {
        const float e3 = ets[(int64_t)((head + 2) & 3) * total + i];
        if (n_ets == 3) {
          m = (23.f * e1 - 16.f * e2 + 5.f * e3) * (1.f / 12.f);
        } else {
          const float e4 = ets[(int64_t)((head + 1) & 3) * total + i];
          m = (1.f / 24.f) * (55.f * e1 - 59.f * e2 + 37.f * e3 - 9.f * e4);
        }
      }
    }
    const float eps = sa * m + sb * x;                         // v-prediction -> epsilon (_get_prev_sample)
    const float xp = sample_coeff * x - (a_prev - a_t) * eps / denom;
    latents[i] = xp;
    latent_in[i] = xp;
    latent_in[total + i] = xp;
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(state + 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    const int next = k + 1;
    const int64_t tn = timesteps[next < num_steps + 1 ? next : num_steps];
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) t_dev[i] = tn;
    if (threadIdx.x == 0) state[0] = next, state[1] = 0, state[2] = n_ets, state[3] = head;
  }
}

// ---------------------------------------------------------------- forward diffusion prep
__global__ void diffusion_prep_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                      const int64_t* __restrict__ t, const float* __restrict__ sa,
                                      const float* __restrict__ sb, float* __restrict__ noisy, float* __restrict__ vt,
                                      int batch, int64_t n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t total = (int64_t)batch * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / n);
    float a = sa[t[b]], s = sb[t[b]];
    float x = x0[i], e = noise[i];
    if (noisy) noisy[i] = a * x + s * e;
    if (vt) vt[i] = a * e - s * x;
  }
}

}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)

extern "C" {

int b200pdm_geglu_fwd(const void* proj, int64_t ldp, void* out, int64_t ldo, int64_t rows, int F,
                      b200pdm_stream_t stream) {
  if (F % 8 || ldp % 8 || ldo % 8) {
    set_err("geglu: F and pitches must be multiples of 8", "");
    return B200PDM_ERR_UNSUPPORTED;
  }
  const int fvec = F / 8;
  launch_pdl(geglu_fwd_kernel, grid_for(rows * fvec, 256), 256, 0, STREAM, CBF(proj), ldp, BF(out), ldo, rows, F, fvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_geglu_bwd(const void* dout, int64_t lddo, const void* proj, int64_t ldp, void* dproj, int64_t lddp,
                      int64_t rows, int F, b200pdm_stream_t stream) {
  if (F % 8 || ldp % 8 || lddo % 8 || lddp % 8) return B200PDM_ERR_UNSUPPORTED;
  const int fvec = F / 8;
  launch_pdl(geglu_bwd_kernel, grid_for(rows * fvec, 256), 256, 0, STREAM, CBF(dout), lddo, CBF(proj), ldp, BF(dproj), lddp, rows,
                                                                  F, fvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_gate_scale(const void* x, int64_t ldx, const float* gate, int ldg, void* y, int64_t ldy, int64_t rows, int C,
                       int rows_per_sample, int period, int group_size, int gate_batch, b200pdm_stream_t stream) {
  if (!x || !gate || !y || rows_per_sample <= 0 || period <= 0 || group_size <= 0 || gate_batch <= 0 || ldx % 8 || ldy % 8)
    return B200PDM_ERR_ARG;
  launch_pdl(gate_scale_kernel, grid_for(rows * ((C + 7) / 8), 256), 256, 0, STREAM, CBF(x), ldx, gate, ldg, BF(y), ldy, rows, C,
             rows_per_sample, period, group_size, gate_batch);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_gate_grad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dgate, int ldg, int64_t rows, int C,
                      int rows_per_sample, int period, int group_size, int gate_batch, b200pdm_stream_t stream) {
  if (!dy || !x || !dgate || rows_per_sample <= 0 || rows % rows_per_sample || period <= 0 || group_size <= 0 || gate_batch <= 0)
    return B200PDM_ERR_ARG;
  const int B = (int)(rows / rows_per_sample);
  int slices = rows_per_sample >= 1024 ? 8 : (rows_per_sample >= 128 ? 2 : 1);
  dim3 grid((C + 255) / 256, B, slices);
  launch_pdl(gate_grad_kernel, grid, 256, 0, STREAM, CBF(dy), lddy, CBF(x), ldx, dgate, ldg, C, rows_per_sample, period, group_size,
             gate_batch, slices);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_depth_blend(const void* inp, int64_t ldi, const void* out, int64_t ldo, const float* gate, void* y, int64_t ldy,
                        int64_t rows, int C, int rows_per_sample, int gate_batch, b200pdm_stream_t stream) {
  if (!inp || !out || !gate || !y || C % 8 || ldi % 8 || ldo % 8 || ldy % 8 || rows_per_sample <= 0 || gate_batch <= 0)
    return B200PDM_ERR_ARG;
  launch_pdl(depth_blend_kernel, grid_for(rows * (C / 8), 256), 256, 0, STREAM, CBF(inp), ldi, CBF(out), ldo, gate, BF(y), ldy, rows, C,
             rows_per_sample, gate_batch);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_depth_blend_bwd(const void* dy, int64_t lddy, const void* inp, int64_t ldi, const void* out, int64_t ldo,
                            const float* gate, void* d_inp, int64_t ldgi, void* d_out, int64_t ldgo, float* dgate, int64_t rows,
                            int C, int rows_per_sample, int gate_batch, b200pdm_stream_t stream) {
  if (!dy || !inp || !out || !gate || !d_inp || !d_out || C % 8 || rows_per_sample <= 0 || rows % rows_per_sample || gate_batch <= 0)
    return B200PDM_ERR_ARG;
  const int B = (int)(rows / rows_per_sample);
  const int slices = rows_per_sample >= 1024 ? 16 : (rows_per_sample >= 64 ? 4 : 1);
  dim3 grid(slices, B);
  launch_pdl(depth_blend_bwd_kernel, grid, 256, 0, STREAM, CBF(dy), lddy, CBF(inp), ldi, CBF(out), ldo, gate, BF(d_inp), ldgi,
             BF(d_out), ldgo, dgate, C, rows_per_sample, gate_batch, slices);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_gelu(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int C, b200pdm_stream_t stream) {
  if (C % 8 || ldx % 8 || ldy % 8) return B200PDM_ERR_UNSUPPORTED;
  launch_pdl(gelu_kernel, grid_for(rows * (C / 8), 256), 256, 0, STREAM, CBF(x), ldx, BF(y), ldy, rows, C / 8);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_clip_embed(const int64_t* ids, const float* token_embedding, const float* position_embedding, void* out,
                       int64_t ldo, int64_t rows, int seq_len, int C, int vocab, b200pdm_stream_t stream) {
  if (!ids || !token_embedding || !position_embedding || !out || C % 4 || ldo % 4 || seq_len <= 0) return B200PDM_ERR_ARG;
  launch_pdl(clip_embed_kernel, grid_for(rows * (C / 4), 256), 256, 0, STREAM, reinterpret_cast<const long long*>(ids),
             token_embedding, position_embedding, BF(out), ldo, rows, seq_len, C, vocab);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_vae_sample(const void* moments, int64_t ldm, const float* eps, float* latents, float* mean_out, int batch,
                       int latent_channels, int hw, float scaling_factor, b200pdm_stream_t stream) {
  if (!moments || !latents || batch <= 0 || latent_channels <= 0 || hw <= 0) return B200PDM_ERR_ARG;
  launch_pdl(vae_sample_kernel, grid_for((int64_t)batch * latent_channels * hw, 256), 256, 0, STREAM, CBF(moments), ldm, eps,
             latents, mean_out, batch, latent_channels, hw, scaling_factor);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_softmax_fwd(const float* s, int64_t lds, void* p, int64_t ldp, int64_t rows, int cols, float scale,
                        b200pdm_stream_t stream) {
  if (rows <= 0) return B200PDM_OK;
  int threads = cols >= 1024 ? 256 : 128;
  launch_pdl(softmax_fwd_kernel, (unsigned)rows, threads, 0, STREAM, s, lds, BF(p), ldp, cols, scale);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int cols, b200pdm_stream_t stream) {
  const bool vec = (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  int col_blocks = vec ? (cols + 255) / 256 : (cols + 31) / 32;
  int slices = (148 * 4 + col_blocks - 1) / col_blocks;
  int64_t rps = (rows + slices - 1) / slices;
  if (rps < 64) rps = 64;
  slices = (int)((rows + rps - 1) / rps);
  dim3 grid(col_blocks, slices), block(32, 8);
  if (vec)
    launch_pdl(colsum_kernel, grid, block, 0, STREAM, CBF(x), ldx, out, rows, cols, (int)rps);
  else
    launch_pdl(colsum_scalar_kernel, grid, block, 0, STREAM, CBF(x), ldx, out, rows, cols, (int)rps);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_colsum_grouped(const void* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int cols,
                           int rows_per_group, b200pdm_stream_t stream) {
  if (rows_per_group <= 0 || rows % rows_per_group) return B200PDM_ERR_ARG;
  const int groups = (int)(rows / rows_per_group);
  int col_blocks = (cols + 31) / 32;
  int slices = (148 * 4 + col_blocks * groups - 1) / (col_blocks * groups);
  int rps = (rows_per_group + slices - 1) / slices;
  if (rps < 32) rps = 32;
  slices = (rows_per_group + rps - 1) / rps;
  dim3 grid(col_blocks, groups * slices), block(32, 8);
  launch_pdl(colsum_grouped_kernel, grid, block, 0, STREAM, CBF(x), ldx, out, ldo, cols, rows_per_group, slices, rps);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_cast_f32_to_bf16(const float* x, void* y, int64_t n, b200pdm_stream_t stream) {
  launch_pdl(cast_f32_to_bf16_kernel, grid_for(n, 256), 256, 0, STREAM, x, BF(y), n);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_add(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows, int C,
                b200pdm_stream_t stream) {
  if (lda % 8 || ldb % 8 || ldo % 8) return B200PDM_ERR_UNSUPPORTED;
  const int cvec = (C + 7) / 8;
  launch_pdl(add_kernel, grid_for(rows * cvec, 256), 256, 0, STREAM, CBF(a), lda, CBF(b), ldb, BF(out), ldo, rows, C, cvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_copy2d(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int C, b200pdm_stream_t stream) {
  if (lds % 8 || ldd % 8) return B200PDM_ERR_UNSUPPORTED;
  const int cvec = (C + 7) / 8;
  launch_pdl(copy2d_kernel, grid_for(rows * cvec, 256), 256, 0, STREAM, CBF(src), lds, BF(dst), ldd, rows, C, cvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_silu_f32_to_bf16(const float* x, void* y, int64_t n, b200pdm_stream_t stream) {
  launch_pdl(silu_f32_to_bf16_kernel, grid_for(n, 256), 256, 0, STREAM, x, BF(y), n);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_silu_bwd(const void* dy, const float* x, void* dx, int64_t n, b200pdm_stream_t stream) {
  launch_pdl(silu_bwd_kernel, grid_for(n, 256), 256, 0, STREAM, CBF(dy), x, BF(dx), n);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int batch, int h, int w, int C,
                           b200pdm_stream_t stream) {
  if (ldx % 8 || ldy % 8) return B200PDM_ERR_UNSUPPORTED;
  const int cvec = (C + 7) / 8;
  launch_pdl(resample2x_kernel<0>, grid_for((int64_t)batch * 4 * h * w * cvec, 256), 256, 0, STREAM, CBF(x), ldx, BF(y), ldy,
                                                                                             batch, h, w, C, cvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int batch, int h, int w, int C,
                           b200pdm_stream_t stream) {
  if (lddy % 8 || lddx % 8) return B200PDM_ERR_UNSUPPORTED;
  const int cvec = (C + 7) / 8;
  launch_pdl(resample2x_kernel<1>, grid_for((int64_t)batch * h * w * cvec, 256), 256, 0, STREAM, CBF(dy), lddy, BF(dx), lddx,
                                                                                         batch, h, w, C, cvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_zero_insert2x(const void* x, int64_t ldx, void* y, int64_t ldy, int batch, int h, int w, int C,
                          b200pdm_stream_t stream) {
  if (ldx % 8 || ldy % 8) return B200PDM_ERR_UNSUPPORTED;
  const int cvec = (C + 7) / 8;
  launch_pdl(resample2x_kernel<2>, grid_for((int64_t)batch * 4 * h * w * cvec, 256), 256, 0, STREAM, CBF(x), ldx, BF(y), ldy,
                                                                                             batch, h, w, C, cvec);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t ldy, int batch, int C, int hw,
                                  b200pdm_stream_t stream) {
  launch_pdl(nchw_to_nhwc_kernel, grid_for((int64_t)batch * C * hw, 256), 256, 0, STREAM, x, BF(y), ldy, batch, C, hw);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_nhwc_bf16_to_nchw_f32(const void* x, int64_t ldx, float* y, int batch, int C, int hw,
                                  b200pdm_stream_t stream) {
  launch_pdl(nhwc_to_nchw_kernel, grid_for((int64_t)batch * C * hw, 256), 256, 0, STREAM, CBF(x), ldx, y, batch, C, hw);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_timestep_embedding(const int64_t* t, void* out, int64_t ldo, int batch, int dim, b200pdm_stream_t stream) {
  if (dim % 2) return B200PDM_ERR_ARG;
  launch_pdl(timestep_embedding_kernel, grid_for((int64_t)batch * dim / 2, 128), 128, 0, STREAM, t, BF(out), ldo, batch, dim);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_cfg_ddim_step(const float* model_out, float* latents, float* latent_in, const float* alphas_cumprod,
                          const int64_t* timesteps, int* state, int64_t* t_dev, int n, int64_t chw, int num_steps,
                          int train_timesteps, float guidance_scale, b200pdm_stream_t stream) {
  if (!model_out || !latents || !latent_in || !alphas_cumprod || !timesteps || !state || !t_dev || n <= 0 || chw <= 0 ||
      num_steps <= 0 || train_timesteps < num_steps)
    return B200PDM_ERR_ARG;
  launch_pdl(cfg_ddim_step_kernel, grid_for((int64_t)n * chw, 256), 256, 0, STREAM, 
      model_out, latents, latent_in, alphas_cumprod, timesteps, state, t_dev, n, chw, num_steps, train_timesteps, guidance_scale,
      reinterpret_cast<unsigned int*>(state + 1));
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_cfg_pndm_step(const float* model_out, float* latents, float* latent_in, const float* alphas_cumprod,
                          const int64_t* timesteps, int* state, int64_t* t_dev, float* ets, float* cur_sample, int n, int64_t chw,
                          int num_steps, int train_timesteps, float guidance_scale, b200pdm_stream_t stream) {
  if (!model_out || !latents || !latent_in || !alphas_cumprod || !timesteps || !state || !t_dev || !ets || !cur_sample || n <= 0 ||
      chw <= 0 || num_steps <= 0 || train_timesteps < num_steps)
    return B200PDM_ERR_ARG;
  launch_pdl(cfg_pndm_step_kernel, grid_for((int64_t)n * chw, 256), 256, 0, STREAM, model_out, latents, latent_in, alphas_cumprod,
             timesteps, state, t_dev, ets, cur_sample, n, chw, num_steps, train_timesteps, guidance_scale);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}
int b200pdm_diffusion_prep(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp,
                           const float* sqrt_1macp, float* noisy, float* vtarget, int batch, int64_t n_per_sample,
                           b200pdm_stream_t stream) {
  launch_pdl(diffusion_prep_kernel, grid_for((int64_t)batch * n_per_sample, 256), 256, 0, STREAM, 
      x0, noise, t, sqrt_acp, sqrt_1macp, noisy, vtarget, batch, n_per_sample);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

}  // extern "C"
