// Fused distillation loss (value + gradient in one pass) and the flat multi-tensor AdamW.
//
// Loss follows pdm/training/trainer.py:2451-2486 (reference):
//   loss_ddpm  = mean_b( w_b * mean_chw (pred - target)^2 ),  w_b = min(snr_b + 1, gamma) / (snr_b + 1)   (v-prediction)
//   loss_kd    = mean( (pred - teacher_pred)^2 )
//   loss_block = (1/n_maps) * sum_k mean( (f_s^k - f_t^k)^2 )
//   total      = w_diff * loss_ddpm + w_block * loss_block + w_kd * loss_kd
// AdamW follows torch.optim.AdamW (non-amsgrad) as configured at trainer.py:265-284.
#include "common.cuh"
#include "../../include/b200pdm.h"

#include <atomic>

namespace b200 {
extern std::atomic<uint64_t> g_launches;
void set_err(const char* fmt, const char* a);

// pred-level losses: B * n elements (n = 4*64*64 = 16384): tiny, one pass, fp32.
__global__ void pred_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                 const float* __restrict__ teacher, const float* __restrict__ snr_w,
                                 float* __restrict__ dpred, float* __restrict__ sums, int batch, int64_t n,
                                 float w_diff, float w_kd) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float wb = snr_w ? snr_w[b] : 1.f;
  const float inv_bn = 1.f / ((float)batch * (float)n);
  float sd = 0.f, sk = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = (int64_t)b * n + i;
    const float p = pred[idx];
    float g = 0.f;
    if (target) {
      const float d = p - target[idx];
      sd += d * d;
      g += w_diff * 2.f * wb * d * inv_bn;
    }
    if (teacher) {
      const float d = p - teacher[idx];
      sk += d * d;
      g += w_kd * 2.f * d * inv_bn;
    }
    if (dpred) dpred[idx] = g;
  }
  sd = block_sum(sd, red);
  sk = block_sum(sk, red);
  if (threadIdx.x == 0) {
    float tot = 0.f;
    if (target) {
      atomicAdd(&sums[0], sd * wb * inv_bn);
      tot += w_diff * sd * wb * inv_bn;
    }
    if (teacher) {
      atomicAdd(&sums[1], sk * inv_bn);
      tot += w_kd * sk * inv_bn;
    }
    atomicAdd(&sums[3], tot);  // running weighted total (trainer.py:2473-2486)
  }
}

// feature-KD: bf16 student/teacher maps (contiguous, numel % 8 == 0 on the vector path).
__global__ void feature_loss_kernel(const bf16* __restrict__ s, const bf16* __restrict__ t, bf16* __restrict__ ds,
                                    float* __restrict__ sums, int64_t numel, float inv_maps, float gscale,
                                    float w_block) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  __shared__ float red[32];
  const int64_t nvec = numel >> 3;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8], g[8];
    unpack8(reinterpret_cast<const bf16x8*>(s)[i], a);
    unpack8(reinterpret_cast<const bf16x8*>(t)[i], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = a[j] - b[j];
      acc += d * d;
      g[j] = gscale * d;
    }
    if (ds) reinterpret_cast<bf16x8*>(ds)[i] = pack8(g);
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 3) + threadIdx.x; i < numel; i += blockDim.x) {
      const float d = __bfloat162float(s[i]) - __bfloat162float(t[i]);
      acc += d * d;
      if (ds) ds[i] = __float2bfloat16(gscale * d);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float c = acc * inv_maps / (float)numel;
    atomicAdd(&sums[2], c);
    atomicAdd(&sums[3], w_block * c);
  }
}

// Flat AdamW: one thread handles 4 consecutive parameters (float4 I/O), bf16 shadow written as 8 bytes.
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             bf16* __restrict__ shadow, int64_t n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, float grad_scale, int zero_grad,
                             const float* __restrict__ dyn) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  if (dyn) {   // step-dependent scalars from device memory: the launch can sit in a replayed CUDA graph
    lr = dyn[0], bc1 = dyn[1], bc2_sqrt = dyn[2];
  }
  const int64_t nvec = n >> 2;
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * grad_scale;
      pp[j] *= (1.f - lr * wd);
      mm[j] = beta1 * mm[j] + (1.f - beta1) * gr;
      vq[j] = beta2 * vq[j] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(vq[j]) / bc2_sqrt + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = pk;
    }
  }
  if (blockIdx.x == 0) {  // tail (n % 4)
    for (int64_t i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float gr = g[i] * grad_scale;
      float pp = p[i] * (1.f - lr * wd);
      const float mm = beta1 * m[i] + (1.f - beta1) * gr;
      const float vq = beta2 * v[i] + (1.f - beta2) * gr * gr;
      pp -= step_size * (mm / (sqrtf(vq) / bc2_sqrt + eps));
      p[i] = pp, m[i] = mm, v[i] = vq;
      if (zero_grad) g[i] = 0.f;
      if (shadow) shadow[i] = __float2bfloat16(pp);
    }
  }
}

__global__ void refresh_shadow_kernel(const float* __restrict__ p, bf16* __restrict__ shadow, int64_t n) {
  pdl_trigger();   // PDL: successors may start their prologue while this grid runs
  const int64_t nvec = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<const float4*>(p)[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(shadow)[i] = pk;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 2) + threadIdx.x; i < n; i += blockDim.x) shadow[i] = __float2bfloat16(p[i]);
}

}  // namespace b200

using namespace b200;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int b200pdm_pred_loss(const float* pred, const float* target, const float* teacher, const float* snr_w, float* dpred,
                      float* sums, int batch, int64_t n_per_sample, float w_diff, float w_kd, b200pdm_stream_t stream) {
  if (!pred || !sums || batch <= 0) return B200PDM_ERR_ARG;
  int bx = (int)((n_per_sample + 255) / 256);
  if (bx > 64) bx = 64;
  dim3 grid(bx, batch);
  launch_pdl(pred_loss_kernel, grid, 256, 0, STREAM, pred, target, teacher, snr_w, dpred, sums, batch, n_per_sample, w_diff, w_kd);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

int b200pdm_feature_loss(const void* s, const void* t, void* ds, float* sums, int64_t numel, float inv_maps,
                         float w_block, b200pdm_stream_t stream) {
  if (!s || !t || !sums || numel <= 0) return B200PDM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(ds)) & 15)
    return B200PDM_ERR_ARG;
  int64_t blocks = ((numel >> 3) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const float gscale = w_block * inv_maps * 2.f / (float)numel;
  launch_pdl(feature_loss_kernel, (int)blocks, 256, 0, STREAM, reinterpret_cast<const bf16*>(s), reinterpret_cast<const bf16*>(t),
                                                      reinterpret_cast<bf16*>(ds), sums, numel, inv_maps, gscale, w_block);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

int b200pdm_adamw_step(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int64_t step, float grad_scale, int zero_grad,
                       b200pdm_stream_t stream) {
  if (!p || !g || !m || !v || n <= 0 || step <= 0) return B200PDM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return B200PDM_ERR_ARG;
  // bias corrections in double on the host (torch computes them as python floats)
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(adamw_kernel, (int)blocks, 256, 0, STREAM, p, g, m, v, reinterpret_cast<bf16*>(shadow_bf16), n, lr, beta1, beta2, eps,
                                               weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, zero_grad, nullptr);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

int b200pdm_adamw_step_dyn(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* dyn,
                           float beta1, float beta2, float eps, float weight_decay, float grad_scale, int zero_grad,
                           b200pdm_stream_t stream) {
  if (!p || !g || !m || !v || !dyn || n <= 0) return B200PDM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return B200PDM_ERR_ARG;
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(adamw_kernel, (int)blocks, 256, 0, STREAM, p, g, m, v, reinterpret_cast<bf16*>(shadow_bf16), n, 0.f, beta1, beta2, eps,
                                               weight_decay, 1.f, 1.f, grad_scale, zero_grad, dyn);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

int b200pdm_refresh_shadow(const float* p, void* shadow_bf16, int64_t n, b200pdm_stream_t stream) {
  if (!p || !shadow_bf16 || n <= 0) return B200PDM_ERR_ARG;
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(refresh_shadow_kernel, (int)blocks, 256, 0, STREAM, p, reinterpret_cast<bf16*>(shadow_bf16), n);
  B200_CHECK_LAUNCH();
  g_launches++;
  return B200PDM_OK;
}

}  // extern "C"
