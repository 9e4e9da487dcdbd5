/*
 * b200pdm.h -- C ABI of the B200-native (sm_100a) hot path for rezashkv/unlearn-ft.
 *
 * The reference (pure Python, `pdm` package) has no FFI: its hot path is the chain
 *   pdm/training/trainer.py:2403-2488 (UnetFineTuner.step) -> pdm/models/unet/unet_2d_conditional.py:1417-1728
 *   -> pdm/models/unet/blocks.py (gated/pruned blocks) -> diffusers -> torch -> cuDNN/cuBLAS/SDPA/ATen.
 * Every entry point below names the reference call site (file:line under /root/reference) whose device work it
 * replaces.  The Python mirror of the reference's class surface (unlearn_ft_b200/pdm/...) binds these with ctypes.
 *
 * Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless noted; no torch types;
 *   - every call enqueues on `stream` and returns immediately: 0 on success, a negative B200PDM_ERR_* otherwise;
 *     nothing here allocates device memory, synchronises the device, or throws.  Calls that need scratch take a
 *     caller-owned `workspace` of at least b200pdm_<op>_workspace(...) bytes (32-byte aligned; contents irrelevant on entry;
 *     it must stay alive until the call's work has run -- allocate it from the stream-ordered / graph-pool allocator).
 *     A NULL workspace is legal for the GEMM-class calls: the planner then restricts itself to plans without scratch;
 *   - activations are bf16, channels-last ("NHWC"): a [B,C,H,W] tensor is a row-major [B*H*W, ld] matrix whose
 *     first C columns are valid, ld % 8 == 0 (16-byte rows; TMA requirement).  ld % 16 == 0 (rows on 32-byte sectors) lets the
 *     GEMM epilogue use 256-bit row accesses for every width -- the host mirror allocates that way;
 *   - weights used as tensor-core operands are bf16 "shadow" copies of the fp32 masters, kept in the layout
 *     [C_out][kh*kw][C_in_ld] for convolutions and [N][K] for linears (same element order as the masters);
 *   - gradients of parameters are fp32 and are ACCUMULATED into the caller's buffers.
 */
#ifndef B200PDM_H_
#define B200PDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200pdm_stream_t; /* cudaStream_t */

#define B200PDM_OK 0
#define B200PDM_ERR_ARG -1
#define B200PDM_ERR_CUDA -2
#define B200PDM_ERR_UNSUPPORTED -3
#define B200PDM_ERR_DRIVER -4

/* Library/ABI version and last CUDA error string (host pointers). */
int b200pdm_version(void);
const char* b200pdm_last_error(void);
/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
uint64_t b200pdm_launch_count(void);
/* Host-only: the tile plan b200pdm_gemm would pick (no device work; usable without a GPU).  n = N per group, tiles_m =
 * ceil(M / 128), kblocks = number of 64-wide K blocks; can_split: the output may be split along K (fp32 accumulate, or a
 * scratch + finalize pass when split_needs_finalize).  out[7] = {block_n, splits, pair, m_sub, stages, tiles, slots}. */
int b200pdm_gemm_plan(int64_t n, int n_groups, int b_mn, int tiles_m, int Z, int kblocks, int can_split,
                      int split_needs_finalize, int* out);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core core: one persistent, warp-specialised tcgen05 kernel (TMA -> smem ring -> tcgen05.mma -> TMEM ->
 * epilogue warps) driven by an operand "addressing program".  All GEMM-shaped reference call sites map onto it.
 * ------------------------------------------------------------------------------------------------------------ */
enum {
  B200PDM_OP_K2D = 0,        /* K-major matrix  elem(mn,k) = ptr[z2*bs2 + z1*bs1 + mn*ld + k]                  */
  B200PDM_OP_MN2D = 1,       /* MN-major matrix elem(mn,k) = ptr[z2*bs2 + z1*bs1 + k*ld + mn]                  */
  B200PDM_OP_CONV_ACT = 2,   /* A only: implicit-GEMM rows = output pixels, K = taps x channels, NHWC source    */
  B200PDM_OP_CONV_W = 3,     /* B only: conv weight [O][taps][I_ld], N = O, K = taps x I                       */
  B200PDM_OP_CONV_WT = 4,    /* B only: same memory, transposed use (dgrad): N = I, K = taps x O               */
  B200PDM_OP_CONV_ACT_MN = 5 /* B only: wgrad: N = (tap, channel) of the NHWC source, K = output pixels        */
};

typedef struct {
  int mode;
  const void* ptr; /* bf16 */
  int64_t ld;      /* row pitch in elements (2D modes), pixel pitch (ACT modes), I_ld (W modes)           */
  int64_t bs1, bs2;/* batch strides in elements (2D modes)                                                */
  /* convolution geometry (ACT/W modes) */
  int batch, h_in, w_in, channels; /* NHWC source: [batch, h_in, w_in, channels]; W modes: channels = I  */
  int h_out, w_out;                /* output grid the GEMM rows / K pixels enumerate                       */
  int stride;                      /* 1 or 2                                                               */
  int taps;                        /* 9 (3x3, pad 1) or 1                                                  */
  int flip;                        /* 1: use tap (2-kh, 2-kw) for the shift (dgrad)                        */
  int out_channels;                /* W modes: O                                                           */
  int no_pad;                      /* ACT modes: 1 = no leading zero padding (window starts AT the pixel; trailing
                                      out-of-bounds taps still read zeros): diffusers Downsample2D(padding=0), which pads
                                      (0, 1, 0, 1) and convolves with stride 2.  0 = the usual one-pixel "same" padding  */
} b200pdm_operand;

typedef struct {
  b200pdm_operand a, b;
  int64_t M, N, K;  /* logical GEMM extents; for conv modes K = taps * reduced channels (informational)   */
  int Z1, Z2;       /* batch grid, z = z2*Z1 + z1; use 1,1 for a plain GEMM                                */
  void* out;        /* bf16 or fp32 [M, N] row-major, pitch ldo, batch strides obs1/obs2                   */
  int out_fp32;
  int64_t ldo, obs1, obs2;
  const float* bias;        /* fp32 [N] or NULL                                                           */
  const float* rowbias;     /* fp32 [M / rows_per_group, ld_rowbias] or NULL (time-embedding broadcast)   */
  int64_t ld_rowbias;
  int rows_per_group;
  const void* residual;     /* bf16 [M, N] pitch ldr (+ rbs1/rbs2 batch strides) or NULL                  */
  int64_t ldr, rbs1, rbs2;
  float alpha;              /* out = alpha*acc + bias + rowbias + residual                                */
  int accumulate;           /* fp32 out only: out += ... (vector atomic adds; split-K allowed)             */
  int splits;               /* accumulate only: fixed split-K factor, 0/1 = planner's choice               */
  int block_n;              /* 0 = choose                                                                  */
  int geglu;                /* 1: fused GEGLU epilogue.  b = K-major [2N, K] (value rows, then gate rows), bias [2N];
                               out[M, N] (bf16) = (acc_v + bias_v) * gelu_erf(acc_g + bias_g); no rowbias / residual  */
  void* aux;                /* geglu: optional bf16 [M, 2N] pre-activations (value | gate), pitch ld_aux, or NULL  */
  int64_t ld_aux;
} b200pdm_gemm_desc;

/* Diagnostics: with B200PDM_GEMM_TRACE=1 in the environment every GEMM launch is timed (serialising); this writes the
 * per-shape table (tab separated) to `path` (host string) and clears it. */
int b200pdm_gemm_trace_dump(const char* path);
/* The same switch at run time (clears the table when turned on), and the table's totals: summed per-launch CUDA-event time
 * (ms), algorithmic FLOPs (2 M N K) and launch count -- bench.py's launch-weighted GEMM-class roofline. */
int b200pdm_gemm_trace_enable(int on);
int b200pdm_gemm_trace_totals(double* ms, double* flops, int64_t* launches);
/* Generic launch.  Non-accumulating outputs of small-M problems may be split along K: every split writes its partial tile
 * into its own fp32 slab of `workspace`, a finalize kernel adds the slabs in split order (deterministic) and applies the
 * epilogue.  b200pdm_gemm_workspace() = bytes that plan needs (0 when the planner does not split). */
size_t b200pdm_gemm_workspace(const b200pdm_gemm_desc* desc);
int b200pdm_gemm(const b200pdm_gemm_desc* desc, void* workspace, size_t ws_bytes, b200pdm_stream_t stream);

/* out[M,N] = x[M,K] . w[N,K]^T + bias + residual.            Replaces F.linear at pdm/models/unet/blocks.py:49
 * (GEGLU proj), :244,:251-252,:283 (attention projections), diffusers Transformer2DModel.proj_in/proj_out
 * (called at blocks.py:1172,1221), FeedForward.net[2], TimestepEmbedding (unet_2d_conditional.py:1521) and
 * time_emb_proj (blocks.py:337,538).                                                                       */
size_t b200pdm_linear_fwd_workspace(int64_t M, int64_t N, int64_t K, int out_fp32);
int b200pdm_linear_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                       const void* residual, int64_t ldr, void* out, int64_t ldo, int out_fp32, int64_t M,
                       int64_t N, int64_t K, void* workspace, size_t ws_bytes, b200pdm_stream_t stream);
/* GEGLU feed-forward input projection with the activation fused into the GEMM epilogue (GEGLUGated.forward,
 * pdm/models/unet/blocks.py:44-59: `hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)`):
 *   out[M, F] = (x . w[:F]^T + bias[:F]) * gelu_erf(x . w[F:]^T + bias[F:])          w: [2F, K], bias: [2F]
 * pre (optional, bf16 [M, 2F], pitch ldp): the two pre-activations, saved for b200pdm_geglu_bwd; without it (frozen teacher,
 * sampling) the projection's 2F-wide output never reaches HBM.                                               */
int b200pdm_linear_geglu_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* out, int64_t ldo,
                             void* pre, int64_t ldp, int64_t M, int64_t F, int64_t K, b200pdm_stream_t stream);
/* dx[M,K] = dy[M,N] . w[N,K] (+ residual[M,K]); autograd backward of the call sites above.                 */
size_t b200pdm_linear_dgrad_workspace(int64_t M, int64_t N, int64_t K);
int b200pdm_linear_dgrad(const void* dy, int64_t lddy, const void* w, int64_t ldw, const void* residual,
                         int64_t ldr, void* dx, int64_t lddx, int64_t M, int64_t N, int64_t K, void* workspace,
                         size_t ws_bytes, b200pdm_stream_t stream);
/* dw[N,K] += dy[M,N]^T . x[M,K]   (fp32 accumulate into the gradient arena, split-K through vector atomics).  */
int b200pdm_linear_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw,
                         int64_t M, int64_t N, int64_t K, b200pdm_stream_t stream);

/* 3x3 (pad 1, stride 1|2) or 1x1 convolution as implicit GEMM over NHWC bf16:
 *   out[b,ho,wo,:Cout] = sum_tap x[b, ho*s+kh-1, wo*s+kw-1, :Cin] . w[:, tap, :Cin]^T + bias + rowbias[b] + residual
 * Replaces conv1/conv2/conv_shortcut at blocks.py:332,374,377,533,575,578 (rowbias = the time-embedding add of
 * blocks.py:339-341), conv_in/conv_out at unet_2d_conditional.py:1616,1723 and the Down/Upsample2D convs.   */
size_t b200pdm_conv_fwd_workspace(int batch, int h_in, int w_in, int c_in, int c_out, int ksize, int stride);
int b200pdm_conv_fwd(const void* x, int64_t ldx, const void* w, int64_t w_ild, const float* bias,
                     const float* rowbias, int64_t ld_rowbias, const void* residual, int64_t ldr, void* out,
                     int64_t ldo, int batch, int h_in, int w_in, int c_in, int c_out, int ksize, int stride,
                     void* workspace, size_t ws_bytes, b200pdm_stream_t stream);
/* Same, 3x3, with the window anchored at the pixel instead of one before it:
 *   out[b,ho,wo,:] = sum_tap x[b, ho*s+kh, wo*s+kw, :] . w[:, tap, :]^T + bias      (zeros beyond the right / bottom edge)
 * = F.pad(x, (0, 1, 0, 1)) followed by a padding-0 convolution: the VAE encoder's Downsample2D (SURVEY 8f-2,
 * pdm/training/trainer.py:2405).  Forward only (the VAE is frozen).                                           */
int b200pdm_conv_fwd_nopad(const void* x, int64_t ldx, const void* w, int64_t w_ild, const float* bias, void* out, int64_t ldo,
                           int batch, int h_in, int w_in, int c_in, int c_out, int stride, b200pdm_stream_t stream);
/* dx[b,h,w,:Cin] = sum_tap dy[b, h+1-kh, w+1-kw, :Cout] . w[:, tap, :Cin] (+ residual)  (stride 1 only).    */
size_t b200pdm_conv_dgrad_workspace(int batch, int h, int w_sp, int c_in, int c_out, int ksize);
int b200pdm_conv_dgrad(const void* dy, int64_t lddy, const void* w, int64_t w_ild, const void* residual,
                       int64_t ldr, void* dx, int64_t lddx, int batch, int h, int w_sp, int c_in, int c_out,
                       int ksize, void* workspace, size_t ws_bytes, b200pdm_stream_t stream);
/* dw[Cout][tap][:Cin] += sum_pixels dy[p,:Cout]^T . x[shift_tap(p), :Cin]   (fp32 accumulate, split-K).      */
int b200pdm_conv_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t w_ild,
                       int batch, int h_in, int w_in, int c_in, int c_out, int ksize, int stride,
                       b200pdm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound kernels
 * ------------------------------------------------------------------------------------------------------------ */
/* GroupNorm (+ optional SiLU) over NHWC bf16.  groups*cpg == C.  Saves stats = per-(sample, channel) coefficient
 * tables scale/shift/rstd/mean*rstd, fp32 [4][B][round8(C)].  scratch: b200pdm_groupnorm_scratch_floats() fp32 (per-slice
 * partial sums, written with plain stores and added in slice order: the result does not depend on block scheduling).
 * Replaces norm1+nonlinearity / norm2+nonlinearity at blocks.py:318-319,348,371,519-520,549,572,
 * conv_norm_out+conv_act at unet_2d_conditional.py:1720-1722, and Transformer2DModel.norm (silu=0).        */
size_t b200pdm_groupnorm_scratch_floats(int batch, int hw, int C);
int b200pdm_groupnorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* stats, float* scratch, int batch, int hw, int C, int groups, float eps, int silu,
                          b200pdm_stream_t stream);
/* dx (+ optional residual: a second gradient stream merged in the same pass), and dgamma/dbeta += (fp32).
 * workspace: b200pdm_groupnorm_bwd_workspace_floats() fp32.                                                 */
size_t b200pdm_groupnorm_bwd_workspace_floats(int batch, int hw, int C);
int b200pdm_groupnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* beta, const float* stats, const void* residual, int64_t ldr, void* dx,
                          int64_t lddx, float* dgamma, float* dbeta, float* workspace, int batch, int hw, int C,
                          int groups, int silu, b200pdm_stream_t stream);
/* LayerNorm over the last dim of [rows, C] bf16 (diffusers BasicTransformerBlock.norm1/2/3).                */
int b200pdm_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                          float* mean, float* rstd, int64_t rows, int C, float eps, b200pdm_stream_t stream);
int b200pdm_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                          const float* mean, const float* rstd, const void* residual, int64_t ldr, void* dx,
                          int64_t lddx, float* dgamma, float* dbeta, int64_t rows, int C, b200pdm_stream_t stream);
/* GEGLU: out[r, f] = p[r, f] * gelu_erf(p[r, F + f]) for p = proj output [rows, 2F] (blocks.py:54-59).      */
int b200pdm_geglu_fwd(const void* proj, int64_t ldp, void* out, int64_t ldo, int64_t rows, int F,
                      b200pdm_stream_t stream);
int b200pdm_geglu_bwd(const void* dout, int64_t lddo, const void* proj, int64_t ldp, void* dproj, int64_t lddp,
                      int64_t rows, int F, b200pdm_stream_t stream);
/* Row softmax of fp32 scores -> bf16 probabilities: p = softmax(scale * s) per row (used between two batched
 * b200pdm_gemm calls where the fused head_dim-64 attention below does not apply).                           */
int b200pdm_softmax_fwd(const float* s, int64_t lds, void* p, int64_t ldp, int64_t rows, int cols, float scale,
                        b200pdm_stream_t stream);
/* Fused flash-style attention forward on tcgen05, head_dim 64, no mask (blocks.py:275-277 = SDPA): scores and
 * probabilities stay in TMEM / shared memory.  q: [B*Lq, >= H*64] pitch ldq (head h = columns [64h, 64h+64));
 * k, v: [B*Lk, ...] pitches ldk/ldv; out like q (pitch ldo); lse: fp32 [B, H, Lq] = log2-domain log-sum-exp of
 * scale*log2(e)*scores (or NULL).                                                                           */
int b200pdm_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          void* out, int64_t ldo, float* lse, int batch, int heads, int lq, int lk, float scale,
                          b200pdm_stream_t stream);
/* Backward of the fused attention (flash style: probabilities are recomputed per 128x128 tile from q, k and the
 * forward's lse; dV/dK accumulate in TMEM, dQ through fp32 vector atomics).  out/dout: forward output and its
 * gradient; dq/dk/dv: bf16 outputs with pitches; workspace: fp32 [roundup4(B*H*Lq) + B*Lq*H*64].               */
int b200pdm_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse, void* dq,
                          int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, float* workspace, int batch,
                          int heads, int lq, int lk, float scale, b200pdm_stream_t stream);
/* Fused attention forward with an optional causal mask (query i attends keys <= i): the CLIP text encoder of the
 * step-front producers (pdm/utils/data_utils.py:155-191 -> transformers CLIPTextModel).          */
int b200pdm_attention_fwd_ex(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* out, int64_t ldo, float* lse, int batch, int heads, int lq, int lk, float scale,
                             int causal, b200pdm_stream_t stream);
/* erf-GELU over a [rows, C] bf16 matrix (CLIP text MLP).                                                    */
int b200pdm_gelu(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int C, b200pdm_stream_t stream);
/* CLIPTextEmbeddings: out[b*L + l, :] = token_embedding[ids[b, l], :] + position_embedding[l, :] (fp32 tables, bf16 out). */
int b200pdm_clip_embed(const int64_t* ids, const float* token_embedding, const float* position_embedding, void* out,
                       int64_t ldo, int64_t rows, int seq_len, int C, int vocab, b200pdm_stream_t stream);
/* vae.encode(x).latent_dist.sample() * scaling_factor (pdm/training/trainer.py:2405-2406; diffusers
 * DiagonalGaussianDistribution): moments = NHWC bf16 [B*hw, 2*Cz] (mean | logvar, logvar clamped to [-30, 20]);
 * latents (NCHW fp32) = (mean + exp(logvar / 2) * eps) * scaling_factor; eps NCHW fp32 or NULL (mode); mean_out optional. */
int b200pdm_vae_sample(const void* moments, int64_t ldm, const float* eps, float* latents, float* mean_out, int batch,
                       int latent_channels, int hw, float scaling_factor, b200pdm_stream_t stream);
/* Runtime gates of the UN-PRUNED network (the pruning phase's mode).  Width gates (pdm/models/gates.py:15-28,
 * 56-62; applied at blocks.py:56-58,267-272,343-348): y[r, c] = x[r, c] * gate[(r / rows_per_sample) % gate_batch][(c % period)
 * / group_size] -- gate is fp32 [gate_batch, widths] with row pitch ldg (per-sample architectures, or one shared row), `period`
 * lets q|k|v or value|gate column blocks share one gate row.  gate_grad accumulates d gate += sum dy * x over the gated
 * elements (the gradient the hypernetwork is trained with, trainer.py:1159-1321); dx is gate_scale applied to dy.          */
int b200pdm_gate_scale(const void* x, int64_t ldx, const float* gate, int ldg, void* y, int64_t ldy, int64_t rows, int C,
                       int rows_per_sample, int period, int group_size, int gate_batch, b200pdm_stream_t stream);
int b200pdm_gate_grad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dgate, int ldg, int64_t rows, int C,
                      int rows_per_sample, int period, int group_size, int gate_batch, b200pdm_stream_t stream);
/* Depth gate (gates.py:43-49; blocks.py:582-587,1241-1244): y = (1 - m) * inp + m * out with m = gate[b % gate_batch], and
 * its backward: d_inp = (1 - m) dy, d_out = m dy, dgate[b % gate_batch] += sum dy * (out - inp)  (dgate may be NULL).     */
int b200pdm_depth_blend(const void* inp, int64_t ldi, const void* out, int64_t ldo, const float* gate, void* y, int64_t ldy,
                        int64_t rows, int C, int rows_per_sample, int gate_batch, b200pdm_stream_t stream);
int b200pdm_depth_blend_bwd(const void* dy, int64_t lddy, const void* inp, int64_t ldi, const void* out, int64_t ldo,
                            const float* gate, void* d_inp, int64_t ldgi, void* d_out, int64_t ldgo, float* dgate, int64_t rows,
                            int C, int rows_per_sample, int gate_batch, b200pdm_stream_t stream);
/* Column sums: out[n] += sum_m x[m, n]  (bias gradients).                                                   */
int b200pdm_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int cols, b200pdm_stream_t stream);
/* Per-sample column sums: out[r / rows_per_group, n] += x[r, n]  (gradient of the time-embedding broadcast add,
 * blocks.py:339-341; summing it over samples gives the conv bias gradient).                                 */
int b200pdm_colsum_grouped(const void* x, int64_t ldx, float* out, int64_t ldo, int64_t rows, int cols,
                           int rows_per_group, b200pdm_stream_t stream);
int b200pdm_cast_f32_to_bf16(const float* x, void* y, int64_t n, b200pdm_stream_t stream);
/* Elementwise helpers over [rows, C] bf16 matrices with pitches. */
int b200pdm_add(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows, int C,
                b200pdm_stream_t stream);
int b200pdm_copy2d(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int C,
                   b200pdm_stream_t stream);
int b200pdm_silu_f32_to_bf16(const float* x, void* y, int64_t n, b200pdm_stream_t stream);
/* dx (bf16) = dy (bf16) * silu'(x) with x the saved fp32 pre-activation (contiguous). */
int b200pdm_silu_bwd(const void* dy, const float* x, void* dx, int64_t n, b200pdm_stream_t stream);
/* Nearest 2x upsample NHWC (diffusers Upsample2D, F.interpolate(scale=2, nearest)) and its adjoint.         */
int b200pdm_upsample2x_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int batch, int h, int w, int C,
                           b200pdm_stream_t stream);
int b200pdm_upsample2x_bwd(const void* dy, int64_t lddy, void* dx, int64_t lddx, int batch, int h, int w, int C,
                           b200pdm_stream_t stream);
/* Zero-insertion 2x (adjoint of stride-2 subsampling): y[b,2h,2w,:] = x[b,h,w,:], zeros elsewhere.          */
int b200pdm_zero_insert2x(const void* x, int64_t ldx, void* y, int64_t ldy, int batch, int h, int w, int C,
                          b200pdm_stream_t stream);
/* NCHW fp32 <-> NHWC bf16 (model boundary: sample in, .sample out; unet_2d_conditional.py:1417,1725).       */
int b200pdm_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t ldy, int batch, int C, int hw,
                                  b200pdm_stream_t stream);
int b200pdm_nhwc_bf16_to_nchw_f32(const void* x, int64_t ldx, float* y, int batch, int C, int hw,
                                  b200pdm_stream_t stream);
/* Sinusoidal timestep embedding, flip_sin_to_cos=True, shift 0 (diffusers Timesteps; unet...py:1514-1519):
 * out[b, :half] = cos(t*f), out[b, half:] = sin(t*f), f_i = exp(-ln(10000) * i / half); bf16 out.           */
int b200pdm_timestep_embedding(const int64_t* t, void* out, int64_t ldo, int batch, int dim,
                               b200pdm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused distillation loss (trainer.py:2451-2486): min-SNR weighted MSE(pred, target) + output-KD MSE(pred, teacher)
 * + mean over feature maps of MSE(student_feat, teacher_feat), forward AND gradients in a single kernel launch.
 * ------------------------------------------------------------------------------------------------------------ */
/* ONE launch for the whole loss, value AND gradients (north star (c); replaces the 11 F.mse_loss calls + autograd of
 * trainer.py:2451-2486).  pred/target/teacher: fp32 [batch, n_per_sample] (the reference's explicit .float() casts,
 * :2452,2468,2485; target or teacher may be NULL to drop that term).  Per-sample min-SNR weights (:2457-2466): either
 * snr_w (fp32 [batch]) or, when snr_w is NULL, computed in the kernel from alphas_cumprod (fp32 table) and timesteps
 * (int64 [batch]): snr = acp/(1-acp) (+1 first for v_prediction), w = min(snr, snr_gamma)/snr; all three NULL = weight 1.
 * pairs: HOST array of n_pairs (<= 16) dense bf16 feature maps (student, teacher, gradient out or NULL, numel) -- the hook
 * outputs of trainer.py:557-572, which the reference does not upcast (:2478).
 *   sums[0] = mean_b(w_b * mean_chw (pred - target)^2)          diff_loss
 *   sums[1] = mean((pred - teacher)^2)                          distillation_loss
 *   sums[2] = (1/n_pairs) * sum_k mean((s_k - t_k)^2)           block_loss
 *   sums[3] = w_diff*sums[0] + w_block*sums[2] + w_kd*sums[1]   loss
 *   dpred   = d sums[3] / d pred (fp32, may be NULL);  pairs[k].ds = d sums[3] / d s_k (bf16)
 * Two-stage reduction inside the launch (per-block partial rows in `workspace`, summed in block order by the block that
 * arrives last): bit-identical from run to run, no floating-point atomics.  workspace: b200pdm_kd_loss_workspace() bytes,
 * 16-byte aligned, contents irrelevant. */
typedef struct {
  const void* s;  /* student feature map, bf16, dense */
  const void* t;  /* teacher feature map, bf16, same layout */
  void* ds;       /* gradient w.r.t. s (bf16) or NULL */
  int64_t numel;
} b200pdm_feature_pair;
size_t b200pdm_kd_loss_workspace(int n_pairs);
int b200pdm_kd_loss_fused(const float* pred, const float* target, const float* teacher, const float* snr_w,
                          const float* alphas_cumprod, const int64_t* timesteps, float snr_gamma, int v_prediction,
                          float* dpred, int batch, int64_t n_per_sample, float w_diff, float w_kd,
                          const b200pdm_feature_pair* pairs, int n_pairs, float w_block, float* sums, void* workspace,
                          size_t ws_bytes, b200pdm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-tensor AdamW over the flat parameter arena (torch.optim.AdamW semantics; trainer.py:265-284,2327,2814).
 * One launch updates p, m, v (fp32), writes the bf16 shadow copy and optionally zeroes the gradient.
 * The bilevel trainer (trainer.py:2795-2816) calls it with a second (m, v, step) state set.
 * ------------------------------------------------------------------------------------------------------------ */
int b200pdm_adamw_step(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int64_t step, float grad_scale, int zero_grad,
                       b200pdm_stream_t stream);
/* Same update with the step-dependent scalars read from device memory: dyn = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}
 * (fp32[3]).  The launch itself is then step-independent and can live in a replayed CUDA graph of the whole training
 * step (trainer.py:2316-2329); the host refreshes `dyn` before each replay.                                        */
int b200pdm_adamw_step_dyn(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* dyn,
                           float beta1, float beta2, float eps, float weight_decay, float grad_scale, int zero_grad,
                           b200pdm_stream_t stream);
/* shadow = bf16(p) (after load_state_dict / a foreign optimizer touched the masters). */
int b200pdm_refresh_shadow(const float* p, void* shadow_bf16, int64_t n, b200pdm_stream_t stream);
/* Same pass, also zeroing `zero[0..n)` (fp32, may be NULL).  Sharded data-parallel update ("ZeRO-1 style sharding of
 * the AdamW update + all-gather"): a rank runs b200pdm_adamw_step on its own 1/world slice of a gradient bucket only; for the
 * slices it received by all-gather it needs the bf16 shadow of the new masters and a zeroed gradient -- 10 B/param instead of
 * AdamW's 34. */
int b200pdm_refresh_shadow_zero(const float* p, void* shadow_bf16, float* zero, int64_t n, b200pdm_stream_t stream);

/* Sampling: one denoising step of the classifier-free-guidance loop, pdm/pipelines/pruning_pipelines.py:957-978
 * = chunk(2) + guidance combine (:972-974) + diffusers DDIMScheduler.step (eta 0, v_prediction, "leading" timesteps,
 * set_alpha_to_one False) (:981), fused.  model_out = U-Net output [2n, chw] fp32 (unconditional half first, :931);
 * latents [n, chw] is updated in place and copied into both halves of latent_in [2n, chw] (the next U-Net input, :961).
 * state = int32[2] {step index, internal counter} (zero it before a loop), t_dev = int64[2n] U-Net timestep tensor that
 * is advanced to timesteps[step + 1]: all step-dependent values live on the device, so the launches of one step can be
 * replayed from a CUDA graph.                                                                                          */
int b200pdm_cfg_ddim_step(const float* model_out, float* latents, float* latent_in, const float* alphas_cumprod,
                          const int64_t* timesteps, int* state, int64_t* t_dev, int n, int64_t chw, int num_steps,
                          int train_timesteps, float guidance_scale, b200pdm_stream_t stream);

/* The same fused step for diffusers PNDMScheduler(skip_prk_steps=True, steps_offset=1, set_alpha_to_one=False, v_prediction):
 * the scheduler scripts/metrics/generate_fid_images.py:113 loads for FID image generation.  Linear multistep over the last
 * four guided model outputs (step_plms): timesteps = int64[num_steps + 1] as PNDMScheduler.set_timesteps builds them (the
 * second entry repeated: num_steps + 1 U-Net evaluations); ets = fp32 [4][n*chw] ring of model outputs, cur_sample = fp32
 * [n*chw] (the sample the warm-up evaluation returns to); state = int32[4] {counter, done, #ets, ring head}, zeroed before a
 * loop.  Everything step-dependent lives on the device: one captured step replays num_steps + 1 times.                  */
int b200pdm_cfg_pndm_step(const float* model_out, float* latents, float* latent_in, const float* alphas_cumprod,
                          const int64_t* timesteps, int* state, int64_t* t_dev, float* ets, float* cur_sample, int n, int64_t chw,
                          int num_steps, int train_timesteps, float guidance_scale, b200pdm_stream_t stream);

/* Forward diffusion (diffusers DDIMScheduler.add_noise / get_velocity at trainer.py:2430,2443):
 * noisy = sa[t_b] x0 + sb[t_b] eps ; target = sa[t_b] eps - sb[t_b] x0 ; fp32.                             */
int b200pdm_diffusion_prep(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp,
                           const float* sqrt_1macp, float* noisy, float* vtarget, int batch, int64_t n_per_sample,
                           b200pdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200PDM_H_ */
