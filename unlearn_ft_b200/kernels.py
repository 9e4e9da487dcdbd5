"""Tensor-level wrappers over the C ABI (``include/b200pdm.h``).

Convention: an activation matrix is a 2-D bf16 ``torch.Tensor`` of shape ``[rows, C]`` with strides ``(ld, 1)``,
``ld % 8 == 0`` and a 16-byte aligned base -- i.e. a channels-last feature map ``[B, H, W, C]`` flattened over pixels.
``alloc2d`` creates such tensors; column slices of them (``t[:, a:b]`` with ``a % 8 == 0``) stay valid.
PyTorch is used for device memory and the current CUDA stream only; every numerical operation is a call into
``libb200pdm.so``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmDesc, Operand, check

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def round8(n: int) -> int:
    return (n + 7) // 8 * 8


def round16(n: int) -> int:
    return (n + 15) // 16 * 16


def alloc2d(rows: int, cols: int, device=None, dtype=BF16, zero: bool = False) -> torch.Tensor:
    """[rows, cols] matrix whose row pitch is a multiple of 16 elements: every row of a bf16 matrix starts on a 32-byte sector
    (the GEMM epilogue's 256-bit row accesses need that; TMA needs 16 bytes), also for pruned widths such as 170 / 340 / 680."""
    ld = round16(cols)
    buf = (torch.zeros if zero else torch.empty)(rows, ld, device=device or "cuda", dtype=dtype)
    return buf[:, :cols] if ld != cols else buf


_WS_BYTES = {}


def _workspace(key, query, device):
    """(tensor | None, bytes): caller-owned scratch for a GEMM-class call, sized by the library's *_workspace() twin (cached per
    shape).  It comes from torch's stream-ordered allocator, so inside a CUDA-graph capture it lives in the graph's pool."""
    n = _WS_BYTES.get(key)
    if n is None:
        n = _WS_BYTES[key] = int(query())
    if n == 0:
        return None, 0
    return torch.empty(n, device=device, dtype=torch.uint8), n


def _chk2d(t: torch.Tensor, name: str, dtype=BF16):
    if t.dim() != 2 or t.dtype != dtype or t.stride(1) != 1 or not t.is_cuda:
        raise ValueError(f"{name}: expected a 2-D {dtype} CUDA tensor with unit inner stride, got {tuple(t.shape)} "
                         f"{t.dtype} strides {t.stride()}")
    if dtype == BF16 and (t.stride(0) % 8 or t.data_ptr() % 16):
        raise ValueError(f"{name}: row pitch must be a multiple of 8 elements and the base 16-byte aligned")


# ------------------------------------------------------------------------------------------------ GEMM-class
def linear_fwd(x, w, bias=None, residual=None, out=None, out_fp32=False):
    """out[M,N] = x[M,K] @ w[N,K]^T + bias + residual  (F.linear; reference blocks.py:49,244,251-252,283)."""
    _chk2d(x, "x"), _chk2d(w, "w")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"linear_fwd: K mismatch {x.shape} vs {w.shape}")
    if out is None:
        out = alloc2d(M, N, x.device, F32 if out_fp32 else BF16)
    if residual is not None:
        _chk2d(residual, "residual")
    L = _lib.lib()
    ws, nws = _workspace(("lf", M, N, K, out.dtype == F32),
                         lambda: L.b200pdm_linear_fwd_workspace(M, N, K, int(out.dtype == F32)), x.device)
    check(L.b200pdm_linear_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                               _ptr(residual), residual.stride(0) if residual is not None else 0,
                               out.data_ptr(), out.stride(0), int(out.dtype == F32), M, N, K, _ptr(ws), nws, _stream()),
          "linear_fwd")
    return out


def linear_geglu_fwd(x, w, bias, save_pre=False, out=None):
    """GEGLU input projection with the activation in the GEMM epilogue (blocks.py:44-59): w [2F, K], bias [2F] ->
    (out [M, F], pre [M, 2F] | None).  `pre` = the (value | gate) pre-activations geglu_bwd needs; None for frozen models."""
    _chk2d(x, "x"), _chk2d(w, "w")
    M, K = x.shape
    F2 = w.shape[0]
    if w.shape[1] != K or F2 % 2:
        raise ValueError(f"linear_geglu_fwd: bad weight shape {tuple(w.shape)} for x {tuple(x.shape)}")
    Fh = F2 // 2
    if out is None:
        out = alloc2d(M, Fh, x.device)
    pre = alloc2d(M, F2, x.device) if save_pre else None
    check(_lib.lib().b200pdm_linear_geglu_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), out.data_ptr(),
                                              out.stride(0), _ptr(pre), pre.stride(0) if pre is not None else 0, M, Fh, K,
                                              _stream()), "linear_geglu_fwd")
    return out, pre


def linear_dgrad(dy, w, residual=None, out=None):
    """dx[M,K] = dy[M,N] @ w[N,K] (+ residual)."""
    _chk2d(dy, "dy"), _chk2d(w, "w")
    M, N = dy.shape
    K = w.shape[1]
    if out is None:
        out = alloc2d(M, K, dy.device)
    L = _lib.lib()
    ws, nws = _workspace(("ld", M, N, K), lambda: L.b200pdm_linear_dgrad_workspace(M, N, K), dy.device)
    check(L.b200pdm_linear_dgrad(dy.data_ptr(), dy.stride(0), w.data_ptr(), w.stride(0), _ptr(residual),
                                 residual.stride(0) if residual is not None else 0, out.data_ptr(),
                                 out.stride(0), M, N, K, _ptr(ws), nws, _stream()), "linear_dgrad")
    return out


def linear_wgrad(dy, x, dw):
    """dw[N,K] (fp32) += dy[M,N]^T @ x[M,K]."""
    _chk2d(dy, "dy"), _chk2d(x, "x"), _chk2d(dw, "dw", F32)
    M, N = dy.shape
    K = x.shape[1]
    check(_lib.lib().b200pdm_linear_wgrad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(),
                                          dw.stride(0), M, N, K, _stream()), "linear_wgrad")
    return dw


def conv_fwd(x, w, B, H, W, c_out, ksize=3, stride=1, bias=None, rowbias=None, residual=None, out=None):
    """NHWC implicit-GEMM convolution. x: [B*H*W, Cin]; w: bf16 [Cout, taps, Cin] view with strides (taps*ild, ild, 1)."""
    _chk2d(x, "x")
    c_in = x.shape[1]
    Ho, Wo = H // stride, W // stride
    if out is None:
        out = alloc2d(B * Ho * Wo, c_out, x.device)
    ild = w.stride(1) if w.dim() == 3 else w.stride(0)
    L = _lib.lib()
    ws, nws = _workspace(("cf", B, H, W, c_in, c_out, ksize, stride),
                         lambda: L.b200pdm_conv_fwd_workspace(B, H, W, c_in, c_out, ksize, stride), x.device)
    check(L.b200pdm_conv_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), ild, _ptr(bias), _ptr(rowbias),
                             rowbias.stride(0) if rowbias is not None else 0, _ptr(residual),
                             residual.stride(0) if residual is not None else 0, out.data_ptr(),
                             out.stride(0), B, H, W, c_in, c_out, ksize, stride, _ptr(ws), nws, _stream()), "conv_fwd")
    return out


def conv_fwd_nopad(x, w, B, H, W, c_out, stride=2, bias=None, out=None):
    """3x3 convolution whose window starts AT the pixel (zeros beyond the right / bottom edge): F.pad(x, (0, 1, 0, 1)) +
    padding-0 conv = diffusers Downsample2D(padding=0) of the VAE encoder.  Forward only."""
    _chk2d(x, "x")
    c_in = x.shape[1]
    Ho, Wo = H // stride, W // stride
    if out is None:
        out = alloc2d(B * Ho * Wo, c_out, x.device)
    ild = w.stride(1) if w.dim() == 3 else w.stride(0)
    check(_lib.lib().b200pdm_conv_fwd_nopad(x.data_ptr(), x.stride(0), w.data_ptr(), ild, _ptr(bias), out.data_ptr(),
                                            out.stride(0), B, H, W, c_in, c_out, stride, _stream()), "conv_fwd_nopad")
    return out


def conv_dgrad(dy, w, B, H, W, c_in, ksize=3, residual=None, out=None):
    """dx[B*H*W, Cin] for a stride-1 convolution (dy at the same resolution)."""
    _chk2d(dy, "dy")
    c_out = dy.shape[1]
    if out is None:
        out = alloc2d(B * H * W, c_in, dy.device)
    ild = w.stride(1) if w.dim() == 3 else w.stride(0)
    L = _lib.lib()
    ws, nws = _workspace(("cd", B, H, W, c_in, c_out, ksize),
                         lambda: L.b200pdm_conv_dgrad_workspace(B, H, W, c_in, c_out, ksize), dy.device)
    check(L.b200pdm_conv_dgrad(dy.data_ptr(), dy.stride(0), w.data_ptr(), ild, _ptr(residual),
                               residual.stride(0) if residual is not None else 0, out.data_ptr(),
                               out.stride(0), B, H, W, c_in, c_out, ksize, _ptr(ws), nws, _stream()), "conv_dgrad")
    return out


def conv_wgrad(dy, x, dw, B, H, W, ksize=3, stride=1):
    """dw (fp32 [Cout, taps, Cin] view, strides (taps*ild, ild, 1)) += conv weight gradient."""
    _chk2d(dy, "dy"), _chk2d(x, "x")
    c_out, c_in = dy.shape[1], x.shape[1]
    ild = dw.stride(1) if dw.dim() == 3 else dw.stride(0)
    check(_lib.lib().b200pdm_conv_wgrad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(), ild, B,
                                        H, W, c_in, c_out, ksize, stride, _stream()), "conv_wgrad")
    return dw


def gemm(desc: GemmDesc, device=None):
    L = _lib.lib()
    n = int(L.b200pdm_gemm_workspace(C.byref(desc)))
    ws = torch.empty(n, device=device or "cuda", dtype=torch.uint8) if n else None
    check(L.b200pdm_gemm(C.byref(desc), _ptr(ws), n, _stream()), "gemm")


def bmm(a, b, out, *, a_mn=False, b_mn=False, M, N, K, Z1=1, Z2=1, a_ld, a_bs=(0, 0), b_ld, b_bs=(0, 0), o_ld,
        o_bs=(0, 0), alpha=1.0):
    """Batched GEMM over raw strided bf16 operands (attention). Pointers are taken from the tensors' data_ptr()."""
    d = GemmDesc()
    d.a = Operand(mode=_lib.OP_MN2D if a_mn else _lib.OP_K2D, ptr=a.data_ptr(), ld=a_ld, bs1=a_bs[0], bs2=a_bs[1])
    d.b = Operand(mode=_lib.OP_MN2D if b_mn else _lib.OP_K2D, ptr=b.data_ptr(), ld=b_ld, bs1=b_bs[0], bs2=b_bs[1])
    d.M, d.N, d.K, d.Z1, d.Z2 = M, N, K, Z1, Z2
    d.out, d.out_fp32, d.ldo, d.obs1, d.obs2 = out.data_ptr(), int(out.dtype == F32), o_ld, o_bs[0], o_bs[1]
    d.alpha = alpha
    gemm(d, out.device)
    return out


# ------------------------------------------------------------------------------------------------ norms
def groupnorm_fwd(x, gamma, beta, B, hw, groups, eps, silu, out=None):
    _chk2d(x, "x")
    Cn = x.shape[1]
    if out is None:
        out = alloc2d(B * hw, Cn, x.device)
    L = _lib.lib()
    stats = torch.empty(4 * B * round8(Cn), device=x.device, dtype=F32)  # coefficient tables (kept for the backward pass)
    key = ("gns", B, hw, Cn)
    n = _WS_BYTES.get(key)
    if n is None:
        n = _WS_BYTES[key] = int(L.b200pdm_groupnorm_scratch_floats(B, hw, Cn))
    scratch = torch.empty(n, device=x.device, dtype=F32)
    check(L.b200pdm_groupnorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                  out.stride(0), stats.data_ptr(), scratch.data_ptr(), B, hw, Cn, groups, eps, int(silu),
                                  _stream()), "groupnorm_fwd")
    return out, stats


def groupnorm_bwd(dy, x, gamma, beta, stats, dgamma, dbeta, B, hw, groups, silu, out=None, residual=None):
    _chk2d(dy, "dy"), _chk2d(x, "x")
    Cn = x.shape[1]
    if out is None:
        out = alloc2d(B * hw, Cn, x.device)
    key = ("gnb", B, hw, Cn)
    n = _WS_BYTES.get(key)
    if n is None:
        n = _WS_BYTES[key] = int(_lib.lib().b200pdm_groupnorm_bwd_workspace_floats(B, hw, Cn))
    ws = torch.empty(n, device=x.device, dtype=F32)
    check(_lib.lib().b200pdm_groupnorm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gamma.data_ptr(),
                                           beta.data_ptr(), stats.data_ptr(), _ptr(residual),
                                           residual.stride(0) if residual is not None else 0, out.data_ptr(),
                                           out.stride(0), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), B, hw, Cn,
                                           groups,
                                           int(silu), _stream()), "groupnorm_bwd")
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-5, save=True, out=None):
    _chk2d(x, "x")
    rows, Cn = x.shape
    if out is None:
        out = alloc2d(rows, Cn, x.device)
    mean = torch.empty(rows, device=x.device, dtype=F32) if save else None
    rstd = torch.empty(rows, device=x.device, dtype=F32) if save else None
    check(_lib.lib().b200pdm_layernorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                           out.stride(0), _ptr(mean), _ptr(rstd), rows, Cn, eps, _stream()),
          "layernorm_fwd")
    return out, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, out=None, residual=None):
    _chk2d(dy, "dy"), _chk2d(x, "x")
    rows, Cn = x.shape
    if out is None:
        out = alloc2d(rows, Cn, x.device)
    check(_lib.lib().b200pdm_layernorm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gamma.data_ptr(),
                                           mean.data_ptr(), rstd.data_ptr(), _ptr(residual),
                                           residual.stride(0) if residual is not None else 0, out.data_ptr(),
                                           out.stride(0), dgamma.data_ptr(), dbeta.data_ptr(), rows, Cn, _stream()),
          "layernorm_bwd")
    return out


# ------------------------------------------------------------------------------------------------ elementwise
def geglu_fwd(proj, out=None):
    _chk2d(proj, "proj")
    rows, F2 = proj.shape
    Fh = F2 // 2
    if out is None:
        out = alloc2d(rows, Fh, proj.device)
    check(_lib.lib().b200pdm_geglu_fwd(proj.data_ptr(), proj.stride(0), out.data_ptr(), out.stride(0), rows, Fh,
                                       _stream()), "geglu_fwd")
    return out


def geglu_bwd(dout, proj, out=None):
    rows, F2 = proj.shape
    if out is None:
        out = alloc2d(rows, F2, proj.device)
    check(_lib.lib().b200pdm_geglu_bwd(dout.data_ptr(), dout.stride(0), proj.data_ptr(), proj.stride(0), out.data_ptr(),
                                       out.stride(0), rows, F2 // 2, _stream()), "geglu_bwd")
    return out


def softmax_fwd(s, p, rows, cols, scale):
    check(_lib.lib().b200pdm_softmax_fwd(s.data_ptr(), s.stride(-2), p.data_ptr(), p.stride(-2), rows, cols, scale,
                                         _stream()), "softmax_fwd")
    return p


def attention_fwd(q, k, v, B, heads, Lq, Lk, scale, out=None, want_lse=False, causal=False):
    """Fused attention forward (head_dim 64). q: [B*Lq, >=heads*64] view, k/v: [B*Lk, ...] views.  causal: query i attends
    keys <= i (CLIP text encoder)."""
    if out is None:
        out = alloc2d(B * Lq, heads * 64, q.device)
    lse = torch.empty(B * heads * Lq, device=q.device, dtype=F32) if want_lse else None
    check(_lib.lib().b200pdm_attention_fwd_ex(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                              v.stride(0), out.data_ptr(), out.stride(0), _ptr(lse), B, heads, Lq, Lk,
                                              scale, int(causal), _stream()), "attention_fwd")
    return out, lse


def gate_scale(x, gate, rows_per_sample, period, group_size, out=None):
    """Width gate: y[r, c] = x[r, c] * gate[(r // rows_per_sample) % Bg][(c % period) // group_size]; gate fp32 [Bg, width] view."""
    _chk2d(x, "x")
    if out is None:
        out = alloc2d(x.shape[0], x.shape[1], x.device)
    check(_lib.lib().b200pdm_gate_scale(x.data_ptr(), x.stride(0), gate.data_ptr(), gate.stride(0), out.data_ptr(), out.stride(0),
                                        x.shape[0], x.shape[1], rows_per_sample, period, group_size, gate.shape[0], _stream()),
          "gate_scale")
    return out


def gate_grad(dy, x, dgate, rows_per_sample, period, group_size):
    """dgate[bg, g] (fp32 view, accumulated) += sum dy * x over the elements gate[bg, g] multiplies."""
    check(_lib.lib().b200pdm_gate_grad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dgate.data_ptr(), dgate.stride(0),
                                       x.shape[0], x.shape[1], rows_per_sample, period, group_size, dgate.shape[0], _stream()),
          "gate_grad")


def depth_blend(inp, out_, gate, rows_per_sample):
    """(1 - m) * inp + m * out_ per sample; gate fp32 [Bg] (contiguous)."""
    y = alloc2d(inp.shape[0], inp.shape[1], inp.device)
    check(_lib.lib().b200pdm_depth_blend(inp.data_ptr(), inp.stride(0), out_.data_ptr(), out_.stride(0), gate.data_ptr(),
                                         y.data_ptr(), y.stride(0), inp.shape[0], inp.shape[1], rows_per_sample, gate.numel(),
                                         _stream()), "depth_blend")
    return y


def depth_blend_bwd(dy, inp, out_, gate, dgate, rows_per_sample):
    d_inp, d_out = alloc2d(inp.shape[0], inp.shape[1], inp.device), alloc2d(inp.shape[0], inp.shape[1], inp.device)
    check(_lib.lib().b200pdm_depth_blend_bwd(dy.data_ptr(), dy.stride(0), inp.data_ptr(), inp.stride(0), out_.data_ptr(),
                                             out_.stride(0), gate.data_ptr(), d_inp.data_ptr(), d_inp.stride(0), d_out.data_ptr(),
                                             d_out.stride(0), _ptr(dgate), inp.shape[0], inp.shape[1], rows_per_sample,
                                             gate.numel(), _stream()), "depth_blend_bwd")
    return d_inp, d_out


def gelu(x, out=None):
    _chk2d(x, "x")
    if out is None:
        out = alloc2d(x.shape[0], x.shape[1], x.device)
    check(_lib.lib().b200pdm_gelu(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], _stream()),
          "gelu")
    return out


def clip_embed(ids, token_embedding, position_embedding):
    """ids int64 [B, L]; fp32 tables [vocab, C], [>= L, C] -> bf16 [B*L, C]."""
    B, L = ids.shape
    vocab, Cn = token_embedding.shape
    ids = ids.to(torch.int64).contiguous()
    out = alloc2d(B * L, Cn, ids.device)
    check(_lib.lib().b200pdm_clip_embed(ids.data_ptr(), token_embedding.data_ptr(), position_embedding.data_ptr(), out.data_ptr(),
                                        out.stride(0), B * L, L, Cn, vocab, _stream()), "clip_embed")
    return out


def vae_sample(moments, B, hw, latent_channels, scaling_factor, eps=None, want_mean=False):
    """moments bf16 [B*hw, 2*Cz] -> latents fp32 [B, Cz, hw] (flat spatial), optional mean."""
    z = torch.empty(B, latent_channels, hw, device=moments.device, dtype=F32)
    mean = torch.empty_like(z) if want_mean else None
    check(_lib.lib().b200pdm_vae_sample(moments.data_ptr(), moments.stride(0), _ptr(eps), z.data_ptr(), _ptr(mean), B,
                                        latent_channels, hw, float(scaling_factor), _stream()), "vae_sample")
    return z, mean


def attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, B, heads, Lq, Lk, scale):
    """Fused attention backward; dq/dk/dv are bf16 output views (column slices allowed)."""
    n_delta = (B * heads * Lq + 3) // 4 * 4
    ws = torch.empty(n_delta + B * Lq * heads * 64, device=q.device, dtype=F32)
    check(_lib.lib().b200pdm_attention_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                           v.stride(0), out.data_ptr(), out.stride(0), dout.data_ptr(), dout.stride(0),
                                           lse.data_ptr(), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0),
                                           dv.data_ptr(), dv.stride(0), ws.data_ptr(), B, heads, Lq, Lk, scale,
                                           _stream()), "attention_bwd")


def colsum(x, out):
    """out[n] (fp32) += sum_m x[m, n]."""
    _chk2d(x, "x")
    check(_lib.lib().b200pdm_colsum(x.data_ptr(), x.stride(0), out.data_ptr(), x.shape[0], x.shape[1], _stream()),
          "colsum")
    return out


def colsum_grouped(x, out, rows_per_group):
    """out[g, n] (fp32, zero-initialised by the caller) += sum of the rows of group g."""
    _chk2d(x, "x")
    check(_lib.lib().b200pdm_colsum_grouped(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0],
                                            x.shape[1], rows_per_group, _stream()), "colsum_grouped")
    return out


def cast_f32_to_bf16(x, out=None):
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(_lib.lib().b200pdm_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "cast")
    return out


def cast2d_f32_to_bf16(x):
    """fp32 [rows, C] made by alloc2d (pitch round16(C)) -> bf16 matrix with the same pitch."""
    rows, cols = x.shape
    ld = x.stride(0)
    if ld != round16(cols) or x.stride(1) != 1:
        raise ValueError("cast2d_f32_to_bf16 expects an alloc2d-style fp32 matrix")
    out = torch.empty(rows, ld, device=x.device, dtype=BF16)
    check(_lib.lib().b200pdm_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), rows * ld - (ld - cols), _stream()), "cast2d")
    return out[:, :cols] if ld != cols else out


def add(a, b, out=None):
    _chk2d(a, "a"), _chk2d(b, "b")
    if out is None:
        out = alloc2d(a.shape[0], a.shape[1], a.device)
    check(_lib.lib().b200pdm_add(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0),
                                 a.shape[0], a.shape[1], _stream()), "add")
    return out


def copy2d(src, dst):
    _chk2d(src, "src"), _chk2d(dst, "dst")
    check(_lib.lib().b200pdm_copy2d(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), src.shape[0],
                                    src.shape[1], _stream()), "copy2d")
    return dst


def silu_f32_to_bf16(x):
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(_lib.lib().b200pdm_silu_f32_to_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "silu")
    return y


def silu_bwd(dy, x):
    """dy: bf16 contiguous, x: fp32 contiguous (saved pre-activation) -> dx bf16."""
    dx = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(_lib.lib().b200pdm_silu_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), _stream()), "silu_bwd")
    return dx


def upsample2x_fwd(x, B, H, W):
    out = alloc2d(B * 4 * H * W, x.shape[1], x.device)
    check(_lib.lib().b200pdm_upsample2x_fwd(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), B, H, W,
                                            x.shape[1], _stream()), "upsample2x_fwd")
    return out


def upsample2x_bwd(dy, B, H, W):
    """(H, W) is the LOW resolution."""
    out = alloc2d(B * H * W, dy.shape[1], dy.device)
    check(_lib.lib().b200pdm_upsample2x_bwd(dy.data_ptr(), dy.stride(0), out.data_ptr(), out.stride(0), B, H, W,
                                            dy.shape[1], _stream()), "upsample2x_bwd")
    return out


def zero_insert2x(x, B, H, W):
    out = alloc2d(B * 4 * H * W, x.shape[1], x.device)
    check(_lib.lib().b200pdm_zero_insert2x(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), B, H, W,
                                           x.shape[1], _stream()), "zero_insert2x")
    return out


def nchw_f32_to_nhwc_bf16(x):
    B, Cn, H, W = x.shape
    x = x.contiguous().float()
    out = alloc2d(B * H * W, Cn, x.device)
    check(_lib.lib().b200pdm_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), out.stride(0), B, Cn, H * W, _stream()),
          "nchw_to_nhwc")
    return out


def nhwc_bf16_to_nchw_f32(x, B, H, W):
    Cn = x.shape[1]
    out = torch.empty(B, Cn, H, W, device=x.device, dtype=F32)
    check(_lib.lib().b200pdm_nhwc_bf16_to_nchw_f32(x.data_ptr(), x.stride(0), out.data_ptr(), B, Cn, H * W, _stream()),
          "nhwc_to_nchw")
    return out


def timestep_embedding(t, dim):
    t = t.to(torch.int64).contiguous()
    out = alloc2d(t.shape[0], dim, t.device)
    check(_lib.lib().b200pdm_timestep_embedding(t.data_ptr(), out.data_ptr(), out.stride(0), t.shape[0], dim, _stream()),
          "timestep_embedding")
    return out


# ------------------------------------------------------------------------------------------------ loss / optimiser
def kd_loss_fused(pred, target, teacher, feats_s, feats_t, w_diff, w_kd, w_block, snr_w=None, alphas_cumprod=None,
                  timesteps=None, snr_gamma=5.0, v_prediction=True, want_grad=True):
    """The whole loss of trainer.py:2451-2486 in ONE launch (see include/b200pdm.h).  pred/target/teacher fp32 [B, ...]
    contiguous; feats_*: lists of dense bf16 maps.  Returns (sums fp32[4] = diff, kd, block, total; dpred; [ds_k])."""
    B = pred.shape[0]
    n = pred.numel() // B
    n_pairs = len(feats_s)
    dpred = torch.empty_like(pred) if want_grad else None
    dfeats = [torch.empty_like(s) for s in feats_s] if want_grad else [None] * n_pairs
    pairs = (_lib.FeaturePair * max(n_pairs, 1))()
    for i, (s, t) in enumerate(zip(feats_s, feats_t)):
        if s.dtype != BF16 or t.dtype != BF16 or s.shape != t.shape or s.stride() != t.stride():
            raise ValueError("kd_loss_fused: feature pairs must be bf16 maps of identical shape and layout")
        pairs[i].s, pairs[i].t, pairs[i].ds, pairs[i].numel = s.data_ptr(), t.data_ptr(), _ptr(dfeats[i]), s.numel()
    ws_bytes = _lib.lib().b200pdm_kd_loss_workspace(n_pairs)
    ws = torch.empty(ws_bytes // 4, device=pred.device, dtype=F32)
    sums = torch.empty(4, device=pred.device, dtype=F32)
    check(_lib.lib().b200pdm_kd_loss_fused(pred.data_ptr(), _ptr(target), _ptr(teacher), _ptr(snr_w), _ptr(alphas_cumprod),
                                           _ptr(timesteps), float(snr_gamma if snr_gamma is not None else 0.0),
                                           int(v_prediction), _ptr(dpred), B, n, w_diff, w_kd, pairs, n_pairs, w_block,
                                           sums.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "kd_loss_fused")
    return sums, dpred, dfeats


def adamw_step(p, g, m, v, shadow, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, zero_grad=True):
    check(_lib.lib().b200pdm_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(shadow), p.numel(),
                                        lr, beta1, beta2, eps, wd, step, grad_scale, int(zero_grad), _stream()),
          "adamw_step")


def adamw_step_dyn(p, g, m, v, shadow, dyn, beta1, beta2, eps, wd, grad_scale=1.0, zero_grad=True):
    """AdamW with {lr, bias corrections} read from the device tensor `dyn` (fp32[3]); CUDA-graph replayable."""
    check(_lib.lib().b200pdm_adamw_step_dyn(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(shadow), p.numel(),
                                            dyn.data_ptr(), beta1, beta2, eps, wd, grad_scale, int(zero_grad), _stream()),
          "adamw_step_dyn")


def refresh_shadow_zero(p, shadow, grad):
    """shadow = bf16(p) and grad = 0 in one pass (received slices of a sharded optimizer step)."""
    check(_lib.lib().b200pdm_refresh_shadow_zero(p.data_ptr(), shadow.data_ptr(), _ptr(grad), p.numel(), _stream()),
          "refresh_shadow_zero")


def refresh_shadow(p, shadow):
    check(_lib.lib().b200pdm_refresh_shadow(p.data_ptr(), shadow.data_ptr(), p.numel(), _stream()), "refresh_shadow")


def cfg_ddim_step(model_out, latents, latent_in, alphas_cumprod, timesteps, state, t_dev, num_steps, train_timesteps, guidance):
    """Fused CFG combine + DDIM step (see include/b200pdm.h); all tensors fp32 / int64 / int32 on the device, in place."""
    n = latents.shape[0]
    chw = latents.numel() // n
    check(_lib.lib().b200pdm_cfg_ddim_step(model_out.data_ptr(), latents.data_ptr(), latent_in.data_ptr(),
                                           alphas_cumprod.data_ptr(), timesteps.data_ptr(), state.data_ptr(), t_dev.data_ptr(),
                                           n, chw, num_steps, train_timesteps, float(guidance), _stream()), "cfg_ddim_step")


def cfg_pndm_step(model_out, latents, latent_in, alphas_cumprod, timesteps, state, t_dev, ets, cur_sample, num_steps,
                  train_timesteps, guidance):
    """Fused CFG combine + PNDM (PLMS) step (see include/b200pdm.h); in place, all state on the device."""
    n = latents.shape[0]
    chw = latents.numel() // n
    check(_lib.lib().b200pdm_cfg_pndm_step(model_out.data_ptr(), latents.data_ptr(), latent_in.data_ptr(), alphas_cumprod.data_ptr(),
                                           timesteps.data_ptr(), state.data_ptr(), t_dev.data_ptr(), ets.data_ptr(),
                                           cur_sample.data_ptr(), n, chw, num_steps, train_timesteps, float(guidance), _stream()),
          "cfg_pndm_step")


def diffusion_prep(x0, noise, t, sqrt_acp, sqrt_1macp):
    B = x0.shape[0]
    noisy, vt = torch.empty_like(x0), torch.empty_like(x0)
    check(_lib.lib().b200pdm_diffusion_prep(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), sqrt_acp.data_ptr(),
                                            sqrt_1macp.data_ptr(), noisy.data_ptr(), vt.data_ptr(), B,
                                            x0.numel() // B, _stream()), "diffusion_prep")
    return noisy, vt
