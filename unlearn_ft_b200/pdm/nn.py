"""Parameter arena + leaf parameter modules + closure-style differentiable ops for the B200 path.

Design
------
* All parameters of a model live in ONE flat fp32 buffer (``ParamArena.master``), with a same-layout bf16 *shadow*
  (tensor-core operands), a flat fp32 gradient buffer and (in the optimiser) flat AdamW moments.  One AdamW launch and
  a handful of NCCL all-reduces then cover the whole model.
* ``nn.Parameter``s are *views* into the master buffer carrying the reference/diffusers shapes, so ``state_dict()``
  keys and shapes match the reference.  Convolution weights are stored ``[O][kh][kw][I_ld]``
  (``I_ld = I`` rounded up to 8, pad lanes are zero and stay zero) and exposed as a strided ``[O, I, kh, kw]`` view.
* Every op is ``f(inputs) -> (output, bwd)`` where ``bwd(dy, ...)`` launches the hand-written backward kernels,
  ACCUMULATES parameter gradients straight into the arena's gradient buffer and returns the input gradient.
  No ATen arithmetic is involved; torch only owns memory.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional, Tuple

import torch
from torch import nn

from .. import _lib
from .. import kernels as K

BF16, F32 = torch.bfloat16, torch.float32
_ALIGN = 64  # elements; keeps every tensor 128-byte (bf16) / 256-byte (fp32) aligned


def _round(n, m):
    return (n + m - 1) // m * m


# ----------------------------------------------------------------------------------------------------------------
# leaf parameter holders (attribute names mirror torch.nn so that state-dict keys are the diffusers keys)
# ----------------------------------------------------------------------------------------------------------------
class PModule(nn.Module):
    """A module owning parameters inside a ParamArena. `_pspecs` = [(attr, logical_shape, kind)]."""

    _pspecs: List[Tuple[str, tuple, str]]

    def extra_repr(self):
        return ", ".join(f"{a}={s}" for a, s, _ in self._pspecs)


class PConv2d(PModule):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.ksize, self.stride_ = kernel_size, stride
        self.kernel_size, self.stride, self.padding = (kernel_size,) * 2, (stride,) * 2, (kernel_size // 2,) * 2
        self._pspecs = [("weight", (out_channels, in_channels, kernel_size, kernel_size), "conv"),
                        ("bias", (out_channels,), "vec")]


class PLinear(PModule):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self._pspecs = [("weight", (out_features, in_features), "mat")]
        if bias:
            self._pspecs.append(("bias", (out_features,), "vec"))
        else:
            self.bias = None


class PGroupNorm(PModule):
    def __init__(self, num_groups, num_channels, eps=1e-5):
        super().__init__()
        self.num_groups, self.num_channels, self.eps, self.affine = num_groups, num_channels, eps, True
        self._pspecs = [("weight", (num_channels,), "ones"), ("bias", (num_channels,), "vec")]


class PLayerNorm(PModule):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.normalized_shape, self.eps = (dim,), eps
        self._pspecs = [("weight", (dim,), "ones"), ("bias", (dim,), "vec")]


class ParamArena:
    """Flat storage for all parameters of a module tree (see module docstring)."""

    def __init__(self, root: nn.Module, device, trainable: bool = True, seed: Optional[int] = None):
        self.device = torch.device(device)
        self.trainable = trainable
        entries = []
        off = 0
        for mod_name, mod in root.named_modules():
            if not isinstance(mod, PModule):
                continue
            for attr, shape, kind in mod._pspecs:
                if kind == "conv":
                    O, I, kh, kw = shape
                    ild = _round(I, 8)
                    n_alloc = O * kh * kw * ild
                else:
                    n_alloc = 1
                    for s in shape:
                        n_alloc *= s
                entries.append((mod, mod_name, attr, shape, kind, off, n_alloc))
                off += _round(n_alloc, _ALIGN)
        self.numel = off
        self.logical_numel = sum(int(torch.Size(e[3]).numel()) for e in entries)
        self.master = torch.zeros(off, device=self.device, dtype=F32)
        self.shadow = torch.zeros(off, device=self.device, dtype=BF16)
        self.grad = torch.zeros(off, device=self.device, dtype=F32) if trainable else None
        self.entries = entries
        self.shadow_fresh = False
        self.ranges = {}  # module-name prefix -> (start, end) filled by block_range()
        for mod, mod_name, attr, shape, kind, o, n in entries:
            p = nn.Parameter(self._view(self.master, shape, kind, o), requires_grad=trainable)
            if trainable:
                p.grad = self._view(self.grad, shape, kind, o)
            setattr(mod, attr, p)
            # kernel-facing views
            if kind == "conv":
                O, I, kh, kw = shape
                ild = _round(I, 8)
                setattr(mod, "w16", self.shadow[o:o + n].view(O, kh * kw, ild)[:, :, :I])
                if trainable:
                    setattr(mod, "gw", self.grad[o:o + n].view(O, kh * kw, ild)[:, :, :I])
            elif kind == "mat":
                setattr(mod, "w16", self.shadow[o:o + n].view(shape))
                if trainable:
                    setattr(mod, "gw", self.grad[o:o + n].view(shape))
            mod.__dict__.setdefault("_arena_off", {})[attr] = (o, n)
        if seed is not None:
            self.init_default(seed)

    @staticmethod
    def _view(flat, shape, kind, o):
        if kind == "conv":
            O, I, kh, kw = shape
            ild = _round(I, 8)
            return flat.as_strided((O, I, kh, kw), (kh * kw * ild, 1, kw * ild, ild), o)
        n = 1
        for s in shape:
            n *= s
        return flat[o:o + n].view(shape)

    def module_range(self, root: nn.Module, sub: nn.Module) -> Tuple[int, int]:
        """[start, end) of the flat buffers covered by `sub`'s parameters (contiguous by construction order)."""
        ids = {id(m) for m in sub.modules()}
        offs = [(o, o + _round(n, _ALIGN)) for (m, _, _, _, _, o, n) in self.entries if id(m) in ids]
        return (min(a for a, _ in offs), max(b for _, b in offs)) if offs else (0, 0)

    @torch.no_grad()
    def init_default(self, seed: int):
        """torch/diffusers default initialisation (kaiming-uniform(a=sqrt(5)) weights, U(-1/sqrt(fan_in), ..) biases,
        norm weight 1 / bias 0), generated on the host with a seeded generator; used for random_init models."""
        g = torch.Generator().manual_seed(seed)
        fan_in_of = {}
        for mod, mod_name, attr, shape, kind, o, n in self.entries:
            p = getattr(mod, attr)
            if kind in ("conv", "mat"):
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                bound = (1.0 / fan_in) ** 0.5  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                p.copy_((torch.rand(shape, generator=g) * 2 - 1) * bound)
                fan_in_of[id(mod)] = fan_in
            elif kind == "ones":
                p.fill_(1.0)
            elif attr == "bias" and id(mod) in fan_in_of:
                bound = (1.0 / fan_in_of[id(mod)]) ** 0.5
                p.copy_((torch.rand(shape, generator=g) * 2 - 1) * bound)
            else:
                p.zero_()
        self.shadow_fresh = False

    def refresh_shadow(self):
        K.refresh_shadow(self.master, self.shadow)
        self.shadow_fresh = True

    def ensure_shadow(self):
        if not self.shadow_fresh:
            self.refresh_shadow()

    def reattach_grads(self):
        """After a foreign ``zero_grad(set_to_none=True)``: point every ``.grad`` back into the flat buffer (zeroed)."""
        if not self.trainable:
            return
        self.grad.zero_()
        for mod, mod_name, attr, shape, kind, o, n in self.entries:
            getattr(mod, attr).grad = self._view(self.grad, shape, kind, o)


# ----------------------------------------------------------------------------------------------------------------
# 4-D <-> 2-D views
# ----------------------------------------------------------------------------------------------------------------
def to4d(t2d: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """[B*H*W, C] (pitch ld) -> logical NCHW view with channels-last strides (no copy)."""
    C, ld = t2d.shape[1], t2d.stride(0)
    return t2d.as_strided((B, C, H, W), (H * W * ld, 1, W * ld, ld), t2d.storage_offset())


def as2d(t: torch.Tensor) -> torch.Tensor:
    """NCHW-shaped tensor -> [B*H*W, C] bf16 matrix view when it already is channels-last with a legal pitch;
    otherwise one layout copy (only for foreign inputs / autograd-accumulated gradients in another layout)."""
    B, C, H, W = t.shape
    if t.dtype != BF16:
        t = t.to(BF16)
    ld = t.stride(3) if W > 1 else (t.stride(2) if H > 1 else C)
    ok = (t.stride(1) == 1 and ld >= C and ld % 8 == 0 and (W == 1 or t.stride(3) == ld) and
          (H == 1 or t.stride(2) == W * ld) and (B == 1 or t.stride(0) == H * W * ld) and
          (t.data_ptr() % 16 == 0))
    if not ok:
        if C % 8 == 0:
            t = t.contiguous(memory_format=torch.channels_last)
            ld = C
        else:
            buf = K.alloc2d(B * H * W, C, t.device)
            buf.copy_(t.permute(0, 2, 3, 1).reshape(B * H * W, C))
            return buf
    return t.as_strided((B * H * W, C), (ld, 1), t.storage_offset())


# ----------------------------------------------------------------------------------------------------------------
# differentiable ops: f(...) -> (out, bwd)
# ----------------------------------------------------------------------------------------------------------------
def gn(x, m: PGroupNorm, B, hw, silu, need_bwd):
    y, stats = K.groupnorm_fwd(x, m.weight, m.bias, B, hw, m.num_groups, m.eps, silu)
    if not need_bwd:
        return y, None

    def bwd(dy, residual=None):
        return K.groupnorm_bwd(dy, x, m.weight, m.bias, stats, m.weight.grad, m.bias.grad, B, hw, m.num_groups, silu,
                               residual=residual)

    return y, bwd


def ln(x, m: PLayerNorm, need_bwd):
    y, mean, rstd = K.layernorm_fwd(x, m.weight, m.bias, m.eps, save=need_bwd)
    if not need_bwd:
        return y, None

    def bwd(dy, residual=None):
        return K.layernorm_bwd(dy, x, m.weight, mean, rstd, m.weight.grad, m.bias.grad, residual=residual)

    return y, bwd


# ----------------------------------------------------------------------------------------------------------------
# Weight-gradient branch.  A layer's wgrad (and bias column sum) only feeds the gradient arena, nothing downstream in the
# backward pass waits for it, so it is enqueued on a second stream: in the captured graph it becomes a parallel branch whose
# kernels fill the SMs the dgrad chain leaves idle (small 8x8 / 16x16 levels, tails of persistent kernels).  The branch is
# joined at the end of every top-level block's backward (blocks._BlockFn.backward), before the block's all-reduce is fired;
# the operands are kept alive until then so the allocator cannot recycle them under the side stream.
# ----------------------------------------------------------------------------------------------------------------
class _Side:
    enabled = not os.environ.get("B200PDM_SERIAL_WGRAD")     # (env: A/B measurement of the serial order)
    streams: dict = {}      # main stream -> its side stream (the teacher's stream and the student's each get their own)
    keep: dict = {}         # main stream -> operands kept alive until the join
    open_: set = set()
    inside = False


@contextlib.contextmanager
def side_branch(*operands):
    """Enqueue the body on the current stream's side stream (after everything enqueued on the current stream so far)."""
    if not _Side.enabled or _Side.inside or not torch.cuda.is_available() or (operands and not operands[0].is_cuda):
        yield               # (nested use runs inline on the side stream it is already on)
        return
    main = torch.cuda.current_stream()
    side = _Side.streams.get(main)
    if side is None:
        side = _Side.streams[main] = torch.cuda.Stream()
    side.wait_stream(main)
    _Side.keep.setdefault(main, []).extend(operands)
    _Side.open_.add(main)
    _Side.inside = True
    try:
        with torch.cuda.stream(side):
            yield
    finally:
        _Side.inside = False


wgrad_branch = side_branch


def wgrad_join():
    """Current stream waits for everything enqueued on its side stream; kept operands may be released afterwards."""
    main = torch.cuda.current_stream()
    if main in _Side.open_:
        main.wait_stream(_Side.streams[main])
        _Side.keep[main].clear()
        _Side.open_.discard(main)


side_join = wgrad_join


def linear(x, m: PLinear, need_bwd, residual=None, out_fp32=False, w16=None, gw=None, bias="own"):
    """y = x @ W^T + b (+ residual).  `w16`/`gw` override lets several adjacent parameters act as one fused matrix
    (e.g. to_q|to_k|to_v stacked in the arena)."""
    w = m.w16 if w16 is None else w16
    b = m.bias if bias == "own" else bias
    y = K.linear_fwd(x, w, b, residual, out_fp32=out_fp32)
    if not need_bwd:
        return y, None
    g = (m.gw if gw is None else gw)

    def bwd(dy, residual=None, need_dx=True):
        with wgrad_branch(dy, x):
            K.linear_wgrad(dy, x, g)
            if b is not None:
                K.colsum(dy, b.grad)
        return K.linear_dgrad(dy, w, residual=residual) if need_dx else None

    return y, bwd


def conv(x, m: PConv2d, B, H, W, need_bwd, rowbias=None, residual=None):
    """NHWC conv (3x3 pad 1 / 1x1; stride 1|2).  bwd(dy, residual=None, want_rowbias=False, need_dx=True) ->
    (dx, d_rowbias fp32 [B, Cout] | None)."""
    st = m.stride_
    y = K.conv_fwd(x, m.w16, B, H, W, m.out_channels, m.ksize, st, bias=m.bias, rowbias=rowbias, residual=residual)
    if not need_bwd:
        return y, None
    Ho, Wo = H // st, W // st

    def bwd(dy, residual=None, want_rowbias=False, need_dx=True):
        with wgrad_branch(dy, x):
            K.conv_wgrad(dy, x, m.gw, B, H, W, m.ksize, st)
            K.colsum(dy, m.bias.grad)
        drb = None
        if want_rowbias:
            drb = K.alloc2d(B, m.out_channels, dy.device, F32, zero=True)
            K.colsum_grouped(dy, drb, Ho * Wo)
        dx = None
        if need_dx:
            if st == 1:
                dx = K.conv_dgrad(dy, m.w16, B, H, W, m.in_channels, m.ksize, residual=residual)
            else:
                dyu = K.zero_insert2x(dy, B, Ho, Wo)
                dx = K.conv_dgrad(dyu, m.w16, B, H, W, m.in_channels, m.ksize, residual=residual)
        return dx, drb

    return y, bwd


def linear_geglu(x, m: PLinear, need_bwd):
    """GEGLUGated.forward (blocks.py:44-59) as ONE GEMM: y = (x W_v^T + b_v) * gelu_erf(x W_g^T + b_g) with the activation in
    the epilogue.  Frozen models never write the 2F-wide projection; trainable ones save it (bf16) for the backward pass:
    bwd(dy) -> dx through geglu_bwd (d value | d gate) and the projection's dgrad / wgrad / bias column sum."""
    y, pre = K.linear_geglu_fwd(x, m.w16, m.bias, save_pre=need_bwd)
    if not need_bwd:
        return y, None

    def bwd(dy):
        dp = K.geglu_bwd(dy, pre)
        with wgrad_branch(dp, x):
            K.linear_wgrad(dp, x, m.gw)
            K.colsum(dp, m.bias.grad)
        return K.linear_dgrad(dp, m.w16)

    return y, bwd


def geglu(p, need_bwd):
    y = K.geglu_fwd(p)
    if not need_bwd:
        return y, None
    return y, (lambda dy: K.geglu_bwd(dy, p))


def attention(q, k, v, B, heads, Lq, Lk, need_bwd, out=None):
    """softmax(q k^T / 8) v per (sample, head), head_dim 64, no mask (reference blocks.py:275-277).

    q: [B*Lq, >=heads*64] view, k/v: [B*Lk, ...] views (column slices of fused projection outputs are fine).
    Forward and backward are the fused tcgen05 flash-style kernels (csrc/attention.cu): scores and probabilities
    never reach HBM; the backward recomputes them per tile from q, k and the saved log-sum-exp.
    bwd(do, dq, dk, dv) writes the gradients into the given bf16 views.
    """
    scale = 64 ** -0.5
    o, lse = K.attention_fwd(q, k, v, B, heads, Lq, Lk, scale, out=out, want_lse=need_bwd)
    if not need_bwd:
        return o, None

    def bwd(do, dq, dk, dv):
        K.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, heads, Lq, Lk, scale)

    return o, bwd


def concat_channels(a, b):
    """cat([a, b], dim=channels) for [M, C] matrices; the backward is two views of the incoming gradient."""
    M, Ca, Cb = a.shape[0], a.shape[1], b.shape[1]
    out = K.alloc2d(M, Ca + Cb, a.device)
    K.copy2d(a, out[:, :Ca])
    K.copy2d(b, out[:, Ca:])
    return out
