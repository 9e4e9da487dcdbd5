"""B200-native mirror of ``pdm/models/unet/blocks.py`` (reference) -- the PRUNED forms of the gated blocks.

The reference builds full-width diffusers blocks, attaches gates and then physically slices the weights in ``prune()``
(blocks.py:62-76,131-138,163-196,435-475,647-702,1324-1334).  Here every block is constructed directly at its pruned
width from the per-gate keep-masks (``keep`` = ascending surviving indices, exactly the boolean-mask selection of the
reference), holds its parameters in the model's flat arena, and runs forward/backward through the hand-written
sm_100a kernels (``unlearn_ft_b200.kernels``).  Class names, sub-module names and therefore state-dict keys follow the
reference/diffusers.

Each block module's ``forward`` is one ``torch.autograd.Function`` (so the reference trainer's forward hooks on
``down_blocks[i]`` / ``mid_block`` / ``up_blocks[i]`` see graph-connected outputs, trainer.py:557-572) whose backward
is a hand-scheduled chain of kernel launches with all gradient merges fused into kernel epilogues.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from ... import nn as bnn
from ...nn import PConv2d, PGroupNorm, PLayerNorm, PLinear, as2d, to4d
from .... import kernels as K

BF16, F32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------------------------------------------
# generic block-level autograd bridge
# ----------------------------------------------------------------------------------------------------------------
class _BlockFn(torch.autograd.Function):
    """forward(runner, need_bwd, anchor, *tensors): runner(need_bwd, *tensors) -> (outputs, bwd).
    `anchor` is any tensor that requires grad (the arena's master buffer) so that a graph is recorded even when no
    activation input requires grad (the network input does not)."""

    @staticmethod
    def forward(ctx, runner, need_bwd, anchor, owner, *tensors):
        outs, bwd = runner(need_bwd, *tensors)
        ctx.bwd = bwd
        ctx.n_in = len(tensors)
        ctx.owner = owner
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.bwd is None:
            raise RuntimeError("backward through a block that ran without gradient bookkeeping")
        in_grads = ctx.bwd(*grads)
        bnn.wgrad_join()  # the block's weight gradients ran on the side stream (nn.wgrad_branch): final from here on
        ctx.bwd = None  # release saved activations
        # this block's parameter gradients are final now: let the data-parallel reducer start on them while the rest of
        # the backward pass runs (GradReducer installs the callbacks; None on a single GPU)
        cb = None if ctx.owner is None else getattr(ctx.owner[0], ctx.owner[1], None)
        if cb is not None:
            cb()
        in_grads = tuple(in_grads) + (None,) * (ctx.n_in - len(in_grads))
        return (None, None, None, None) + in_grads


def run_block(runner, anchor, *tensors, owner=None):
    """`owner` = (module, attribute name) of an optional zero-argument callback fired after the block's backward."""
    need_bwd = torch.is_grad_enabled() and anchor is not None and anchor.requires_grad
    if not need_bwd:
        outs, _ = runner(False, *tensors)
        return tuple(outs)
    return _BlockFn.apply(runner, True, anchor, owner, *tensors)


# ----------------------------------------------------------------------------------------------------------------
# runtime gates of the un-pruned network (reference gates.py + the `if not self.pruned:` branches of blocks.py)
# ----------------------------------------------------------------------------------------------------------------
class RuntimeGates:
    """Gate values of one forward: `g` fp32 [Bg, n_gates] on the device (all width gates in get_structure() order, then the
    depth gates -- the arch-vector layout of hypernet.py:101-126) and `dg`, the same-shaped accumulator of their gradients.
    Gated modules hold the column range(s) they own (`gate_cols`, `depth_col`) and read / accumulate views of these."""

    def __init__(self, g: torch.Tensor):
        self.g = g.detach().to(F32).contiguous()
        self.dg = torch.zeros_like(self.g)

    def cols(self, rng):
        return self.g[:, rng[0]:rng[0] + rng[1]], self.dg[:, rng[0]:rng[0] + rng[1]]


def width_gate(x, rt: Optional[RuntimeGates], rng, rows_per_sample, period, group_size, need_bwd):
    """y = x * gate (gates.py:15-28,56-62).  Returns (y, bwd) with bwd(dy) -> dx, accumulating d gate."""
    if rt is None or rng is None:
        return x, (lambda dy: dy)
    g, dg = rt.cols(rng)
    y = K.gate_scale(x, g, rows_per_sample, period, group_size)
    if not need_bwd:
        return y, None

    def bwd(dy):
        K.gate_grad(dy, x, dg, rows_per_sample, period, group_size)
        return K.gate_scale(dy, g, rows_per_sample, period, group_size)

    return y, bwd


def depth_gate(inp, out, rt: Optional[RuntimeGates], col, rows_per_sample, need_bwd):
    """(1 - m) * inp + m * out (gates.py:43-49).  bwd(dy) -> (d_inp, d_out), accumulating d m."""
    if rt is None or col is None:
        return out, None
    g, dg = rt.g[:, col].contiguous(), rt.dg[:, col]
    y = K.depth_blend(inp, out, g, rows_per_sample)
    if not need_bwd:
        return y, None

    def bwd(dy):
        tmp = torch.zeros_like(g)
        d_inp, d_out = K.depth_blend_bwd(dy, inp, out, g, tmp, rows_per_sample)
        dg.add_(tmp)                      # (column view of the accumulator is strided: one tiny add)
        return d_inp, d_out

    return y, bwd


# ----------------------------------------------------------------------------------------------------------------
# leaf composites
# ----------------------------------------------------------------------------------------------------------------
class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=1, name="op"): 3x3 stride-2 conv (key: downsamplers.0.conv)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = PConv2d(channels, channels, 3, stride=2)

    def run(self, x, B, H, W, need_bwd):
        y, b = bnn.conv(x, self.conv, B, H, W, need_bwd)
        return y, (None if b is None else (lambda dy: b(dy)[0]))


class Upsample2D(nn.Module):
    """diffusers Upsample2D(use_conv=True): nearest 2x then 3x3 conv (key: upsamplers.0.conv)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = PConv2d(channels, channels, 3)

    def run(self, x, B, H, W, need_bwd):
        up = K.upsample2x_fwd(x, B, H, W)
        y, b = bnn.conv(up, self.conv, B, 2 * H, 2 * W, need_bwd)
        if b is None:
            return y, None
        return y, (lambda dy: K.upsample2x_bwd(b(dy)[0], B, H, W))


class ResnetBlock2DWidthGated(nn.Module):
    """Pruned form of reference blocks.py:298-475 (and :478-702 when ``depth_gated``).

    keep_groups: ascending indices of surviving norm2 groups; C_mid = len(keep) * (C_out / 32)  (blocks.py:439-441).
    dropped   : depth gate < 0.5 -> the block returns ``input[:, :C_out]`` (blocks.py:502-515,651-663); no parameters.
    """

    def __init__(self, in_channels, out_channels, temb_channels, keep_groups: Sequence[int], eps=1e-5, groups=32,
                 depth_gated=False, dropped=False, is_input_concatenated=False, skip_connection_dim=None, gate_cols=None,
                 depth_col=None):
        super().__init__()
        self.gate_cols, self.depth_col, self._rt = gate_cols, depth_col, None   # runtime gating (un-pruned mode), see RuntimeGates
        self.in_channels, self.out_channels = in_channels, out_channels
        self.depth_gated, self.dropped, self.pruned = depth_gated, dropped, True
        self.is_input_concatenated, self.skip_connection_dim = is_input_concatenated, skip_connection_dim
        self.keep_groups = list(keep_groups)
        self.group_dim = out_channels // groups
        self.mid_channels = len(self.keep_groups) * self.group_dim
        if dropped:
            return
        self.norm1 = PGroupNorm(groups, in_channels, eps)
        self.conv1 = PConv2d(in_channels, self.mid_channels, 3)
        self.time_emb_proj = PLinear(temb_channels, self.mid_channels)
        self.norm2 = PGroupNorm(len(self.keep_groups), self.mid_channels, eps)
        self.conv2 = PConv2d(self.mid_channels, out_channels, 3)
        if in_channels != out_channels:
            self.conv_shortcut = PConv2d(in_channels, out_channels, 1)
        else:
            self.conv_shortcut = None

    def keep_channels(self) -> torch.Tensor:
        g = torch.tensor(self.keep_groups, dtype=torch.long)
        return (g[:, None] * self.group_dim + torch.arange(self.group_dim)[None, :]).reshape(-1)

    def prune(self):
        """Reference blocks.py:435-475 / :647-702.  Modules here exist only in pruned form (the model-level `prune()` creates
        them that way), so the reference's per-module call has nothing left to do."""
        return self

    def run(self, x, temb_act, B, H, W, need_bwd):
        """x: [B*H*W, C_in]; temb_act = SiLU(emb) bf16 [B, 1280]. Returns (y, bwd) with bwd(dy) -> (dx, d_temb_act)."""
        if self.dropped:
            cin = x.shape[1]
            keep = cin - self.skip_connection_dim if (self.is_input_concatenated and self.depth_gated) else cin
            y = x[:, :keep]
            if not need_bwd:
                return y, None

            def bwd_drop(dy):
                if keep == cin:
                    return dy, None
                dx = K.alloc2d(x.shape[0], cin, x.device, zero=True)
                K.copy2d(dy, dx[:, :keep])
                return dx, None

            return y, bwd_drop
        hw = H * W
        with bnn.side_branch(temb_act):   # 16-row GEMM on a handful of CTAs: next to GroupNorm instead of in front of conv1
            t, b_t = bnn.linear(temb_act, self.time_emb_proj, need_bwd, out_fp32=True)  # blocks.py:334-337
        h1, b_n1 = bnn.gn(x, self.norm1, B, hw, True, need_bwd)                       # blocks.py:318-319
        bnn.side_join()
        h2, b_c1 = bnn.conv(h1, self.conv1, B, H, W, need_bwd, rowbias=t)               # blocks.py:332,339-341
        rt = self._rt
        h2, b_g = width_gate(h2, rt, self.gate_cols, hw, self.mid_channels, self.group_dim, need_bwd)   # blocks.py:343-346
        h3, b_n2 = bnn.gn(h2, self.norm2, B, hw, True, need_bwd)                       # blocks.py:348,371
        if self.conv_shortcut is not None:
            sc, b_sc = bnn.conv(x, self.conv_shortcut, B, H, W, need_bwd)               # blocks.py:376-377
        else:
            sc, b_sc = x, None
        y, b_c2 = bnn.conv(h3, self.conv2, B, H, W, need_bwd, residual=sc)              # blocks.py:374,379
        b_d = None
        if rt is not None and self.depth_gated and self.depth_col is not None:          # blocks.py:582-587
            cin = x.shape[1]
            keep = cin - self.skip_connection_dim if self.is_input_concatenated else cin
            inp_h = x[:, :keep]
            y, b_d = depth_gate(inp_h, y, rt, self.depth_col, hw, need_bwd)
        if not need_bwd:
            return y, None

        def bwd(dy):
            d_inp = None
            if b_d is not None:
                d_inp, dy = b_d(dy)
            dh3, _ = b_c2(dy)
            dh2 = b_g(b_n2(dh3))
            dh1, dt = b_c1(dh2, want_rowbias=True)
            with bnn.side_branch(dt):     # time-embedding gradient chain: off the dgrad chain, joined at the block's end
                dtemb = b_t(K.cast2d_f32_to_bf16(dt))
            if b_sc is not None:
                dsc, _ = b_sc(dy)
            else:
                dsc = dy
            dx = b_n1(dh1, residual=dsc)     # merges the shortcut-branch gradient in the GroupNorm-backward epilogue
            if d_inp is not None:            # the depth gate's bypass branch: gradient of input[:, :keep]
                K.add(dx[:, :d_inp.shape[1]], d_inp, out=dx[:, :d_inp.shape[1]])
            return dx, dtemb

        return y, bwd


class ResnetBlock2DWidthDepthGated(ResnetBlock2DWidthGated):
    """Reference blocks.py:478-702."""

    def __init__(self, *a, **k):
        k.setdefault("depth_gated", True)
        super().__init__(*a, **k)


class GatedAttention(nn.Module):
    """Pruned form of reference blocks.py:141-196 (+ processor :199-295): `heads` surviving 64-wide heads."""

    def __init__(self, query_dim, heads, dim_head=64, cross_attention_dim=None, keep_heads: Optional[Sequence[int]] = None,
                 orig_heads: Optional[int] = None, gate_cols=None):
        super().__init__()
        self.gate_cols, self._rt = gate_cols, None
        if dim_head != 64:
            raise ValueError("the sm_100a attention path is specialised for head_dim 64 (SD-2.1)")
        self.heads, self.dim_head, self.query_dim = heads, dim_head, query_dim
        self.keep_heads = list(range(heads)) if keep_heads is None else list(keep_heads)
        self.orig_heads = orig_heads or heads
        self.is_cross = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if self.is_cross else query_dim
        inner = heads * dim_head
        self.to_q = PLinear(query_dim, inner, bias=False)
        self.to_k = PLinear(self.cross_attention_dim, inner, bias=False)
        self.to_v = PLinear(self.cross_attention_dim, inner, bias=False)
        self.to_out = nn.ModuleList([PLinear(inner, query_dim, bias=True), nn.Identity()])
        self.pruned = True

    def prune(self):
        """Reference blocks.py:163-196: already in pruned form (see ResnetBlock2DWidthGated.prune)."""
        return self

    def _fused(self, names):
        """One [sum N, K] bf16 weight (and fp32 grad) view over parameters that are adjacent in the arena."""
        mods = [getattr(self, n) for n in names]
        first = mods[0]
        n_rows = sum(m.out_features for m in mods)
        kdim = first.in_features
        w = first.w16.as_strided((n_rows, kdim), (kdim, 1), first.w16.storage_offset())
        g = None
        if hasattr(first, "gw"):
            g = first.gw.as_strided((n_rows, kdim), (kdim, 1), first.gw.storage_offset())
        for a, b in zip(mods[:-1], mods[1:]):
            assert b.w16.storage_offset() == a.w16.storage_offset() + a.out_features * kdim, "q/k/v not adjacent"
        return w, g

    def run(self, x, residual, ctx2d, B, L, Lctx, need_bwd):
        """x: normalised tokens [B*L, C]; returns to_out(attn(x)) + residual and bwd(dy) -> dx (gradient w.r.t. x only;
        the residual branch is merged by the caller's LayerNorm backward)."""
        inner = self.heads * 64
        rt = self._rt
        b_gq = b_gkv = (lambda d: d)
        if not self.is_cross:
            w, g = self._fused(("to_q", "to_k", "to_v"))
            qkv, b_qkv = bnn.linear(x, self.to_q, need_bwd, w16=w, gw=g, bias=None)    # blocks.py:244,251-252
            qkv, b_gq = width_gate(qkv, rt, self.gate_cols, L, inner, 64, need_bwd)    # blocks.py:267-272: head gate on q, k, v
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
            Lk = L
        else:
            q, b_q = bnn.linear(x, self.to_q, need_bwd)
            q, b_gq = width_gate(q, rt, self.gate_cols, L, inner, 64, need_bwd)
            w, g = self._fused(("to_k", "to_v"))
            kv, b_kv = bnn.linear(ctx2d, self.to_k, need_bwd, w16=w, gw=g, bias=None)
            kv, b_gkv = width_gate(kv, rt, self.gate_cols, Lctx, inner, 64, need_bwd)
            k, v = kv[:, :inner], kv[:, inner:]
            Lk = Lctx
        a, b_att = bnn.attention(q, k, v, B, self.heads, L, Lk, need_bwd)             # blocks.py:275-277
        y, b_o = bnn.linear(a, self.to_out[0], need_bwd, residual=residual)            # blocks.py:283 (+ residual add)
        if not need_bwd:
            return y, None

        def bwd(dy):
            da = b_o(dy)
            if not self.is_cross:
                dqkv = K.alloc2d(B * L, 3 * inner, dy.device)
                b_att(da, dqkv[:, :inner], dqkv[:, inner:2 * inner], dqkv[:, 2 * inner:])
                return b_qkv(b_gq(dqkv))
            dq = K.alloc2d(B * L, inner, dy.device)
            dkv = K.alloc2d(B * Lk, 2 * inner, dy.device)
            b_att(da, dq, dkv[:, :inner], dkv[:, inner:])
            b_kv(b_gkv(dkv), need_dx=False)     # text-encoder context is frozen (trainer.py:2433)
            return b_q(b_gq(dq))

        return y, bwd


class GEGLUGated(nn.Module):
    """Pruned form of reference blocks.py:27-76: proj has 2*F' rows (value half then gate half, same mask)."""

    def __init__(self, dim_in, dim_out_kept):
        super().__init__()
        self.proj = PLinear(dim_in, 2 * dim_out_kept)
        self.pruned = True


class FeedForwardWidthGated(nn.Module):
    """Pruned form of reference blocks.py:79-138: net = [GEGLU(proj), Dropout(0), Linear]."""

    def __init__(self, dim, keep_groups: Sequence[int], gate_width=32, mult=4, gate_cols=None):
        super().__init__()
        self.gate_cols, self._rt = gate_cols, None
        self.keep_groups = list(keep_groups)
        self.group_dim = dim * mult // gate_width
        self.inner_full = dim * mult
        inner = len(self.keep_groups) * self.group_dim
        self.inner = inner
        self.net = nn.ModuleList([GEGLUGated(dim, inner), nn.Identity(), PLinear(inner, dim)])

    def prune(self):
        """Reference blocks.py:131-138 (+ GEGLUGated.prune_gate :62-76): already in pruned form."""
        return self

    def keep_units(self) -> torch.Tensor:
        g = torch.tensor(self.keep_groups, dtype=torch.long)
        return (g[:, None] * self.group_dim + torch.arange(self.group_dim)[None, :]).reshape(-1)

    def run(self, x, residual, need_bwd, rows_per_sample=None):
        if self._rt is not None and self.gate_cols is not None:
            # un-pruned mode: the gate sits BETWEEN the projection and the activation (blocks.py:54-58), on both halves
            p, b_p = bnn.linear(x, self.net[0].proj, need_bwd)
            p, b_gt = width_gate(p, self._rt, self.gate_cols, rows_per_sample, self.inner, self.group_dim, need_bwd)
            gl, b_g = bnn.geglu(p, need_bwd)
            y, b_2 = bnn.linear(gl, self.net[2], need_bwd, residual=residual)
            if not need_bwd:
                return y, None
            return y, (lambda dy: b_p(b_gt(b_g(b_2(dy)))))
        gl, b_pg = bnn.linear_geglu(x, self.net[0].proj, need_bwd)                      # blocks.py:49,54-59 in one GEMM
        y, b_2 = bnn.linear(gl, self.net[2], need_bwd, residual=residual)
        if not need_bwd:
            return y, None
        return y, (lambda dy: b_pg(b_2(dy)))


class BasicTransformerBlockWidthGated(nn.Module):
    """Pruned form of reference blocks.py:705-868 (diffusers BasicTransformerBlock data flow)."""

    def __init__(self, dim, cross_attention_dim, keep1, keep2, keep_ff, orig_heads, ff_gate_width=32, gate_cols=(None, None, None)):
        super().__init__()
        self.norm1 = PLayerNorm(dim)
        self.attn1 = GatedAttention(dim, len(keep1), 64, None, keep1, orig_heads, gate_cols=gate_cols[0])
        self.norm2 = PLayerNorm(dim)
        self.attn2 = GatedAttention(dim, len(keep2), 64, cross_attention_dim, keep2, orig_heads, gate_cols=gate_cols[1])
        self.norm3 = PLayerNorm(dim)
        self.ff = FeedForwardWidthGated(dim, keep_ff, ff_gate_width, gate_cols=gate_cols[2])

    def run(self, x0, ctx2d, B, L, Lctx, need_bwd):
        n1, b_n1 = bnn.ln(x0, self.norm1, need_bwd)
        x1, b_a1 = self.attn1.run(n1, x0, None, B, L, 0, need_bwd)
        n2, b_n2 = bnn.ln(x1, self.norm2, need_bwd)
        x2, b_a2 = self.attn2.run(n2, x1, ctx2d, B, L, Lctx, need_bwd)
        n3, b_n3 = bnn.ln(x2, self.norm3, need_bwd)
        x3, b_ff = self.ff.run(n3, x2, need_bwd, L)
        if not need_bwd:
            return x3, None

        def bwd(dx3):
            dx2 = b_n3(b_ff(dx3), residual=dx3)      # residual stream merged in the LayerNorm-backward epilogue
            dx1 = b_n2(b_a2(dx2), residual=dx2)
            dx0 = b_n1(b_a1(dx1), residual=dx1)
            return dx0

        return x3, bwd


class Transformer2DModelWidthGated(nn.Module):
    """Pruned form of reference blocks.py:870-1003 / :1006-1334 (continuous input, use_linear_projection=True)."""

    def __init__(self, num_attention_heads, in_channels, cross_attention_dim, keep1, keep2, keep_ff, norm_num_groups=32,
                 depth_gated=False, dropped=False, ff_gate_width=32, gate_cols=(None, None, None), depth_col=None):
        super().__init__()
        self.depth_col, self._rt = depth_col, None
        self.in_channels = in_channels
        self.depth_gated, self.dropped, self.pruned = depth_gated, dropped, True
        if dropped:
            self.transformer_blocks = nn.ModuleList([nn.Identity()])
            return
        self.norm = PGroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = PLinear(in_channels, in_channels)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlockWidthGated(
            in_channels, cross_attention_dim, keep1, keep2, keep_ff, num_attention_heads, ff_gate_width, gate_cols=gate_cols)])
        self.proj_out = PLinear(in_channels, in_channels)

    def prune_module(self):
        """Reference blocks.py:1324-1334 (depth-dropped transformer -> Identity): already applied at construction."""
        return self

    def run(self, x, ctx2d, B, H, W, Lctx, need_bwd):
        if self.dropped:                                                                # blocks.py:1134-1138
            return x, (None if not need_bwd else (lambda dy: dy))
        L = H * W
        n, b_n = bnn.gn(x, self.norm, B, L, False, need_bwd)                            # GroupNorm eps 1e-6, no act
        t0, b_pi = bnn.linear(n, self.proj_in, need_bwd)
        t1, b_tb = self.transformer_blocks[0].run(t0, ctx2d, B, L, Lctx, need_bwd)
        y, b_po = bnn.linear(t1, self.proj_out, need_bwd, residual=x)                   # + residual (blocks.py:1221-1228)
        b_d = None
        if self._rt is not None and self.depth_gated and self.depth_col is not None:    # blocks.py:1241-1244
            y, b_d = depth_gate(x, y, self._rt, self.depth_col, L, need_bwd)
        if not need_bwd:
            return y, None
        if b_d is None:
            return y, (lambda dy: b_n(b_pi(b_tb(b_po(dy))), residual=dy))

        def bwd(dy):
            d_inp, d_out = b_d(dy)
            dx = b_n(b_pi(b_tb(b_po(d_out))), residual=d_out)
            return K.add(dx, d_inp)

        return y, bwd


class Transformer2DModelWidthDepthGated(Transformer2DModelWidthGated):
    def __init__(self, *a, **k):
        k.setdefault("depth_gated", True)
        super().__init__(*a, **k)


# ----------------------------------------------------------------------------------------------------------------
# U-Net blocks
# ----------------------------------------------------------------------------------------------------------------
class _UNetBlock(nn.Module):
    has_cross_attention = False
    _anchor = None  # set by the model: arena master buffer (requires_grad) or None for frozen models

    def _pairs(self):
        atts = list(getattr(self, "attentions", [])) or [None] * len(self.resnets)
        return list(zip(self.resnets, atts))


class CrossAttnDownBlock2DWidthHalfDepthGated(_UNetBlock):
    """Reference blocks.py:1573-1706; forward = diffusers CrossAttnDownBlock2D.forward."""
    has_cross_attention = True

    def __init__(self, resnets, attentions, out_channels, add_downsample):
        super().__init__()
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, **kw):
        B, _, H, W = hidden_states.shape
        Lctx = encoder_hidden_states.shape[0] // B if encoder_hidden_states is not None else 0

        def runner(need_bwd, x4, temb_act, ctx2d):
            x = as2d(x4)
            outs, bwds = [], []
            for resnet, attn in self._pairs():
                x, b_r = resnet.run(x, temb_act, B, H, W, need_bwd)
                b_a = None
                if attn is not None:
                    x, b_a = attn.run(x, ctx2d, B, H, W, Lctx, need_bwd)
                outs.append(x)
                bwds.append((b_r, b_a))
            b_d = None
            res4 = [to4d(o, B, H, W) for o in outs]
            if self.downsamplers is not None:
                x, b_d = self.downsamplers[0].run(x, B, H, W, need_bwd)
                res4.append(to4d(x, B, H // 2, W // 2))
            if not need_bwd:
                return res4, None

            def bwd(*gouts):
                g = [None if t is None else as2d(t) for t in gouts]
                dtemb = None
                cur = None
                if self.downsamplers is not None:
                    cur = b_d(g[-1]) if g[-1] is not None else None
                    g = g[:-1]
                for i in reversed(range(len(bwds))):
                    cur = _merge(cur, g[i])
                    b_r, b_a = bwds[i]
                    if cur is None:
                        continue
                    if b_a is not None:
                        cur = b_a(cur)
                    cur, dt = b_r(cur)
                    dtemb = _merge_small(dtemb, dt)
                return (None if cur is None else to4d(cur, B, H, W)), dtemb, None

            return res4, bwd

        outs = run_block(runner, self._anchor, hidden_states, temb, encoder_hidden_states, owner=(self, "_grad_ready"))
        return outs[-1], tuple(outs)


class DownBlock2DWidthHalfDepthGated(CrossAttnDownBlock2DWidthHalfDepthGated):
    """Reference blocks.py:2187-2247 (no attentions)."""
    has_cross_attention = False

    def __init__(self, resnets, out_channels, add_downsample):
        nn.Module.__init__(self)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class UNetMidBlock2DCrossAttnWidthGated(_UNetBlock):
    """Reference blocks.py:2450-2544; forward = diffusers UNetMidBlock2DCrossAttn.forward."""
    has_cross_attention = True

    def __init__(self, resnets, attentions):
        super().__init__()
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, **kw):
        B, _, H, W = hidden_states.shape
        Lctx = encoder_hidden_states.shape[0] // B

        def runner(need_bwd, x4, temb_act, ctx2d):
            x = as2d(x4)
            x, b0 = self.resnets[0].run(x, temb_act, B, H, W, need_bwd)
            chain = []
            for attn, resnet in zip(self.attentions, self.resnets[1:]):
                x, b_a = attn.run(x, ctx2d, B, H, W, Lctx, need_bwd)
                x, b_r = resnet.run(x, temb_act, B, H, W, need_bwd)
                chain.append((b_a, b_r))
            out = [to4d(x, B, H, W)]
            if not need_bwd:
                return out, None

            def bwd(gy):
                cur = as2d(gy)
                dtemb = None
                for b_a, b_r in reversed(chain):
                    cur, dt = b_r(cur)
                    dtemb = _merge_small(dtemb, dt)
                    cur = b_a(cur)
                cur, dt = b0(cur)
                dtemb = _merge_small(dtemb, dt)
                return to4d(cur, B, H, W), dtemb, None

            return out, bwd

        return run_block(runner, self._anchor, hidden_states, temb, encoder_hidden_states, owner=(self, "_grad_ready"))[0]


class CrossAttnUpBlock2DWidthHalfDepthGated(_UNetBlock):
    """Reference blocks.py:1900-2039; forward = diffusers CrossAttnUpBlock2D.forward (skip concat per layer)."""
    has_cross_attention = True

    def __init__(self, resnets, attentions, out_channels, add_upsample):
        super().__init__()
        if attentions is not None:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None, **kw):
        B, _, H, W = hidden_states.shape
        Lctx = encoder_hidden_states.shape[0] // B if encoder_hidden_states is not None else 0
        n_layers = len(self.resnets)
        skips = list(res_hidden_states_tuple)[-n_layers:][::-1]   # consumed last-first (diffusers pops from the end)

        def runner(need_bwd, x4, temb_act, ctx2d, *skip4):
            x = as2d(x4)
            steps = []
            for (resnet, attn), s4 in zip(self._pairs(), skip4):
                c_h = x.shape[1]
                if resnet.dropped:
                    # depth-dropped: returns the non-skip part of the concat == x itself (blocks.py:502-515)
                    b_r, cat_used = None, False
                else:
                    xc = bnn.concat_channels(x, as2d(s4))
                    x, b_r = resnet.run(xc, temb_act, B, H, W, need_bwd)
                    cat_used = True
                b_a = None
                if attn is not None:
                    x, b_a = attn.run(x, ctx2d, B, H, W, Lctx, need_bwd)
                steps.append((b_r, b_a, cat_used, c_h))
            b_u = None
            h, w = H, W
            if self.upsamplers is not None:
                x, b_u = self.upsamplers[0].run(x, B, H, W, need_bwd)
                h, w = 2 * H, 2 * W
            out = [to4d(x, B, h, w)]
            if not need_bwd:
                return out, None

            def bwd(gy):
                cur = as2d(gy)
                if b_u is not None:
                    cur = b_u(cur)
                dtemb = None
                dskips = [None] * len(steps)
                for i in reversed(range(len(steps))):
                    b_r, b_a, cat_used, c_h = steps[i]
                    if b_a is not None:
                        cur = b_a(cur)
                    if cat_used:
                        dcat, dt = b_r(cur)
                        dtemb = _merge_small(dtemb, dt)
                        cur = dcat[:, :c_h]
                        dskips[i] = to4d(dcat[:, c_h:], B, H, W)
                return (to4d(cur, B, H, W), dtemb, None) + tuple(dskips)

            return out, bwd

        return run_block(runner, self._anchor, hidden_states, temb, encoder_hidden_states, *skips, owner=(self, "_grad_ready"))[0]


class UpBlock2DWidthHalfDepthGated(CrossAttnUpBlock2DWidthHalfDepthGated):
    """Reference blocks.py:2316-2381 (no attentions)."""
    has_cross_attention = False

    def __init__(self, resnets, out_channels, add_upsample):
        super().__init__(resnets, None, out_channels, add_upsample)


def _merge(a, b):
    """Sum of two bf16 gradient matrices (either may be None) with the add kernel."""
    if a is None:
        return b
    if b is None:
        return a
    return K.add(a, b)


def _merge_small(a, b):
    """Accumulate the (tiny) [B, 1280] time-embedding gradients of the resnets inside one block -- on the side stream, where
    the resnets' backward produces them (ResnetBlock2DWidthGated.run); the block's backward joins it before returning."""
    if a is None or b is None:
        return _merge(a, b)
    with bnn.side_branch(a, b):
        return K.add(a, b)
