"""Which kernel is not bit-reproducible?  Every module / kernel of the forward is run twice on identical inputs and the two
results are compared bitwise (full-size teacher shapes, batch 2).  Diagnostics only (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K
from unlearn_ft_b200.pdm import nn as bnn
from unlearn_ft_b200.pdm.models import UNet2DConditionModel

torch.manual_seed(0)
m = UNet2DConditionModel(seed=1)
m.arena.ensure_shadow()
B = 2


def diff(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item(), int((a != b).sum())


def x2d(rows, C, mean=0.0):
    t = K.alloc2d(rows, C)
    t.copy_(torch.randn(rows, C, device="cuda") + mean)
    return t


temb = torch.randn(B, 1280, device="cuda").bfloat16()
ctx = torch.randn(B * 77, 1024, device="cuda").bfloat16()
for name, mod in m.named_modules():
    kind = type(mod).__name__
    if kind.startswith("ResnetBlock2D"):
        H = {320: 64, 640: 32, 1280: 16}.get(mod.out_channels, 16)
        x = x2d(B * H * H, mod.in_channels, 0.5)
        y1, _ = mod.run(x, temb, B, H, H, False)
        y2, _ = mod.run(x, temb, B, H, H, False)
    elif kind.startswith("Transformer2DModel"):
        H = {320: 64, 640: 32, 1280: 16}.get(mod.in_channels, 16)
        x = x2d(B * H * H, mod.in_channels, 0.5)
        y1, _ = mod.run(x, ctx, B, H, H, 77, False)
        y2, _ = mod.run(x, ctx, B, H, H, 77, False)
    else:
        continue
    d, n = diff(y1, y2)
    if n:
        print(f"{name:40s} {kind[:20]:20s} max-rel {d:.2e}  differing elements {n}", flush=True)
print("-- kernels on one input, twice")
H = 64
x = x2d(B * H * H, 320, 3.0)
gn = m.down_blocks[0].resnets[0].norm1
a, _ = bnn.gn(x, gn, B, H * H, True, False)
b, _ = bnn.gn(x, gn, B, H * H, True, False)
print("groupnorm+silu (mean 3, std 1)", diff(a, b))
x = x2d(B * H * H, 320, 30.0)
a, _ = bnn.gn(x, gn, B, H * H, True, False)
b, _ = bnn.gn(x, gn, B, H * H, True, False)
ref = torch.nn.functional.silu(torch.nn.functional.group_norm(x.float().view(B, H * H, 320).permute(0, 2, 1), 32, gn.weight, gn.bias, 1e-5))
print("groupnorm+silu (mean 30, std 1)", diff(a, b), "| vs torch", diff(a.float().view(B, H * H, 320).permute(0, 2, 1), ref))
conv = m.down_blocks[0].resnets[0].conv1
x = x2d(B * H * H, 320)
print("conv3x3", diff(bnn.conv(x, conv, B, H, H, False)[0], bnn.conv(x, conv, B, H, H, False)[0]))
q = x2d(B * 4096, 960)
o1, _ = K.attention_fwd(q[:, :320], q[:, 320:640], q[:, 640:], B, 5, 4096, 4096, 0.125)
o2, _ = K.attention_fwd(q[:, :320], q[:, 320:640], q[:, 640:], B, 5, 4096, 4096, 0.125)
print("attention", diff(o1, o2))
ln = m.down_blocks[0].attentions[0].transformer_blocks[0].norm1
print("layernorm", diff(bnn.ln(x, ln, False)[0], bnn.ln(x, ln, False)[0]))
lin = m.down_blocks[0].attentions[0].proj_in
print("linear", diff(bnn.linear(x, lin, False)[0], bnn.linear(x, lin, False)[0]))
xs = x2d(B * 64, 1280)
c2 = m.mid_block.resnets[0].conv1
print("conv3x3 8x8 (split-K)", diff(bnn.conv(xs, c2, B, 8, 8, False)[0], bnn.conv(xs, c2, B, 8, 8, False)[0]))
