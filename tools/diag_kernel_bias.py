"""Is any kernel BIASED?  Each forward kernel on realistic random inputs against torch fp32 on the same bf16-rounded inputs:
scale = <ours, ref> / <ref, ref> (a correctly rounded kernel gives |scale - 1| ~ 1e-5 over millions of outputs), mean error in
units of the reference's std, rms-relative error.  Diagnostics only (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from unlearn_ft_b200 import kernels as K


def report(name, a, b):
    a, b = a.double().flatten(), b.double().flatten()
    e = a - b
    print(f"{name:34s} scale-1 {((a @ b) / (b @ b) - 1).item():+.2e}  bias/std {(e.mean() / b.std()).item():+.2e}  "
          f"rms_rel {(e.norm() / b.norm()).item():.2e}  ideal-bf16-rounding rms_rel {((b.float().bfloat16().double() - b).norm() / b.norm()).item():.2e}",
          flush=True)


def nhwc(x):
    B, C, H, W = x.shape
    t = K.alloc2d(B * H * W, C)
    t.copy_(x.permute(0, 2, 3, 1).reshape(B * H * W, C))
    return t


def from2d(t, B, H, W):
    return t.float().reshape(B, H, W, -1).permute(0, 3, 1, 2)


def pack_w(w):
    O, I, kh, kw = w.shape
    buf = torch.zeros(O, kh * kw, K.round8(I), device="cuda", dtype=torch.bfloat16)
    buf[:, :, :I] = w.permute(0, 2, 3, 1).reshape(O, kh * kw, I)
    return buf[:, :, :I]


g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
B, H, W = 2, 64, 64

# conv 3x3 (+bias, +rowbias, +residual)
for Ci, Co in ((320, 320), (960, 170), (4, 320), (320, 4)):
    x = rn(B, Ci, H, W).bfloat16().float()
    w = (rn(Co, Ci, 3, 3) / (9 * Ci) ** 0.5).bfloat16().float()
    bias = rn(Co) * 0.05
    y = K.conv_fwd(nhwc(x), pack_w(w), B, H, W, Co, 3, 1, bias=bias)
    report(f"conv3x3 {Ci}->{Co}", from2d(y, B, H, W), F.conv2d(x, w, bias, padding=1))
x = rn(B, 320, H, W).bfloat16().float()
w = (rn(320, 320, 3, 3) / (9 * 320) ** 0.5).bfloat16().float()
rb = rn(B, 320) * 0.3
res = rn(B, 320, H, W).bfloat16().float()
y = K.conv_fwd(nhwc(x), pack_w(w), B, H, W, 320, 3, 1, bias=None, rowbias=rb, residual=nhwc(res))
report("conv3x3 +rowbias +residual", from2d(y, B, H, W), F.conv2d(x, w, None, padding=1) + rb[:, :, None, None] + res)
y = K.conv_fwd(nhwc(x), pack_w(w), B, H, W, 320, 3, 2)
report("conv3x3 stride 2", from2d(y, B, H // 2, W // 2), F.conv2d(x, w, None, padding=1, stride=2))

# linear
for M, N, Kd in ((8192, 320, 320), (8192, 2560, 320), (8192, 320, 1280), (154, 640, 1024)):
    x = rn(M, Kd).bfloat16()
    w = (rn(N, Kd) / Kd ** 0.5).bfloat16()
    b = rn(N) * 0.05
    r = rn(M, N).bfloat16()
    xx = K.alloc2d(M, Kd); xx.copy_(x)
    rr = K.alloc2d(M, N); rr.copy_(r)
    y = K.linear_fwd(xx, w, bias=b, residual=rr)
    report(f"linear {M}x{N}x{Kd} +bias +res", y, x.float() @ w.float().t() + b + r.float())

# groupnorm (+silu)
for C, G, eps, silu in ((320, 32, 1e-5, True), (170, 17, 1e-5, True), (1280, 32, 1e-5, True), (320, 32, 1e-6, False), (2560, 32, 1e-5, True)):
    hw = 64 * 64 if C < 1000 else 16 * 16
    x = (rn(B, C, hw) * 1.3 + 0.2).bfloat16().float()
    gm, bt = 1 + 0.1 * rn(C), 0.05 * rn(C)
    x2 = K.alloc2d(B * hw, C); x2.copy_(x.permute(0, 2, 1).reshape(B * hw, C))
    y, _ = K.groupnorm_fwd(x2, gm, bt, B, hw, G, eps, silu)
    ref = F.group_norm(x, G, gm, bt, eps)
    if silu:
        ref = F.silu(ref)
    report(f"groupnorm C={C} G={G} silu={int(silu)}", y.float().reshape(B, hw, C).permute(0, 2, 1), ref)

# layernorm
for C in (320, 640, 1280):
    x = (rn(8192, C) * 1.5 + 0.1).bfloat16()
    gm, bt = 1 + 0.1 * rn(C), 0.05 * rn(C)
    xx = K.alloc2d(8192, C); xx.copy_(x)
    y, _, _ = K.layernorm_fwd(xx, gm, bt)
    report(f"layernorm C={C}", y, F.layer_norm(x.float(), (C,), gm, bt))

# geglu
p = rn(8192, 2560).bfloat16()
pp = K.alloc2d(8192, 2560); pp.copy_(p)
y = K.geglu_fwd(pp)
h, gt = p.float().chunk(2, -1)
report("geglu", y, h * F.gelu(gt))

# silu (time embedding)
e = rn(16, 1280)
report("silu_f32_to_bf16", K.silu_f32_to_bf16(e), F.silu(e))

# attention
for heads, Lq, Lk in ((5, 4096, 4096), (10, 1024, 1024), (20, 256, 256), (20, 64, 64), (5, 4096, 77), (20, 64, 77)):
    q, k, v = (rn(B * Lq, heads * 64).bfloat16(), rn(B * Lk, heads * 64).bfloat16(), rn(B * Lk, heads * 64).bfloat16())
    qq = K.alloc2d(B * Lq, heads * 64); qq.copy_(q)
    kk = K.alloc2d(B * Lk, heads * 64); kk.copy_(k)
    vv = K.alloc2d(B * Lk, heads * 64); vv.copy_(v)
    o, _ = K.attention_fwd(qq, kk, vv, B, heads, Lq, Lk, 0.125)
    qf = q.float().view(B, Lq, heads, 64).transpose(1, 2)
    kf = k.float().view(B, Lk, heads, 64).transpose(1, 2)
    vf = v.float().view(B, Lk, heads, 64).transpose(1, 2)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, -1) @ vf
    report(f"attention h={heads} Lq={Lq} Lk={Lk}", o.view(B, Lq, heads, 64).transpose(1, 2), ref)
    # peaked softmax (trained-network-like logits)
    q2 = (q.float() * 3).bfloat16()
    qq.copy_(q2)
    o, _ = K.attention_fwd(qq, kk, vv, B, heads, Lq, Lk, 0.125)
    ref = torch.softmax((q2.float().view(B, Lq, heads, 64).transpose(1, 2)) @ kf.transpose(-1, -2) * 0.125, -1) @ vf
    report(f"   same, 3x sharper logits", o.view(B, Lq, heads, 64).transpose(1, 2), ref)

# upsample + layout kernels
x = rn(B, 640, 32, 32).bfloat16().float()
y = K.upsample2x_fwd(nhwc(x), B, 32, 32)
report("upsample2x", from2d(y, B, 64, 64), F.interpolate(x, scale_factor=2.0, mode="nearest"))
x = rn(B, 4, 64, 64)
report("nchw_f32_to_nhwc_bf16", from2d(K.nchw_f32_to_nhwc_bf16(x), B, 64, 64), x)
t = torch.tensor([37, 861], device="cuda")
emb = K.timestep_embedding(t, 320).float()
import math
half = 160
freq = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
arg = t[:, None].float() * freq[None]
report("timestep_embedding", emb, torch.cat([torch.cos(arg), torch.sin(arg)], -1))
