"""(1) Is the network-level gain error of the CUDA path systematic?  Teacher (full SD-2.1 U-Net) on several input draws:
scale-1 of the prediction against the fp32 oracle, for this library, torch autocast and torch pure-bf16.
(2) Per-module probe: every resnet / transformer / down- / up-sampler of OUR teacher is fed the fp32 oracle's own input of that
module and compared with the oracle's output of that module (scale, bias, rms error) -- localises a biased stage.
Diagnostics only (GPU box)."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import diffusers_restated as D
from oracle.make_golden import deterministic_fill
from unlearn_ft_b200 import kernels as K
from unlearn_ft_b200.pdm.models import UNet2DConditionModel
from unlearn_ft_b200.pdm.nn import as2d


def stats(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    e = a - b
    return ((a @ b) / (b @ b) - 1).item(), (e.mean() / b.std()).item(), (e.norm() / b.norm()).item()


with torch.device("meta"):
    t_o = D.UNet2DConditionModel(**D.SD21_UNET_CONFIG)
t_o = t_o.to_empty(device="cpu")
deterministic_fill(t_o, 5)
teacher = UNet2DConditionModel(seed=None)
teacher.load_state_dict(t_o.state_dict())
t_o = t_o.eval().requires_grad_(False).cuda()
t_bf = copy.deepcopy(t_o).to(torch.bfloat16)
B = 2
sched = D.DDIMSchedulerLite()
print("== (1) prediction gain over input draws: scale-1 | bias/std | rms_rel")
for seed in range(6):
    g = torch.Generator().manual_seed(100 + seed)
    lat, noise = torch.randn(B, 4, 64, 64, generator=g).cuda(), torch.randn(B, 4, 64, 64, generator=g).cuda()
    ts = torch.randint(0, 1000, (B,), generator=g).cuda()
    ctx = torch.randn(B, 77, 1024, generator=g).cuda()
    x = sched.add_noise(lat, noise, ts)
    with torch.no_grad():
        y32 = t_o(x, ts, ctx).sample
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yac = t_o(x, ts, ctx).sample.float()
        ybf = t_bf(x.bfloat16(), ts, ctx.bfloat16()).sample.float()
        ym = teacher(x, ts, ctx).sample
    print(f"seed {seed} t={ts.tolist()}: b200 %+.2e %+.2e %.2e | autocast %+.2e %+.2e %.2e | pure-bf16 %+.2e %+.2e %.2e"
          % (stats(ym, y32) + stats(yac, y32) + stats(ybf, y32)), flush=True)
del t_bf

print("== (2) per-module probe (our module on the oracle's fp32 input): scale-1 | bias/std | rms_rel")
rec = {}


def hook(name):
    def f(mod, args, kwargs, out):
        rec[name] = (args, kwargs, out)
    return f


names = []
for name, mod in t_o.named_modules():
    if isinstance(mod, (D.ResnetBlock2D, D.Transformer2DModel, D.Downsample2D, D.Upsample2D)):
        mod.register_forward_hook(hook(name), with_kwargs=True)
        names.append(name)
g = torch.Generator().manual_seed(11)
lat, noise = torch.randn(B, 4, 64, 64, generator=g).cuda(), torch.randn(B, 4, 64, 64, generator=g).cuda()
ts = torch.tensor([37, 861]).cuda()
ctx = torch.randn(B, 77, 1024, generator=g).cuda()
x = sched.add_noise(lat, noise, ts)
with torch.no_grad():
    t_o(x, ts, ctx)
ctx2d = teacher._context(ctx)
teacher.arena.ensure_shadow()
mods = dict(teacher.named_modules())


def to2d(x4):
    return as2d(x4.to(memory_format=torch.channels_last).bfloat16())


def back(y2d, Bn, H, W):
    return y2d.float().reshape(Bn, H, W, -1).permute(0, 3, 1, 2)


with torch.no_grad():
    for name in names:
        args, kwargs, out = rec[name]
        o = mods[name]
        xin = args[0]
        Bn, C, H, W = xin.shape
        if isinstance(out, tuple):
            out = out[0]
        kind = type(o).__name__
        if "Resnet" in kind:
            temb_act = F.silu(args[1]).bfloat16().contiguous()
            y, _ = o.run(to2d(xin), temb_act, Bn, H, W, False)
        elif "Transformer" in kind:
            y, _ = o.run(to2d(xin), ctx2d, Bn, H, W, 77, False)
        elif "Down" in kind:
            y, _ = o.run(to2d(xin), Bn, H, W, False)
            H, W = H // 2, W // 2
        else:
            y, _ = o.run(to2d(xin), Bn, H, W, False)
            H, W = 2 * H, 2 * W
        s = stats(back(y, Bn, H, W), out)
        print(f"{name:38s} {kind[:14]:14s} %+.2e %+.2e %.2e" % s, flush=True)

print("== (3) real forward: gain of the nine hook features and the prediction (ours | torch autocast), scale-1 / rms_rel")
from oracle import pdm_restated as P
from unlearn_ft_b200.pdm.training import cast_block_act_hooks
for h in list(t_o.modules()):
    h._forward_hooks.clear()


def feats_of(model, hooks_fn, autocast=False):
    st = {}
    hs = hooks_fn(model, st)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = model(x, ts, ctx).sample.float()
    for h in hs:
        h.remove()
    st = {k: v.detach().float() for k, v in st.items()}
    st["pred"] = y
    return st


f32 = feats_of(t_o, P.cast_block_act_hooks)
fac = feats_of(t_o, P.cast_block_act_hooks, True)
fm = feats_of(teacher, cast_block_act_hooks)
for k in f32:
    a, b = stats(fm[k], f32[k]), stats(fac[k], f32[k])
    print(f"{k:5s} b200 %+.2e / %.2e   autocast %+.2e / %.2e" % (a[0], a[2], b[0], b[2]))

print("== (4) head in isolation: conv_norm_out -> SiLU -> conv_out on the fp32 oracle's u3 and on our own u3")
with torch.no_grad():
    def head_ref(u):
        h = F.silu(F.group_norm(u, 32, t_o.conv_norm_out.weight, t_o.conv_norm_out.bias, 1e-5))
        return h, F.conv2d(h, t_o.conv_out.weight, t_o.conv_out.bias, padding=1)

    for tag, u in (("oracle u3", f32["u3"]), ("our u3", fm["u3"]), ("autocast u3", fac["u3"])):
        h_ref, y_ref = head_ref(u.bfloat16().float())
        x2 = to2d(u)
        h1, _ = __import__("unlearn_ft_b200.pdm.nn", fromlist=["gn"]).gn(x2, teacher.conv_norm_out, B, 64 * 64, True, False)
        print(f"  [{tag}] GN+SiLU    %+.2e %+.2e %.2e" % stats(back(h1, B, 64, 64), h_ref))
        y1 = K.conv_fwd(h1, teacher.conv_out.w16, B, 64, 64, 4, 3, 1, bias=teacher.conv_out.bias)
        h1f = back(h1, B, 64, 64)
        print(f"  [{tag}] conv_out on our GN output vs torch fp32 conv of the same tensor %+.2e %+.2e %.2e"
              % stats(back(y1, B, 64, 64), F.conv2d(h1f, t_o.conv_out.weight.bfloat16().float(), t_o.conv_out.bias, padding=1)))
        print(f"  [{tag}] ... vs fp32-weight conv %+.2e %+.2e %.2e" % stats(back(y1, B, 64, 64), F.conv2d(h1f, t_o.conv_out.weight, t_o.conv_out.bias, padding=1)))
        print(f"  [{tag}] head total vs fp32 head of the same u3 %+.2e %+.2e %.2e" % stats(back(y1, B, 64, 64), y_ref))
        print(f"  [{tag}] fp32 head of this u3 vs fp32 prediction %+.2e %+.2e %.2e" % stats(head_ref(u)[1], f32["pred"]))
    print("  conv_out weight: mean %.3e rms %.3e ; bf16 rounding gain of w: %+.2e" % (
        t_o.conv_out.weight.mean().item(), t_o.conv_out.weight.pow(2).mean().sqrt().item(),
        stats(t_o.conv_out.weight.bfloat16().float(), t_o.conv_out.weight)[0]))
