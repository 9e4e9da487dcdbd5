"""Where does the CUDA path's numerical error come from?  Full-size (SD-2.1) teacher and r=0.55 student, batch 2:
per-model error of the prediction and of the nine hook features against the fp32 oracle (rms-relative, scale, bias), for
this library and for torch's own bf16-autocast evaluation of the oracle.  Diagnostics only (GPU box)."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import diffusers_restated as D
from oracle import pdm_restated as P
from oracle.make_golden import deterministic_fill, make_arch_vector
from unlearn_ft_b200.pdm.models import UNet2DConditionModel, UNet2DConditionModelPruned
from unlearn_ft_b200.pdm.training import cast_block_act_hooks


def stats(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    e = a - b
    return dict(rms_rel=(e.norm() / b.norm()).item(), scale=((a @ b) / (b @ b)).item(), bias=(e.mean() / b.std()).item(),
                maxrel=(e.abs().max() / b.abs().max()).item())


def run(model, hooks_fn, x, t, ctx, autocast=False):
    feats = {}
    hs = hooks_fn(model, feats)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = model(x, t, ctx).sample.float()
    for h in hs:
        h.remove()
    return y, {k: v.detach().float() for k, v in feats.items()}


with torch.device("meta"):
    t_o = D.UNet2DConditionModel(**D.SD21_UNET_CONFIG)
    full = P.UNetGated()
t_o, full = t_o.to_empty(device="cpu"), full.to_empty(device="cpu")
deterministic_fill(t_o, 5)
deterministic_fill(full, 3)
av = make_arch_vector(full.get_structure(), 0.55, 21, ())
student = UNet2DConditionModelPruned(arch_vector=av, seed=None, trainable=False)
student.load_unpruned_state_dict(full.state_dict())
full.set_structure(P.transform_arch_vector(av, full.get_structure()))
full.prune()
s_o = full.eval().cuda()
teacher = UNet2DConditionModel(seed=None)
teacher.load_state_dict(t_o.state_dict())
t_o = t_o.eval().cuda()
g = torch.Generator().manual_seed(11)
B = 2
lat, noise = torch.randn(B, 4, 64, 64, generator=g).cuda(), torch.randn(B, 4, 64, 64, generator=g).cuda()
ts = torch.tensor([37, 861]).cuda()
ctx = torch.randn(B, 77, 1024, generator=g).cuda()
x = D.DDIMSchedulerLite().add_noise(lat, noise, ts)

res = {}
for name, mine, orc in (("teacher", teacher, t_o), ("student", student, s_o)):
    y32, f32 = run(orc, P.cast_block_act_hooks, x, ts, ctx)
    ybf, fbf = run(orc, P.cast_block_act_hooks, x, ts, ctx, autocast=True)
    obf = copy.deepcopy(orc).to(torch.bfloat16)
    ypb, fpb = run(obf, P.cast_block_act_hooks, x.bfloat16(), ts, ctx.bfloat16())
    del obf
    ym, fm = run(mine, cast_block_act_hooks, x, ts, ctx)
    res[name] = (y32, ym, ybf, ypb)
    print(f"== {name}: pred rms {y32.pow(2).mean().sqrt().item():.4f}")
    for tag, yy, ff in (("b200", ym, fm), ("torch autocast", ybf, fbf), ("torch pure-bf16", ypb, fpb)):
        print(f"  {tag:16s} pred {stats(yy, y32)}")
        print("   feats rms_rel: " + " ".join(f"{k}={stats(ff[k], f32[k])['rms_rel']:.4f}" for k in f32))
for tag, i in (("b200", 1), ("torch autocast", 2), ("torch pure-bf16 teacher / autocast student", None)):
    t32, s32 = res["teacher"][0], res["student"][0]
    if i is None:
        tt, ss = res["teacher"][3], res["student"][2]
    else:
        tt, ss = res["teacher"][i], res["student"][i]
    kd32 = (s32 - t32).pow(2).mean().item()
    kd = (ss - tt).pow(2).mean().item()
    es, et = ss - s32, tt - t32
    print(f"{tag}: kd {kd:.6f} vs fp32 {kd32:.6f} rel {(kd - kd32) / kd32:+.2e}; |es|^2 {es.pow(2).mean().item():.3e} "
          f"|et|^2 {et.pow(2).mean().item():.3e} |es-et|^2 {(es - et).pow(2).mean().item():.3e} "
          f"cross {2 * ((s32 - t32) * (es - et)).mean().item():+.3e}")
