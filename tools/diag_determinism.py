"""Run-to-run noise floor of one training step (same weights, same batch, two fresh tuners): losses, prediction, gradient
(first-step AdamW moment).  Diagnostics only (GPU box).  Env toggles: B200PDM_SERIAL_WGRAD, B200PDM_SERIAL_TEACHER."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import pdm_restated as P
from oracle.make_golden import SMALL64, deterministic_fill, make_arch_vector
from unlearn_ft_b200.pdm.models import UNet2DConditionModel, UNet2DConditionModelPruned
from unlearn_ft_b200.pdm.training import UnetFineTuner

cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"],
           cross_attention_dim=SMALL64["cross_attention_dim"])
full = P.UNetGated(**SMALL64)
deterministic_fill(full, 3)
av = make_arch_vector(full.get_structure(), 0.55, 21, (2,))
teacher = UNet2DConditionModel(cfg, seed=7)


def student():
    m = UNet2DConditionModelPruned(cfg, arch_vector=av, seed=None)
    m.load_unpruned_state_dict(full.state_dict())
    return m


B = int(os.environ.get("B", 2))
g = torch.Generator().manual_seed(0)
batch = dict(latents=torch.randn(B, 4, 16, 16, generator=g).cuda(), noise=torch.randn(B, 4, 16, 16, generator=g).cuda(),
             timesteps=torch.randint(0, 1000, (B,), generator=g).cuda(),
             prompt_embeds=torch.randn(B, 77, SMALL64["cross_attention_dim"], generator=g).cuda())
runs = []
for i in range(3):
    t = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0)
    out = t.step(batch)
    feats = {k: v.detach().float().clone() for k, v in t.block_act_student.items()}
    losses = [float(v.detach()) for v in out]
    t._backward_and_update(out[0], t.optimizer)
    torch.cuda.synchronize()
    runs.append((losses, feats, t.optimizer.exp_avg.double().clone()))
for i in (1, 2):
    l0, f0, m0 = runs[0]
    l1, f1, m1 = runs[i]
    print(f"run {i} vs 0: losses rel", ["%.1e" % (abs(a - b) / abs(b)) for a, b in zip(l1, l0)],
          "feat max-rel", "%.1e" % max(((f1[k] - f0[k]).abs().max() / f0[k].abs().max()).item() for k in f0),
          "grad rel-L2 %.2e" % ((m1 - m0).norm() / m0.norm()).item(), flush=True)
# per-block breakdown of the gradient difference
t = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0)
m0, m1 = runs[0][2], runs[1][2]
for (mod, attr), (lo, hi) in t.reducer.buckets.items():
    name = [n for n, mm in t.student.named_modules() if mm is mod][0] or attr
    d = (m1[lo:hi] - m0[lo:hi]).norm() / m0[lo:hi].norm().clamp_min(1e-30)
    print(f"  {name:16s} {attr:18s} rel-L2 {d.item():.2e}")
