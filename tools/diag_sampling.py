import os, sys
sys.path.insert(0, "/root/repo")
import torch
from oracle import diffusers_restated as D
from oracle import pdm_restated as P
from unlearn_ft_b200.pdm.pipelines import CFGSampler
from tests.test_unet_gpu import build_pair
def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()
gold = torch.load("tests/golden/reference_golden.pt", weights_only=False)
av = gold["small64_r055"]["arch_vector"]
mine, orc = build_pair(av, trainable=False)
steps, n, g = 6, 2, 7.5
gen = torch.Generator().manual_seed(11)
lat0 = torch.randn(n, 4, 16, 16, generator=gen).cuda()
pos = torch.randn(n, 77, 64, generator=gen).cuda()
neg = torch.randn(1, 77, 64, generator=gen).cuda().expand(n, -1, -1).contiguous()
ref = P.cfg_sample_loop(orc, D.DDIMSchedulerLite(), lat0.clone(), pos, neg, steps, g)
with torch.autocast("cuda", dtype=torch.bfloat16):
    ref_bf = P.cfg_sample_loop(orc, D.DDIMSchedulerLite(), lat0.clone(), pos, neg, steps, g)
print("oracle bf16-autocast vs fp32:", rel(ref_bf, ref))
e = CFGSampler(mine, steps, g, use_cuda_graph=False)
oe = [e.sample(lat0, pos, neg).clone() for _ in range(3)]
print("eager vs ref", [rel(o, ref) for o in oe], "eager run-to-run", rel(oe[1], oe[0]), rel(oe[2], oe[0]))
s = CFGSampler(mine, steps, g, use_cuda_graph=True)
og = [s.sample(lat0, pos, neg).clone() for _ in range(4)]
print("graph vs ref", [rel(o, ref) for o in og])
print("graph vs eager", [rel(o, oe[0]) for o in og])
print("graph run-to-run", [rel(o, og[0]) for o in og])
