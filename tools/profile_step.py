"""One warm-up step + one profiled training step of the bench workload (for ncu / launch lists).
Usage: python tools/profile_step.py [--batch 16] [--ratio 0.55] [--fwd-only]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModel, UNet2DConditionModelPruned
from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
from unlearn_ft_b200.pdm.training import UnetFineTuner

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--ratio", type=float, default=0.55)
ap.add_argument("--warm", type=int, default=1)
args = ap.parse_args()
torch.manual_seed(43)
av = HyperStructure.get_random_arch_vector(args.ratio, structure_from_config(SD21_CONFIG))
student = UNet2DConditionModelPruned(arch_vector=av, seed=43)
teacher = UNet2DConditionModel(seed=44)
tuner = UnetFineTuner(student, teacher)
B = args.batch
g = torch.Generator().manual_seed(0)
batch = dict(latents=torch.randn(B, 4, 64, 64, generator=g).cuda(), noise=torch.randn(B, 4, 64, 64, generator=g).cuda(),
             timesteps=torch.randint(0, 1000, (B,), generator=g).cuda(),
             prompt_embeds=torch.randn(B, 77, 1024, generator=g).bfloat16().cuda())
for _ in range(args.warm):
    tuner.train_step(batch)
torch.cuda.synchronize()
if os.environ.get("B200PDM_GEMM_TRACE"):
    from unlearn_ft_b200 import _lib
    _lib.lib().b200pdm_gemm_trace_dump(b"/dev/null")   # drop warm-up rows
torch.cuda.cudart().cudaProfilerStart()
out = tuner.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(out[0]))
if os.environ.get("B200PDM_GEMM_TRACE"):
    from unlearn_ft_b200 import _lib
    _lib.lib().b200pdm_gemm_trace_dump(b"gpurun_out/gemm_trace.tsv")
