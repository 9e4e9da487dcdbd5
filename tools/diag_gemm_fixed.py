"""Diagnostics: fixed cost of a GEMM launch (prologue / CTA lifetime / event-timed duration) on tiny and small shapes."""
import os
import sys

os.environ["B200PDM_GEMM_DBG"] = "1"
os.environ["B200PDM_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K

for (M, N, Kd) in [(16, 64, 64), (256, 64, 64), (256, 256, 64), (16, 1280, 1280), (4096, 1280, 1280), (65536, 320, 320)]:
    x = K.alloc2d(M, Kd).normal_()
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    out = K.alloc2d(M, N)
    for _ in range(3):
        K.linear_fwd(x, w, out=out)
    torch.cuda.synchronize()
