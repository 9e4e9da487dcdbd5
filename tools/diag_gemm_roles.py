"""Diagnostics: per-role wait/issue cycle counters of CTA 0 (B200PDM_GEMM_DBG=1) for a few shapes."""
import os
import sys

os.environ["B200PDM_GEMM_DBG"] = "1"
os.environ["B200PDM_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K


def lin(M, N, Kd):
    x = K.alloc2d(M, Kd).normal_()
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    out = K.alloc2d(M, N)
    for _ in range(3):
        K.linear_fwd(x, w, out=out)
    torch.cuda.synchronize()


for shape in [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (65536, 2560, 320), (65536, 1360, 320), (16384, 5120, 640),
              (8192, 8192, 8192)]:
    lin(*shape)
