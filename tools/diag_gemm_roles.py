"""Diagnostics: per-role wait/issue cycle counters of CTA 0 (B200PDM_GEMM_DBG=1) under each B200PDM_GEMM_DBGMODE."""
import os
import sys

os.environ["B200PDM_GEMM_DBG"] = "1"
os.environ["B200PDM_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K


def run(name, fn):
    for m in (0, 1, 5, 3, 6, 7):
        os.environ["B200PDM_GEMM_DBGMODE"] = str(m)
        sys.stderr.write(f"--- {name} mode {m}\n")
        sys.stderr.flush()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()


M, N, Kd = 8192, 8192, 8192
x = K.alloc2d(M, Kd).normal_()
w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
out = K.alloc2d(M, N)
run("linear 8192^3", lambda: K.linear_fwd(x, w, out=out))
B, H, W, Ci, Co = 16, 64, 64, 640, 640
xc = K.alloc2d(B * H * W, Ci).normal_()
wc = torch.randn(Co, 9, Ci, device="cuda", dtype=torch.bfloat16) * 0.02
oc = K.alloc2d(B * H * W, Co)
run("conv 640->640@64", lambda: K.conv_fwd(xc, wc, B, H, W, Co, 3, 1, out=oc))
