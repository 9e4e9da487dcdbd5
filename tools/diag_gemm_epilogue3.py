"""(diag build: make -C unlearn_ft_b200/csrc diag; run with B200PDM_LIB=libb200pdm_diag.so)  Decomposition of the small-K (epilogue-bound) GEMM shapes with the kernel's diagnostic modes:
1 = quarter MMA, 2|4 = no operand loads, 8 = epilogue math but no stores, 128 = epilogue skips its chunks entirely."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K


def time_it(fn, reps=10, iters=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        for _ in range(reps): fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) / reps


for (M, N, Kd, res) in [(65536, 2560, 320, False), (65536, 320, 320, False), (65536, 320, 320, True), (16384, 640, 640, True),
                        (4096, 1280, 1280, True), (65536, 1360, 320, False)]:
    x = K.alloc2d(M, Kd).normal_(); w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02; out = K.alloc2d(M, N)
    r = K.alloc2d(M, N).normal_() if res else None
    line = []
    for mode in (0, 8, 128, 7, 7 | 8, 7 | 128):
        os.environ["B200PDM_GEMM_DBGMODE"] = str(mode)
        line.append(f"m{mode}={time_it(lambda: K.linear_fwd(x, w, out=out, residual=r))*1e3:.1f}")
    os.environ["B200PDM_GEMM_DBGMODE"] = "0"
    print(f"M={M} N={N} K={Kd} res={int(res)}: " + " ".join(line) + " us", flush=True)
