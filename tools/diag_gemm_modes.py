"""(diag build: make -C unlearn_ft_b200/csrc diag; run with B200PDM_LIB=libb200pdm_diag.so)  Diagnostics: time GEMM/conv shapes with operand streams / MMA switched off (B200PDM_GEMM_DBGMODE) to see which
pipeline bounds the kernel.  Results of modes != 0 are garbage by construction; only their durations matter."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K

MODES = [(0, "full"), (1, "quarter MMA"), (5, "qMMA, A only"), (3, "qMMA, B only"), (6, "no loads"), (7, "no loads, qMMA")]


def time_it(fn, iters=10):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record()
        fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


def sweep(name, fn, flops):
    out = []
    for m, label in MODES:
        os.environ["B200PDM_GEMM_DBGMODE"] = str(m)
        ms = time_it(fn)
        out.append(f"{label}={ms*1e3:.0f}us")
    os.environ["B200PDM_GEMM_DBGMODE"] = "0"
    ms = time_it(fn)
    print(f"{name}: {flops/ms/1e9:.0f} TF/s | " + "  ".join(out), flush=True)


def conv_case(B, H, W, Ci, Co, nbuf=4):
    xs = [K.alloc2d(B * H * W, Ci).normal_() for _ in range(nbuf)]
    w = torch.randn(Co, 9, Ci, device="cuda", dtype=torch.bfloat16) * 0.02
    out = K.alloc2d(B * H * W, Co)
    sweep(f"conv {Ci}->{Co} @{H}x{W} B{B}", lambda i: K.conv_fwd(xs[i % nbuf], w, B, H, W, Co, 3, 1, out=out),
          2.0 * B * H * W * Co * 9 * Ci)


def lin_case(M, N, Kd, nbuf=4):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    out = K.alloc2d(M, N)
    sweep(f"linear M={M} N={N} K={Kd}", lambda i: K.linear_fwd(xs[i % nbuf], w, out=out), 2.0 * M * N * Kd)


conv_case(16, 64, 64, 960, 170)
lin_case(65536, 176, 8640)
conv_case(16, 64, 64, 320, 320)
lin_case(65536, 320, 2880)
conv_case(16, 64, 64, 640, 640)
conv_case(16, 32, 32, 1280, 1280)
lin_case(16384, 1280, 11520)
conv_case(16, 16, 16, 1280, 1280)
lin_case(8192, 8192, 8192)
lin_case(65536, 320, 320)
lin_case(65536, 2560, 320)
lin_case(16384, 1280, 1280)
