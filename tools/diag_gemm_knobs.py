"""Diagnostics: effect of tile rasterisation (B200PDM_ORDER) and CTA count (B200PDM_GRID) on isolated GEMM/conv shapes."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K


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


CASES = [({}, "base"), ({"B200PDM_ASPLIT": "2"}, "A2"), ({"B200PDM_ASPLIT": "4"}, "A4"), ({"B200PDM_BSPLIT": "2"}, "B2"),
         ({"B200PDM_ASPLIT": "2", "B200PDM_BSPLIT": "2"}, "A2B2"), ({"B200PDM_ASPLIT": "4", "B200PDM_BSPLIT": "4"}, "A4B4"),
         ({"B200PDM_ASPLIT": "8", "B200PDM_BSPLIT": "4"}, "A8B4"),
         ({"B200PDM_GRID": "16"}, "grid16"), ({"B200PDM_GRID": "16", "B200PDM_ASPLIT": "4", "B200PDM_BSPLIT": "4"}, "grid16+A4B4")]
KEYS = ("B200PDM_ORDER", "B200PDM_GRID", "B200PDM_GEMM_DBGMODE", "B200PDM_ASPLIT", "B200PDM_BSPLIT")


def sweep(name, fn, flops):
    out = []
    for env, label in CASES:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        ms = time_it(fn)
        g = int(env.get("B200PDM_GRID", 148))
        out.append(f"{label}={ms*1e3:.0f}us({flops/ms/1e9*148/g:.0f})")
    for k in KEYS:
        os.environ.pop(k, None)
    print(f"{name}: " + "  ".join(out), flush=True)


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


lin_case(8192, 8192, 8192)
lin_case(16384, 1280, 11520)
conv_case(16, 64, 64, 640, 640)
conv_case(16, 32, 32, 1280, 1280)
conv_case(16, 64, 64, 320, 320)
