"""Diagnostics for the small-K projection GEMMs (K = 320 ... 1280, the epilogue-side of the kernel): time each shape with parts
of the kernel switched off (diag build: B200PDM_LIB=unlearn_ft_b200/libb200pdm_diag.so, B200PDM_GEMM_DBGMODE bits: 1 quarter of
the MMAs, 2 no A loads, 4 no B loads, 8 no epilogue stores, 128 no epilogue work).  Durations only; results of modes != 0 are
garbage by construction."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K

MODES = [(0, "full"), (8, "no stores"), (128, "no epilogue"), (6, "no loads"), (6 | 128, "no loads, no epilogue"),
         (6 | 128 | 1, "skeleton")]


def time_it(fn, iters=24):
    """Average duration of one launch inside a replayed CUDA graph of `iters` launches (eager back-to-back launches of kernels
    under ~20 us measure the host's launch rate instead)."""
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


def sweep(name, fn, flops, bytes_):
    out = []
    for m, label in MODES:
        os.environ["B200PDM_GEMM_DBGMODE"] = str(m)
        out.append(f"{label}={time_it(fn)*1e3:.1f}")
    os.environ["B200PDM_GEMM_DBGMODE"] = "0"
    ms = time_it(fn)
    print(f"{name}: {ms*1e3:.1f} us = {flops/ms/1e9:.0f} TF/s, {bytes_/ms/1e6:.0f} GB/s | " + "  ".join(out), flush=True)


def lin_case(M, N, Kd, res=False, bias=False, nbuf=6):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    rs = [K.alloc2d(M, N).normal_() for _ in range(nbuf)] if res else None
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    b = torch.randn(N, device="cuda") if bias else None
    outs = [K.alloc2d(M, N) for _ in range(nbuf)]
    sweep(f"linear M={M} N={N} K={Kd}{' +res' if res else ''}{' +bias' if bias else ''}",
          lambda i: K.linear_fwd(xs[i % nbuf], w, b, residual=rs[i % nbuf] if res else None, out=outs[i % nbuf]),
          2.0 * M * N * Kd, 2.0 * (M * Kd + N * Kd + M * N * (2 if res else 1)))


if __name__ == "__main__":
    lin_case(65536, 320, 320)
    lin_case(65536, 320, 320, res=True, bias=True)
    lin_case(65536, 960, 320)
    lin_case(65536, 2560, 320, bias=True)
    lin_case(65536, 320, 1280, res=True, bias=True)
    lin_case(16384, 640, 640)
    lin_case(16384, 640, 640, res=True, bias=True)
    lin_case(16384, 1920, 640)
    lin_case(16384, 640, 2560, res=True, bias=True)
    lin_case(4096, 1280, 1280)
    lin_case(4096, 1280, 1280, res=True, bias=True)
    lin_case(4096, 3840, 1280)
    lin_case(4096, 1280, 5120, res=True, bias=True)
    lin_case(1024, 1280, 1280, res=True, bias=True)
    lin_case(65536, 320, 128, res=True, bias=True)
    lin_case(16, 1280, 1280, bias=True)
    lin_case(128, 128, 64)
