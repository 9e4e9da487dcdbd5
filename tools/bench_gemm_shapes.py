"""Isolated timing of representative GEMM / conv shapes (CUDA events, rotating inputs > L2).
Usage: python tools/bench_gemm_shapes.py [conv|all]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K


def time_it(fn, iters=12):
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


def conv_case(B, H, W, Ci, Co, nbuf=4):
    xs = [K.alloc2d(B * H * W, Ci).normal_() for _ in range(nbuf)]
    w = (torch.randn(Co, 9, K.round8(Ci), device="cuda", dtype=torch.bfloat16) * 0.02)[:, :, :Ci]
    out = K.alloc2d(B * H * W, Co)
    ms = time_it(lambda i: K.conv_fwd(xs[i % nbuf], w, B, H, W, Co, 3, 1, out=out))
    fl = 2.0 * B * H * W * Co * 9 * Ci
    print(f"conv {Ci:5d}->{Co:5d} @{H:3d}x{W:<3d} B{B}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")


def lin_case(M, N, Kd, nbuf=4):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    out = K.alloc2d(M, N)
    ms = time_it(lambda i: K.linear_fwd(xs[i % nbuf], w, out=out))
    fl = 2.0 * M * N * Kd
    print(f"linear M={M:6d} N={N:5d} K={Kd:5d}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("conv", "all"):
    conv_case(16, 64, 64, 960, 170)
    conv_case(16, 64, 64, 320, 320)
    conv_case(16, 64, 64, 640, 640)
    conv_case(16, 32, 32, 1280, 1280)
    conv_case(16, 16, 16, 1280, 1280)
    conv_case(16, 8, 8, 1280, 1280)
    conv_case(16, 64, 64, 320, 170)
    conv_case(16, 64, 64, 170, 320)
    conv_case(16, 32, 32, 640, 340)
    conv_case(16, 32, 32, 340, 640)
    conv_case(16, 32, 32, 1920, 640)
    conv_case(16, 16, 16, 1280, 680)
    conv_case(16, 16, 16, 2560, 1280)
if which in ("lin", "all"):
    lin_case(8192, 8192, 8192)
    lin_case(65536, 256, 8192)
    lin_case(65536, 176, 8640)
    lin_case(65536, 320, 320)
    lin_case(65536, 2560, 320)
    lin_case(16384, 1280, 1280)
