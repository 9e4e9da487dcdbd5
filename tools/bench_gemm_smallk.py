"""Isolated timing of the small-K linear shapes of the training step (rotating inputs).  Usage: [B200PDM_MSUB=1|2] python ..."""
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


def lin_case(M, N, Kd, nbuf=4):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    b = torch.randn(N, device="cuda")
    out = K.alloc2d(M, N)
    ms = time_it(lambda i: K.linear_fwd(xs[i % nbuf], w, bias=b, out=out))
    fl = 2.0 * M * N * Kd
    byt = 2.0 * (M * Kd + N * Kd + M * N)
    print(f"linear M={M:6d} N={N:5d} K={Kd:5d}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {byt/ms/1e6:7.0f} GB/s", flush=True)


def lin_res_case(M, N, Kd, nbuf=4):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    rs = [K.alloc2d(M, N).normal_() for _ in range(nbuf)]
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    b = torch.randn(N, device="cuda")
    out = K.alloc2d(M, N)
    ms = time_it(lambda i: K.linear_fwd(xs[i % nbuf], w, bias=b, residual=rs[i % nbuf], out=out))
    byt = 2.0 * (M * Kd + N * Kd + 2 * M * N)
    print(f"linear+res M={M:6d} N={N:5d} K={Kd:5d}: {ms*1e3:8.1f} us  {2.0*M*N*Kd/ms/1e9:7.1f} TFLOP/s  {byt/ms/1e6:7.0f} GB/s", flush=True)


def geglu_case(M, Fh, Kd, save, nbuf=4):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    w = torch.randn(2 * Fh, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    b = torch.randn(2 * Fh, device="cuda")
    ms = time_it(lambda i: K.linear_geglu_fwd(xs[i % nbuf], w, b, save_pre=save))
    byt = 2.0 * (M * Kd + 2 * Fh * Kd + M * Fh * (3 if save else 1))
    print(f"geglu-fused M={M:6d} F={Fh:5d} K={Kd:5d} save_pre={int(save)}: {ms*1e3:8.1f} us  {4.0*M*Fh*Kd/ms/1e9:7.1f} TFLOP/s  {byt/ms/1e6:7.0f} GB/s", flush=True)


print("library:", os.environ.get("B200PDM_LIB", "libb200pdm.so"))
for shape in [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (65536, 320, 128), (65536, 320, 1280)]:
    lin_res_case(*shape)
for shape in [(65536, 1280, 320, False), (65536, 680, 320, True), (16384, 2560, 640, False), (16384, 1360, 640, True), (4096, 5120, 1280, False)]:
    geglu_case(*shape)
for shape in [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (65536, 1360, 320), (65536, 2560, 320), (16384, 5120, 640),
              (65536, 960, 320), (16384, 2720, 640), (4096, 5440, 1280), (65536, 320, 1280), (16384, 640, 2560), (4096, 10240, 1280),
              (65536, 320, 128), (1024, 1280, 1280), (1024, 10240, 1280)]:
    lin_case(*shape)
