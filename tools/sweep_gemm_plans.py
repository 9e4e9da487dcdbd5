"""Brute-force check of the GEMM tile planner (diag build only: B200PDM_LIB=libb200pdm_diag.so): every (block_n, m_sub, pair, splits)
the kernel accepts is timed for a set of step shapes (CUDA-graph replays), and the best is printed next to the planner's own choice.
Usage: B200PDM_LIB=libb200pdm_diag.so python tools/sweep_gemm_plans.py [wgrad|lin|conv]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K
from unlearn_ft_b200._lib import B200PdmError


def time_it(fn, iters=12):
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
    return e0.elapsed_time(e1) / (3 * iters) * 1e3


def sweep(name, fn, bns, flops):
    os.environ.pop("B200PDM_PLAN", None)
    K._WS_BYTES.clear()
    base = time_it(fn)
    res = []
    for bn, ms, pair, sp in itertools.product(bns, (1, 2), (0, 1), (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)):
        os.environ["B200PDM_PLAN"] = f"{bn},{ms},{pair},{sp}"
        K._WS_BYTES.clear()
        try:
            res.append((time_it(fn), bn, ms, pair, sp))
        except (B200PdmError, RuntimeError):
            torch.cuda.synchronize()
            continue
    os.environ.pop("B200PDM_PLAN", None)
    K._WS_BYTES.clear()
    res.sort()
    if os.environ.get("SWEEP_TSV"):
        with open(os.environ["SWEEP_TSV"], "a") as f:
            f.write(f"# {name}\tplanner\t{base:.2f}\n")
            for t, bn, ms, pair, sp in res:
                f.write(f"{name}\t{bn}\t{ms}\t{pair}\t{sp}\t{t:.2f}\n")
    top = "  ".join(f"{t:.1f}us(bn={bn} msub={ms} pair={pair} split={sp})" for t, bn, ms, pair, sp in res[:4])
    print(f"{name}: planner {base:.1f} us ({flops/base/1e6:.0f} TF/s) | best {top}", flush=True)


def wgrad_lin(M, N, Kd, nbuf=3):
    dys = [K.alloc2d(M, N).normal_() for _ in range(nbuf)]
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    dw = K.alloc2d(N, Kd, dtype=torch.float32, zero=True)
    sweep(f"wgrad linear pixels={M} dW={N}x{Kd}", lambda i: K.linear_wgrad(dys[i % nbuf], xs[i % nbuf], dw), (64, 128, 192, 256),
          2.0 * M * N * Kd)


def wgrad_conv(B, H, W, Ci, Co, nbuf=3):
    dys = [K.alloc2d(B * H * W, Co).normal_() for _ in range(nbuf)]
    xs = [K.alloc2d(B * H * W, Ci).normal_() for _ in range(nbuf)]
    dw = torch.zeros(Co, 9, K.round8(Ci), device="cuda")[:, :, :Ci]
    sweep(f"wgrad conv {Ci}->{Co} @{H}x{W} B{B}", lambda i: K.conv_wgrad(dys[i % nbuf], xs[i % nbuf], dw, B, H, W, 3, 1),
          (64, 128, 192, 256), 2.0 * B * H * W * Co * 9 * Ci)


def lin(M, N, Kd, res=False, nbuf=3):
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    rs = [K.alloc2d(M, N).normal_() for _ in range(nbuf)] if res else None
    w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
    b = torch.randn(N, device="cuda")
    outs = [K.alloc2d(M, N) for _ in range(nbuf)]
    sweep(f"linear M={M} N={N} K={Kd}{' +res' if res else ''}",
          lambda i: K.linear_fwd(xs[i % nbuf], w, b, residual=rs[i % nbuf] if res else None, out=outs[i % nbuf]),
          (64, 96, 128, 144, 160, 192, 224, 240, 256), 2.0 * M * N * Kd)


def conv(B, H, W, Ci, Co, nbuf=3):
    xs = [K.alloc2d(B * H * W, Ci).normal_() for _ in range(nbuf)]
    w = (torch.randn(Co, 9, K.round8(Ci), device="cuda", dtype=torch.bfloat16) * 0.02)[:, :, :Ci]
    out = K.alloc2d(B * H * W, Co)
    sweep(f"conv {Ci}->{Co} @{H}x{W} B{B}", lambda i: K.conv_fwd(xs[i % nbuf], w, B, H, W, Co, 3, 1, out=out),
          (128, 144, 160, 176, 192, 224, 240, 256), 2.0 * B * H * W * Co * 9 * Ci)


which = sys.argv[1] if len(sys.argv) > 1 else "wgrad"
if which == "wgrad":
    wgrad_lin(65536, 320, 320)
    wgrad_lin(65536, 320, 1360)
    wgrad_lin(16384, 640, 640)
    wgrad_lin(16384, 640, 2720)
    wgrad_lin(4096, 1280, 1280)
    wgrad_lin(4096, 1280, 5440)
    wgrad_conv(16, 64, 64, 170, 320)
    wgrad_conv(16, 64, 64, 320, 170)
    wgrad_conv(16, 64, 64, 640, 640)
    wgrad_conv(16, 32, 32, 340, 640)
    wgrad_conv(16, 32, 32, 640, 340)
    wgrad_conv(16, 16, 16, 680, 1280)
    wgrad_conv(16, 16, 16, 1280, 680)
    wgrad_conv(16, 8, 8, 680, 1280)
    wgrad_conv(16, 64, 64, 320, 320)
    wgrad_conv(16, 32, 32, 640, 640)
    wgrad_conv(16, 32, 32, 1280, 640)
    wgrad_conv(16, 16, 16, 1280, 1280)
    wgrad_conv(16, 16, 16, 2560, 1280)
    wgrad_conv(16, 8, 8, 1280, 1280)
    wgrad_lin(65536, 320, 128)
    wgrad_lin(16384, 640, 320)
    wgrad_lin(4096, 1280, 704)
    wgrad_lin(1232, 1280, 1024)
elif which == "lin":
    lin(65536, 320, 320)
    lin(65536, 320, 320, res=True)
    lin(65536, 960, 320)
    lin(16384, 640, 640, res=True)
    lin(16384, 1920, 640)
    lin(4096, 1280, 1280, res=True)
    lin(4096, 3840, 1280)
    lin(4096, 1280, 5120, res=True)
    lin(1024, 1280, 1280, res=True)
else:
    conv(16, 64, 64, 960, 170)
    conv(16, 64, 64, 320, 320)
    conv(16, 64, 64, 640, 640)
    conv(16, 32, 32, 1280, 1280)
    conv(16, 32, 32, 640, 640)
    conv(16, 16, 16, 1280, 1280)
    conv(16, 8, 8, 1280, 1280)
