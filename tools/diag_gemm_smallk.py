"""Diagnostics for the small-K projection GEMMs (K = 320 ... 1280, the epilogue-side of the kernel): time each shape with parts
of the kernel switched off (diag build: B200PDM_LIB=libb200pdm_diag.so, B200PDM_GEMM_DBGMODE bits: 1 quarter of
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


def wgrad_lin_case(M, N, Kd, nbuf=4):
    """dw[N, Kd] += dy[M, N]^T x[M, Kd]  (GEMM M=N, N=Kd, K=M pixels)."""
    dys = [K.alloc2d(M, N).normal_() for _ in range(nbuf)]
    xs = [K.alloc2d(M, Kd).normal_() for _ in range(nbuf)]
    dw = K.alloc2d(N, Kd, dtype=torch.float32, zero=True)
    sweep(f"wgrad linear pixels={M} dW={N}x{Kd}", lambda i: K.linear_wgrad(dys[i % nbuf], xs[i % nbuf], dw),
          2.0 * M * N * Kd, 2.0 * (M * Kd + M * N) + 4.0 * N * Kd)


def wgrad_conv_case(B, H, W, Ci, Co, nbuf=4):
    dys = [K.alloc2d(B * H * W, Co).normal_() for _ in range(nbuf)]
    xs = [K.alloc2d(B * H * W, Ci).normal_() for _ in range(nbuf)]
    dw = torch.zeros(Co, 9, K.round8(Ci), device="cuda")[:, :, :Ci]
    sweep(f"wgrad conv {Ci}->{Co} @{H}x{W} B{B}", lambda i: K.conv_wgrad(dys[i % nbuf], xs[i % nbuf], dw, B, H, W, 3, 1),
          2.0 * B * H * W * Co * 9 * Ci, 2.0 * B * H * W * (Ci + Co) + 4.0 * Co * 9 * Ci)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "wgrad":
        MODES[:] = [(0, "full"), (1, "quarter MMA"), (2, "no A loads"), (4, "no B loads"), (6, "no loads"), (128, "no epilogue"),
                    (6 | 128, "no loads, no epilogue")]
        wgrad_lin_case(65536, 320, 320)
        wgrad_lin_case(65536, 320, 1360)
        wgrad_lin_case(16384, 640, 640)
        wgrad_lin_case(4096, 1280, 1280)
        wgrad_lin_case(4096, 1280, 5440)
        wgrad_conv_case(16, 64, 64, 170, 320)
        wgrad_conv_case(16, 64, 64, 320, 170)
        wgrad_conv_case(16, 64, 64, 640, 640)
        wgrad_conv_case(16, 32, 32, 340, 640)
        wgrad_conv_case(16, 32, 32, 1280, 1280)
        wgrad_conv_case(16, 16, 16, 680, 1280)
        wgrad_conv_case(16, 8, 8, 680, 1280)
        sys.exit(0)
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
