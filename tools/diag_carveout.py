"""Diagnostics: does alternating a 0-smem elementwise kernel with the 227 KB-smem GEMM cost a carve-out reconfiguration?
Times a CUDA graph of 50 x [add, small GEMM]; run with and without B200PDM_CARVEOUT_TEST=1 (add_kernel prefers max smem)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K

M, N, Kd = 4096, 320, 320
x = K.alloc2d(M, Kd).normal_(); y = K.alloc2d(M, Kd).normal_(); z = K.alloc2d(M, Kd)
w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02
out = K.alloc2d(M, N)


def body(with_add, with_gemm, reps=50):
    for _ in range(reps):
        if with_add:
            K.add(x, y, out=z)
        if with_gemm:
            K.linear_fwd(z, w, out=out)


def timed(with_add, with_gemm):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body(with_add, with_gemm, 3)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body(with_add, with_gemm)
    g.replay(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) / 50 * 1e3


print(f"carveout test={os.environ.get('B200PDM_CARVEOUT_TEST')}: add only {timed(True, False):.2f} us, gemm only {timed(False, True):.2f} us, "
      f"add+gemm {timed(True, True):.2f} us per pair")
