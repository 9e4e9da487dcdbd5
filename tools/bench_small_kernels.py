"""Graph-timed column sums (bias gradients) and a split-K conv with its finalize pass."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unlearn_ft_b200 import kernels as K
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (2 * iters) * 1e3
for rows, C in ((65536, 320), (65536, 170), (16384, 640), (4096, 1280)):
    xs = [K.alloc2d(rows, C).normal_() for _ in range(4)]
    out = torch.zeros(C, device="cuda")
    i = [0]
    def f():
        i[0] += 1
        K.colsum(xs[i[0] % 4], out)
    us = t(f)
    print(f"colsum rows={rows} C={C}: {us:.1f} us  {rows*C*2/us/1e3:.0f} GB/s")
# split-K conv at 8x8 (slab path)
B, H = 16, 8
x = K.alloc2d(B * H * H, 1280).normal_()
w = torch.randn(1280, 9, 1280, device="cuda", dtype=torch.bfloat16) * 0.02
out = K.alloc2d(B * H * H, 1280)
print("conv 1280->1280 @8x8 (split-K + finalize): %.1f us" % t(lambda: K.conv_fwd(x, w, B, H, H, 1280, 3, 1, out=out)))
