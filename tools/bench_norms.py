"""Isolated timing of LayerNorm / GroupNorm kernels on the model's shapes (achieved HBM bandwidth vs algorithmic bytes)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K


def t(fn, reps=20, iters=5):
    """Kernel-only time: `reps` calls captured into a CUDA graph (the Python wrappers cost more than these kernels run)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) / reps


for rows, C in [(65536, 320), (16384, 640), (4096, 1280)]:
    x = K.alloc2d(rows, C).normal_(); dy = K.alloc2d(rows, C).normal_()
    g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    y, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    f = t(lambda: K.layernorm_fwd(x, g, b, 1e-5))
    bw = t(lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dg, db))
    nb = rows * C * 2
    print(f"LN rows={rows} C={C}: fwd {f*1e3:6.1f} us {2*nb/f/1e6:6.0f} GB/s | bwd {bw*1e3:6.1f} us {3*nb/bw/1e6:6.0f} GB/s", flush=True)

for B, hw, C, G in [(16, 4096, 320, 32), (16, 4096, 960, 32), (16, 4096, 170, 17), (16, 1024, 640, 32), (16, 256, 1280, 32)]:
    x = K.alloc2d(B * hw, C).normal_(); dy = K.alloc2d(B * hw, C).normal_()
    g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    y, stats = K.groupnorm_fwd(x, g, b, B, hw, G, 1e-5, True)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    f = t(lambda: K.groupnorm_fwd(x, g, b, B, hw, G, 1e-5, True))
    bw = t(lambda: K.groupnorm_bwd(dy, x, g, b, stats, dg, db, B, hw, G, True))
    nb = B * hw * C * 2
    print(f"GN B={B} hw={hw} C={C}: fwd {f*1e3:6.1f} us ({3*nb/f/1e6:6.0f} GB/s at 2R+1W) | bwd {bw*1e3:6.1f} us ({5*nb/bw/1e6:6.0f} GB/s at 4R+1W)", flush=True)
