"""Isolated timing of the fused attention kernels on the model's shapes."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K

def t(fn, iters=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)

for (B, H, Lq, Lk) in [(16, 5, 4096, 4096), (16, 2, 4096, 4096), (16, 10, 1024, 1024), (16, 20, 256, 256), (16, 5, 4096, 77)]:
    D = 64
    q = torch.randn(B * Lq, H * D, device="cuda").bfloat16(); k = torch.randn(B * Lk, H * D, device="cuda").bfloat16()
    v = torch.randn(B * Lk, H * D, device="cuda").bfloat16(); do = torch.randn(B * Lq, H * D, device="cuda").bfloat16()
    out, lse = K.attention_fwd(q, k, v, B, H, Lq, Lk, 0.125, want_lse=True)
    ms = t(lambda: K.attention_fwd(q, k, v, B, H, Lq, Lk, 0.125, out=out))
    fl = 4.0 * B * H * Lq * Lk * D
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    msb = t(lambda: K.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, Lq, Lk, 0.125))
    print(f"attn B{B} H{H} Lq{Lq} Lk{Lk}: fwd {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TFLOP/s | bwd {msb*1e3:8.1f} us {2.5*fl/msb/1e9:7.1f} TFLOP/s")
