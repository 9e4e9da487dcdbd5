"""Race hunt without a sanitizer: attention forward / backward (dK, dV accumulate in TMEM in a fixed order) and the row-shared-tap conv are
run 40 times, every other time with competing GEMMs on a second stream, and must be bit-identical -- a shared-memory / TMEM hazard in the
barrier protocols shows up as run-to-run differences."""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unlearn_ft_b200 import kernels as K
torch.manual_seed(0)
B, H, L = 4, 5, 4096
q = torch.randn(B * L, H * 64, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
out, lse = K.attention_fwd(q, k, v, B, H, L, L, 0.125, want_lse=True)
side = torch.cuda.Stream()
noise_a = torch.randn(8192, 8192, device="cuda").bfloat16(); noise_b = torch.randn_like(noise_a)
ref = None
bad = 0
for it in range(40):
    if it % 2:                      # competing work on another stream: changes CTA scheduling / timing
        with torch.cuda.stream(side):
            for _ in range(3): noise_a @ noise_b
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    K.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, L, L, 0.125)
    o2, _ = K.attention_fwd(q, k, v, B, H, L, L, 0.125, want_lse=True)
    torch.cuda.synchronize()
    cur = (dk.clone(), dv.clone(), o2.clone())
    if ref is None: ref = cur
    else:
        for a, b, n in zip(cur, ref, ("dk", "dv", "out")):
            if not torch.equal(a, b): bad += 1; print("iteration", it, n, "differs: max", (a.float() - b.float()).abs().max().item())
print("attention: dk / dv / out bitwise stable over 40 runs" if bad == 0 else f"attention: {bad} mismatches")
# kh3 conv fprop / dgrad
Bc, Hc, Ci, Co = 8, 64, 320, 320
x = K.alloc2d(Bc * Hc * Hc, Ci).normal_(); w = (torch.randn(Co, 9, Ci, device="cuda", dtype=torch.bfloat16) * 0.02)
ref = None; bad = 0
for it in range(40):
    if it % 2:
        with torch.cuda.stream(side):
            for _ in range(3): noise_a @ noise_b
    y = K.conv_fwd(x, w, Bc, Hc, Hc, Co, 3, 1)
    dx = K.conv_dgrad(y, w, Bc, Hc, Hc, Ci, 3)
    torch.cuda.synchronize()
    cur = (y.clone(), dx.clone())
    if ref is None: ref = cur
    else:
        for a, b, n in zip(cur, ref, ("y", "dx")):
            if not torch.equal(a, b): bad += 1; print("iteration", it, n, "differs")
print("conv (row-shared taps): y / dx bitwise stable over 40 runs" if bad == 0 else f"conv: {bad} mismatches")
