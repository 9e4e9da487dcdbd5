import os, sys, statistics
sys.path.insert(0, "/root/repo")
import torch
from unlearn_ft_b200 import kernels as K
def time_it(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
for (M, N, Kd) in [(65536, 2560, 320), (65536, 320, 320)]:
    x = K.alloc2d(M, Kd).normal_(); w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02; out = K.alloc2d(M, N)
    res = []
    for mode in (0, 8, 16, 32):
        os.environ["B200PDM_GEMM_DBGMODE"] = str(mode)
        res.append(f"mode{mode}={time_it(lambda: K.linear_fwd(x, w, out=out))*1e3:.0f}us")
    os.environ["B200PDM_GEMM_DBGMODE"] = "0"
    print(M, N, Kd, " ".join(res), flush=True)
