import os, sys
os.environ["B200PDM_GEMM_DBG"] = "1"
os.environ["B200PDM_GEMM_TRACE"] = "1"
sys.path.insert(0, "/root/repo")
import torch
from unlearn_ft_b200 import kernels as K
M, N, Kd = 65536, 2560, 320
x = K.alloc2d(M, Kd).normal_(); w = torch.randn(N, Kd, device="cuda", dtype=torch.bfloat16) * 0.02; out = K.alloc2d(M, N)
for mode in (0, 8, 32, 64, 128):
    os.environ["B200PDM_GEMM_DBGMODE"] = str(mode)
    sys.stderr.write(f"--- mode {mode}\n"); sys.stderr.flush()
    for _ in range(2):
        K.linear_fwd(x, w, out=out)
    torch.cuda.synchronize()
