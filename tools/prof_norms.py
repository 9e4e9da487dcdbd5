"""One launch of every norm kernel on the 64x64-level shape (for ncu --set full; see profiles/r1_ncu_norms_*.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, hw, C, G = 16, 4096, 320, 32
x = K.alloc2d(B * hw, C).normal_(); dy = K.alloc2d(B * hw, C).normal_()
g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(reps):
    y, stats = K.groupnorm_fwd(x, g, b, B, hw, G, 1e-5, True)
    K.groupnorm_bwd(dy, x, g, b, stats, dg, db, B, hw, G, True)
    y2, mean, rstd = K.layernorm_fwd(x, g, b, 1e-5)
    K.layernorm_bwd(dy, x, g, mean, rstd, dg, db)
torch.cuda.synchronize()
