import os, sys
os.environ["B200PDM_ATTN_DBG"] = "1"
sys.path.insert(0, "/root/repo")
import torch
from unlearn_ft_b200 import kernels as K
B, H, Lq, Lk, D = 16, 5, 4096, 4096, 64
q = torch.randn(B * Lq, H * D, device="cuda").bfloat16(); k = torch.randn(B * Lk, H * D, device="cuda").bfloat16()
v = torch.randn(B * Lk, H * D, device="cuda").bfloat16()
for _ in range(3):
    K.attention_fwd(q, k, v, B, H, Lq, Lk, 0.125)
torch.cuda.synchronize()
do = torch.randn(B * Lq, H * D, device="cuda").bfloat16()
out, lse = K.attention_fwd(q, k, v, B, H, Lq, Lk, 0.125, want_lse=True)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(2):
    K.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, Lq, Lk, 0.125)
torch.cuda.synchronize()
