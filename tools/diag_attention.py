"""Phase timers of the attention kernels (diag build: `make -C unlearn_ft_b200/csrc diag`, then
B200PDM_LIB=libb200pdm_diag.so B200PDM_ATTN_DBG=1 python tools/diag_attention.py): the library prints, for CTA 0, the cycles its softmax
warps and its MMA warp spent in each phase (forward: wait for S, TMEM read, exponentials, P store; backward: wait for S/dP, softmax,
dQ drain), summed over the key / query blocks, at L = 4096, 5 heads, batch 16.  Durations only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unlearn_ft_b200 import kernels as K
B, H, L = 16, 5, 4096
q = torch.randn(B * L, H * 64, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(2):
    K.attention_fwd(q, k, v, B, H, L, L, 0.125)
torch.cuda.synchronize()
# backward phase timers of CTA 0 (diag build, B200PDM_ATTN_DBG=1): printed by the library on stderr
do = torch.randn_like(q)
out, lse = K.attention_fwd(q, k, v, B, H, L, L, 0.125, want_lse=True)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(2):
    K.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, L, L, 0.125)
torch.cuda.synchronize()
