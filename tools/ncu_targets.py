"""One launch each of the kernels whose `ncu --set full` captures are committed per round (profiles/): the roofline conv
(960 -> 170 @ 64x64, batch 16), a small-K projection with residual (M = 65536, N = K = 320), the GEGLU-fused projection
(M = 65536, F = 1280, K = 320), attention forward / backward at L = 4096 (5 heads), GroupNorm+SiLU and LayerNorm at 64x64.
Run under:  ncu --set full --clock-control none --import-source on --profile-from-start off -o <file> python tools/ncu_targets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unlearn_ft_b200 import kernels as K

B, H = 16, 64
M = B * H * H
x960 = K.alloc2d(M, 960).normal_()
w = torch.randn(170, 9, 960, device="cuda", dtype=torch.bfloat16) * 0.02
bias = torch.zeros(170, device="cuda")
x320 = K.alloc2d(M, 320).normal_()
res = K.alloc2d(M, 320).normal_()
w320 = torch.randn(320, 320, device="cuda", dtype=torch.bfloat16) * 0.05
b320 = torch.zeros(320, device="cuda")
wff = torch.randn(2560, 320, device="cuda", dtype=torch.bfloat16) * 0.05
bff = torch.zeros(2560, device="cuda")
qkv = K.alloc2d(B * 4096, 960).normal_()
do = K.alloc2d(B * 4096, 320).normal_()
g320, be320 = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")


def all_once():
    K.conv_fwd(x960, w, B, H, H, 170, 3, 1, bias=bias)
    K.linear_fwd(x320, w320, b320, res)
    K.linear_geglu_fwd(x320, wff, bff, save_pre=False)
    q, k, v = qkv[:, :320], qkv[:, 320:640], qkv[:, 640:]
    o, lse = K.attention_fwd(q, k, v, B, 5, 4096, 4096, 0.125, want_lse=True)
    dq, dk, dv = (K.alloc2d(B * 4096, 320) for _ in range(3))
    K.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, 5, 4096, 4096, 0.125)
    K.groupnorm_fwd(x320, g320, be320, B, H * H, 32, 1e-5, True)
    K.layernorm_fwd(x320, g320, be320, 1e-5)


for _ in range(2):
    all_once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
all_once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
