import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from unlearn_ft_b200 import kernels as k
sys.path.insert(0, 'tests')
from test_gemm_gpu import nhwc, pack_w, from2d

def run(B, H, W, Ci, Co, ks=3, st=1):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(Co, Ci, ks, ks, device="cuda", generator=g) * 0.05).bfloat16().float()
    out = k.conv_fwd(nhwc(x), pack_w(w), B, H, W, Co, ks, st)
    ref = F.conv2d(x, w, None, stride=st, padding=ks // 2)
    o = from2d(out, B, H // st, W // st)
    err = (o - ref).abs()
    scale = ref.abs().max()
    bad = err > 0.02 * scale
    print(f"case B{B} H{H} W{W} Ci{Ci} Co{Co}: max rel {err.max().item()/scale.item():.4f} bad frac {bad.float().mean().item():.4f}")
    if bad.any():
        # which channels / pixels
        ch = bad.any(dim=0).any(dim=-1).any(dim=-1).nonzero().flatten()
        print("  bad channels: n=", ch.numel(), ch[:20].tolist(), "...", ch[-5:].tolist())
        px = bad.any(dim=1).nonzero()
        print("  bad pixels n=", px.shape[0], px[:10].tolist())
for case in [(4,8,8,1280,680), (4,8,8,1280,176), (4,8,8,1280,352), (4,8,8,640,680), (4,8,8,64,680), (2,8,8,1280,680),
             (4,16,16,1280,680), (4,8,8,1280,256), (4,8,8,1280,512), (4,8,8,1280,128), (4,8,8,2560,680), (16,8,8,1280,1280)]:
    run(*case)
