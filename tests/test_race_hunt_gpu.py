"""Race hunt without a sanitizer (compute-sanitizer is not available on the GPU pool): kernels whose results do not involve
floating-point atomics must be BIT-identical from run to run, also while other work competes for the SMs on a second stream.  A hazard
in the mbarrier / tensor-memory protocols of the attention kernels (S / dP overwritten early, P / dS re-written before dV / dK / dQ have
read them) or of the row-shared-tap convolution (A box released before its third tap has been read) shows up as a difference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _noise(side, a, b):
    with torch.cuda.stream(side):
        for _ in range(2):
            a @ b


def test_attention_and_conv_bit_stable_under_concurrent_load():
    from unlearn_ft_b200 import kernels as K
    torch.manual_seed(0)
    side = torch.cuda.Stream()
    na = torch.randn(4096, 4096, device="cuda").bfloat16()
    nb = torch.randn_like(na)
    B, H, L = 2, 5, 2048
    q = torch.randn(B * L, H * 64, device="cuda").bfloat16()
    k, v, do = torch.randn_like(q), torch.randn_like(q), torch.randn_like(q)
    out, lse = K.attention_fwd(q, k, v, B, H, L, L, 0.125, want_lse=True)
    Bc, Hc, Ci, Co = 4, 64, 192, 320
    x = K.alloc2d(Bc * Hc * Hc, Ci).normal_()
    w = torch.randn(Co, 9, Ci, device="cuda", dtype=torch.bfloat16) * 0.02
    ref = None
    for it in range(12):
        if it % 2:
            _noise(side, na, nb)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        K.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, L, L, 0.125)
        o2, _ = K.attention_fwd(q, k, v, B, H, L, L, 0.125, want_lse=True)
        y = K.conv_fwd(x, w, Bc, Hc, Hc, Co, 3, 1)
        dx = K.conv_dgrad(y, w, Bc, Hc, Hc, Ci, 3)
        torch.cuda.synchronize()
        cur = dict(dk=dk.clone(), dv=dv.clone(), out=o2.clone(), y=y.clone(), dx=dx.clone())
        if ref is None:
            ref = cur
            continue
        for name in cur:
            assert torch.equal(cur[name], ref[name]), f"{name} differs in run {it}"
