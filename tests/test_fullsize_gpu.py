"""Parity at BASELINE.json's FULL sizes (SD-2.1 dimensions, batch 16, 64x64 latents, r = 0.55 student + teacher).

The CPU oracle needs seconds per sample at these sizes, so here every kernel family is checked (a) directly against torch's
own GPU ops on the same bf16-exact inputs -- torch's cuDNN / cuBLAS / SDPA results are an independent implementation of the
very calls the reference makes (blocks.py:49,244-283,318-381) -- and (b) through size-independent properties: adjointness of
fprop / dgrad / wgrad, softmax rows summing to one, gated == sliced networks, AdamW == torch.optim.AdamW, finite loss and
a decreasing loss on a repeated batch for the whole step.  Tolerances: 2e-2 relative (north star) for bf16 outputs,
2e-3 for fp32-accumulated gradients.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

B_FULL = 16


def _k():
    from unlearn_ft_b200 import kernels
    return kernels


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def nhwc(x):
    k = _k()
    B, C, H, W = x.shape
    t = k.alloc2d(B * H * W, C)
    t.copy_(x.permute(0, 2, 3, 1).reshape(B * H * W, C))
    return t


def from2d(t, B, H, W):
    return t.float().reshape(B, H, W, -1).permute(0, 3, 1, 2)


def pack_w(w):
    k = _k()
    O, I, kh, kw = w.shape
    buf = torch.zeros(O, kh * kw, k.round8(I), device="cuda", dtype=torch.bfloat16)
    buf[:, :, :I] = w.permute(0, 2, 3, 1).reshape(O, kh * kw, I)
    return buf[:, :, :I]


# the largest-FLOP student conv of the step (up_blocks.3.resnets.0.conv1), a pruned mid-width pair, and the stride-2 downsampler
@pytest.mark.parametrize("H,Ci,Co,st", [(64, 960, 170, 1), (64, 170, 320, 1), (32, 1280, 340, 1), (64, 320, 320, 2)])
def test_conv_fullsize_vs_cudnn_and_adjoint(H, Ci, Co, st):
    k = _k()
    B, W = B_FULL, H
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) * 0.03).bfloat16()
    bias = torch.randn(Co, device="cuda", generator=g)
    dy = torch.randn(B, Co, H // st, W // st, device="cuda", generator=g).bfloat16()
    x2, dy2, w2 = nhwc(x.float()), nhwc(dy.float()), pack_w(w.float())
    y = k.conv_fwd(x2, w2, B, H, W, Co, 3, st, bias=bias)
    ref = F.conv2d(x.float(), w.float(), bias, stride=st, padding=1)          # fp32 cuDNN on bf16-exact inputs
    assert rel_err(from2d(y, B, H // st, W // st), ref) < 1e-2
    dw = torch.zeros(Co, 9, k.round8(Ci), device="cuda")[:, :, :Ci]
    k.conv_wgrad(dy2, x2, dw, B, H, W, 3, st)
    dw_ref = torch.nn.grad.conv2d_weight(x.float(), (Co, Ci, 3, 3), dy.float(), stride=st, padding=1)
    assert rel_err(dw, dw_ref.permute(0, 2, 3, 1).reshape(Co, 9, Ci)) < 2e-3
    # adjointness, independent of any reference: <conv(x, w), dy> == <w, wgrad(x, dy)> (== <x, dgrad(dy, w)> for stride 1)
    y_nobias = k.conv_fwd(x2, w2, B, H, W, Co, 3, st)
    lhs = (y_nobias.float() * dy2.float()).sum().item()
    rhs_w = (dw.double() * w2.double()).sum().item()
    assert abs(lhs - rhs_w) <= 1e-2 * abs(rhs_w)                               # y is rounded to bf16, dw is fp32
    if st == 1:
        dx = k.conv_dgrad(dy2, w2, B, H, W, Ci, 3)
        dx_ref = torch.nn.grad.conv2d_input((B, Ci, H, W), w.float(), dy.float(), padding=1)
        assert rel_err(from2d(dx, B, H, W), dx_ref) < 1e-2
        rhs_x = (dx.float() * x2.float()).sum().item()
        assert abs(lhs - rhs_x) <= 2e-2 * abs(lhs)


@pytest.mark.parametrize("M,N,K", [(65536, 2560, 320), (65536, 320, 1280), (16384, 5120, 640), (16 * 77, 680, 1024)])
def test_linear_fullsize_vs_cublas(M, N, K):
    k = _k()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = k.alloc2d(M, K)
    x.copy_(torch.randn(M, K, device="cuda", generator=g))
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = k.alloc2d(M, N)
    res.copy_(torch.randn(M, N, device="cuda", generator=g))
    y = k.linear_fwd(x, w, bias=bias, residual=res)
    ref = x.float() @ w.float().t() + bias + res.float()
    assert rel_err(y, ref) < 1e-2
    dy = k.alloc2d(M, N)
    dy.copy_(torch.randn(M, N, device="cuda", generator=g))
    dw = torch.zeros(N, K, device="cuda")
    k.linear_wgrad(dy, x, dw)
    assert rel_err(dw, dy.float().t() @ x.float()) < 2e-3
    dx = k.linear_dgrad(dy, w)
    assert rel_err(dx, dy.float() @ w.float()) < 1e-2


@pytest.mark.parametrize("H,Lq,Lk", [(5, 4096, 4096), (10, 1024, 1024), (5, 4096, 77), (20, 64, 77)])
def test_attention_fullsize(H, Lq, Lk):
    """Self / cross attention at the step's shapes: vs SDPA, rows of the (implicit) softmax sum to one (V = 1 -> O = 1), and
    the backward against autograd through SDPA."""
    k = _k()
    B, D = B_FULL, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B * Lq, H * D, device="cuda", generator=g).bfloat16()
    kk = torch.randn(B * Lk, H * D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B * Lk, H * D, device="cuda", generator=g).bfloat16()
    out, lse = k.attention_fwd(q, kk, v, B, H, Lq, Lk, D ** -0.5, want_lse=True)
    qh, kh, vh = (t.reshape(B, -1, H, D).transpose(1, 2) for t in (q, kk, v))
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (qh, kh, vh))
    ref = F.scaled_dot_product_attention(qr, kr, vr)                          # bf16 flash / mem-efficient SDPA
    assert rel_err(out, ref.transpose(1, 2).reshape(B * Lq, H * D)) < 2e-2
    ones, _ = k.attention_fwd(q, kk, torch.ones_like(v), B, H, Lq, Lk, D ** -0.5, want_lse=True)
    assert (ones.float() - 1.0).abs().max().item() < 1e-2
    do = torch.randn(B * Lq, H * D, device="cuda", generator=g).bfloat16()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)
    k.attention_bwd(q, kk, v, out, do, lse, dq, dk, dv, B, H, Lq, Lk, D ** -0.5)
    ref.backward(do.reshape(B, Lq, H, D).transpose(1, 2))
    for mine, r in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert rel_err(mine, r.transpose(1, 2).reshape(mine.shape)) < 3e-2    # two bf16 pipelines


def test_norms_fullsize():
    k = _k()
    B, hw, C = B_FULL, 4096, 320
    g = torch.Generator(device="cuda").manual_seed(4)
    x = k.alloc2d(B * hw, C)
    x.copy_(torch.randn(B * hw, C, device="cuda", generator=g) * 2 + 0.3)
    gamma, beta = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    y, stats = k.groupnorm_fwd(x, gamma, beta, B, hw, 32, 1e-5, True)
    xr = x.float().reshape(B, hw, C).permute(0, 2, 1)
    ref = F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(B * hw, C)
    assert rel_err(y, ref) < 1e-2
    yl, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-5)
    assert rel_err(yl, F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)) < 1e-2


def test_adamw_full_arena_matches_torch():
    """One fused step over a 508 M-element arena (the r = 0.55 student) vs torch.optim.AdamW on the same tensors."""
    k = _k()
    n = 508_224_076 // 4 * 4
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.randn(n, device="cuda", generator=g) * 0.02
    grad = torch.randn(n, device="cuda", generator=g) * 1e-3
    ref_p = p.clone().requires_grad_(True)
    ref_p.grad = grad.clone()
    opt = torch.optim.AdamW([ref_p], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v, sh = torch.zeros_like(p), torch.zeros_like(p), torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in (1, 2):
        opt.step()
        k.adamw_step(p, grad.clone(), m, v, sh, 1e-4, 0.9, 0.999, 1e-8, 0.01, step, zero_grad=False)
    assert rel_err(p, ref_p.detach()) < 1e-5
    assert torch.equal(sh, p.bfloat16())


@pytest.fixture(scope="module")
def full_models():
    from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModel, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
    torch.manual_seed(43)
    av = HyperStructure.get_random_arch_vector(0.55, structure_from_config(SD21_CONFIG))
    student = UNet2DConditionModelPruned(arch_vector=av, seed=43)
    teacher = UNet2DConditionModel(seed=44)
    return av, student, teacher


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(latents=torch.randn(B, 4, 64, 64, generator=g).cuda(), noise=torch.randn(B, 4, 64, 64, generator=g).cuda(),
                timesteps=torch.randint(0, 1000, (B,), generator=g).cuda(),
                prompt_embeds=torch.randn(B, 77, 1024, generator=g).bfloat16().cuda())


def test_fullsize_structure_and_step(full_models):
    """BASELINE config 2 end to end: parameter counts of SURVEY 8(c), a finite four-part loss, gradients in every block, and
    a loss that goes down when the same batch is repeated (lr raised to 1e-4 so 6 steps are enough to see it)."""
    from unlearn_ft_b200.pdm.training import UnetFineTuner
    av, student, teacher = full_models
    assert teacher.num_parameters() == 865_910_724
    assert student.num_parameters() == 508_224_076
    tuner = UnetFineTuner(student, teacher, lr=1e-4, warmup_steps=0)
    batch = _batch(B_FULL, 0)
    loss, diff, kd, blk = tuner.step(batch)
    vals = [float(v.detach()) for v in (loss, diff, kd, blk)]
    assert all(torch.isfinite(torch.tensor(vals))), vals
    assert abs(vals[0] - (vals[1] + 2.0 * vals[2] + 0.1 * vals[3])) < 1e-3 * abs(vals[0])   # weights 1 / 2 / 0.1
    loss.backward()
    gabs = student.arena.grad.abs()
    for (mod, attr), (lo, hi) in tuner.reducer.buckets.items():
        assert float(gabs[lo:hi].max()) > 0, attr
    student.arena.grad.zero_()
    first = None
    for i in range(6):
        out = tuner.train_step(batch)
        cur = float(out[0].detach())
        first = cur if first is None else first
    assert cur < first, (first, cur)


def test_fullsize_batch_rows_are_independent(full_models):
    """Size-independent property of the forward: sample b of a batch-16 call equals the same sample run in a batch of 2
    (no cross-sample leakage through tiles, GroupNorm statistics or attention batching)."""
    av, student, teacher = full_models
    b16 = _batch(B_FULL, 7)
    with torch.no_grad():
        full = teacher(b16["latents"], b16["timesteps"], b16["prompt_embeds"]).sample
        sub = {k: v[5:7].contiguous() for k, v in b16.items()}
        part = teacher(sub["latents"], sub["timesteps"], sub["prompt_embeds"]).sample
    # different batch sizes take different tile / split-K plans (fp32 atomics order, bf16 rounding points): the same 2e-2
    # budget as any two bf16 evaluations of this 30-layer network, far below the O(1) error of a cross-sample leak
    err = rel_err(full[5:7], part)
    print("batch-16 vs batch-2 rows: max-abs relative difference", err)
    assert err < 2e-2
