"""Un-pruned ("gated") training mode of the pruning phase: a FULL-width U-Net whose width / depth gates
multiply activations at run time (reference pdm/models/gates.py:15-62, applied at blocks.py:56-58,267-272,343-348,582-587,
1241-1244), per sample when one gate row per sample is given, with d loss / d gate -- the gradient the hypernetwork is
trained with (trainer.py:1159-1321) -- returned to the tensor handed to `set_structure()`.

Oracle: `oracle.pdm_restated.UNetGated` with `set_structure()` and NO `prune()` (its gated forward is pinned bit for bit to
the reference's own blocks by oracle/live_check.py), fp32 autograd on the same device.  Tolerances: prediction 2e-2 (north
star, bf16), gradients by relative L2 error."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _models():
    from oracle import pdm_restated as P
    from oracle.make_golden import SMALL64, deterministic_fill
    from unlearn_ft_b200.pdm.models import UNet2DConditionModelGated
    orc = P.UNetGated(**SMALL64)
    deterministic_fill(orc, 3)
    cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"],
               cross_attention_dim=SMALL64["cross_attention_dim"])
    mine = UNet2DConditionModelGated(cfg, arch_vector=None, seed=None)
    mine.load_state_dict(orc.state_dict())
    return mine, orc.eval().cuda(), SMALL64


def _gates(n, rows, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(rows, n, generator=g) * 0.9 + 0.1           # soft gates in [0.1, 1]
    hard = torch.rand(rows, n, generator=g)
    a = torch.where(hard < 0.15, torch.zeros_like(a), a)       # some closed ...
    a = torch.where(hard > 0.85, torch.ones_like(a), a)        # ... and some fully open (hard-concrete saturates at both ends)
    return a


@pytest.mark.parametrize("rows", [1, 2])
def test_gated_unpruned_forward_and_gate_gradients_match_oracle(rows):
    from oracle import pdm_restated as P
    mine, orc, SMALL64 = _models()
    n = mine.arch_vector.shape[1]
    B = 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, 16, 16, generator=g).cuda()
    t = torch.tensor([40, 700]).cuda()
    ctx = torch.randn(B, 77, SMALL64["cross_attention_dim"], generator=g).cuda()
    tgt = torch.randn(B, 4, 16, 16, generator=g).cuda()
    a0 = _gates(n, rows, 5)
    a_m = a0.clone().cuda().requires_grad_(True)
    a_o = a0.clone().cuda().requires_grad_(True)
    # oracle: set_structure without prune -> runtime gates (per sample for rows == B)
    orc.set_structure(P.transform_arch_vector(a_o, orc.get_structure()))
    ref = orc(x, t, ctx).sample
    F.mse_loss(ref, tgt).backward()
    # ours: the same call sequence on the full-width model
    mine.set_structure(a_m)
    assert not mine.is_pruned()
    out = mine(x, t, ctx).sample
    print("gated forward rel", rel(out, ref))
    assert rel(out, ref) < 2e-2
    F.mse_loss(out, tgt).backward()
    torch.cuda.synchronize()
    assert a_m.grad is not None and a_m.grad.shape == a_o.grad.shape
    e = rel_l2(a_m.grad, a_o.grad)
    print("d gate rel-L2", e, "| nonzero", int((a_o.grad != 0).sum()), "of", a_o.grad.numel())
    assert e < 5e-2
    params = dict(mine.named_parameters())
    worst = 0.0
    for k, p in orc.named_parameters():
        if p.grad is None or p.grad.abs().max() < 1e-10:
            continue
        err = rel_l2(params[k].grad, p.grad)
        worst = max(worst, err)
        assert err < 8e-2, (k, err)
    print("worst parameter-gradient rel-L2", worst)


def test_open_gates_equal_the_ungated_network_and_closed_depth_gates_bypass():
    """All gates 1 -> exactly the plain forward; a closed depth gate makes the block return its (non-skip) input, which is
    what prune() then materialises by dropping the block (blocks.py:502-515,1134-1138)."""
    mine, orc, SMALL64 = _models()
    n = mine.arch_vector.shape[1]
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 16, 16, generator=g).cuda()
    t = torch.tensor([5, 500]).cuda()
    ctx = torch.randn(2, 77, SMALL64["cross_attention_dim"], generator=g).cuda()
    with torch.no_grad():
        plain = mine(x, t, ctx).sample
        mine.set_structure(torch.ones(1, n))
        ones = mine(x, t, ctx).sample
        assert rel(ones, plain) < 2e-2                    # (gated path: un-fused GEGLU, an extra bf16 rounding per gate)
        a = torch.ones(1, n)
        a[0, -3] = 0.0
        a[0, -6] = 0.0                                    # two depth gates closed, widths open
        mine.set_structure(a)
        gated = mine(x, t, ctx).sample
        from unlearn_ft_b200.pdm.models import UNet2DConditionModelPruned
        pruned = UNet2DConditionModelPruned(dict(mine._config), arch_vector=a, trainable=False, seed=None)
        pruned.load_unpruned_state_dict(mine.state_dict())
        assert rel(gated, pruned(x, t, ctx).sample) < 2e-2
