"""CPU: pin the oracle (oracle/pdm_restated.py) against golden vectors produced by the REFERENCE'S OWN CODE
(oracle/make_golden.py, run in the builder container where /root/reference exists)."""
import hashlib
import os

import pytest
import torch

from oracle import diffusers_restated as D
from oracle import pdm_restated as P
from oracle.make_golden import SMALL64, TINY, deterministic_fill, tensor_digest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_structure_and_param_counts(gold):
    with torch.device("meta"):
        m = P.UNetGated()
        teacher = D.UNet2DConditionModel(**D.SD21_UNET_CONFIG)
    st = m.get_structure()
    assert st == gold["sd21_structure"]
    assert sum(w for s in st["width"] for w in s) == 1606 and sum(d for s in st["depth"] for d in s) == 14
    n = sum(p.numel() for p in m.parameters())
    assert n == gold["sd21_gated_params"] == 865_910_724          # published SD-2.x U-Net size
    assert sum(p.numel() for p in teacher.parameters()) == 865_910_724
    keys = list(m.state_dict().keys())
    assert hashlib.sha1("\n".join(keys).encode()).hexdigest() == gold["sd21_state_keys_sha1"]
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == gold["sd21_state_shapes"]
    assert list(teacher.state_dict().keys()) == keys               # teacher and student share diffusers key names


def test_arch_vector_plumbing(gold):
    st = gold["sd21_structure"]
    torch.manual_seed(1234)
    av = P.get_random_arch_vector(0.55, st)
    assert torch.equal(av, gold["av_full_seed1234_r055"])          # bit-exact incl. RNG consumption order
    assert av.shape == (1, 1620)
    sep = P.transform_arch_vector(av, st)
    assert [int(w.shape[1]) for w in sep["width"]] == gold["av_full_split_lens"]
    assert [float(w.sum()) for w in sep["width"]] == gold["av_full_split_sums"]
    assert [float(d) for d in sep["depth"]] == gold["av_full_depth"]


def test_masks_gates_snr(gold):
    assert torch.equal(P.hard_concrete(gold["hard_concrete_in"]), gold["hard_concrete_out"])
    assert torch.equal(P.width_gate(gold["width_gate_in"], torch.tensor([[1.0, 0.0, 1.0, 0.0]])), gold["width_gate_out"])
    assert torch.equal(P.linear_width_gate(gold["linear_gate_in"], torch.tensor([[0.0, 1.0, 1.0, 0.0]])),
                       gold["linear_gate_out"])
    a, b = gold["depth_gate_in"]
    assert torch.equal(P.depth_gate(a, b, torch.tensor([0.25])), gold["depth_gate_out"])
    acp = D.DDIMSchedulerLite().alphas_cumprod
    assert torch.equal(P.compute_snr(acp, gold["snr_timesteps"]), gold["snr"])


@pytest.mark.parametrize("case", ["r055", "r070_drop"])
def test_pruned_network_matches_reference(gold, case):
    g = gold[f"tiny_{case}"]
    m = P.UNetGated(**TINY)
    deterministic_fill(m, 3)
    m.set_structure(P.transform_arch_vector(g["arch_vector"], m.get_structure()))
    m.prune()
    m.eval()
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == g["shapes"]           # same keys, same pruned shapes
    assert {k: tensor_digest(v) for k, v in sd.items()} == g["digests"]       # bit-exact index selection
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]
    feats = {}
    P.cast_block_act_hooks(m, feats)
    inp = gold["tiny_inputs"]
    with torch.no_grad():
        y = m(inp["sample"], inp["timesteps"], inp["ctx"]).sample
    assert torch.allclose(y, g["sample"], rtol=1e-5, atol=1e-6)
    for k in P.BLOCK_KEYS:
        assert torch.allclose(feats[k], g["feats"][k], rtol=1e-5, atol=1e-6), k


def test_gated_equals_sliced():
    """Independent check of the slicing logic: with binary gates and norm2.bias == 0 the
    multiplicative-gate path (blocks.py:56-58,267-272,343-348,582-587) equals the physically pruned network."""
    import copy
    torch.manual_seed(1)
    gated = P.UNetGated(**TINY).eval()
    st = gated.get_structure()
    torch.manual_seed(5)
    av = P.get_random_arch_vector(0.6, st)
    n_w = sum(w for s in st["width"] for w in s)
    av[0, n_w + 2] = 0.0
    av[0, n_w + 11] = 0.0
    pruned = copy.deepcopy(gated)
    gated.set_structure(P.transform_arch_vector((av >= 0.5).float(), st))
    pruned.set_structure(P.transform_arch_vector(av, st))
    pruned.prune()
    x, t, ctx = torch.randn(2, 4, 16, 16), torch.tensor([3, 700]), torch.randn(2, 5, TINY["cross_attention_dim"])
    with torch.no_grad():
        assert torch.allclose(gated(x, t, ctx).sample, pruned(x, t, ctx).sample, atol=2e-5)


def test_step_loss_terms():
    """finetune_step == explicit formula of trainer.py:2457-2486 on a tiny model."""
    torch.manual_seed(0)
    student = P.build_pruned_unet(None, **TINY)
    teacher = P.build_pruned_unet(None, **TINY)
    fs, ft = {}, {}
    P.cast_block_act_hooks(student, fs), P.cast_block_act_hooks(teacher, ft)
    sched = D.DDIMSchedulerLite()
    x, n = torch.randn(2, 4, 16, 16), torch.randn(2, 4, 16, 16)
    t, ctx = torch.tensor([10, 900]), torch.randn(2, 5, TINY["cross_attention_dim"])
    loss, diff, kd, blk = P.finetune_step(student, teacher, sched, x, n, t, ctx, fs, ft)
    noisy, target = sched.add_noise(x, n, t), sched.get_velocity(x, n, t)
    with torch.no_grad():
        p_s, p_t = student(noisy, t, ctx).sample, teacher(noisy, t, ctx).sample
    snr = P.compute_snr(sched.alphas_cumprod, t) + 1
    w = torch.minimum(snr, torch.full_like(snr, 5.0)) / snr
    exp_diff = (((p_s - target) ** 2).mean(dim=(1, 2, 3)) * w).mean()
    exp_kd = ((p_s - p_t) ** 2).mean()
    assert torch.allclose(diff, exp_diff, rtol=1e-5) and torch.allclose(kd, exp_kd, rtol=1e-5)
    assert torch.allclose(loss.detach(), exp_diff + 0.1 * blk + 2.0 * exp_kd, rtol=1e-5)
    assert len(fs) == 9 and set(fs) == set(P.BLOCK_KEYS)


def test_step_and_upper_step_match_reference_trainer_golden():
    """tests/golden/reference_step_golden.pt holds the loss terms the reference TRAINER's own `step()` / `upper_step()` source
    (trainer.py:2403-2488, 2904-3001, run by oracle/make_step_golden.py on the reference's own pruned U-Net) produced, with the
    noise / timesteps it drew: the oracle reproduces them on the same sample."""
    from oracle.make_step_golden import teacher_model
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_step_golden.pt"), weights_only=False)
    m = P.UNetGated(**SMALL64)
    deterministic_fill(m, g["student_seed"])
    m.set_structure(P.transform_arch_vector(g["arch_vector"], m.get_structure()))
    m.prune()
    m.eval()
    teacher = teacher_model()
    fs, ft = {}, {}
    P.cast_block_act_hooks(m, fs), P.cast_block_act_hooks(teacher, ft)
    for case in g["cases"]:
        with torch.no_grad():
            out = P.finetune_step(m, teacher, D.DDIMSchedulerLite(), g["latents"], case["noise"], case["timesteps"],
                                  g["prompt_embeds"], fs, ft)
            up = P.upper_step(m, teacher, D.DDIMSchedulerLite(), g["latents"], case["noise"], case["timesteps"],
                              g["prompt_embeds"], g["empty_prompt_embeds"])
        for a, b in zip(out, case["step"]):
            assert abs(float(a) - b) <= 1e-6 * abs(b), (case["rng_seed"], [float(v) for v in out], case["step"])
        assert abs(float(up) - case["upper_step"][0]) <= 1e-6 * case["upper_step"][0]
        assert case["upper_step"][1] == 0.0 and case["upper_step"][3] == 0.0       # shipped upper weights: 0 / 1 / 0
