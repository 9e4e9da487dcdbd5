"""TEST INFRASTRUCTURE (builder container only): live cross-check of the oracle against the REFERENCE'S OWN CODE.

    python -m oracle.live_check            # needs /root/reference; exits 0 when every check passes

Unlike tests/golden (recorded once), this re-runs the reference's gated blocks / prune() / forward (through oracle/refshim)
on arch vectors and seeds that are NOT in the golden file, and additionally compares GRADIENTS: d(loss)/d(parameter) of the
reference's forward under torch autograd against the oracle's, for every parameter of the pruned network.
Run by tests/test_oracle_vs_reference.py in a subprocess (the shim puts a fake `diffusers` into sys.modules), skipped
where /root/reference does not exist (the GPU box).
"""
from __future__ import annotations

import sys

import torch

from oracle import pdm_restated as P
from oracle import refshim
from oracle.make_golden import TINY, deterministic_fill, make_arch_vector, ref_pruned_model, tensor_digest

CASES = [dict(ratio=0.5, seed=31, drop=()), dict(ratio=0.85, seed=32, drop=(2, 3, 7, 10))]   # (TINY has 2 heads: ratios below 0.5 would keep none)


def main() -> int:
    if not refshim.available():
        print("reference tree not present")
        return 2
    ref = refshim.load_reference()
    g = torch.Generator().manual_seed(2024)
    sample = torch.randn(2, 4, 16, 16, generator=g)
    tsteps = torch.tensor([5, 642])
    ctx = torch.randn(2, 77, TINY["cross_attention_dim"], generator=g)
    target = torch.randn(2, 4, 16, 16, generator=g)
    worst_out, worst_grad = 0.0, 0.0
    for case in CASES:
        orc = P.UNetGated(**TINY)
        deterministic_fill(orc, 9)
        av = make_arch_vector(orc.get_structure(), case["ratio"], case["seed"], case["drop"])
        orc.set_structure(P.transform_arch_vector(av, orc.get_structure()))
        orc.prune()
        orc.eval()
        rm = ref_pruned_model(ref, TINY, av, 9)
        sd_o, sd_r = orc.state_dict(), rm.state_dict()
        assert list(sd_o.keys()) == list(sd_r.keys()), "state-dict keys / order differ"
        for k in sd_o:                                     # bit-exact index selection of the pruned weights
            assert sd_o[k].shape == sd_r[k].shape and tensor_digest(sd_o[k]) == tensor_digest(sd_r[k]), k
        y_o = orc(sample, tsteps, ctx).sample
        y_r = rm(sample, tsteps, ctx).sample
        err = float(((y_o - y_r).abs().max() / y_r.abs().max()).detach())
        worst_out = max(worst_out, err)
        assert err < 1e-5, ("forward", case, err)
        (y_o - target).pow(2).mean().backward()
        (y_r - target).pow(2).mean().backward()
        po, pr = dict(orc.named_parameters()), dict(rm.named_parameters())
        for k, p in pr.items():
            if p.grad is None:                             # parameters of depth-dropped blocks take no part in the forward
                assert po[k].grad is None or float(po[k].grad.abs().max()) == 0.0, k
                continue
            scale = float(p.grad.abs().max())
            if scale == 0.0:
                continue
            e = float((po[k].grad - p.grad).abs().max()) / scale
            worst_grad = max(worst_grad, e)
            assert e < 1e-4, ("gradient", case, k, e)
    print(f"live check ok: {len(CASES)} pruned networks, worst output error {worst_out:.2e}, worst gradient error {worst_grad:.2e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
