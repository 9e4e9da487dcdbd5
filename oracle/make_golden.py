"""TEST INFRASTRUCTURE (builder container only): writes tests/golden/*.pt by running the REFERENCE'S OWN CODE.

    python -m oracle.make_golden            # needs /root/reference; not runnable on the GPU box

What is executed from /root/reference (through oracle/refshim, i.e. on top of the restated diffusers base classes):
  * pdm/models/unet/unet_2d_conditional.py : UNet2DConditionModelGated.__init__/get_structure/set_structure/forward
  * pdm/models/unet/blocks.py              : every gated block, every prune()/prune_module()
  * pdm/models/hypernet.py                 : HyperStructure.transform_arch_vector / get_random_arch_vector
  * pdm/models/gates.py, pdm/utils/estimation_utils.py (hard_concrete), pdm/utils/metric_utils.py (compute_snr)
The pruning sequence replicated here is unet_2d_conditional.py:2444-2459 (from_pretrained needs hub files).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TINY = dict(block_out_channels=(32, 64, 64, 64), heads=(2, 2, 2, 2), cross_attention_dim=48)
# smallest configuration with head_dim 64 (what the sm_100a attention path is specialised for), for GPU parity
SMALL64 = dict(block_out_channels=(128, 128, 256, 256), heads=(2, 2, 4, 4), cross_attention_dim=64)
REF_BLOCKS = dict(down_block_types=("CrossAttnDownBlock2DHalfGated",) * 3 + ("DownBlock2DHalfGated",),
                  mid_block_type="UNetMidBlock2DCrossAttnWidthGated",
                  up_block_types=("UpBlock2DHalfGated",) + ("CrossAttnUpBlock2DHalfGated",) * 3)


def deterministic_fill(model: torch.nn.Module, seed: int) -> None:
    """Weights as a pure function of (state-dict key, shape, seed) so that the reference model (here) and the oracle /
    B200 model (in the tests) hold identical parameters without shipping them."""
    with torch.no_grad():
        for name, p in model.state_dict().items():
            h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:4], "little")
            g = torch.Generator().manual_seed(h)
            r = torch.randn(p.shape, generator=g)
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(r / fan_in ** 0.5)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.05 * r)


def make_arch_vector(structure, ratio, seed, drop_depth=()):
    from oracle.pdm_restated import get_random_arch_vector

    torch.manual_seed(seed)
    av = get_random_arch_vector(ratio, structure)
    n_w = sum(w for s in structure["width"] for w in s)
    for i in drop_depth:
        av[0, n_w + i] = 0.1
    return av


def tensor_digest(t: torch.Tensor) -> str:
    return hashlib.sha1(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def ref_pruned_model(ref, cfg, arch_vector, seed):
    m = ref.unet.UNet2DConditionModelGated(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                           attention_head_dim=cfg["heads"],
                                           cross_attention_dim=cfg["cross_attention_dim"], use_linear_projection=True,
                                           norm_eps=1e-5, gated_ff=True, ff_gate_width=32, **REF_BLOCKS)
    deterministic_fill(m, seed)
    # unet_2d_conditional.py:2444-2459
    sep = ref.hypernet.HyperStructure.transform_arch_vector(arch_vector, m.get_structure())
    m.set_structure(sep)
    for _, mod in m.named_modules():
        if hasattr(mod, "prune"):
            mod.prune()
    for mod in m.modules():
        if hasattr(mod, "prune_module"):
            mod.prune_module()
    return m.eval()


def main():
    sys.path.insert(0, ROOT)
    from oracle import refshim
    from oracle import pdm_restated as P

    ref = refshim.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    out = {}

    # 1. full-size SD-2.1 structure from the reference ctor (meta device: no memory)
    with torch.device("meta"):
        full = ref.unet.UNet2DConditionModelGated(sample_size=96, block_out_channels=(320, 640, 1280, 1280),
                                                  attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                                                  use_linear_projection=True, **REF_BLOCKS)
    st = full.get_structure()
    out["sd21_structure"] = {"width": st["width"], "depth": st["depth"]}
    out["sd21_gated_params"] = sum(p.numel() for p in full.parameters())
    out["sd21_state_keys_sha1"] = hashlib.sha1("\n".join(full.state_dict().keys()).encode()).hexdigest()
    out["sd21_state_shapes"] = {k: list(v.shape) for k, v in full.state_dict().items()}

    # 2. arch-vector plumbing from the reference classmethods
    torch.manual_seed(1234)
    av_full = ref.hypernet.HyperStructure.get_random_arch_vector(0.55, st)
    out["av_full_seed1234_r055"] = av_full
    sep = ref.hypernet.HyperStructure.transform_arch_vector(av_full, st)
    out["av_full_split_lens"] = [int(w.shape[1]) for w in sep["width"]]
    out["av_full_split_sums"] = [float(w.sum()) for w in sep["width"]]
    out["av_full_depth"] = [float(d) for d in sep["depth"]]

    # 3. masks / gates / snr on seeded inputs, from the reference functions
    g = torch.Generator().manual_seed(7)
    x = torch.rand(3, 32, generator=g)
    out["hard_concrete_in"], out["hard_concrete_out"] = x, ref.estimation.hard_concrete(x)
    wg = ref.gates.WidthGate(4)
    wg.set_structure_value(torch.tensor([[1.0, 0.0, 1.0, 0.0]]))
    t4 = torch.randn(2, 8, 3, 3, generator=g)
    out["width_gate_in"], out["width_gate_out"] = t4, wg(t4)
    lg = ref.gates.LinearWidthGate(4)
    lg.set_structure_value(torch.tensor([[0.0, 1.0, 1.0, 0.0]]))
    t3 = torch.randn(2, 5, 8, generator=g)
    out["linear_gate_in"], out["linear_gate_out"] = t3, lg(t3)
    dg = ref.gates.DepthGate(1)
    dg.set_structure_value(torch.tensor([0.25]))
    a4, b4 = torch.randn(2, 3, 2, 2, generator=g), torch.randn(2, 3, 2, 2, generator=g)
    out["depth_gate_in"], out["depth_gate_out"] = (a4, b4), dg((a4, b4))

    class _S:
        pass

    s = _S()
    s.alphas_cumprod = torch.cumprod(1 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2, 0)
    ts = torch.tensor([0, 1, 17, 250, 500, 998, 999])
    out["snr_timesteps"], out["snr"] = ts, ref.metric.compute_snr(s, ts)

    # 4. tiny pruned networks: reference model outputs, pruned shapes and weight digests
    cases = {"r055": dict(ratio=0.55, seed=11, drop=()), "r070_drop": dict(ratio=0.7, seed=12, drop=(0, 1, 5, 6, 9, 12, 13))}
    gi = torch.Generator().manual_seed(99)
    sample = torch.randn(2, 4, 16, 16, generator=gi)
    tsteps = torch.tensor([37, 801])
    ctx = torch.randn(2, 5, TINY["cross_attention_dim"], generator=gi)
    out["tiny_inputs"] = dict(sample=sample, timesteps=tsteps, ctx=ctx)
    for name, c in cases.items():
        probe = P.UNetGated(**TINY)
        av = make_arch_vector(probe.get_structure(), c["ratio"], c["seed"], c["drop"])
        m = ref_pruned_model(ref, TINY, av, seed=3)
        feats = {}
        P.cast_block_act_hooks(m, feats)
        with torch.no_grad():
            y = m(sample, tsteps, ctx).sample
        out[f"tiny_{name}"] = dict(
            arch_vector=av, sample=y, feats={k: v.clone() for k, v in feats.items()},
            shapes={k: list(v.shape) for k, v in m.state_dict().items()},
            digests={k: tensor_digest(v) for k, v in m.state_dict().items()},
            n_params=sum(p.numel() for p in m.parameters()))
        print(name, "params", out[f"tiny_{name}"]["n_params"], "out abs max", float(y.abs().max()))

    # 5. head_dim-64 network (B200 path vs the reference's own pruned model); outputs only, weights are a function of
    #    deterministic_fill(seed=3)
    gi = torch.Generator().manual_seed(77)
    sample64 = torch.randn(2, 4, 16, 16, generator=gi)
    ctx64 = torch.randn(2, 77, SMALL64["cross_attention_dim"], generator=gi)
    t64 = torch.tensor([123, 940])
    out["small64_inputs"] = dict(sample=sample64, timesteps=t64, ctx=ctx64)
    for name, c in {"r055": dict(ratio=0.55, seed=21, drop=()), "r082_drop": dict(ratio=0.82, seed=22, drop=(1, 2, 8, 11))}.items():
        probe = P.UNetGated(**SMALL64)
        av = make_arch_vector(probe.get_structure(), c["ratio"], c["seed"], c["drop"])
        m = ref_pruned_model(ref, SMALL64, av, seed=3)
        feats = {}
        P.cast_block_act_hooks(m, feats)
        with torch.no_grad():
            y = m(sample64, t64, ctx64).sample
        out[f"small64_{name}"] = dict(arch_vector=av, sample=y, feats={k: v.clone() for k, v in feats.items()},
                                      shapes={k: list(v.shape) for k, v in m.state_dict().items()},
                                      n_params=sum(p.numel() for p in m.parameters()))
        print("small64", name, "params", out[f"small64_{name}"]["n_params"], "out abs max", float(y.abs().max()))

    torch.save(out, os.path.join(GOLD, "reference_golden.pt"))
    meta = {"generated_by": "oracle/make_golden.py", "reference": "rezashkv/unlearn-ft @ /root/reference",
            "torch": torch.__version__, "keys": sorted(out.keys())}
    with open(os.path.join(GOLD, "reference_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(GOLD, "reference_golden.pt"), os.path.getsize(os.path.join(GOLD, "reference_golden.pt")))


if __name__ == "__main__":
    main()
