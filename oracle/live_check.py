"""TEST INFRASTRUCTURE (builder container only): live cross-check of the oracle against the REFERENCE'S OWN CODE.

    python -m oracle.live_check            # needs /root/reference; exits 0 when every check passes

Unlike tests/golden (recorded once), this re-runs the reference's gated blocks / prune() / forward (through oracle/refshim)
on arch vectors and seeds that are NOT in the golden file, and additionally compares GRADIENTS: d(loss)/d(parameter) of the
reference's forward under torch autograd against the oracle's, for every parameter of the pruned network.  It also executes
the reference TRAINER's own `step()` and `upper_step()` source (lifted out of pdm/training/trainer.py by name, see
`_reference_trainer_methods`) against the oracle's `finetune_step` / `upper_step`: loss terms and gradients; the reference's
gated (un-pruned) forward; and the reference pipeline's own `generate_samples` CFG loop against `cfg_sample_loop`.
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
    worst_gated = check_gated_mode(ref, sample, tsteps, ctx)
    worst_step = check_training_steps(ref)
    worst_loop = check_sampling_loop(ref)
    print(f"live check ok: {len(CASES)} pruned networks, worst output error {worst_out:.2e}, worst gradient error "
          f"{worst_grad:.2e}; gated (un-pruned, multiplicative gates) forward: {worst_gated:.2e}; step()/upper_step() of the "
          f"reference trainer vs oracle: worst loss-term error {worst_step:.2e}; generate_samples() CFG loop of the reference "
          f"pipeline vs oracle: {worst_loop:.2e}")
    return 0


def check_gated_mode(ref, sample, tsteps, ctx) -> float:
    """The reference's OTHER code path for the same arch vector: gates applied multiplicatively at run time, nothing sliced
    (`pruned=False`; blocks.py:56-58,267-272,343-348,582-587 -- what the pruning phase trains with).  Reference gated model
    vs the oracle's gated model, structure set, prune() NOT called."""
    from oracle.make_golden import REF_BLOCKS
    worst = 0.0
    for case in CASES:
        orc = P.UNetGated(**TINY)
        deterministic_fill(orc, 13)
        av = make_arch_vector(orc.get_structure(), case["ratio"], case["seed"] + 100, case["drop"])
        orc.set_structure(P.transform_arch_vector(av, orc.get_structure()))
        orc.eval()
        rm = ref.unet.UNet2DConditionModelGated(sample_size=16, block_out_channels=TINY["block_out_channels"],
                                                attention_head_dim=TINY["heads"],
                                                cross_attention_dim=TINY["cross_attention_dim"], use_linear_projection=True,
                                                norm_eps=1e-5, gated_ff=True, ff_gate_width=32, **REF_BLOCKS)
        deterministic_fill(rm, 13)
        rm.set_structure(ref.hypernet.HyperStructure.transform_arch_vector(av, rm.get_structure()))
        rm.eval()
        with torch.no_grad():
            y_o = orc(sample, tsteps, ctx).sample
            y_r = rm(sample, tsteps, ctx).sample
        err = float((y_o - y_r).abs().max() / y_r.abs().max())
        worst = max(worst, err)
        assert err < 1e-5, ("gated forward", case, err)
    return worst


def check_sampling_loop(ref) -> float:
    """The reference pipeline's OWN `generate_samples` (pdm/pipelines/pruning_pipelines.py:867-1010), lifted out of the file by
    name like the trainer methods, on a stub `self` (reference pruned U-Net; restated DDIM scheduler; given text states and
    initial latents; output_type="latent") against oracle.pdm_restated.cfg_sample_loop: pins the classifier-free-guidance
    loop -- batch doubling, scalar-timestep U-Net call with return_dict=False, chunk + guidance combine, scheduler hand-off."""
    import ast
    import contextlib
    import os
    from types import SimpleNamespace as NS
    from typing import Any, Callable, Dict, List, Optional, Union

    from oracle import diffusers_restated as D

    path = os.path.join(refshim.REFERENCE_ROOT, "pdm", "pipelines", "pruning_pipelines.py")
    tree = ast.parse(open(path).read())
    fn = None
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for f in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            if f.name == "generate_samples" and fn is None:
                fn = f
    assert fn is not None, "generate_samples not found in the reference pipelines"
    fn.decorator_list = []
    ns = {"torch": torch, "Union": Union, "List": List, "Optional": Optional, "Callable": Callable, "Dict": Dict, "Any": Any,
          "StableDiffusionPipelineOutput": lambda images, nsfw_content_detected: NS(images=images),
          "rescale_noise_cfg": None}
    exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), path, "exec"), ns)
    generate_samples = ns["generate_samples"]

    class Sched(D.DDIMSchedulerLite):
        order = 1

        def step(self, model_output, timestep, sample, return_dict=False, **kw):       # diffusers returns a tuple
            return (super().step(model_output, timestep, sample, **kw),)

    orc = P.UNetGated(**TINY)
    deterministic_fill(orc, 8)
    av = make_arch_vector(orc.get_structure(), 0.7, 51, (4,))
    orc.set_structure(P.transform_arch_vector(av, orc.get_structure()))
    orc.prune()
    orc.eval()
    rm = ref_pruned_model(ref, TINY, av, 8)
    g = torch.Generator().manual_seed(5)
    lat0 = torch.randn(2, 4, 16, 16, generator=g)
    pe = torch.randn(2, 77, TINY["cross_attention_dim"], generator=g)
    ne = torch.randn(1, 77, TINY["cross_attention_dim"], generator=g).expand(2, -1, -1).contiguous()
    sched = Sched()
    me = NS(unet=rm, vae_scale_factor=8, check_inputs=lambda *a, **k: None, _execution_device=torch.device("cpu"),
            encode_prompt=lambda prompt, device, n, cfg, neg, prompt_embeds=None, negative_prompt_embeds=None, lora_scale=None:
            (prompt_embeds, negative_prompt_embeds),
            scheduler=sched,
            prepare_latents=lambda b, c, h, w, dtype, device, generator, latents: latents * sched.init_noise_sigma,
            prepare_extra_step_kwargs=lambda generator, eta: {},
            progress_bar=lambda total=None: contextlib.nullcontext(NS(update=lambda: None)),
            image_processor=NS(postprocess=lambda image, output_type=None, do_denormalize=None: image),
            maybe_free_model_hooks=lambda: None, vae=None)
    worst = 0.0
    for steps, scale in ((4, 7.5), (5, 1.5)):
        with torch.no_grad():
            out_r = generate_samples(me, prompt=None, num_inference_steps=steps, guidance_scale=scale, latents=lat0.clone(),
                                     prompt_embeds=pe, negative_prompt_embeds=ne, output_type="latent").images
            out_o = P.cfg_sample_loop(orc, D.DDIMSchedulerLite(), lat0.clone(), pe, ne, num_inference_steps=steps,
                                      guidance_scale=scale)
        err = float((out_o - out_r).abs().max() / out_r.abs().max())
        worst = max(worst, err)
        assert err < 1e-5, ("cfg loop", steps, scale, err)
    return worst


def _reference_trainer_methods(names):
    """The reference's OWN `step`, `upper_step`, `cast_block_act_hooks` (pdm/training/trainer.py:557-572, 2403-2488,
    2904-3001), lifted out of the file by name with `ast` and compiled as plain functions: importing trainer.py would pull in
    accelerate / diffusers pipelines / wandb, none of which are installed, but these three bodies only need torch, F and
    compute_snr."""
    import ast
    import os
    import torch.nn.functional as F_

    path = os.path.join(refshim.REFERENCE_ROOT, "pdm", "training", "trainer.py")
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            key = (cls.name, fn.name)
            if key in names:
                fn.decorator_list = []
                mod = ast.Module(body=[fn], type_ignores=[])
                ns = {"torch": torch, "F": F_, "compute_snr": refshim.load_reference().metric.compute_snr}
                exec(compile(ast.fix_missing_locations(mod), path, "exec"), ns)
                found[key] = ns[fn.name]
    missing = set(names) - set(found)
    assert not missing, f"not found in the reference trainer: {missing}"
    return found


def check_training_steps(ref) -> float:
    """Reference trainer `step()` / `upper_step()` (their own source, see above) on a stub `self` whose models are the
    reference's own pruned U-Net + the teacher, against oracle.pdm_restated.finetune_step / upper_step on the same inputs:
    the four loss terms and every parameter gradient."""
    from types import SimpleNamespace as NS

    from oracle import diffusers_restated as D

    fns = _reference_trainer_methods({("UnetFineTuner", "step"), ("BilevelUnetFineTuner", "upper_step"),
                                      ("Trainer", "cast_block_act_hooks")})
    ref_step, ref_upper = fns[("UnetFineTuner", "step")], fns[("BilevelUnetFineTuner", "upper_step")]
    ref_hooks = fns[("Trainer", "cast_block_act_hooks")]

    class Sched(D.DDIMSchedulerLite):
        def register_to_config(self, **kw):                      # diffusers ConfigMixin.register_to_config
            for k, v in kw.items():
                setattr(self.config, k, v)

    orc = P.UNetGated(**TINY)
    deterministic_fill(orc, 4)
    av = make_arch_vector(orc.get_structure(), 0.6, 41, (1,))
    orc.set_structure(P.transform_arch_vector(av, orc.get_structure()))
    orc.prune()
    orc.eval()
    rm = ref_pruned_model(ref, TINY, av, 4)
    teacher = D.UNet2DConditionModel(**{**D.SD21_UNET_CONFIG, "block_out_channels": TINY["block_out_channels"],
                                        "attention_head_dim": TINY["heads"],
                                        "cross_attention_dim": TINY["cross_attention_dim"]}).eval()
    deterministic_fill(teacher, 6)
    teacher.requires_grad_(False)

    g = torch.Generator().manual_seed(77)
    latents = torch.randn(2, 4, 16, 16, generator=g)
    ehs = torch.randn(2, 77, TINY["cross_attention_dim"], generator=g)
    empty = torch.randn(1, 77, TINY["cross_attention_dim"], generator=g).expand(2, -1, -1).contiguous()
    losses_cfg = NS(diffusion_loss=NS(snr_gamma=5.0, weight=1.0), block_loss=NS(weight=0.1, upper_weight=0.0),
                    distillation_loss=NS(weight=2.0, upper_weight=1.0))      # the shipped YAML weights
    me = NS(vae=NS(encode=lambda x: NS(latent_dist=NS(sample=lambda: x)), config=NS(scaling_factor=1.0)),
            weight_dtype=torch.float32, accelerator=NS(device=torch.device("cpu"), unwrap_model=lambda m: m),
            config=NS(model=NS(prediction_model=NS(noise_offset=0, input_perturbation=0, max_scheduler_steps=None,
                                                   prediction_type="v_prediction")),
                      training=NS(losses=losses_cfg)),
            noise_scheduler=Sched(), teacher_model=teacher, prediction_model=rm, block_act_student={}, block_act_teacher={})
    ref_hooks(me, rm, me.block_act_student)
    ref_hooks(me, teacher, me.block_act_teacher)
    fs, ft = {}, {}
    P.cast_block_act_hooks(orc, fs)
    P.cast_block_act_hooks(teacher, ft)
    batch = {"pixel_values": latents, "prompt_embeds": ehs, "empty_prompt_embeds": empty}
    worst = 0.0

    def grads_equal(tag):
        nonlocal worst
        po, pr = dict(orc.named_parameters()), dict(rm.named_parameters())
        compared = 0
        for k, p in pr.items():
            if p.grad is None or float(p.grad.abs().max()) == 0.0:
                continue
            e = float((po[k].grad - p.grad).abs().max() / p.grad.abs().max())
            worst = max(worst, e)
            assert e < 1e-4, (tag, k, e)
            compared += 1
        assert compared > 0.8 * len(pr), (tag, compared, len(pr))   # (a depth-dropped block's parameters get no gradient)
        for p in list(po.values()) + list(pr.values()):
            p.grad = None

    for seed in (123, 124):
        # lower step: the reference draws noise and timesteps itself (trainer.py:2409,2421); same RNG calls, same order
        torch.manual_seed(seed)
        out_r = ref_step(me, batch)
        torch.manual_seed(seed)
        noise = torch.randn_like(latents)
        timesteps = torch.randint(0, 1000, (latents.shape[0],)).long()
        out_o = P.finetune_step(orc, teacher, D.DDIMSchedulerLite(), latents, noise, timesteps, ehs, fs, ft)
        for a, b, name in zip(out_o, out_r, ("loss", "diff_loss", "distillation_loss", "block_loss")):
            e = abs(float(a.detach()) - float(b.detach())) / max(abs(float(b.detach())), 1e-12)
            worst = max(worst, e)
            assert e < 1e-6, ("step", name, float(a), float(b))
        out_o[0].backward(), out_r[0].backward()
        grads_equal("step")
        # upper (concept-suppression) step
        torch.manual_seed(seed)
        up_r = ref_upper(me, batch)
        torch.manual_seed(seed)
        noise = torch.randn_like(latents)
        timesteps = torch.randint(0, 1000, (latents.shape[0],)).long()
        up_o = P.upper_step(orc, teacher, D.DDIMSchedulerLite(), latents, noise, timesteps, ehs, empty)
        e = abs(float(up_o.detach()) - float(up_r[0].detach())) / abs(float(up_r[0].detach()))
        worst = max(worst, e)
        assert e < 1e-6 and float(up_r[1]) == 0.0 and float(up_r[3]) == 0.0, ("upper_step", float(up_o), [float(v) for v in up_r])
        up_o.backward(), up_r[0].backward()
        grads_equal("upper_step")
    return worst


if __name__ == "__main__":
    sys.exit(main())
