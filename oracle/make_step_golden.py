"""TEST INFRASTRUCTURE (builder container only): golden values of the reference TRAINER's own `step()` / `upper_step()`.

    python -m oracle.make_step_golden       # needs /root/reference; writes tests/golden/reference_step_golden.pt

The two method bodies are lifted out of /root/reference/pdm/training/trainer.py by name (oracle.live_check.
_reference_trainer_methods: importing that file needs accelerate / diffusers pipelines / wandb) and run on a stub `self` whose
student is the reference's OWN pruned U-Net (pdm/models/unet/*.py through oracle/refshim) at the SMALL64 size the GPU parity
tests use.  The noise and timesteps the reference draws inside `step()` (trainer.py:2409,2421) are recorded next to the
loss terms, so the oracle (CPU, tests/test_oracle_golden.py) and the CUDA path (tests/test_unet_gpu.py) can be fed the same
sample on a box where /root/reference does not exist.
"""
from __future__ import annotations

import os
from types import SimpleNamespace as NS

import torch

from oracle import diffusers_restated as D
from oracle import pdm_restated as P
from oracle import refshim
from oracle.live_check import _reference_trainer_methods
from oracle.make_golden import GOLD, SMALL64, deterministic_fill, make_arch_vector, ref_pruned_model

STUDENT_SEED, TEACHER_SEED, AV = 3, 5, dict(ratio=0.55, seed=21, drop=(2,))     # = __graft_entry__.smoke()'s networks


def teacher_model():
    t = D.UNet2DConditionModel(**{**D.SD21_UNET_CONFIG, "block_out_channels": SMALL64["block_out_channels"],
                                  "attention_head_dim": SMALL64["heads"],
                                  "cross_attention_dim": SMALL64["cross_attention_dim"]}).eval()
    deterministic_fill(t, TEACHER_SEED)
    return t.requires_grad_(False)


def main():
    ref = refshim.load_reference()
    fns = _reference_trainer_methods({("UnetFineTuner", "step"), ("BilevelUnetFineTuner", "upper_step"),
                                      ("Trainer", "cast_block_act_hooks")})

    class Sched(D.DDIMSchedulerLite):
        def register_to_config(self, **kw):
            for k, v in kw.items():
                setattr(self.config, k, v)

    av = make_arch_vector(P.UNetGated(**SMALL64).get_structure(), AV["ratio"], AV["seed"], AV["drop"])
    rm = ref_pruned_model(ref, SMALL64, av, STUDENT_SEED)
    teacher = teacher_model()
    g = torch.Generator().manual_seed(2025)
    latents = torch.randn(2, 4, 16, 16, generator=g)
    ehs = torch.randn(2, 77, SMALL64["cross_attention_dim"], generator=g)
    empty = torch.randn(1, 77, SMALL64["cross_attention_dim"], generator=g).expand(2, -1, -1).contiguous()
    losses_cfg = NS(diffusion_loss=NS(snr_gamma=5.0, weight=1.0), block_loss=NS(weight=0.1, upper_weight=0.0),
                    distillation_loss=NS(weight=2.0, upper_weight=1.0))
    me = NS(vae=NS(encode=lambda x: NS(latent_dist=NS(sample=lambda: x)), config=NS(scaling_factor=1.0)),
            weight_dtype=torch.float32, accelerator=NS(device=torch.device("cpu"), unwrap_model=lambda m: m),
            config=NS(model=NS(prediction_model=NS(noise_offset=0, input_perturbation=0, max_scheduler_steps=None,
                                                   prediction_type="v_prediction")),
                      training=NS(losses=losses_cfg)),
            noise_scheduler=Sched(), teacher_model=teacher, prediction_model=rm, block_act_student={}, block_act_teacher={})
    fns[("Trainer", "cast_block_act_hooks")](me, rm, me.block_act_student)
    fns[("Trainer", "cast_block_act_hooks")](me, teacher, me.block_act_teacher)
    batch = {"pixel_values": latents, "prompt_embeds": ehs, "empty_prompt_embeds": empty}
    out = {"arch_vector": av, "student_seed": STUDENT_SEED, "teacher_seed": TEACHER_SEED, "latents": latents,
           "prompt_embeds": ehs, "empty_prompt_embeds": empty, "cases": []}
    for seed in (501, 502):
        torch.manual_seed(seed)
        noise = torch.randn_like(latents)                               # what step() is about to draw (:2409, :2421)
        timesteps = torch.randint(0, 1000, (latents.shape[0],)).long()
        torch.manual_seed(seed)
        with torch.no_grad():
            step = [float(v) for v in fns[("UnetFineTuner", "step")](me, batch)]
        torch.manual_seed(seed)
        with torch.no_grad():
            upper = [float(v) for v in fns[("BilevelUnetFineTuner", "upper_step")](me, batch)]
        out["cases"].append({"rng_seed": seed, "noise": noise, "timesteps": timesteps, "step": step, "upper_step": upper})
        print(seed, timesteps.tolist(), "step", step, "upper", upper)
    path = os.path.join(GOLD, "reference_step_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
