"""Classifier-free-guidance sampling loop of the reference pipeline (`generate_samples`,
pdm/pipelines/pruning_pipelines.py:867-1010, output_type="latent"; SURVEY 8f-1, BASELINE config 5) on the B200 path.

Per denoising step: one forward of the (pruned) U-Net at batch 2N (unconditional + text halves, :959-969) followed by ONE
fused kernel that does the guidance combine (:972-974) and the DDIM update (:981, diffusers DDIMScheduler.step with the
SD-2.1 scheduler config: eta 0, v-prediction, "leading" timesteps with steps_offset 1, set_alpha_to_one False), refreshes the
duplicated latent batch and advances the device-side timestep tensor.  Nothing step-dependent stays on the host, so the
launch sequence of one step (~700 kernels) is captured once into a CUDA graph and replayed `num_inference_steps` times.
VAE decode / text encoding are outside this path ("U-Net only").
"""
from __future__ import annotations

import torch

from ... import kernels as K


class CFGSampler:
    def __init__(self, unet, num_inference_steps: int = 50, guidance_scale: float = 7.5, num_train_timesteps: int = 1000,
                 beta_start: float = 0.00085, beta_end: float = 0.012, use_cuda_graph: bool = True, scheduler: str = "ddim"):
        """scheduler: "ddim" (DDIMScheduler, what the trainers' pipelines use) or "pndm" (PNDMScheduler with skip_prk_steps,
        what scripts/metrics/generate_fid_images.py:113 loads: num_inference_steps + 1 U-Net evaluations)."""
        if guidance_scale <= 1.0:
            raise ValueError("CFGSampler implements the guided branch of the pipeline (guidance_scale > 1)")
        if scheduler not in ("ddim", "pndm"):
            raise ValueError("scheduler must be 'ddim' or 'pndm'")
        self.scheduler = scheduler
        self.unet = unet
        self.device = unet.device
        self.steps, self.guidance, self.T = int(num_inference_steps), float(guidance_scale), int(num_train_timesteps)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, self.T, dtype=torch.float32) ** 2     # scaled_linear
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).to(self.device)
        ratio = self.T // self.steps
        base = (torch.arange(0, self.steps) * ratio).round().to(torch.int64) + 1           # "leading" spacing, steps_offset 1
        if scheduler == "pndm":      # PNDMScheduler.set_timesteps, skip_prk_steps: [..., t_{N-2}, t_{N-2}, t_{N-1}] reversed
            base = torch.cat([base[:-1], base[-2:-1], base[-1:]])
        self.timesteps = base.flip(0).contiguous().to(self.device)
        self.evals = int(self.timesteps.numel())                                          # U-Net evaluations per sample() call
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._shape = None

    # ------------------------------------------------------------------------------------------------------------
    def _alloc(self, n, c, h, w, ctx):
        dev = self.device
        self._lat = torch.empty(n, c, h, w, device=dev, dtype=torch.float32)
        self._lat_in = torch.empty(2 * n, c, h, w, device=dev, dtype=torch.float32)
        self._t_dev = torch.empty(2 * n, device=dev, dtype=torch.int64)
        self._state = torch.zeros(4, device=dev, dtype=torch.int32)
        if self.scheduler == "pndm":
            self._ets = torch.zeros(4, n, c, h, w, device=dev, dtype=torch.float32)
            self._cur = torch.zeros(n, c, h, w, device=dev, dtype=torch.float32)
        self._ctx = torch.empty(2 * n, *ctx, device=dev, dtype=torch.bfloat16)
        self._shape = (n, c, h, w, tuple(ctx))
        self._graph = None

    def _one_step(self):
        out = self.unet(self._lat_in, self._t_dev, self._ctx).sample                          # fp32 [2N, C, H, W]
        if self.scheduler == "pndm":
            K.cfg_pndm_step(out, self._lat, self._lat_in, self.alphas_cumprod, self.timesteps, self._state, self._t_dev,
                            self._ets, self._cur, self.steps, self.T, self.guidance)
        else:
            K.cfg_ddim_step(out, self._lat, self._lat_in, self.alphas_cumprod, self.timesteps, self._state, self._t_dev,
                            self.steps, self.T, self.guidance)

    def _reset(self, latents, prompt_embeds, negative_prompt_embeds):
        n = latents.shape[0]
        self._lat.copy_(latents, non_blocking=True)
        self._lat_in[:n].copy_(latents, non_blocking=True)
        self._lat_in[n:].copy_(latents, non_blocking=True)
        self._ctx[:n].copy_(negative_prompt_embeds, non_blocking=True)                        # unconditional half first (:931)
        self._ctx[n:].copy_(prompt_embeds, non_blocking=True)
        self._t_dev.fill_(int(self.timesteps[0].item()) if self._t0 is None else self._t0)
        self._state.zero_()

    _t0 = None

    @torch.no_grad()
    def sample(self, latents, prompt_embeds, negative_prompt_embeds):
        """latents: [N, 4, H, W] initial noise (init_noise_sigma = 1); prompt / negative embeds: [N, 77, 1024].
        Returns the final latents [N, 4, H, W] fp32 (a graph-owned buffer when CUDA graphs are on: clone to keep)."""
        n, c, h, w = latents.shape
        ctx = tuple(prompt_embeds.shape[1:])
        if self._shape != (n, c, h, w, ctx):
            self._alloc(n, c, h, w, ctx)
        if self._t0 is None:
            self._t0 = int(self.timesteps[0].item())
        self._reset(latents, prompt_embeds, negative_prompt_embeds)
        if not self.use_cuda_graph:
            for _ in range(self.evals):
                self._one_step()
            return self._lat
        if self._graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._one_step()                                                                # warm-up (scratch, pools)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step()
            self._graph = g
            self._reset(latents, prompt_embeds, negative_prompt_embeds)                         # the warm-up consumed a step
        for _ in range(self.evals):
            self._graph.replay()
        return self._lat
