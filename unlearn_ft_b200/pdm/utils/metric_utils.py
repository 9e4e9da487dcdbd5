"""Mirror of reference pdm/utils/metric_utils.py."""
import torch


def compute_snr(noise_scheduler, timesteps: torch.Tensor) -> torch.Tensor:
    """Reference metric_utils.py:3-26: snr_t = (sqrt(acp_t) / sqrt(1 - acp_t))^2, a gather over a 1000-entry table
    (host-side plumbing; the weights it feeds are consumed by the fused loss kernel)."""
    acp = noise_scheduler.alphas_cumprod
    a = (acp ** 0.5).to(device=timesteps.device)[timesteps].float()
    s = ((1.0 - acp) ** 0.5).to(device=timesteps.device)[timesteps].float()
    return (a / s) ** 2
