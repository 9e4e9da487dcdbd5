"""Signature mirror of reference pdm/losses/resource_loss.py:5-23.  Pruning-phase scalar arithmetic (not on the
fine-tuning hot path) -- kept as host-side torch scalars, no kernel."""
import torch
from torch import nn


class ResourceLoss(nn.Module):
    def __init__(self, p=0.9, loss_type="log"):
        super().__init__()
        if loss_type not in ("log", "mae", "mse"):
            raise AssertionError(f"Unknown loss type {loss_type}")
        self.p, self.loss_type = p, loss_type

    def forward(self, resource_ratio):
        if self.loss_type == "mae":
            return torch.abs(resource_ratio - self.p)
        if self.loss_type == "mse":
            return (resource_ratio - self.p) ** 2
        hi, lo = (resource_ratio, self.p) if resource_ratio > self.p else (self.p, resource_ratio)
        return torch.log(hi / lo)
