"""Signature mirror of reference pdm/losses/contrastive_loss.py:5-22.  Pruning-phase [B, B] similarity matching (not
on the fine-tuning hot path) -- host-side torch on tiny matrices, no kernel."""
import torch.nn.functional as F
from torch import nn


class ContrastiveLoss(nn.Module):
    def __init__(self, arch_vector_temperature=1.0, prompt_embedding_temperature=1.0):
        super().__init__()
        self.arch_vector_temperature = arch_vector_temperature
        self.prompt_embedding_temperature = prompt_embedding_temperature

    @staticmethod
    def _self_similarity(x, temperature):
        x = x / x.norm(dim=1, keepdim=True)
        return F.softmax((x @ x.T) / temperature, dim=-1)

    def forward(self, prompt_embeddings, arch_vectors, return_similarity=False):
        arch_sim = self._self_similarity(arch_vectors, self.arch_vector_temperature)
        text_sim = self._self_similarity(prompt_embeddings, self.prompt_embedding_temperature)
        loss = F.binary_cross_entropy(arch_sim.T, text_sim.T, reduction="mean")
        if return_similarity:
            return loss, arch_sim.detach().float().cpu().numpy()
        return loss
