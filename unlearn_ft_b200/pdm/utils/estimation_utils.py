"""Mirror of reference pdm/utils/estimation_utils.py (hot-path subset)."""
import torch


def hard_concrete(out: torch.Tensor) -> torch.Tensor:
    """Reference estimation_utils.py:67-75: 1 where out >= 0.5 else 0 (straight-through estimator for autograd)."""
    hard = (out >= 0.5).to(torch.float32).to(out.device)
    return (hard - out).detach() + out


def keep_indices(gate_f: torch.Tensor) -> list:
    """Ascending indices of the surviving gate entries of a [1, width] gate (boolean-mask selection order of the
    reference's prune() methods, blocks.py:64-72,169-177,444-473). Integer work: exact."""
    assert gate_f.dim() == 2 and gate_f.shape[0] == 1, "Pruning is only supported for single batch size"
    return torch.nonzero(gate_f[0] >= 0.5).flatten().tolist()
