"""Mirror of the fine-tune-time classmethods of reference pdm/models/hypernet.py (the hypernetwork itself belongs to
the pruning phase and is out of scope)."""
import torch


class HyperStructure:
    @classmethod
    def transform_arch_vector(cls, inputs, structure, force_width_non_zero=False):
        """Reference hypernet.py:101-126: flat [1, sum(widths)+sum(depths)] -> {'width': [...], 'depth': [...]}."""
        width_list = [w for sub in structure["width"] for w in sub]
        depth_list = [d for sub in structure["depth"] for d in sub]
        assert inputs.shape[1] == (sum(width_list) + sum(depth_list))
        width_vectors, depth_vectors = inputs[:, :sum(width_list)], inputs[:, sum(width_list):]
        w_list, start = [], 0
        for w in width_list:
            sub = width_vectors[:, start:start + w]
            if force_width_non_zero and not (sub >= 0.5).any(dim=1).all():
                sub = sub.clone()
                ind = ~(sub >= 0.5).any(dim=1)
                sub[ind, 0] = sub[ind, 0] + 0.5
            w_list.append(sub)
            start += w
        d_list = [depth_vectors[:, i] for i in range(sum(depth_list))]
        return {"width": w_list, "depth": d_list}

    @classmethod
    def get_random_arch_vector(cls, target_ratio, structure):
        """Reference hypernet.py:129-150 (same RNG consumption: one torch.randperm per width gate, in order)."""
        width_list = [w for sub in structure["width"] for w in sub]
        depth_list = [d for sub in structure["depth"] for d in sub]
        parts = []
        for w in width_list:
            v = torch.zeros(1, w)
            v[0, torch.randperm(w)[:int(target_ratio * w)]] = 0.9
            parts.append(v)
        parts += [torch.tensor([[0.9]]) for _ in range(sum(depth_list))]
        return torch.cat(parts, dim=1)
