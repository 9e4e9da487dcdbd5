"""B200-native mirror of ``pdm/models/unet/unet_2d_conditional.py`` (reference): the APTP-pruned SD-2.1 U-Net.

Drop-in surface kept:
  * ``UNet2DConditionModelPruned.from_pretrained(path, subfolder=..., down_block_types=..., mid_block_type=...,
    up_block_types=..., gated_ff=..., ff_gate_width=..., arch_vector=Tensor[1, 1620], random_init=...,
    random_pruning_ratio=..., checkpoint_loading=...)``  (reference :2185-2495; kwargs :2304-2310)
  * ``from_config`` / ``.config`` / ``get_structure()`` (reference :1334-1365) / ``state_dict()`` with diffusers keys
  * ``model(sample[B,4,H,W], timestep[B], encoder_hidden_states[B,77,1024], return_dict=True).sample`` (:1417-1728)
  * ``down_blocks`` / ``mid_block`` / ``up_blocks`` are hook-able ``nn.Module``s returning
    ``(hidden, residual_tuple)`` / tensor / tensor (trainer.py:557-572)

Differences by design: the model is built directly at its pruned widths (no build-full-then-slice pass), parameters
live in one flat arena (fp32 masters + bf16 shadows), activations are bf16 channels-last, and all arithmetic runs in
``libb200pdm.so``.  Construction requires a CUDA device: there is no CPU fallback.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
from torch import nn

from .... import kernels as K
from ... import nn as bnn
from ...nn import ParamArena, PConv2d, PGroupNorm, PLinear, as2d, to4d
from ...utils.estimation_utils import keep_indices
from ..hypernet import HyperStructure
from .blocks import (CrossAttnDownBlock2DWidthHalfDepthGated, CrossAttnUpBlock2DWidthHalfDepthGated,
                     DownBlock2DWidthHalfDepthGated, ResnetBlock2DWidthGated, RuntimeGates, Transformer2DModelWidthGated,
                     UNetMidBlock2DCrossAttnWidthGated, UpBlock2DWidthHalfDepthGated, run_block)

BF16, F32 = torch.bfloat16, torch.float32

# stabilityai/stable-diffusion-2-1 unet/config.json merged with the reference's gated block types
# (configs/baselines/sd-2-1_coco_aptp_both_512.yaml:11-26)
SD21_CONFIG = dict(
    sample_size=96, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2DHalfGated", "CrossAttnDownBlock2DHalfGated", "CrossAttnDownBlock2DHalfGated",
                      "DownBlock2DHalfGated"),
    mid_block_type="UNetMidBlock2DCrossAttnWidthGated",
    up_block_types=("UpBlock2DHalfGated", "CrossAttnUpBlock2DHalfGated", "CrossAttnUpBlock2DHalfGated",
                    "CrossAttnUpBlock2DHalfGated"),
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
    cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20), use_linear_projection=True, gated_ff=True,
    ff_gate_width=32, prediction_type="v_prediction",
)
_DOWN_TYPES = {"CrossAttnDownBlock2DHalfGated": True, "DownBlock2DHalfGated": False,
               "CrossAttnDownBlock2D": True, "DownBlock2D": False}
_UP_TYPES = {"CrossAttnUpBlock2DHalfGated": True, "UpBlock2DHalfGated": False,
             "CrossAttnUpBlock2D": True, "UpBlock2D": False}


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor = None


class _TimestepEmbedding(nn.Module):
    """diffusers TimestepEmbedding: Linear(320,1280) -> SiLU -> Linear(1280,1280) (keys time_embedding.linear_1/2)."""

    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = PLinear(cin, dim)
        self.linear_2 = PLinear(dim, dim)


def structure_from_config(cfg: dict) -> Dict[str, list]:
    """Reference get_structure() (:1334-1365) evaluated from the config alone: per U-Net block, resnet gates first
    ([norm groups]) then transformer gates ([heads, heads, ff_gate_width]); depth [1] for the last layer of each
    down/up block (blocks.py:1573-1706,1900-2039,2187-2247,2316-2381), none in the mid block (:2450-2544)."""
    ch, heads = cfg["block_out_channels"], cfg["attention_head_dim"]
    if isinstance(heads, int):
        heads = (heads,) * len(ch)
    ffw, G = cfg.get("ff_gate_width", 32), cfg.get("norm_num_groups", 32)
    width, depth = [], []

    def block(n_layers, has_attn, h, last_depth):
        for i in range(n_layers):
            width.append([G])
            depth.append([1] if (last_depth and i == n_layers - 1) else [0])
        if has_attn:
            for i in range(n_layers):
                width.append([h, h, ffw])
                depth.append([1] if (last_depth and i == n_layers - 1) else [0])

    L = cfg.get("layers_per_block", 2)
    for i, t in enumerate(cfg["down_block_types"]):
        block(L, _DOWN_TYPES[t], heads[i], True)
    width.append([G]), depth.append([0])                      # mid: resnets (2) then attentions (1)
    width.append([G]), depth.append([0])
    width.append([heads[-1], heads[-1], ffw]), depth.append([0])
    rheads = list(reversed(heads))
    for i, t in enumerate(cfg["up_block_types"]):
        block(L + 1, _UP_TYPES[t], rheads[i], True)
    return {"width": width, "depth": depth}


class _IndexOnOwnDevice(dict):
    """state-dict view whose tensors accept host-built index tensors in index_select."""

    class _T:
        def __init__(self, t):
            self.t = t

        def index_select(self, dim, idx):
            return self.t.index_select(dim, idx.to(self.t.device))

    def __getitem__(self, k):
        return self._T(dict.__getitem__(self, k))

    def raw(self, k):
        return dict.__getitem__(self, k)


class UNet2DConditionModelGated(nn.Module):
    """See module docstring.  ``arch_vector=None`` builds the un-pruned network (all gates open)."""

    def __init__(self, config: Optional[dict] = None, arch_vector: Optional[torch.Tensor] = None, device=None,
                 trainable: bool = True, seed: Optional[int] = 0, **overrides):
        super().__init__()
        cfg = dict(SD21_CONFIG)
        cfg.update(config or {})
        cfg.update(overrides)
        self._config = cfg
        self.config = SimpleNamespace(**cfg)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise RuntimeError("unlearn_ft_b200 models need a CUDA (sm_100a) device: there is no CPU fallback "
                               "(pass device='meta' to inspect structure / state-dict shapes without a GPU)")
        self.structure = structure_from_config(cfg)
        n_total = sum(w for s in self.structure["width"] for w in s) + sum(d for s in self.structure["depth"] for d in s)
        if arch_vector is None:
            arch_vector = torch.ones(1, n_total)
        self.arch_vector = arch_vector.detach().float().cpu().clone()
        self._pending_arch = None
        self._gate_src = None
        self._materialise(device, trainable, seed)
        self.eval()

    def _materialise(self, device, trainable, seed):
        """Blocks at the widths `self.arch_vector` prescribes + their flat parameter arena."""
        self._build(self._config)
        self.arena = ParamArena(self, device, trainable=trainable, seed=seed)
        anchor = self.arena.master if trainable else None
        if trainable:
            self.arena.master.requires_grad_(True)
        for blk in list(self.down_blocks) + [self.mid_block] + list(self.up_blocks):
            blk._anchor = anchor
        self._anchor = anchor

    # ------------------------------------------------------------------------------------------------ construction
    def _build(self, cfg):
        ch = cfg["block_out_channels"]
        heads = cfg["attention_head_dim"]
        heads = (heads,) * len(ch) if isinstance(heads, int) else tuple(heads)
        ctx, eps, G, L = cfg["cross_attention_dim"], cfg["norm_eps"], cfg["norm_num_groups"], cfg["layers_per_block"]
        ffw = cfg["ff_gate_width"]
        temb = ch[0] * 4
        sep = HyperStructure.transform_arch_vector(self.arch_vector, self.structure)
        wv, dv = list(sep["width"]), list(sep["depth"])
        flat_depth = [d for s in self.structure["depth"] for d in s]
        it_w, it_d = iter(wv), iter(dv)
        it_has_depth = iter(flat_depth)

        # column of every gate inside the flat arch vector (all widths in get_structure() order, then the depth gates:
        # hypernet.py:101-126) -- what a gated module reads at run time in the un-pruned mode (blocks.RuntimeGates)
        n_width_total = sum(w for s_ in self.structure["width"] for w in s_)
        cursor = {"w": 0, "d": n_width_total}

        def next_gate():
            w = next(it_w)
            cols = (cursor["w"], int(w.shape[-1]))
            cursor["w"] += int(w.shape[-1])
            return keep_indices(w), cols

        def next_depth():
            has = next(it_has_depth)
            if not has:
                return False, False, None
            d = next(it_d)
            col = cursor["d"]
            cursor["d"] += 1
            return True, bool(float(d[0]) < 0.5), col      # hard_concrete(depth) == 0 -> dropped (blocks.py:649)

        def make_layers(n_layers, cins, cout, has_attn, h, concat_skips=None):
            """Gates are consumed resnets-first then attentions, as in get_structure()/set_gate_structure()."""
            r_specs = [(next_gate(),) + next_depth() for _ in range(n_layers)]
            a_specs = [((next_gate(), next_gate(), next_gate()),) + next_depth() for _ in range(n_layers)] if has_attn else []
            resnets, atts = [], []
            for i, ((keep, cols), dg, drop, dcol) in enumerate(r_specs):
                skip = concat_skips[i] if concat_skips is not None else None
                resnets.append(ResnetBlock2DWidthGated(cins[i], cout, temb, keep, eps, G, depth_gated=dg, dropped=drop,
                                                       is_input_concatenated=concat_skips is not None,
                                                       skip_connection_dim=skip if dg else None, gate_cols=cols, depth_col=dcol))
            for ((k1, c1), (k2, c2), (kf, cf)), dg, drop, dcol in a_specs:
                atts.append(Transformer2DModelWidthGated(h, cout, ctx, k1, k2, kf, G, depth_gated=dg, dropped=drop,
                                                         ff_gate_width=ffw, gate_cols=(c1, c2, cf), depth_col=dcol))
            return resnets, atts

        self.conv_in = PConv2d(cfg["in_channels"], ch[0], 3)
        self.time_embedding = _TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, t in enumerate(cfg["down_block_types"]):
            in_c, out_c = out_c, ch[i]
            final = i == len(ch) - 1
            resnets, atts = make_layers(L, [in_c] + [out_c] * (L - 1), out_c, _DOWN_TYPES[t], heads[i])
            if _DOWN_TYPES[t]:
                self.down_blocks.append(CrossAttnDownBlock2DWidthHalfDepthGated(resnets, atts, out_c, not final))
            else:
                self.down_blocks.append(DownBlock2DWidthHalfDepthGated(resnets, out_c, not final))
        # mid block: resnets[0], resnets[1] gates, then the attention's (get_gate_structure order, blocks.py:2548-2565)
        (k_r0, c_r0), (k_r1, c_r1) = next_gate(), next_gate()
        next(it_has_depth), next(it_has_depth)
        (k_a1, c_a1), (k_a2, c_a2), (k_af, c_af) = next_gate(), next_gate(), next_gate()
        next(it_has_depth)
        c = ch[-1]
        self.mid_block = UNetMidBlock2DCrossAttnWidthGated(
            [ResnetBlock2DWidthGated(c, c, temb, k_r0, eps, G, gate_cols=c_r0),
             ResnetBlock2DWidthGated(c, c, temb, k_r1, eps, G, gate_cols=c_r1)],
            [Transformer2DModelWidthGated(heads[-1], c, ctx, k_a1, k_a2, k_af, G, ff_gate_width=ffw, gate_cols=(c_a1, c_a2, c_af))])
        self.up_blocks = nn.ModuleList()
        rch, rheads = list(reversed(ch)), list(reversed(heads))
        out_c = rch[0]
        n = len(ch)
        for i, t in enumerate(cfg["up_block_types"]):
            prev, out_c = out_c, rch[i]
            in_c = rch[min(i + 1, n - 1)]
            nl = L + 1
            skips = [in_c if j == nl - 1 else out_c for j in range(nl)]
            cins = [(prev if j == 0 else out_c) + skips[j] for j in range(nl)]
            resnets, atts = make_layers(nl, cins, out_c, _UP_TYPES[t], rheads[i], concat_skips=skips)
            if _UP_TYPES[t]:
                self.up_blocks.append(CrossAttnUpBlock2DWidthHalfDepthGated(resnets, atts, out_c, i < n - 1))
            else:
                self.up_blocks.append(UpBlock2DWidthHalfDepthGated(resnets, out_c, i < n - 1))
        self.conv_norm_out = PGroupNorm(G, ch[0], eps)
        self.conv_out = PConv2d(ch[0], cfg["out_channels"], 3)
        assert next(it_w, None) is None and next(it_d, None) is None, "arch vector not fully consumed"

    # ------------------------------------------------------------------------------------------------ reference API
    def get_structure(self):
        return self.structure

    def is_pruned(self) -> bool:
        return bool((self.arch_vector < 0.5).any())

    def set_structure(self, arch_vectors):
        """Reference :1367-1415 -- first half of its two-phase API (build full -> set_structure -> prune).  `arch_vectors` is
        what `HyperStructure.transform_arch_vector(arch_vector, model.get_structure())` returns ({'width': [T[1, w], ...],
        'depth': [T[1, 1], ...]}, consumed block by block in get_structure() order) or the flat [1, 1620] vector itself.
        The gate values are recorded; `prune()` materialises them, and until then a forward pass applies them as runtime
        multiplicative gates (rows = 1, or one row per sample)."""
        if torch.is_tensor(arch_vectors):
            live = arch_vectors if arch_vectors.dim() == 2 else arch_vectors.reshape(1, -1)
        else:
            widths, depths = list(arch_vectors["width"]), list(arch_vectors["depth"])
            want_w = [w for s in self.structure["width"] for w in s]
            if [int(t.shape[-1]) for t in widths] != want_w:
                raise ValueError("set_structure: width vectors do not match get_structure()")
            if len(depths) != sum(d for s in self.structure["depth"] for d in s):
                raise ValueError("set_structure: depth vectors do not match get_structure()")
            rows = widths[0].shape[0] if widths[0].dim() == 2 else 1
            live = torch.cat([t.reshape(rows, -1) for t in widths + depths], dim=1)      # depth gates arrive 1-D (hypernet.py:124)
        if live.shape[1] != self.arch_vector.shape[1]:
            raise ValueError(f"set_structure: expected {self.arch_vector.shape[1]} gate values per row, got {tuple(live.shape)}")
        self._pending_arch = live.detach().float().cpu().clone()
        # Un-pruned ("gated") mode of the pruning phase: until prune() is called the recorded values act as
        # MULTIPLICATIVE gates at run time -- per sample when one row per sample is given -- and a tensor that requires
        # grad receives d loss / d gate (what the hypernetwork is trained with, trainer.py:1159-1321).
        self._gate_src = live

    @torch.no_grad()
    def prune(self):
        """Second half of the reference's two-phase API: its `for m in model.modules(): m.prune()` / `m.prune_module()` pass
        (:2455-2461; per-module bodies blocks.py:62-76,131-138,163-196,435-475,647-702,1324-1334).  Here it is one
        operation on the model: the blocks are re-created at the widths the recorded gates keep and every weight is
        selected out of the current full-width tensors with the reference's boolean-mask (ascending index) selections.
        Idempotent once applied, so the reference's own module loop -- which reaches this method first (the root module)
        and then the leaf modules' no-op `prune()` -- can be run unchanged."""
        if self._pending_arch is None:
            return self
        self._gate_src = None
        if self._pending_arch.shape[0] != 1:
            raise ValueError("prune(): one architecture (a [1, n] arch vector) can be materialised, got per-sample gates")
        if self.is_pruned():
            raise RuntimeError("prune(): this network is already pruned (the reference prunes a full-width model once)")
        full_sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
        device, trainable = self.arena.device, self.arena.trainable
        self.arch_vector, self._pending_arch = self._pending_arch, None
        for name in ("conv_in", "time_embedding", "down_blocks", "mid_block", "up_blocks", "conv_norm_out", "conv_out"):
            delattr(self, name)
        self._materialise(device, trainable, None)
        self.load_unpruned_state_dict(full_sd)
        return self

    @property
    def device(self):
        return self.arena.device

    @property
    def dtype(self):
        return F32

    def num_parameters(self):
        return self.arena.logical_numel

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config) if not isinstance(config, SimpleNamespace) else dict(vars(config))
        return cls(cfg, **kwargs)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, **kwargs):
        """Reference :2185-2495.  Offline build: `random_init=True` (the BASELINE configs) needs no files; otherwise
        `pretrained_model_name_or_path[/subfolder]` must hold `config.json` + a full-width diffusers state dict
        (`diffusion_pytorch_model.safetensors`), which is sliced by the arch vector while loading; with
        `checkpoint_loading=True` the file holds already-pruned weights (trainer.py:314-346)."""
        subfolder = kwargs.pop("subfolder", None)
        arch_vector = kwargs.pop("arch_vector", None)
        random_init = kwargs.pop("random_init", False)
        random_pruning_ratio = kwargs.pop("random_pruning_ratio", None)
        checkpoint_loading = kwargs.pop("checkpoint_loading", False)
        for k in ("revision", "torch_dtype", "cache_dir", "variant", "use_safetensors", "low_cpu_mem_usage"):
            kwargs.pop(k, None)
        device, trainable, seed = kwargs.pop("device", None), kwargs.pop("trainable", True), kwargs.pop("seed", 0)
        cfg = {}
        root = None
        if pretrained_model_name_or_path is not None:
            root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
            cpath = os.path.join(root, "config.json")
            if os.path.exists(cpath):
                with open(cpath) as f:
                    cfg = {k: v for k, v in json.load(f).items() if k in SD21_CONFIG}
            elif not random_init:
                raise FileNotFoundError(f"{cpath} not found (no network access: weights must be local)")
            apath = os.path.join(pretrained_model_name_or_path, "arch_vector.pt")
            if arch_vector is None and os.path.exists(apath):                       # reference :2429-2441
                arch_vector = torch.load(apath, map_location="cpu")
        for k in ("down_block_types", "mid_block_type", "up_block_types", "gated_ff", "ff_gate_width"):
            if k in kwargs and kwargs[k] is not None:
                cfg[k] = kwargs.pop(k)
            else:
                kwargs.pop(k, None)
        if random_pruning_ratio is not None:                                        # reference :2444-2446
            full = dict(SD21_CONFIG)
            full.update(cfg)
            arch_vector = HyperStructure.get_random_arch_vector(random_pruning_ratio, structure_from_config(full))
        model = cls(cfg, arch_vector=arch_vector, device=device, trainable=trainable, seed=seed)
        if not random_init:
            from safetensors.torch import load_file
            sd = load_file(os.path.join(root, "diffusion_pytorch_model.safetensors"))
            if checkpoint_loading:
                model.load_state_dict(sd)
            else:
                model.load_unpruned_state_dict(sd)
        return model

    def save_pretrained(self, save_directory, safe_serialization: bool = True, save_arch_vector: bool = True, **_):
        """diffusers `ModelMixin.save_pretrained` layout, as the reference's checkpoint hook uses it (trainer.py:314-327):
        `<dir>/config.json` + `<dir>/diffusion_pytorch_model.safetensors` with the diffusers key names and the PRUNED shapes
        (fp32 master weights).  The arch vector that explains those shapes goes to `<parent>/arch_vector.pt`, where the
        reference keeps it next to the `unet` folder (trainer.py:2163,2366-2368; read back at :2429-2441), so
        `from_pretrained(parent, subfolder=basename, checkpoint_loading=True)` restores the model from these files."""
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self._config.items()}
        cfg.update({"_class_name": type(self).__name__, "_diffusers_version": "0.30.3"})
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "diffusion_pytorch_model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "diffusion_pytorch_model.bin"))
        if save_arch_vector:
            parent = os.path.dirname(os.path.abspath(save_directory))
            torch.save(self.arch_vector.clone(), os.path.join(parent, "arch_vector.pt"))

    # ------------------------------------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.arena.shadow_fresh = False
        return out

    @torch.no_grad()
    def load_unpruned_state_dict(self, full_sd: Dict[str, torch.Tensor]):
        """Slice a FULL-width diffusers state dict into this (pruned) model: the index selection of the reference's
        prune() methods (ascending surviving indices; blocks.py:64-72,133-138,169-185,444-473). Bit-exact gathers."""
        own = dict(self.named_parameters())
        used = set()
        full_sd = _IndexOnOwnDevice(full_sd)          # (index tensors are built on the host; the weights may live anywhere)
        for name, mod in self.named_modules():
            pre = name + "." if name else ""
            if isinstance(mod, ResnetBlock2DWidthGated) and not mod.dropped:
                keep = mod.keep_channels()
                for k, dim in (("conv1.weight", 0), ("conv1.bias", 0), ("time_emb_proj.weight", 0),
                               ("time_emb_proj.bias", 0), ("norm2.weight", 0), ("norm2.bias", 0), ("conv2.weight", 1)):
                    own[pre + k].copy_(full_sd[pre + k].index_select(dim, keep))
                    used.add(pre + k)
            elif hasattr(mod, "keep_heads") and hasattr(mod, "to_q"):
                hk = torch.tensor(mod.keep_heads, dtype=torch.long)
                rows = (hk[:, None] * 64 + torch.arange(64)[None, :]).reshape(-1)
                for k in ("to_q.weight", "to_k.weight", "to_v.weight"):
                    own[pre + k].copy_(full_sd[pre + k].index_select(0, rows))
                    used.add(pre + k)
                own[pre + "to_out.0.weight"].copy_(full_sd[pre + "to_out.0.weight"].index_select(1, rows))
                used.add(pre + "to_out.0.weight")
            elif hasattr(mod, "keep_units"):
                units = mod.keep_units()
                both = torch.cat([units, units + mod.inner_full])
                own[pre + "net.0.proj.weight"].copy_(full_sd[pre + "net.0.proj.weight"].index_select(0, both))
                own[pre + "net.0.proj.bias"].copy_(full_sd[pre + "net.0.proj.bias"].index_select(0, both))
                own[pre + "net.2.weight"].copy_(full_sd[pre + "net.2.weight"].index_select(1, units))
                used.update({pre + "net.0.proj.weight", pre + "net.0.proj.bias", pre + "net.2.weight"})
        for k, p in own.items():
            if k not in used:
                p.copy_(full_sd.raw(k))
        self.arena.shadow_fresh = False

    # ------------------------------------------------------------------------------------------------ forward
    def _stem(self, sample, timesteps, gate_in=None):
        B, _, H, W = sample.shape
        te = self.time_embedding
        rt = getattr(self, "_rt_live", None)

        def runner(need_bwd, sample_, t_, *gates_):
            x0 = K.nchw_f32_to_nhwc_bf16(sample_)
            h0, b_ci = bnn.conv(x0, self.conv_in, B, H, W, need_bwd)                          # reference :1616
            t_emb = K.timestep_embedding(t_, self.conv_in.out_channels)                        # reference :1514-1519
            e1, b_l1 = bnn.linear(t_emb, te.linear_1, need_bwd, out_fp32=True)                # reference :1521
            a1 = K.silu_f32_to_bf16(e1)
            emb, b_l2 = bnn.linear(a1, te.linear_2, need_bwd, out_fp32=True)
            temb_act = K.silu_f32_to_bf16(emb)     # SiLU(emb): shared by every resnet (blocks.py:334-336)
            outs = [to4d(h0, B, H, W), temb_act]
            if not need_bwd:
                return outs, None

            def bwd(g_h0, g_temb):
                if g_h0 is not None:
                    b_ci(as2d(g_h0), need_dx=False)
                if g_temb is not None:
                    d_emb = K.silu_bwd(g_temb.contiguous(), emb)
                    d_a1 = b_l2(d_emb)
                    d_e1 = K.silu_bwd(d_a1, e1)
                    b_l1(d_e1, need_dx=False)
                if gates_:                       # every gated module has accumulated its d gate by now
                    return (None, None, rt.dg)
                return (None, None)

            return outs, bwd

        extra = () if gate_in is None else (gate_in,)
        return run_block(runner, self._anchor, sample, timesteps, *extra, owner=(self, "_grad_ready_stem"))

    @staticmethod
    @torch.no_grad()
    def _context(ehs):
        """Text-encoder states [B, 77, 1024] -> bf16 matrix [B*77, 1024] (frozen input: no gradient, trainer.py:2433)."""
        if ehs.dtype == BF16 and ehs.is_contiguous():
            return ehs.view(-1, ehs.shape[-1])
        return K.cast_f32_to_bf16(ehs.contiguous().float()).view(-1, ehs.shape[-1])

    def _head(self, x4):
        B, _, H, W = x4.shape

        def runner(need_bwd, x4_):
            x = as2d(x4_)
            h, b_n = bnn.gn(x, self.conv_norm_out, B, H * W, True, need_bwd)                   # reference :1720-1722
            y, b_c = bnn.conv(h, self.conv_out, B, H, W, need_bwd)                             # reference :1723
            out = [K.nhwc_bf16_to_nchw_f32(y, B, H, W)]
            if not need_bwd:
                return out, None

            def bwd(g):
                dy = K.nchw_f32_to_nhwc_bf16(g)
                dh, _ = b_c(dy)
                return (to4d(b_n(dh), B, H, W),)

            return out, bwd

        return run_block(runner, self._anchor, x4, owner=(self, "_grad_ready_head"))[0]

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True, **kwargs):
        """Reference :1417-1728 for the SD-2.1 configuration."""
        if self.arena.device.type != "cuda":
            raise RuntimeError("forward needs a CUDA (sm_100a) device: there is no CPU fallback")
        if self.arena.trainable and self.conv_in.weight.grad is None:
            self.arena.reattach_grads()                     # a foreign zero_grad(set_to_none=True) detached them
        self.arena.ensure_shadow()
        if self.arena.trainable and torch.is_grad_enabled():
            # Per-forward graph anchor: a fresh leaf makes autograd record the block functions even though no activation
            # input requires grad.  (A long-lived leaf would pin its AccumulateGrad node to whatever stream first used it,
            # and the engine's end-of-backward sync with that stream breaks CUDA-graph capture of the training step.)
            anchor = torch.empty(1, device=sample.device, requires_grad=True)
            for blk in list(self.down_blocks) + [self.mid_block] + list(self.up_blocks):
                blk._anchor = anchor
            self._anchor = anchor
        rt, gate_in = None, None
        if self._gate_src is not None and not self.is_pruned():
            # un-pruned mode: distribute this forward's gate values; their gradient comes back through the stem's node,
            # the last one of the backward pass (every other block consumes the stem's outputs)
            gate_in = self._gate_src.to(device=sample.device, dtype=F32)
            if sample.shape[0] % gate_in.shape[0]:
                raise ValueError("gate rows must divide the batch (gates.py:23-25 repeats them over the batch)")
            rt = RuntimeGates(gate_in)
        for m in self.modules():
            if hasattr(m, "_rt"):
                m._rt = rt
        self._rt_live = rt
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0]).to(torch.int64).contiguous()
        h, temb_act = self._stem(sample, timesteps, gate_in)
        ctx2d = self._context(encoder_hidden_states)
        res = (h,)
        for blk in self.down_blocks:                                                            # reference :1631-1653
            h, r = blk(hidden_states=h, temb=temb_act, encoder_hidden_states=ctx2d)
            res += r
        h = self.mid_block(h, temb=temb_act, encoder_hidden_states=ctx2d)                       # reference :1667
        for blk in self.up_blocks:                                                              # reference :1688-1717
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            h = blk(hidden_states=h, res_hidden_states_tuple=r, temb=temb_act, encoder_hidden_states=ctx2d)
        out = self._head(h)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)


class UNet2DConditionModelPruned(UNet2DConditionModelGated):
    """Reference :2183 -- same network; the class the trainer instantiates (trainer.py:2165-2176)."""


class UNet2DConditionModel(UNet2DConditionModelGated):
    """Frozen full-width teacher (reference: stock diffusers.UNet2DConditionModel, trainer.py:2145-2149,2187)."""

    def __init__(self, config=None, device=None, seed: Optional[int] = 1, **kw):
        kw.pop("arch_vector", None)
        kw.pop("trainable", None)
        super().__init__(config, arch_vector=None, device=device, trainable=False, seed=seed, **kw)
