"""Step-front producers of the training step, forward only, on the same sm_100a kernels as the U-Net:

* ``AutoencoderKL``  -- the ENCODER half of diffusers' SD-2.1 VAE, what ``vae.encode(pixel_values).latent_dist.sample() *
  vae.config.scaling_factor`` runs in front of every step (reference pdm/training/trainer.py:2405-2406; the model is loaded at
  :2126-2134 and frozen, :2187).
* ``CLIPTextModel``  -- the SD-2.1 text encoder (transformers CLIPTextModel: OpenCLIP ViT-H/14 text tower, 23 layers, width
  1024, 16 heads of 64, causal mask, erf-GELU MLP), what the dataset transform calls as ``text_encoder(input_ids)[0]``
  (reference pdm/utils/data_utils.py:155-191,247-276).

Both hold their parameters in a flat arena under the dependency's own state-dict keys (so real checkpoints load with
``load_state_dict``), are frozen (``requires_grad`` False, no backward), and need a CUDA device: there is no CPU fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from ... import kernels as K
from .. import nn as bnn
from ..nn import ParamArena, PConv2d, PGroupNorm, PLayerNorm, PLinear, PModule

BF16, F32 = torch.bfloat16, torch.float32

# stabilityai/stable-diffusion-2-1 vae/config.json (diffusers AutoencoderKL)
SD21_VAE_CONFIG = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, act_fn="silu", scaling_factor=0.18215, sample_size=768)
# stabilityai/stable-diffusion-2-1 text_encoder/config.json (transformers CLIPTextConfig)
SD21_TEXT_CONFIG = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                        num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5)


class _Frozen(nn.Module):
    def _finish(self, device, seed):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise RuntimeError("unlearn_ft_b200 encoders need a CUDA (sm_100a) device: there is no CPU fallback "
                               "(pass device='meta' to inspect state-dict shapes without a GPU)")
        self.arena = ParamArena(self, device, trainable=False, seed=seed)
        self.eval()

    @property
    def device(self):
        return self.arena.device

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in state_dict.items() if k in own}       # (a full VAE checkpoint also carries the decoder)
        missing = own - set(sd)
        if strict and missing:
            raise KeyError(f"missing keys: {sorted(missing)[:5]} ...")
        out = super().load_state_dict(sd, strict=False, assign=False)
        self.arena.shadow_fresh = False
        return out

    _WEIGHTS = ()              # file names tried in order inside the checkpoint folder (set by the subclasses)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, device=None, **kwargs):
        """Local checkpoint folder in the dependency's own layout: ``<root>/<subfolder>/config.json`` + the safetensors (or
        .bin) weight file -- the call the reference makes at pdm/training/trainer.py:2126-2143 (`AutoencoderKL.from_pretrained(
        path, subfolder="vae")`, `CLIPTextModel.from_pretrained(path, subfolder="text_encoder")`).  Hub-only keyword arguments
        (revision, variant, torch_dtype ...) are accepted and ignored; there is no download path."""
        import json
        import os
        root = os.path.join(str(pretrained_model_name_or_path), subfolder) if subfolder else str(pretrained_model_name_or_path)
        with open(os.path.join(root, "config.json")) as f:
            raw = json.load(f)
        known = cls._default_config()
        cfg = {k: (tuple(v) if isinstance(v, list) else v) for k, v in raw.items() if k in known}
        model = cls(cfg, device=device, seed=None)
        for name in cls._WEIGHTS:
            path = os.path.join(root, name)
            if os.path.exists(path):
                if name.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(path)
                else:
                    sd = torch.load(path, map_location="cpu", weights_only=True)
                model.load_state_dict(sd)
                return model
        raise FileNotFoundError(f"none of {cls._WEIGHTS} under {root}")

    def save_pretrained(self, save_directory):
        """config.json + the first of `_WEIGHTS` (safetensors, fp32, the dependency's key names)."""
        import json
        import os
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in self._config.items()}, f, indent=1)
        sd = {k: v.detach().float().cpu().contiguous() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, self._WEIGHTS[0]), metadata={"format": "pt"})

    def _check(self):
        if self.arena.device.type != "cuda":
            raise RuntimeError("forward needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.arena.ensure_shadow()


# ----------------------------------------------------------------------------------------------------------------
# VAE encoder
# ----------------------------------------------------------------------------------------------------------------
class _VaeResnet(nn.Module):
    """diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6): GN -> SiLU -> conv -> GN -> SiLU -> conv (+ 1x1 shortcut)."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = PGroupNorm(groups, cin, 1e-6)
        self.conv1 = PConv2d(cin, cout, 3)
        self.norm2 = PGroupNorm(groups, cout, 1e-6)
        self.conv2 = PConv2d(cout, cout, 3)
        self.conv_shortcut = PConv2d(cin, cout, 1) if cin != cout else None

    def run(self, x, B, H, W):
        h, _ = bnn.gn(x, self.norm1, B, H * W, True, False)
        h, _ = bnn.conv(h, self.conv1, B, H, W, False)
        h, _ = bnn.gn(h, self.norm2, B, H * W, True, False)
        sc = x if self.conv_shortcut is None else bnn.conv(x, self.conv_shortcut, B, H, W, False)[0]
        return bnn.conv(h, self.conv2, B, H, W, False, residual=sc)[0]


class _VaeDownsample(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=0): F.pad(x, (0, 1, 0, 1)) then Conv2d(C, C, 3, stride=2, padding=0)."""

    def __init__(self, c):
        super().__init__()
        self.conv = PConv2d(c, c, 3, stride=2)

    def run(self, x, B, H, W):
        return K.conv_fwd_nopad(x, self.conv.w16, B, H, W, self.conv.out_channels, 2, bias=self.conv.bias)


class _VaeAttention(nn.Module):
    """diffusers Attention(heads=1, dim_head=C, norm_num_groups=32, eps=1e-6, bias=True, residual_connection=True) of
    UNetMidBlock2D: one head of width C = 512 over the H*W tokens.  Head widths other than 64 go through two batched GEMMs
    with a row softmax between them (scores in fp32); the fused tcgen05 attention is specialised for the U-Net's 64."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = PGroupNorm(groups, c, 1e-6)
        self.to_q = PLinear(c, c)
        self.to_k = PLinear(c, c)
        self.to_v = PLinear(c, c)
        self.to_out = nn.ModuleList([PLinear(c, c), nn.Identity()])

    def run(self, x, B, H, W):
        L, C = H * W, x.shape[1]
        n, _ = bnn.gn(x, self.group_norm, B, L, False, False)
        q = K.linear_fwd(n, self.to_q.w16, self.to_q.bias)
        k = K.linear_fwd(n, self.to_k.w16, self.to_k.bias)
        v = K.linear_fwd(n, self.to_v.w16, self.to_v.bias)
        Lp = K.round8(L)
        out = K.alloc2d(B * L, C, x.device)
        step = max(1, min(B, (1 << 30) // (L * Lp * 4)))              # <= 1 GiB of fp32 scores at a time
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            sl = slice(b0 * L, (b0 + nb) * L)
            s = torch.empty(nb * L * Lp, device=x.device, dtype=F32)
            K.bmm(q[sl], k[sl], s, M=L, N=L, K=C, Z1=1, Z2=nb, a_ld=q.stride(0), a_bs=(0, L * q.stride(0)), b_ld=k.stride(0),
                  b_bs=(0, L * k.stride(0)), o_ld=Lp, o_bs=(0, L * Lp))
            p = torch.empty(nb * L * Lp, device=x.device, dtype=BF16)
            K.softmax_fwd(s.view(-1, Lp), p.view(-1, Lp), nb * L, L, C ** -0.5)
            del s
            K.bmm(p, v[sl], out[sl], b_mn=True, M=L, N=C, K=L, Z1=1, Z2=nb, a_ld=Lp, a_bs=(0, L * Lp), b_ld=v.stride(0),
                  b_bs=(0, L * v.stride(0)), o_ld=out.stride(0), o_bs=(0, L * out.stride(0)))
        return K.linear_fwd(out, self.to_out[0].w16, self.to_out[0].bias, residual=x)


class _VaeDownBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([_VaeResnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_VaeDownsample(cout)]) if add_downsample else None


class _VaeMidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([_VaeAttention(c, groups)])
        self.resnets = nn.ModuleList([_VaeResnet(c, c, groups), _VaeResnet(c, c, groups)])


class _VaeEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, G = cfg["block_out_channels"], cfg["norm_num_groups"]
        self.conv_in = PConv2d(cfg["in_channels"], ch[0], 3)
        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, c in enumerate(ch):
            in_c, out_c = out_c, c
            self.down_blocks.append(_VaeDownBlock(in_c, out_c, cfg["layers_per_block"], G, i < len(ch) - 1))
        self.mid_block = _VaeMidBlock(ch[-1], G)
        self.conv_norm_out = PGroupNorm(G, ch[-1], 1e-6)
        self.conv_out = PConv2d(ch[-1], 2 * cfg["latent_channels"], 3)


class _LatentDist:
    """diffusers DiagonalGaussianDistribution over the encoder's moments (kept on the device as NHWC bf16)."""

    def __init__(self, moments, B, H, W, cz):
        self._m, self._B, self._H, self._W, self._cz = moments, B, H, W, cz

    def sample(self, generator=None, noise=None, scale=1.0):
        if noise is None:
            noise = torch.randn(self._B, self._cz, self._H, self._W, device=self._m.device, dtype=F32, generator=generator)
        z, _ = K.vae_sample(self._m, self._B, self._H * self._W, self._cz, scale, eps=noise.contiguous().float())
        return z.view(self._B, self._cz, self._H, self._W)

    def mode(self, scale=1.0):
        z, _ = K.vae_sample(self._m, self._B, self._H * self._W, self._cz, scale, eps=None)
        return z.view(self._B, self._cz, self._H, self._W)

    @property
    def mean(self):
        return self.mode()


class AutoencoderKL(_Frozen):
    """Encoder half of diffusers.AutoencoderKL (keys ``encoder.*`` and ``quant_conv.*``; decoder keys are ignored on load)."""
    _WEIGHTS = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")

    @staticmethod
    def _default_config():
        return SD21_VAE_CONFIG

    def __init__(self, config: Optional[dict] = None, device=None, seed: Optional[int] = 2, **overrides):
        super().__init__()
        cfg = dict(SD21_VAE_CONFIG)
        cfg.update(config or {})
        cfg.update(overrides)
        self._config, self.config = cfg, SimpleNamespace(**cfg)
        self.encoder = _VaeEncoder(cfg)
        self.quant_conv = PConv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)
        self._finish(device, seed)

    @torch.no_grad()
    def encode(self, x, return_dict: bool = True):
        """pixel_values [B, 3, H, W] (H, W multiples of 8 * 2^(levels-1)) -> object with ``.latent_dist`` (diffusers API)."""
        self._check()
        B, _, H, W = x.shape
        e = self.encoder
        h = K.nchw_f32_to_nhwc_bf16(x)
        h, _ = bnn.conv(h, e.conv_in, B, H, W, False)
        for blk in e.down_blocks:
            for r in blk.resnets:
                h = r.run(h, B, H, W)
            if blk.downsamplers is not None:
                h = blk.downsamplers[0].run(h, B, H, W)
                H, W = H // 2, W // 2
        m = e.mid_block
        h = m.resnets[0].run(h, B, H, W)
        h = m.attentions[0].run(h, B, H, W)
        h = m.resnets[1].run(h, B, H, W)
        h, _ = bnn.gn(h, e.conv_norm_out, B, H * W, True, False)
        h, _ = bnn.conv(h, e.conv_out, B, H, W, False)
        mom, _ = bnn.conv(h, self.quant_conv, B, H, W, False)
        dist = _LatentDist(mom, B, H, W, self._config["latent_channels"])
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def encode_latents(self, pixel_values, generator=None, noise=None):
        """trainer.py:2405-2406 in one call: ``vae.encode(x).latent_dist.sample() * vae.config.scaling_factor``, with the
        scaling folded into the sampling kernel."""
        return self.encode(pixel_values).latent_dist.sample(generator=generator, noise=noise,
                                                            scale=self._config["scaling_factor"])


# ----------------------------------------------------------------------------------------------------------------
# CLIP text encoder
# ----------------------------------------------------------------------------------------------------------------
class _Embedding(PModule):
    def __init__(self, n, dim):
        super().__init__()
        self.num_embeddings, self.embedding_dim = n, dim
        self._pspecs = [("weight", (n, dim), "mat")]


class _ClipAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.q_proj = PLinear(dim, dim)
        self.k_proj = PLinear(dim, dim)
        self.v_proj = PLinear(dim, dim)
        self.out_proj = PLinear(dim, dim)


class _ClipMLP(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.fc1 = PLinear(dim, inner)
        self.fc2 = PLinear(inner, dim)


class _ClipLayer(nn.Module):
    def __init__(self, dim, heads, inner, eps):
        super().__init__()
        self.self_attn = _ClipAttention(dim, heads)
        self.layer_norm1 = PLayerNorm(dim, eps)
        self.mlp = _ClipMLP(dim, inner)
        self.layer_norm2 = PLayerNorm(dim, eps)


class _ClipEmbeddings(nn.Module):
    def __init__(self, vocab, dim, max_pos):
        super().__init__()
        self.token_embedding = _Embedding(vocab, dim)
        self.position_embedding = _Embedding(max_pos, dim)


class _ClipEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_ClipLayer(cfg["hidden_size"], cfg["num_attention_heads"], cfg["intermediate_size"],
                                                cfg["layer_norm_eps"]) for _ in range(cfg["num_hidden_layers"])])


class _ClipTextTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _ClipEmbeddings(cfg["vocab_size"], cfg["hidden_size"], cfg["max_position_embeddings"])
        self.encoder = _ClipEncoder(cfg)
        self.final_layer_norm = PLayerNorm(cfg["hidden_size"], cfg["layer_norm_eps"])


class CLIPTextModel(_Frozen):
    """transformers.CLIPTextModel (keys ``text_model.*``): ``model(input_ids)[0]`` = last hidden state after the final
    LayerNorm, [B, 77, 1024] (bf16), as pdm/utils/data_utils.py:180 takes it."""
    _WEIGHTS = ("model.safetensors", "pytorch_model.bin")

    @staticmethod
    def _default_config():
        return SD21_TEXT_CONFIG

    def __init__(self, config: Optional[dict] = None, device=None, seed: Optional[int] = 3, **overrides):
        super().__init__()
        cfg = dict(SD21_TEXT_CONFIG)
        cfg.update(config or {})
        cfg.update(overrides)
        if cfg["hidden_size"] // cfg["num_attention_heads"] != 64:
            raise ValueError("the sm_100a attention path is specialised for head_dim 64 (SD-2.1 text encoder: 1024 / 16)")
        if cfg["hidden_act"] != "gelu":
            raise ValueError("only the erf-GELU text encoder of SD-2.x is built (hidden_act 'gelu')")
        self._config, self.config = cfg, SimpleNamespace(**cfg)
        self.text_model = _ClipTextTransformer(cfg)
        self._finish(device, seed)

    @property
    def dtype(self):
        return BF16

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, **_):
        self._check()
        tm = self.text_model
        B, L = input_ids.shape
        heads = self._config["num_attention_heads"]
        dim = self._config["hidden_size"]
        x = K.clip_embed(input_ids.to(self.device), tm.embeddings.token_embedding.weight, tm.embeddings.position_embedding.weight)
        for layer in tm.encoder.layers:
            a = layer.self_attn
            n, _, _ = K.layernorm_fwd(x, layer.layer_norm1.weight, layer.layer_norm1.bias, layer.layer_norm1.eps, save=False)
            q = K.linear_fwd(n, a.q_proj.w16, a.q_proj.bias)          # (biased projections: weights are not adjacent in the
            k = K.linear_fwd(n, a.k_proj.w16, a.k_proj.bias)          # arena, and with B * 77 rows three GEMMs cost nothing)
            v = K.linear_fwd(n, a.v_proj.w16, a.v_proj.bias)
            o, _ = K.attention_fwd(q, k, v, B, heads, L, L, 64 ** -0.5, causal=True)
            x = K.linear_fwd(o, a.out_proj.w16, a.out_proj.bias, residual=x)
            n, _, _ = K.layernorm_fwd(x, layer.layer_norm2.weight, layer.layer_norm2.bias, layer.layer_norm2.eps, save=False)
            h = K.gelu(K.linear_fwd(n, layer.mlp.fc1.w16, layer.mlp.fc1.bias))
            x = K.linear_fwd(h, layer.mlp.fc2.w16, layer.mlp.fc2.bias, residual=x)
        y, _, _ = K.layernorm_fwd(x, tm.final_layer_norm.weight, tm.final_layer_norm.bias, tm.final_layer_norm.eps, save=False)
        last = y.view(B, L, dim)
        return _TextOutput(last)


class _TextOutput(tuple):
    """`out[0]` and `out.last_hidden_state`, like transformers' BaseModelOutputWithPooling."""

    def __new__(cls, last):
        return super().__new__(cls, (last,))

    @property
    def last_hidden_state(self):
        return self[0]
