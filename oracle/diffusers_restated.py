"""TEST INFRASTRUCTURE -- torch-only restatement of the diffusers==0.30.3 classes the reference subclasses.

The reference (`/root/reference/pdm/models/unet/blocks.py:7-16`, `unet_2d_conditional.py:8-42`) builds its gated
U-Net on top of diffusers 0.30.3 (`env.yaml:52`), which is NOT vendored in the reference tree and is not installable
in this image.  This module restates, from the published diffusers 0.30.3 semantics
(SURVEY.md Appendix B), exactly the subset the SD-2.1 configuration exercises.  Constructor signatures and attribute
names follow diffusers so that (a) state-dict keys are the diffusers keys and (b) the reference's own subclass code
can be executed on top of these classes through `oracle/refshim` when generating golden fixtures.

PARITY STATUS: the diffusers semantics here are "restated, unpinned" (no diffusers wheel to compare with); the
reference's own code on top of them IS pinned by tests/golden (see oracle/make_golden.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


def get_activation(name: str) -> nn.Module:
    name = name.lower()
    if name in ("swish", "silu"):
        return nn.SiLU()
    if name == "gelu":
        return nn.GELU()
    if name == "relu":
        return nn.ReLU()
    if name == "mish":
        return nn.Mish()
    raise ValueError(f"unsupported activation {name}")


# ----------------------------------------------------------------------------------------------------------------
# diffusers.models.embeddings
# ----------------------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float, scale: int = 1):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos = num_channels, flip_sin_to_cos
        self.downscale_freq_shift, self.scale = downscale_freq_shift, scale

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift,
                                      self.scale)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None,
                 sample_proj_bias=True):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim, sample_proj_bias)
        self.cond_proj = nn.Linear(cond_proj_dim, in_channels, bias=False) if cond_proj_dim is not None else None
        self.act = get_activation(act_fn)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim, sample_proj_bias)
        self.post_act = get_activation(post_act_fn) if post_act_fn is not None else None

    def forward(self, sample, condition=None):
        if condition is not None:
            sample = sample + self.cond_proj(condition)
        sample = self.linear_2(self.act(self.linear_1(sample)))
        if self.post_act is not None:
            sample = self.post_act(sample)
        return sample


# ----------------------------------------------------------------------------------------------------------------
# diffusers.models.resnet / upsampling / downsampling
# ----------------------------------------------------------------------------------------------------------------
class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv",
                 kernel_size=None, padding=1, norm_type=None, eps=None, elementwise_affine=None, bias=True,
                 interpolate=True):
        super().__init__()
        if use_conv_transpose or norm_type is not None:
            raise NotImplementedError
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.name, self.interpolate = use_conv, name, interpolate
        conv = None
        if use_conv:
            conv = nn.Conv2d(self.channels, self.out_channels, kernel_size=kernel_size or 3, padding=padding, bias=bias)
        if name == "conv":
            self.conv = conv
        else:
            self.Conv2d_0 = conv

    def forward(self, hidden_states, output_size=None, *args, **kwargs):
        if self.interpolate:
            if output_size is None:
                hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
            else:
                hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        if self.use_conv:
            hidden_states = self.conv(hidden_states) if self.name == "conv" else self.Conv2d_0(hidden_states)
        return hidden_states


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv", kernel_size=3,
                 norm_type=None, eps=None, elementwise_affine=None, bias=True):
        super().__init__()
        if norm_type is not None:
            raise NotImplementedError
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        if use_conv:
            conv = nn.Conv2d(self.channels, self.out_channels, kernel_size=kernel_size, stride=2, padding=padding,
                             bias=bias)
        else:
            conv = nn.AvgPool2d(kernel_size=2, stride=2)
        if name == "conv":
            self.Conv2d_0 = conv
            self.conv = conv
        else:
            self.conv = conv

    def forward(self, hidden_states, *args, **kwargs):
        if self.use_conv and self.padding == 0:
            hidden_states = F.pad(hidden_states, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(hidden_states)


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        if time_embedding_norm not in ("default", "scale_shift") or up or down or kernel is not None:
            raise NotImplementedError
        self.pre_norm = True
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.up, self.down = up, down
        self.output_scale_factor = output_scale_factor
        self.time_embedding_norm = time_embedding_norm
        self.skip_time_act = skip_time_act
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if temb_channels is not None:
            width = out_channels if time_embedding_norm == "default" else 2 * out_channels
            self.time_emb_proj = nn.Linear(temb_channels, width)
        else:
            self.time_emb_proj = None
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_2d_out_channels = conv_2d_out_channels or out_channels
        self.conv2 = nn.Conv2d(out_channels, conv_2d_out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = get_activation(non_linearity)
        self.upsample = self.downsample = None
        self.use_in_shortcut = self.in_channels != conv_2d_out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = None
        if self.use_in_shortcut:
            self.conv_shortcut = nn.Conv2d(in_channels, conv_2d_out_channels, kernel_size=1, stride=1, padding=0,
                                           bias=conv_shortcut_bias)

    def forward(self, input_tensor, temb, *args, **kwargs):
        hidden_states = self.nonlinearity(self.norm1(input_tensor))
        hidden_states = self.conv1(hidden_states)
        if self.time_emb_proj is not None:
            if not self.skip_time_act:
                temb = self.nonlinearity(temb)
            temb = self.time_emb_proj(temb)[:, :, None, None]
        if self.time_embedding_norm == "default":
            if temb is not None:
                hidden_states = hidden_states + temb
            hidden_states = self.norm2(hidden_states)
        else:
            time_scale, time_shift = torch.chunk(temb, 2, dim=1)
            hidden_states = self.norm2(hidden_states)
            hidden_states = hidden_states * (1 + time_scale) + time_shift
        hidden_states = self.conv2(self.dropout(self.nonlinearity(hidden_states)))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + hidden_states) / self.output_scale_factor


# ----------------------------------------------------------------------------------------------------------------
# diffusers.models.activations / attention / attention_processor
# ----------------------------------------------------------------------------------------------------------------
class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2, bias=bias)

    def gelu(self, gate):
        return F.gelu(gate)

    def forward(self, hidden_states, *args, **kwargs):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * self.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False,
                 inner_dim=None, bias=True):
        super().__init__()
        if activation_fn != "geglu":
            raise NotImplementedError
        inner_dim = int(dim * mult) if inner_dim is None else inner_dim
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner_dim, bias=bias), nn.Dropout(dropout),
                                  nn.Linear(inner_dim, dim_out, bias=bias)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states, *args, **kwargs):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class AttnProcessor2_0:
    """softmax(QK^T / sqrt(d)) V through F.scaled_dot_product_attention (diffusers AttnProcessor2_0)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, *args, **kwargs):
        if attention_mask is not None or attn.spatial_norm is not None or attn.group_norm is not None:
            raise NotImplementedError
        batch_size = hidden_states.shape[0]
        query = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        key, value = attn.to_k(ctx), attn.to_v(ctx)
        head_dim = key.shape[-1] // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
        hidden_states = attn.to_out[1](attn.to_out[0](hidden_states))
        return hidden_states / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, kv_heads=None, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, cross_attention_norm=None,
                 cross_attention_norm_num_groups=32, qk_norm=None, added_kv_proj_dim=None, norm_num_groups=None,
                 spatial_norm_dim=None, out_bias=True, scale_qk=True, only_cross_attention=False, eps=1e-5,
                 rescale_output_factor=1.0, residual_connection=False, _from_deprecated_attn_block=False, processor=None,
                 out_dim=None, context_pre_only=None):
        super().__init__()
        if any(v is not None for v in (kv_heads, cross_attention_norm, qk_norm, added_kv_proj_dim, norm_num_groups,
                                       spatial_norm_dim, out_dim)):
            raise NotImplementedError
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.rescale_output_factor, self.residual_connection = rescale_output_factor, residual_connection
        self.dropout = dropout
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.only_cross_attention = only_cross_attention
        self.group_norm = self.spatial_norm = self.norm_q = self.norm_k = self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = None
        self.set_processor(processor if processor is not None else AttnProcessor2_0())

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, *a, **k):
        raise NotImplementedError

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, upcast_attention=False, norm_elementwise_affine=True,
                 norm_type="layer_norm", norm_eps=1e-5, final_dropout=False, attention_type="default",
                 positional_embeddings=None, num_positional_embeddings=None,
                 ada_norm_continous_conditioning_embedding_dim=None, ada_norm_bias=None, ff_inner_dim=None, ff_bias=True,
                 attention_out_bias=True):
        super().__init__()
        if norm_type != "layer_norm" or positional_embeddings is not None or attention_type != "default":
            raise NotImplementedError
        self.only_cross_attention = only_cross_attention
        self.norm_type = norm_type
        self.pos_embed = None
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine, eps=norm_eps)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias,
                               cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention, out_bias=attention_out_bias)
        if cross_attention_dim is not None or double_self_attention:
            self.norm2 = nn.LayerNorm(dim, norm_eps, norm_elementwise_affine)
            self.attn2 = Attention(query_dim=dim,
                                   cross_attention_dim=cross_attention_dim if not double_self_attention else None,
                                   heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                                   bias=attention_bias, upcast_attention=upcast_attention, out_bias=attention_out_bias)
        else:
            self.norm2 = self.attn2 = None
        self.norm3 = nn.LayerNorm(dim, norm_eps, norm_elementwise_affine)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout,
                              inner_dim=ff_inner_dim, bias=ff_bias)
        self._chunk_size = None
        self._chunk_dim = 0

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None, added_cond_kwargs=None):
        attn_output = self.attn1(self.norm1(hidden_states),
                                 encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                                 attention_mask=attention_mask)
        hidden_states = attn_output + hidden_states
        if self.attn2 is not None:
            attn_output = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states,
                                     attention_mask=encoder_attention_mask)
            hidden_states = attn_output + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


class _Config(dict):
    """Minimal stand-in for diffusers' FrozenDict config (attribute + item access)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def register_to_config(init):
    """Decorator equivalent of diffusers.configuration_utils.register_to_config: records ctor kwargs in self.config."""
    import functools
    import inspect

    sig = inspect.signature(init)

    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
        init(self, *args, **kwargs)
        existing = dict(getattr(self, "_internal_dict", {}))
        existing.update(cfg)  # the outermost (sub)class ctor runs last and wins
        self._internal_dict = _Config(existing)

    return inner


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict

    def register_to_config(self, **kwargs):
        d = dict(getattr(self, "_internal_dict", {}))
        d.update(kwargs)
        self._internal_dict = _Config(d)

    @classmethod
    def from_config(cls, config, **kwargs):
        import inspect

        params = inspect.signature(cls.__init__).parameters
        init_kwargs = {k: v for k, v in dict(config).items() if k in params and not k.startswith("_")}
        init_kwargs.update({k: v for k, v in kwargs.items() if k in params})
        return cls(**init_kwargs)


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype


class Transformer2DModel(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False, sample_size=None,
                 num_vector_embeds=None, patch_size=None, activation_fn="geglu", num_embeds_ada_norm=None,
                 use_linear_projection=False, only_cross_attention=False, double_self_attention=False,
                 upcast_attention=False, norm_type="layer_norm", norm_elementwise_affine=True, norm_eps=1e-5,
                 attention_type="default", caption_channels=None, interpolation_scale=None,
                 use_additional_conditions=None):
        super().__init__()
        if patch_size is not None or num_vector_embeds is not None or in_channels is None:
            raise NotImplementedError("only continuous inputs are restated")
        self.use_linear_projection = use_linear_projection
        self.interpolation_scale, self.caption_channels = interpolation_scale, caption_channels
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        self.inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.gradient_checkpointing = False
        self.is_input_continuous, self.is_input_vectorized, self.is_input_patches = True, False, False
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, self.inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, self.inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(self.inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, double_self_attention=double_self_attention,
                                  upcast_attention=upcast_attention, norm_type=norm_type,
                                  norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps,
                                  attention_type=attention_type) for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(self.inner_dim, self.out_channels)
        else:
            self.proj_out = nn.Conv2d(self.inner_dim, self.out_channels, kernel_size=1, stride=1, padding=0)

    def _operate_on_continuous_inputs(self, hidden_states):
        batch, _, height, width = hidden_states.shape
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)
        return hidden_states, inner_dim

    def _get_output_for_continuous_inputs(self, hidden_states, residual, batch_size, height, width, inner_dim):
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch_size, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch_size, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        return hidden_states + residual

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, added_cond_kwargs=None,
                class_labels=None, cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None,
                return_dict=True):
        if attention_mask is not None or encoder_attention_mask is not None:
            raise NotImplementedError
        batch_size, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states, inner_dim = self._operate_on_continuous_inputs(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask,
                                  encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        output = self._get_output_for_continuous_inputs(hidden_states, residual, batch_size, height, width, inner_dim)
        if not return_dict:
            return (output,)
        return Transformer2DModelOutput(sample=output)


# ----------------------------------------------------------------------------------------------------------------
# diffusers.models.unets.unet_2d_blocks
# ----------------------------------------------------------------------------------------------------------------
class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1,
                 transformer_layers_per_block=1, resnet_eps=1e-6, resnet_time_scale_shift="default",
                 resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True, num_attention_heads=1,
                 cross_attention_dim=1280, output_scale_factor=1.0, downsample_padding=1, add_downsample=True,
                 dual_cross_attention=False, use_linear_projection=False, only_cross_attention=False,
                 upcast_attention=False, attention_type="default"):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions = [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                         temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups,
                                         dropout=dropout, time_embedding_norm=resnet_time_scale_shift,
                                         non_linearity=resnet_act_fn, output_scale_factor=output_scale_factor,
                                         pre_norm=resnet_pre_norm))
            attentions.append(Transformer2DModel(num_attention_heads, out_channels // num_attention_heads,
                                                 in_channels=out_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim, norm_num_groups=resnet_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 only_cross_attention=only_cross_attention,
                                                 upcast_attention=upcast_attention, attention_type=attention_type))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        if add_downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                            padding=downsample_padding, name="op")])
        else:
            self.downsamplers = None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None, additional_residuals=None):
        output_states = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for downsampler in self.downsamplers:
                hidden_states = downsampler(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 output_scale_factor=1.0, add_downsample=True, downsample_padding=1):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                         temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups,
                                         dropout=dropout, time_embedding_norm=resnet_time_scale_shift,
                                         non_linearity=resnet_act_fn, output_scale_factor=output_scale_factor,
                                         pre_norm=resnet_pre_norm))
        self.resnets = nn.ModuleList(resnets)
        if add_downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                            padding=downsample_padding, name="op")])
        else:
            self.downsamplers = None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, *args, **kwargs):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for downsampler in self.downsamplers:
                hidden_states = downsampler(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class CrossAttnUpBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, resolution_idx=None, dropout=0.0,
                 num_layers=1, transformer_layers_per_block=1, resnet_eps=1e-6, resnet_time_scale_shift="default",
                 resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True, num_attention_heads=1,
                 cross_attention_dim=1280, output_scale_factor=1.0, add_upsample=True, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 attention_type="default"):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions = [], []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(in_channels=resnet_in_channels + res_skip_channels, out_channels=out_channels,
                                         temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups,
                                         dropout=dropout, time_embedding_norm=resnet_time_scale_shift,
                                         non_linearity=resnet_act_fn, output_scale_factor=output_scale_factor,
                                         pre_norm=resnet_pre_norm))
            attentions.append(Transformer2DModel(num_attention_heads, out_channels // num_attention_heads,
                                                 in_channels=out_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim, norm_num_groups=resnet_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 only_cross_attention=only_cross_attention,
                                                 upcast_attention=upcast_attention, attention_type=attention_type))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False
        self.resolution_idx = resolution_idx

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None, encoder_attention_mask=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
        if self.upsamplers is not None:
            for upsampler in self.upsamplers:
                hidden_states = upsampler(hidden_states, upsample_size)
        return hidden_states


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, resolution_idx=None, dropout=0.0,
                 num_layers=1, resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish",
                 resnet_groups=32, resnet_pre_norm=True, output_scale_factor=1.0, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(in_channels=resnet_in_channels + res_skip_channels, out_channels=out_channels,
                                         temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups,
                                         dropout=dropout, time_embedding_norm=resnet_time_scale_shift,
                                         non_linearity=resnet_act_fn, output_scale_factor=output_scale_factor,
                                         pre_norm=resnet_pre_norm))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False
        self.resolution_idx = resolution_idx

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, *args, **kwargs):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for upsampler in self.upsamplers:
                hidden_states = upsampler(hidden_states, upsample_size)
        return hidden_states


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, out_channels=None, dropout=0.0, num_layers=1,
                 transformer_layers_per_block=1, resnet_eps=1e-6, resnet_time_scale_shift="default",
                 resnet_act_fn="swish", resnet_groups=32, resnet_groups_out=None, resnet_pre_norm=True,
                 num_attention_heads=1, output_scale_factor=1.0, cross_attention_dim=1280, dual_cross_attention=False,
                 use_linear_projection=False, upcast_attention=False, attention_type="default"):
        super().__init__()
        if dual_cross_attention:
            raise NotImplementedError
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnet_groups_out = resnet_groups_out or resnet_groups

        def mk(cin):
            return ResnetBlock2D(in_channels=cin, out_channels=out_channels, temb_channels=temb_channels, eps=resnet_eps,
                                 groups=resnet_groups, groups_out=resnet_groups_out, dropout=dropout,
                                 time_embedding_norm=resnet_time_scale_shift, non_linearity=resnet_act_fn,
                                 output_scale_factor=output_scale_factor, pre_norm=resnet_pre_norm)

        resnets, attentions = [mk(in_channels)], []
        for _ in range(num_layers):
            attentions.append(Transformer2DModel(num_attention_heads, out_channels // num_attention_heads,
                                                 in_channels=out_channels, num_layers=transformer_layers_per_block,
                                                 cross_attention_dim=cross_attention_dim,
                                                 norm_num_groups=resnet_groups_out,
                                                 use_linear_projection=use_linear_projection,
                                                 upcast_attention=upcast_attention, attention_type=attention_type))
            resnets.append(mk(out_channels))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor = None


# SD-2.1 U-Net hub config (stabilityai/stable-diffusion-2-1 unet/config.json)
SD21_UNET_CONFIG = dict(
    sample_size=96, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1024,
    attention_head_dim=(5, 10, 20, 20), dual_cross_attention=False, use_linear_projection=True,
    upcast_attention=True, resnet_time_scale_shift="default",
)


class UNet2DConditionModel(ModelMixin, ConfigMixin):
    """Stock (teacher) U-Net: the subset of diffusers.UNet2DConditionModel SD-2.1 instantiates
    (reference use: pdm/training/trainer.py:2145-2149,2448)."""

    @register_to_config
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1280, attention_head_dim=8, dual_cross_attention=False,
                 use_linear_projection=False, upcast_attention=False, resnet_time_scale_shift="default"):
        super().__init__()
        n = len(down_block_types)
        heads = attention_head_dim if isinstance(attention_head_dim, (tuple, list)) else (attention_head_dim,) * n
        time_embed_dim = block_out_channels[0] * 4
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn)
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            final = i == n - 1
            if t == "CrossAttnDownBlock2D":
                blk = CrossAttnDownBlock2D(in_channels=input_channel, out_channels=output_channel,
                                           temb_channels=time_embed_dim, num_layers=layers_per_block,
                                           resnet_eps=norm_eps, resnet_act_fn=act_fn, resnet_groups=norm_num_groups,
                                           num_attention_heads=heads[i], cross_attention_dim=cross_attention_dim,
                                           downsample_padding=downsample_padding, add_downsample=not final,
                                           use_linear_projection=use_linear_projection,
                                           upcast_attention=upcast_attention)
            elif t == "DownBlock2D":
                blk = DownBlock2D(in_channels=input_channel, out_channels=output_channel, temb_channels=time_embed_dim,
                                  num_layers=layers_per_block, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                                  resnet_groups=norm_num_groups, add_downsample=not final,
                                  downsample_padding=downsample_padding)
            else:
                raise NotImplementedError(t)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2DCrossAttn(in_channels=block_out_channels[-1], temb_channels=time_embed_dim,
                                                 resnet_eps=norm_eps, resnet_act_fn=act_fn,
                                                 output_scale_factor=mid_block_scale_factor,
                                                 cross_attention_dim=cross_attention_dim,
                                                 num_attention_heads=heads[-1], resnet_groups=norm_num_groups,
                                                 use_linear_projection=use_linear_projection,
                                                 upcast_attention=upcast_attention)
        rev_ch, rev_heads = list(reversed(block_out_channels)), list(reversed(heads))
        output_channel = rev_ch[0]
        for i, t in enumerate(up_block_types):
            final = i == n - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            if t == "CrossAttnUpBlock2D":
                blk = CrossAttnUpBlock2D(in_channels=input_channel, out_channels=output_channel,
                                         prev_output_channel=prev_output_channel, temb_channels=time_embed_dim,
                                         num_layers=layers_per_block + 1, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                                         resnet_groups=norm_num_groups, num_attention_heads=rev_heads[i],
                                         cross_attention_dim=cross_attention_dim, add_upsample=not final,
                                         use_linear_projection=use_linear_projection, upcast_attention=upcast_attention)
            elif t == "UpBlock2D":
                blk = UpBlock2D(in_channels=input_channel, prev_output_channel=prev_output_channel,
                                out_channels=output_channel, temb_channels=time_embed_dim,
                                num_layers=layers_per_block + 1, resnet_eps=norm_eps, resnet_act_fn=act_fn,
                                resnet_groups=norm_num_groups, add_upsample=not final)
            else:
                raise NotImplementedError(t)
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = get_activation(act_fn)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, return_dict=True, **kwargs):
        return unet_forward(self, sample, timestep, encoder_hidden_states, return_dict)


def unet_forward(model, sample, timestep, encoder_hidden_states, return_dict=True):
    """Data flow of UNet2DConditionModel.forward for the SD-2.1 configuration (identical in the reference's gated
    copy, pdm/models/unet/unet_2d_conditional.py:1417-1728)."""
    timesteps = timestep
    if not torch.is_tensor(timesteps):
        timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
    elif timesteps.dim() == 0:
        timesteps = timesteps[None].to(sample.device)
    timesteps = timesteps.expand(sample.shape[0])
    t_emb = model.time_proj(timesteps).to(dtype=sample.dtype)
    emb = model.time_embedding(t_emb)
    sample = model.conv_in(sample)
    down_block_res_samples = (sample,)
    for blk in model.down_blocks:
        if getattr(blk, "has_cross_attention", False):
            sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
        else:
            sample, res = blk(hidden_states=sample, temb=emb)
        down_block_res_samples += res
    sample = model.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
    for blk in model.up_blocks:
        res = down_block_res_samples[-len(blk.resnets):]
        down_block_res_samples = down_block_res_samples[:-len(blk.resnets)]
        if getattr(blk, "has_cross_attention", False):
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                         encoder_hidden_states=encoder_hidden_states)
        else:
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res)
    sample = model.conv_out(model.conv_act(model.conv_norm_out(sample)))
    if not return_dict:
        return (sample,)
    return UNet2DConditionOutput(sample=sample)


# ----------------------------------------------------------------------------------------------------------------
# diffusers.schedulers.DDIMScheduler -- training-time subset (add_noise / get_velocity / alphas_cumprod)
# ----------------------------------------------------------------------------------------------------------------
class DDIMSchedulerLite:
    """SD-2.1 scheduler config: scaled_linear betas 0.00085 -> 0.012, 1000 steps, v_prediction."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, prediction_type="v_prediction"):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)

        class _C:
            pass

        self.config = _C()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.prediction_type = prediction_type

    def _coeffs(self, sample, timesteps):
        acp = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        timesteps = timesteps.to(sample.device)
        a = acp[timesteps] ** 0.5
        s = (1 - acp[timesteps]) ** 0.5
        while a.dim() < sample.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a, s

    def add_noise(self, original_samples, noise, timesteps):
        a, s = self._coeffs(original_samples, timesteps)
        return a * original_samples + s * noise

    def get_velocity(self, sample, noise, timesteps):
        a, s = self._coeffs(sample, timesteps)
        return a * noise - s * sample

    # ---- sampling side (diffusers 0.30.3 DDIMScheduler with the SD-2.1 scheduler_config.json: clip_sample False,
    #      set_alpha_to_one False, steps_offset 1, timestep_spacing "leading", v_prediction; restated, unpinned)
    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        step_ratio = T // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * step_ratio).round().flip(0).to(torch.int64) + 1     # steps_offset = 1
        self.timesteps = ts.to(device) if device is not None else ts
        self.init_noise_sigma = 1.0

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        """DDIMScheduler.step for eta = 0 (the pipeline default, pruning_pipelines.py:876): returns prev_sample."""
        assert eta == 0.0
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        acp = self.alphas_cumprod.to(sample.device)
        a_t = acp[t]
        a_prev = acp[prev_t] if prev_t >= 0 else acp[0]                       # final_alpha_cumprod (set_alpha_to_one False)
        b_t = 1 - a_t
        assert self.config.prediction_type == "v_prediction"
        pred_x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
        pred_eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        direction = (1 - a_prev) ** 0.5 * pred_eps                            # sigma_t = 0
        return a_prev ** 0.5 * pred_x0 + direction


# ----------------------------------------------------------------------------------------------------------------
# diffusers.schedulers.PNDMScheduler -- what scripts/metrics/generate_fid_images.py:113 loads from the SD-2.1 hub folder
# (scheduler_config.json: skip_prk_steps True, steps_offset 1, set_alpha_to_one False, timestep_spacing "leading",
# prediction_type v_prediction).  Restated from diffusers 0.30.3 `scheduling_pndm.py` (set_timesteps / step_plms /
# _get_prev_sample); unpinned like the rest of this file.
# ----------------------------------------------------------------------------------------------------------------
class PNDMSchedulerLite(DDIMSchedulerLite):
    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        step_ratio = T // num_inference_steps
        base = (torch.arange(0, num_inference_steps) * step_ratio).round().to(torch.int64) + 1            # steps_offset = 1
        plms = torch.cat([base[:-1], base[-2:-1], base[-1:]]).flip(0)                                      # skip_prk_steps
        self.timesteps = plms.to(device) if device is not None else plms
        self.init_noise_sigma = 1.0
        self.ets, self.counter, self.cur_sample = [], 0, None

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        """PNDMScheduler.step -> step_plms (the PRK phase is skipped): fourth-order linear multistep."""
        t = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        prev_t = t - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = t
            t = t + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        self.counter += 1
        return self._get_prev_sample(sample, t, prev_t, model_output)

    def _get_prev_sample(self, sample, t, prev_t, model_output):
        acp = self.alphas_cumprod.to(sample.device)
        a_t = acp[t]
        a_prev = acp[prev_t] if prev_t >= 0 else acp[0]                       # final_alpha_cumprod (set_alpha_to_one False)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        assert self.config.prediction_type == "v_prediction"
        model_output = a_t ** 0.5 * model_output + b_t ** 0.5 * sample        # v -> epsilon
        sample_coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        return sample_coeff * sample - (a_prev - a_t) * model_output / denom
