"""ORACLE (test infrastructure, not the product): CPU/torch restatement of the ENCODER half of diffusers 0.30.3
`AutoencoderKL` with the SD-2.1 VAE config, as `vae.encode(x).latent_dist.sample() * scaling_factor` uses it at reference
pdm/training/trainer.py:2405-2406.  diffusers is an un-vendored dependency (env.yaml:52) that cannot be installed offline, so
this layer is **restated, unpinned** (written from the published diffusers source: `models/autoencoders/autoencoder_kl.py`,
`models/autoencoders/vae.py::Encoder`, `models/unets/unet_2d_blocks.py::DownEncoderBlock2D / UNetMidBlock2D`,
`models/resnet.py::ResnetBlock2D`, `models/downsampling.py::Downsample2D`, `models/attention_processor.py::Attention`).
State-dict keys follow diffusers so the same tensors load into unlearn_ft_b200.pdm.models.AutoencoderKL.

The text encoder needs no restatement: transformers (the dependency the reference imports, `CLIPTextModel`) is installed in the
image and serves as the pinned oracle directly (tests/test_encoders_gpu.py)."""
import torch
import torch.nn.functional as F
from torch import nn

SD21_VAE_CONFIG = dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, scaling_factor=0.18215)


class ResnetBlock2D(nn.Module):
    """resnet.py::ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, non_linearity='silu', output_scale_factor=1)."""

    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    """downsampling.py::Downsample2D(use_conv=True, padding=0): pad (0, 1, 0, 1) then a stride-2 padding-0 convolution."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Attention(nn.Module):
    """attention_processor.py::Attention(heads=1, dim_head=C, norm_num_groups=32, eps=1e-6, bias=True,
    residual_connection=True, rescale_output_factor=1) with AttnProcessor2_0, as UNetMidBlock2D builds it for the VAE."""

    def __init__(self, c, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        res = x
        h = x.view(B, C, H * W).transpose(1, 2)
        h = self.group_norm(h.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        q, k, v = (t.view(B, -1, 1, C).transpose(1, 2) for t in (q, k, v))
        h = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        h = h.transpose(1, 2).reshape(B, -1, C)
        h = self.to_out[1](self.to_out[0](h))
        return h.transpose(-1, -2).reshape(B, C, H, W) + res


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(c, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    """vae.py::Encoder(double_z=True)."""

    def __init__(self, cfg):
        super().__init__()
        ch, G = cfg["block_out_channels"], cfg["norm_num_groups"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, c in enumerate(ch):
            in_c, out_c = out_c, c
            self.down_blocks.append(DownEncoderBlock2D(in_c, out_c, cfg["layers_per_block"], G, i < len(ch) - 1))
        self.mid_block = UNetMidBlock2D(ch[-1], G)
        self.conv_norm_out = nn.GroupNorm(G, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoder(nn.Module):
    """autoencoder_kl.py::AutoencoderKL.encode: Encoder -> quant_conv (1x1) -> DiagonalGaussianDistribution(moments)."""

    def __init__(self, **cfg):
        super().__init__()
        c = dict(SD21_VAE_CONFIG)
        c.update(cfg)
        self.cfg = c
        self.encoder = Encoder(c)
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)

    def moments(self, x):
        return self.quant_conv(self.encoder(x))

    def encode_latents(self, x, noise):
        """trainer.py:2405-2406 with the Gaussian draw passed in: (mean + exp(0.5 clamp(logvar, -30, 20)) * noise) * scaling."""
        mean, logvar = self.moments(x).chunk(2, dim=1)
        logvar = logvar.clamp(-30.0, 20.0)
        return (mean + torch.exp(0.5 * logvar) * noise) * self.cfg["scaling_factor"]
