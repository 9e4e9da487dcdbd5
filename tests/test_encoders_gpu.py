"""Step-front producers on the GPU against their oracles:

* CLIPTextModel  vs  transformers.CLIPTextModel itself (the dependency the reference imports -- a PINNED oracle), SD-2.1 text
  config, random init, identical weights: `text_encoder(input_ids)[0]` (pdm/utils/data_utils.py:180);
* AutoencoderKL.encode  vs  oracle/vae_restated.py (diffusers 0.30.3 restated, unpinned), SD-2.1 VAE config:
  moments and `latent_dist.sample() * scaling_factor` with the same Gaussian draw (pdm/training/trainer.py:2405-2406);
* the kernels added for them (padding-0 stride-2 conv, causal attention, erf-GELU, embedding gather, latent sampling);
* the reference batch contract through the tuner: batch['pixel_values'] / batch['input_ids'] instead of latents / embeddings.
Tolerance: 2e-2 relative (north star, bf16 kernels) against fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _k():
    from unlearn_ft_b200 import kernels
    return kernels


def test_conv_nopad_stride2_matches_padded_torch_conv():
    k = _k()
    B, C, H, W, Co = 2, 128, 64, 64, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(Co, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16().float()
    b = torch.randn(Co, device="cuda", generator=g) * 0.1
    x2 = k.alloc2d(B * H * W, C)
    x2.copy_(x.permute(0, 2, 3, 1).reshape(-1, C))
    wp = torch.zeros(Co, 9, C, device="cuda", dtype=torch.bfloat16)
    wp.copy_(w.permute(0, 2, 3, 1).reshape(Co, 9, C))
    y = k.conv_fwd_nopad(x2, wp, B, H, W, Co, 2, bias=b)
    ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2, padding=0)
    assert rel(y.float().reshape(B, H // 2, W // 2, Co).permute(0, 3, 1, 2), ref) < 1e-2


def test_causal_attention_gelu_embed_sample_kernels():
    k = _k()
    B, Hh, L = 3, 16, 77
    g = torch.Generator(device="cuda").manual_seed(1)
    q, kk, v = (k.alloc2d(B * L, Hh * 64).copy_(torch.randn(B * L, Hh * 64, device="cuda", generator=g)) for _ in range(3))
    o, _ = k.attention_fwd(q, kk, v, B, Hh, L, L, 0.125, causal=True)
    qf, kf, vf = (t.float().view(B, L, Hh, 64).transpose(1, 2) for t in (q, kk, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf, is_causal=True)
    assert rel(o.view(B, L, Hh, 64).transpose(1, 2), ref) < 1e-2
    for Lc in (130, 300):                       # more than one key block: rows whose whole block is masked
        q, kk, v = (k.alloc2d(2 * Lc, 128).copy_(torch.randn(2 * Lc, 128, device="cuda", generator=g)) for _ in range(3))
        o, _ = k.attention_fwd(q, kk, v, 2, 2, Lc, Lc, 0.125, causal=True)
        qf, kf, vf = (t.float().view(2, Lc, 2, 64).transpose(1, 2) for t in (q, kk, v))
        assert rel(o.view(2, Lc, 2, 64).transpose(1, 2), F.scaled_dot_product_attention(qf, kf, vf, is_causal=True)) < 1e-2
    x = k.alloc2d(500, 4096).copy_(torch.randn(500, 4096, device="cuda", generator=g) * 2)
    assert rel(k.gelu(x), F.gelu(x.float())) < 1e-2
    tok, pos = torch.randn(1000, 64, device="cuda", generator=g), torch.randn(77, 64, device="cuda", generator=g)
    ids = torch.randint(0, 1000, (4, 77), device="cuda", generator=g)
    e = k.clip_embed(ids, tok, pos)
    assert rel(e.view(4, 77, 64), tok[ids] + pos[None]) < 5e-3
    mom = k.alloc2d(2 * 64, 8).copy_(torch.randn(128, 8, device="cuda", generator=g) * 3)
    eps = torch.randn(2, 4, 64, device="cuda", generator=g)
    z, mean = k.vae_sample(mom, 2, 64, 4, 0.18215, eps=eps, want_mean=True)
    m = mom.float().view(2, 64, 8).permute(0, 2, 1)
    mu, lv = m[:, :4], m[:, 4:].clamp(-30, 20)
    assert rel(z, (mu + torch.exp(0.5 * lv) * eps) * 0.18215) < 1e-5 and rel(mean, mu) < 1e-6


def _hf_text(layers=23):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HF
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=layers,
                         num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5,
                         projection_dim=512)
    torch.manual_seed(0)
    return HF(cfg).eval().cuda()


def test_clip_text_encoder_matches_transformers():
    from unlearn_ft_b200.pdm.models import CLIPTextModel
    from unlearn_ft_b200.pdm.training import encode_prompt
    hf = _hf_text()
    mine = CLIPTextModel(seed=None)
    mine.load_state_dict(hf.state_dict())
    assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in hf.parameters()) == 340_387_840
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 49406, (4, 77), generator=g)
    ids[:, 0] = 49406                                   # <|startoftext|>
    for b, n in enumerate((5, 20, 76, 40)):             # <|endoftext|> then padding, as the tokenizer produces
        ids[b, n:] = 49407
    ids = ids.cuda()
    with torch.no_grad():
        ref = hf(ids)[0]
    out = mine(ids)
    assert out[0].shape == ref.shape and out.last_hidden_state is out[0]
    print("clip rel", rel(out[0], ref))
    assert rel(out[0], ref) < 2e-2
    emb = encode_prompt(None, mine, text_input_ids=ids)  # reference data_utils.py:155-191 with tokenizer=None
    assert torch.equal(emb, out[0]) and emb.dtype == torch.bfloat16


def test_encoders_load_from_checkpoint_folders(tmp_path):
    """`from_pretrained(root, subfolder=...)` (reference trainer.py:2126-2143) on folders in the dependencies' own layout: a
    transformers CLIPTextModel saved with ITS save_pretrained (config.json + model.safetensors) loads and evaluates equal to the
    in-memory load; the VAE round-trips through save_pretrained / from_pretrained bit for bit."""
    from unlearn_ft_b200.pdm.models import AutoencoderKL, CLIPTextModel
    hf = _hf_text()
    hf.save_pretrained(tmp_path / "text_encoder", safe_serialization=True)
    a = CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder", revision=None, variant=None)
    b = CLIPTextModel(seed=None)
    b.load_state_dict(hf.state_dict())
    ids = torch.randint(0, 49406, (2, 77), generator=torch.Generator().manual_seed(5)).cuda()
    assert torch.equal(a(ids)[0], b(ids)[0])
    v = AutoencoderKL(seed=9)
    v.save_pretrained(str(tmp_path / "vae"))
    w = AutoencoderKL.from_pretrained(str(tmp_path), subfolder="vae")
    assert w.config.scaling_factor == 0.18215 and w.config.block_out_channels == (128, 256, 512, 512)
    x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(6)).cuda()
    assert torch.equal(v.encode(x).latent_dist.mean, w.encode(x).latent_dist.mean)


@pytest.mark.parametrize("B,HW", [(2, 256), (1, 512)])
def test_vae_encoder_matches_oracle(B, HW):
    from oracle.make_golden import deterministic_fill
    from oracle.vae_restated import AutoencoderKLEncoder
    from unlearn_ft_b200.pdm.models import AutoencoderKL
    orc = AutoencoderKLEncoder()
    deterministic_fill(orc, 9)
    mine = AutoencoderKL(seed=None)
    mine.load_state_dict(orc.state_dict())
    orc = orc.eval().cuda()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, HW, HW, generator=g).clamp(-1, 1).cuda()
    noise = torch.randn(B, 4, HW // 8, HW // 8, generator=g).cuda()
    with torch.no_grad():
        mom = orc.moments(x)
        ref = orc.encode_latents(x, noise)
    dist = mine.encode(x).latent_dist
    mean_ref = mom[:, :4]
    print("vae mean rel", rel(dist.mode(), mean_ref), "latents rel", rel(mine.encode_latents(x, noise=noise), ref))
    assert rel(dist.mode(), mean_ref) < 2e-2
    assert rel(mine.encode_latents(x, noise=noise), ref) < 2e-2
    assert rel(dist.sample(noise=noise, scale=0.18215), ref) < 2e-2


def test_reference_batch_contract_through_the_tuner():
    """step() fed with the reference's own batch keys -- pixel_values (-> VAE encode, trainer.py:2405-2406) and token ids
    (-> text encoder, data_utils.py:247-276) -- equals step() fed with the latents / embeddings those producers make."""
    from oracle.make_golden import SMALL64
    from unlearn_ft_b200.pdm.models import (AutoencoderKL, CLIPTextModel, HyperStructure, UNet2DConditionModel,
                                            UNet2DConditionModelPruned)
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import structure_from_config, SD21_CONFIG
    from unlearn_ft_b200.pdm.training import UnetFineTuner
    cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"], cross_attention_dim=128)
    torch.manual_seed(2)
    full = dict(SD21_CONFIG)
    full.update(cfg)
    av = HyperStructure.get_random_arch_vector(0.6, structure_from_config(full))
    student = UNet2DConditionModelPruned(cfg, arch_vector=av, seed=1)
    teacher = UNet2DConditionModel(cfg, seed=2)
    vae = AutoencoderKL(dict(block_out_channels=(32, 64), layers_per_block=1), seed=4)            # 2 levels: 32x32 px -> 16x16 latent
    text = CLIPTextModel(dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2), seed=5)
    tuner = UnetFineTuner(student, teacher, vae=vae, text_encoder=text)
    g = torch.Generator().manual_seed(7)
    px = torch.randn(2, 3, 32, 32, generator=g).cuda()
    ids = torch.randint(0, 49408, (2, 77), generator=g).cuda()
    vae_noise = torch.randn(2, 4, 16, 16, generator=g).cuda()
    noise = torch.randn(2, 4, 16, 16, generator=g).cuda()
    t = torch.tensor([10, 900]).cuda()
    with torch.no_grad():
        a = [float(v) for v in tuner.step(dict(pixel_values=px, input_ids=ids, vae_noise=vae_noise, noise=noise, timesteps=t))]
        lat = vae.encode_latents(px, noise=vae_noise)
        emb = text(ids)[0]
        b = [float(v) for v in tuner.step(dict(latents=lat, prompt_embeds=emb, noise=noise, timesteps=t))]
    assert a == b and all(x == x for x in a)      # (the forward pass is bit-reproducible)
