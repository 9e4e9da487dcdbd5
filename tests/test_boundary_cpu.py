"""CPU (no GPU): the drop-in boundary -- C-ABI exports, and the B200 model's structure / state-dict contract against the
reference golden (models are built on the `meta` device: shapes only, no compute)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_cabi_exports_every_declared_symbol():
    from unlearn_ft_b200 import _lib
    header = open(os.path.join(ROOT, "include", "b200pdm.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(b200pdm_\w+)\s*\(", header))
    so = os.path.join(ROOT, "unlearn_ft_b200", "libb200pdm.so")
    if not os.path.exists(so):
        _lib.build()
    h = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/b200pdm.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.lib()
    assert lib.b200pdm_version() >= 100
    assert lib.b200pdm_launch_count() == 0          # nothing launched: no compute without a GPU


def test_full_model_contract(gold):
    from unlearn_ft_b200.pdm.models import UNet2DConditionModel, UNet2DConditionModelPruned
    teacher = UNet2DConditionModel(device="meta", seed=None)
    assert teacher.num_parameters() == 865_910_724 == gold["sd21_gated_params"]
    assert teacher.get_structure() == gold["sd21_structure"]
    assert {k: list(v.shape) for k, v in teacher.state_dict().items()} == gold["sd21_state_shapes"]
    student = UNet2DConditionModelPruned(arch_vector=gold["av_full_seed1234_r055"], device="meta", seed=None)
    assert student.num_parameters() == 508_224_076              # SURVEY App. C (r = 0.55)
    with pytest.raises(RuntimeError):
        student(torch.zeros(1, 4, 64, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 77, 1024))


@pytest.mark.parametrize("case", ["r055", "r082_drop"])
def test_pruned_shapes_match_reference(gold, case):
    from oracle.make_golden import SMALL64
    from unlearn_ft_b200.pdm.models import UNet2DConditionModelPruned
    g = gold[f"small64_{case}"]
    cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"],
               cross_attention_dim=SMALL64["cross_attention_dim"])
    m = UNet2DConditionModelPruned(cfg, arch_vector=g["arch_vector"], device="meta", seed=None)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == g["shapes"]
    assert m.num_parameters() == g["n_params"]


def test_arch_vector_classmethods_match_reference(gold):
    from unlearn_ft_b200.pdm.models import HyperStructure
    st = gold["sd21_structure"]
    torch.manual_seed(1234)
    av = HyperStructure.get_random_arch_vector(0.55, st)
    assert torch.equal(av, gold["av_full_seed1234_r055"])
    sep = HyperStructure.transform_arch_vector(av, st)
    assert [int(w.shape[1]) for w in sep["width"]] == gold["av_full_split_lens"]
    assert [float(d) for d in sep["depth"]] == gold["av_full_depth"]


def test_hard_concrete_and_snr_match_reference(gold):
    from unlearn_ft_b200.pdm.utils import compute_snr, hard_concrete

    class S:
        alphas_cumprod = torch.cumprod(1 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2, 0)

    assert torch.equal(hard_concrete(gold["hard_concrete_in"]), gold["hard_concrete_out"])
    assert torch.equal(compute_snr(S, gold["snr_timesteps"]), gold["snr"])


def test_checkpoint_wire_format_roundtrip(gold, tmp_path):
    """`save_pretrained` writes the reference's checkpoint layout (trainer.py:314-327,2366-2368) -- `<root>/unet/
    config.json` + `diffusion_pytorch_model.safetensors` under the diffusers key names with the pruned shapes, `<root>/
    arch_vector.pt` -- and `from_pretrained(root, subfolder="unet", checkpoint_loading=True)` (reference :2185-2495) restores it
    bit for bit."""
    import json
    import os

    from safetensors.torch import load_file

    from oracle.make_golden import SMALL64
    from unlearn_ft_b200.pdm.models import UNet2DConditionModelPruned
    g = gold["small64_r055"]
    cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"],
               cross_attention_dim=SMALL64["cross_attention_dim"])
    m = UNet2DConditionModelPruned(cfg, arch_vector=g["arch_vector"], device="cpu", seed=5)
    root = str(tmp_path / "checkpoint-10")
    m.save_pretrained(os.path.join(root, "unet"))
    assert sorted(os.listdir(root)) == ["arch_vector.pt", "unet"]
    assert sorted(os.listdir(os.path.join(root, "unet"))) == ["config.json", "diffusion_pytorch_model.safetensors"]
    sd = load_file(os.path.join(root, "unet", "diffusion_pytorch_model.safetensors"))
    assert {k: list(v.shape) for k, v in sd.items()} == g["shapes"]                 # the reference's keys and pruned shapes
    assert all(v.dtype == torch.float32 for v in sd.values())
    cj = json.load(open(os.path.join(root, "unet", "config.json")))
    assert cj["_class_name"] == "UNet2DConditionModelPruned" and cj["block_out_channels"] == list(SMALL64["block_out_channels"])
    assert torch.equal(torch.load(os.path.join(root, "arch_vector.pt")), g["arch_vector"].float())
    back = UNet2DConditionModelPruned.from_pretrained(root, subfolder="unet", checkpoint_loading=True, device="cpu", seed=None)
    assert torch.equal(back.arch_vector, m.arch_vector)
    a, b = m.state_dict(), back.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_gemm_tile_planner_invariants():
    """Host logic of the tensor-core core, queried without a GPU (b200pdm_gemm_plan): for the shape families of the U-Net
    (conv fprop / dgrad / wgrad, attention and feed-forward projections, time-embedding rows) the plan always satisfies the
    kernel's hard limits -- tile width granularity, two TMEM accumulators of <= 256 columns, >= 2 pipeline stages inside
    227 KB of shared memory, split-K only where allowed and never below 8 k-blocks per split -- and fills the machine."""
    from unlearn_ft_b200 import _lib
    shapes = []
    for tiles_m in (1, 2, 8, 32, 128, 512, 2048):                       # M = 16 ... 262144 rows
        for n in (4, 64, 170, 320, 340, 640, 680, 960, 1280, 2560, 5120, 10240):
            for kb in (2, 5, 10, 20, 45, 90, 135, 180, 360, 1024):
                shapes.append((n, tiles_m, kb))
    for n, tiles_m, kb in shapes:
        for b_mn in (False, True):
            for can_split, fin in ((False, True), (True, True), (True, False)):
                p = _lib.gemm_plan(n, tiles_m=tiles_m, kblocks=kb, b_mn=b_mn, can_split=can_split, split_needs_finalize=fin)
                g = 64 if b_mn else 16
                assert p["block_n"] % g == 0 and g <= p["block_n"] <= 256, (n, tiles_m, kb, p)
                assert p["m_sub"] in (1, 2) and p["m_sub"] * p["block_n"] <= 512
                assert p["pair"] in (0, 1) and (p["pair"] == 0 or tiles_m >= 2)
                assert p["m_sub"] == 1 or tiles_m >= 2 * (2 if p["pair"] else 1)
                assert p["splits"] >= 1 and (can_split or p["splits"] == 1)
                assert p["splits"] == 1 or kb // p["splits"] >= 8
                stage = p["m_sub"] * 16384 + (p["block_n"] // 2 if p["pair"] else p["block_n"]) * 128
                assert 2 <= p["stages"] <= 8 and p["stages"] * stage <= 227 * 1024
                assert p["slots"] == (74 if p["pair"] else 148) and p["tiles"] >= 1
                if p["block_n"] > n:                                     # a padded tile only when N itself is small
                    assert p["block_n"] - n < g or n < 64
    # the planner prefers plans that fill whole waves: the 1280 -> 1280 conv at 8x8 (M = 1024) is split along K
    p = _lib.gemm_plan(1280, tiles_m=8, kblocks=180, can_split=True)
    assert p["splits"] > 1 and p["tiles"] >= 0.9 * p["slots"]
    assert _lib.lib().b200pdm_launch_count() == 0


def test_two_phase_prune_api_matches_direct_construction():
    """Reference two-phase API (unet_2d_conditional.py:1367-1415 set_structure, then the module loop calling prune() /
    prune_module(), :2455-2461) run UNCHANGED on a full-width model == constructing at pruned width: same keys, same
    shapes, same parameter count (values: tests/test_unet_gpu.py::test_two_phase_prune_values)."""
    import torch

    from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
    torch.manual_seed(1)
    av = HyperStructure.get_random_arch_vector(0.55, structure_from_config(SD21_CONFIG))
    av[0, -3] = 0.1                                                  # one depth gate closed
    direct = UNet2DConditionModelPruned(arch_vector=av, device="meta", seed=None)
    model = UNet2DConditionModelPruned(arch_vector=None, device="meta", seed=None)
    assert not model.is_pruned() and model.num_parameters() == 865_910_724
    model.set_structure(HyperStructure.transform_arch_vector(av, model.get_structure()))
    for name, m in model.named_modules():                            # the reference's loop, verbatim
        if hasattr(m, "prune"):
            m.prune()
    for m in model.modules():
        if hasattr(m, "prune_module"):
            m.prune_module()
    sa, sb = direct.state_dict(), model.state_dict()
    assert set(sa) == set(sb) and all(sa[k].shape == sb[k].shape for k in sa)
    assert model.is_pruned() and model.num_parameters() == direct.num_parameters()
    assert torch.equal(model.arch_vector, direct.arch_vector)
    import pytest
    with pytest.raises(RuntimeError):
        model.set_structure(av)
        model.prune()                                                # pruning twice is an error, as in the reference


def test_encoder_state_dict_keys_match_their_dependencies():
    """The frozen step-front producers carry the state-dict keys and shapes of the dependency classes the reference loads
    (transformers CLIPTextModel with the SD-2.1 text config; diffusers AutoencoderKL encoder + quant_conv as restated in
    oracle/vae_restated.py), so real checkpoints load by key."""
    import torch
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HF

    from oracle.vae_restated import AutoencoderKLEncoder
    from unlearn_ft_b200.pdm.models import AutoencoderKL, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
                         max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, projection_dim=512)
    with torch.device("meta"):
        hf = HF(cfg)
        vo = AutoencoderKLEncoder()
    mine = CLIPTextModel(device="meta", seed=None)
    assert {k: tuple(v.shape) for k, v in hf.state_dict().items()} == {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    vae = AutoencoderKL(device="meta", seed=None)
    assert {k: tuple(v.shape) for k, v in vo.state_dict().items()} == {k: tuple(v.shape) for k, v in vae.state_dict().items()}
    assert sum(p.numel() for p in vae.parameters()) == 34_163_664


def test_encoder_from_pretrained_reads_the_dependency_folder_layout(tmp_path):
    """`CLIPTextModel.from_pretrained(root, subfolder="text_encoder")` on a folder written by transformers' own save_pretrained
    (config.json with many more keys than we use + model.safetensors): config values are taken over, every key is found (strict
    load).  Meta device: no GPU, no compute."""
    import pytest
    import torch
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HF

    from unlearn_ft_b200.pdm.models import CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                         max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, projection_dim=64)
    HF(cfg).save_pretrained(tmp_path / "text_encoder", safe_serialization=True)
    m = CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder", device="meta", revision=None)
    assert (m.config.hidden_size, m.config.num_hidden_layers, m.config.vocab_size) == (128, 2, 1000)
    assert len(m.text_model.encoder.layers) == 2
    (tmp_path / "text_encoder" / "model.safetensors").unlink()
    with pytest.raises(FileNotFoundError):
        CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder", device="meta")


def test_fused_gelu_constants_match_their_derivation():
    """The one-exponential erf-GELU of the fused GEGLU epilogue (csrc/common.cuh::gelu_erf_fast): its coefficients evaluated in float32
    exactly as the kernel does (Horner, ex2, clamp at 8) stay within 2e-6 of the exact erf-GELU over [-12, 12] -- far below the bf16
    rounding applied to the result.  Guards the constants in the source against drift from tools/fit_gelu.py.  CPU only."""
    import os
    import re

    import numpy as np
    from scipy.special import erf
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unlearn_ft_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_erf_fast(float x)"):]
    body = body[:body.index("\n}\n")]
    assert "fminf(fabsf(x), 8.f)" in body
    c = [np.float32(v) for v in re.findall(r"(-?\d\.\d{6,})f", body)]      # the polynomial, highest power first
    assert len(c) == 5
    c5, c4, c3, c2, c1 = c
    g = np.linspace(-12, 12, 400001).astype(np.float32)
    a = np.minimum(np.abs(g), np.float32(8.0))
    q = a * c5 + c4
    q = a * q + c3
    q = a * q + c2
    q = a * q + c1
    e = np.exp2(-(q * a)).astype(np.float32)
    r = (np.float32(0.5) * g * e).astype(np.float32)
    gelu = np.where(g >= 0, g - r, r)
    ref = 0.5 * g.astype(np.float64) * (1 + erf(g.astype(np.float64) / np.sqrt(2)))
    assert float(np.abs(gelu - ref).max()) < 2e-6


def test_fma_pipe_exp2_accuracy_claim():
    """csrc/common.cuh::exp2_fma (the share of the attention exponentials evaluated on the FMA pipe): Cody-Waite split + degree-3
    polynomial, restated in float32 numpy with the constants read from the source; maximum relative error below 1e-4 on the range the
    softmax uses (x <= 8; far below bf16's 2^-9), exact zero-flush behaviour not required.  CPU only."""
    import os
    import re

    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unlearn_ft_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float exp2_fma(float x)"):]
    body = body[:body.index("\n}\n")]
    k3, k2 = [np.float32(v) for v in re.search(r"fmaf\(f, (\d\.\d+)f, (\d\.\d+)f\)", body).groups()]
    k1, k0 = [np.float32(v) for v in re.findall(r"fmaf\(f, p, (\d\.\d+)f\)", body)]
    x = np.linspace(-100, 8, 2000001).astype(np.float32)
    magic = np.float32(12582912.0)
    t = (x + magic).astype(np.float32)
    f = (x - (t - magic)).astype(np.float32)
    p = (f * k3 + k2).astype(np.float32)
    p = (f * p + k1).astype(np.float32)
    p = (f * p + k0).astype(np.float32)
    y = (p.view(np.int32) + (t.view(np.int32) << 23)).view(np.float32)
    ref = np.exp2(x.astype(np.float64))
    assert float(np.abs(y / ref - 1).max()) < 1e-4
