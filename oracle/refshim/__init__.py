"""TEST INFRASTRUCTURE (builder container only) -- lets the reference's OWN python run without diffusers.

`install()` registers a fake `diffusers` package in `sys.modules` whose classes are the restatements in
`oracle/diffusers_restated.py`, and stub `pdm` / `pdm.models` packages that point at /root/reference WITHOUT running
`pdm/models/__init__.py` (which pulls in FLUX / quantizer code that needs far more of diffusers).  After that,
`import pdm.models.unet.unet_2d_conditional` executes the reference's real gated blocks, `prune()` methods, structure
plumbing and `UNet2DConditionModelPruned.from_pretrained` on top of our base classes.  `oracle/make_golden.py` uses
this to write tests/golden/*.pt.  /root/reference does not exist on the GPU box, so nothing at test time imports this
module except the (skipped-when-absent) live cross-check in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UNLEARN_FT_REFERENCE", "/root/reference")


class _AutoModule(types.ModuleType):
    """Module whose unknown attributes resolve to fresh placeholder classes (for names the reference imports but
    never uses on the SD-2.1 path)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        placeholder = type(name, (), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
            NotImplementedError(f"diffusers.{name} is not restated"))})
        setattr(self, name, placeholder)
        return placeholder


class _Logger:
    def __getattr__(self, name):
        return lambda *a, **k: None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pdm"))


def install():
    from .. import diffusers_restated as D
    import torch
    from torch import nn

    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_is_refshim", False):
        raise RuntimeError("a real diffusers is already imported")

    def mod(name, **attrs):
        m = _AutoModule(name)
        m.__path__ = []  # mark as package so that submodule imports resolve through sys.modules
        m._is_refshim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    logging = types.SimpleNamespace(get_logger=lambda *a, **k: _Logger())

    class UNet2DConditionLoadersMixin:
        pass

    def deprecate(*a, **k):
        return None

    class ModelMixin(D.ModelMixin):
        def _set_gradient_checkpointing(self, *a, **k):
            pass

    mod("diffusers", __version__="0.30.3", ModelMixin=ModelMixin, ConfigMixin=D.ConfigMixin)
    mod("diffusers.configuration_utils", register_to_config=D.register_to_config, ConfigMixin=D.ConfigMixin)
    mod("diffusers.models", Transformer2DModel=D.Transformer2DModel)
    mod("diffusers.models.activations", GEGLU=D.GEGLU, get_activation=D.get_activation)
    mod("diffusers.models.resnet", ResnetBlock2D=D.ResnetBlock2D, Upsample2D=D.Upsample2D, Downsample2D=D.Downsample2D)
    mod("diffusers.models.transformers")
    mod("diffusers.models.transformers.transformer_2d", Transformer2DModelOutput=D.Transformer2DModelOutput)
    mod("diffusers.models.attention", BasicTransformerBlock=D.BasicTransformerBlock, FeedForward=D.FeedForward)
    mod("diffusers.models.unets")
    mod("diffusers.models.unets.unet_2d_blocks", CrossAttnDownBlock2D=D.CrossAttnDownBlock2D,
        CrossAttnUpBlock2D=D.CrossAttnUpBlock2D, DownBlock2D=D.DownBlock2D, UpBlock2D=D.UpBlock2D,
        UNetMidBlock2DCrossAttn=D.UNetMidBlock2DCrossAttn)
    mod("diffusers.models.unets.unet_2d_condition", UNet2DConditionOutput=D.UNet2DConditionOutput,
        UNet2DConditionModel=D.UNet2DConditionModel)
    mod("diffusers.models.attention_processor", AttnProcessor2_0=D.AttnProcessor2_0, Attention=D.Attention,
        ADDED_KV_ATTENTION_PROCESSORS=(), CROSS_ATTENTION_PROCESSORS=())
    mod("diffusers.models.embeddings", TimestepEmbedding=D.TimestepEmbedding, Timesteps=D.Timesteps)
    mod("diffusers.loaders", UNet2DConditionLoadersMixin=UNet2DConditionLoadersMixin)
    mod("diffusers.utils", logging=logging, deprecate=deprecate, is_torch_npu_available=lambda: False,
        is_torch_version=lambda *a: True, _get_model_file=None, _add_variant=None)
    import typing
    mu = mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin, _LOW_CPU_MEM_USAGE_DEFAULT=False, torch=torch,
             nn=nn, os=os, load_state_dict=None)
    for n in ("Any", "Callable", "Dict", "List", "Optional", "Tuple", "Union"):
        setattr(mu, n, getattr(typing, n))
    mu.__all__ = ["ModelMixin", "_LOW_CPU_MEM_USAGE_DEFAULT", "torch", "nn", "os", "load_state_dict", "Any", "Callable",
                  "Dict", "List", "Optional", "Tuple", "Union"]

    # stub packages for the reference so that pdm/models/__init__.py (FLUX, quantizer) is never executed
    for name, rel in (("pdm", "pdm"), ("pdm.models", "pdm/models")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            sys.modules[name] = m


def load_reference():
    """Returns the reference's own modules (executed from /root/reference)."""
    install()
    import importlib

    unet = importlib.import_module("pdm.models.unet.unet_2d_conditional")
    blocks = importlib.import_module("pdm.models.unet.blocks")
    hypernet = importlib.import_module("pdm.models.hypernet")
    gates = importlib.import_module("pdm.models.gates")
    est = importlib.import_module("pdm.utils.estimation_utils")
    metric = importlib.import_module("pdm.utils.metric_utils")
    return types.SimpleNamespace(unet=unet, blocks=blocks, hypernet=hypernet, gates=gates, estimation=est, metric=metric)
