"""B200-native (sm_100a) hot path for rezashkv/unlearn-ft: pruned SD-2.1 U-Net DDPM+KD training step.

Package layout: ``csrc/`` (hand-written CUDA kernels + the C ABI of ``include/b200pdm.h``), ``_lib.py`` (ctypes
binding), ``kernels.py`` (tensor-level wrappers), ``pdm/`` (mirror of the reference's ``pdm`` class surface).
"""
__version__ = "0.1.0"
