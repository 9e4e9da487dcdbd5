"""ctypes binding of the C-ABI library ``libb200pdm.so`` (declared in ``include/b200pdm.h``).

The product path has NO fallback: if the shared library is missing or a call returns a non-zero status this module
raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_SO = _PKG / os.environ.get("B200PDM_LIB", "libb200pdm.so")   # (libb200pdm_diag.so = `make -C csrc diag`, tools/diag_*.py only)

c_p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float
sz = C.c_size_t


class Operand(C.Structure):
    """Mirror of ``b200pdm_operand``."""

    _fields_ = [
        ("mode", i32), ("ptr", c_p), ("ld", i64), ("bs1", i64), ("bs2", i64),
        ("batch", i32), ("h_in", i32), ("w_in", i32), ("channels", i32),
        ("h_out", i32), ("w_out", i32), ("stride", i32), ("taps", i32), ("flip", i32), ("out_channels", i32),
        ("no_pad", i32),
    ]


class GemmDesc(C.Structure):
    """Mirror of ``b200pdm_gemm_desc``."""

    _fields_ = [
        ("a", Operand), ("b", Operand), ("M", i64), ("N", i64), ("K", i64), ("Z1", i32), ("Z2", i32),
        ("out", c_p), ("out_fp32", i32), ("ldo", i64), ("obs1", i64), ("obs2", i64),
        ("bias", c_p), ("rowbias", c_p), ("ld_rowbias", i64), ("rows_per_group", i32),
        ("residual", c_p), ("ldr", i64), ("rbs1", i64), ("rbs2", i64),
        ("alpha", f32), ("accumulate", i32), ("splits", i32), ("block_n", i32),
        ("geglu", i32), ("aux", c_p), ("ld_aux", i64),
    ]


class FeaturePair(C.Structure):
    """Mirror of ``b200pdm_feature_pair``."""

    _fields_ = [("s", c_p), ("t", c_p), ("ds", c_p), ("numel", i64)]


OP_K2D, OP_MN2D, OP_CONV_ACT, OP_CONV_W, OP_CONV_WT, OP_CONV_ACT_MN = range(6)

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGS = {
    "b200pdm_version": [],
    "b200pdm_last_error": [],
    "b200pdm_launch_count": [],
    "b200pdm_gemm_plan": [i64, i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int)],
    "b200pdm_gemm_trace_dump": [C.c_char_p],
    "b200pdm_gemm_trace_enable": [i32],
    "b200pdm_gemm_trace_totals": [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)],
    "b200pdm_gemm_workspace": [C.POINTER(GemmDesc)],
    "b200pdm_gemm": [C.POINTER(GemmDesc), c_p, sz, c_p],
    "b200pdm_linear_fwd_workspace": [i64, i64, i64, i32],
    "b200pdm_linear_fwd": [c_p, i64, c_p, i64, c_p, c_p, i64, c_p, i64, i32, i64, i64, i64, c_p, sz, c_p],
    "b200pdm_linear_geglu_fwd": [c_p, i64, c_p, i64, c_p, c_p, i64, c_p, i64, i64, i64, i64, c_p],
    "b200pdm_linear_dgrad_workspace": [i64, i64, i64],
    "b200pdm_linear_dgrad": [c_p, i64, c_p, i64, c_p, i64, c_p, i64, i64, i64, i64, c_p, sz, c_p],
    "b200pdm_linear_wgrad": [c_p, i64, c_p, i64, c_p, i64, i64, i64, i64, c_p],
    "b200pdm_conv_fwd_workspace": [i32, i32, i32, i32, i32, i32, i32],
    "b200pdm_conv_fwd": [c_p, i64, c_p, i64, c_p, c_p, i64, c_p, i64, c_p, i64, i32, i32, i32, i32, i32, i32, i32, c_p, sz, c_p],
    "b200pdm_conv_fwd_nopad": [c_p, i64, c_p, i64, c_p, c_p, i64, i32, i32, i32, i32, i32, i32, c_p],
    "b200pdm_conv_dgrad_workspace": [i32, i32, i32, i32, i32, i32],
    "b200pdm_conv_dgrad": [c_p, i64, c_p, i64, c_p, i64, c_p, i64, i32, i32, i32, i32, i32, i32, c_p, sz, c_p],
    "b200pdm_conv_wgrad": [c_p, i64, c_p, i64, c_p, i64, i32, i32, i32, i32, i32, i32, i32, c_p],
    "b200pdm_groupnorm_scratch_floats": [i32, i32, i32],
    "b200pdm_groupnorm_bwd_workspace_floats": [i32, i32, i32],
    "b200pdm_groupnorm_fwd": [c_p, i64, c_p, c_p, c_p, i64, c_p, c_p, i32, i32, i32, i32, f32, i32, c_p],
    "b200pdm_groupnorm_bwd": [c_p, i64, c_p, i64, c_p, c_p, c_p, c_p, i64, c_p, i64, c_p, c_p, c_p, i32, i32, i32, i32, i32, c_p],
    "b200pdm_layernorm_fwd": [c_p, i64, c_p, c_p, c_p, i64, c_p, c_p, i64, i32, f32, c_p],
    "b200pdm_layernorm_bwd": [c_p, i64, c_p, i64, c_p, c_p, c_p, c_p, i64, c_p, i64, c_p, c_p, i64, i32, c_p],
    "b200pdm_geglu_fwd": [c_p, i64, c_p, i64, i64, i32, c_p],
    "b200pdm_geglu_bwd": [c_p, i64, c_p, i64, c_p, i64, i64, i32, c_p],
    "b200pdm_softmax_fwd": [c_p, i64, c_p, i64, i64, i32, f32, c_p],
    "b200pdm_attention_fwd": [c_p, i64, c_p, i64, c_p, i64, c_p, i64, c_p, i32, i32, i32, i32, f32, c_p],
    "b200pdm_attention_bwd": [c_p, i64, c_p, i64, c_p, i64, c_p, i64, c_p, i64, c_p, c_p, i64, c_p, i64, c_p, i64, c_p,
                              i32, i32, i32, i32, f32, c_p],
    "b200pdm_attention_fwd_ex": [c_p, i64, c_p, i64, c_p, i64, c_p, i64, c_p, i32, i32, i32, i32, f32, i32, c_p],
    "b200pdm_gate_scale": [c_p, i64, c_p, i32, c_p, i64, i64, i32, i32, i32, i32, i32, c_p],
    "b200pdm_gate_grad": [c_p, i64, c_p, i64, c_p, i32, i64, i32, i32, i32, i32, i32, c_p],
    "b200pdm_depth_blend": [c_p, i64, c_p, i64, c_p, c_p, i64, i64, i32, i32, i32, c_p],
    "b200pdm_depth_blend_bwd": [c_p, i64, c_p, i64, c_p, i64, c_p, c_p, i64, c_p, i64, c_p, i64, i32, i32, i32, c_p],
    "b200pdm_gelu": [c_p, i64, c_p, i64, i64, i32, c_p],
    "b200pdm_clip_embed": [c_p, c_p, c_p, c_p, i64, i64, i32, i32, i32, c_p],
    "b200pdm_vae_sample": [c_p, i64, c_p, c_p, c_p, i32, i32, i32, f32, c_p],
    "b200pdm_colsum": [c_p, i64, c_p, i64, i32, c_p],
    "b200pdm_colsum_grouped": [c_p, i64, c_p, i64, i64, i32, i32, c_p],
    "b200pdm_cast_f32_to_bf16": [c_p, c_p, i64, c_p],
    "b200pdm_add": [c_p, i64, c_p, i64, c_p, i64, i64, i32, c_p],
    "b200pdm_copy2d": [c_p, i64, c_p, i64, i64, i32, c_p],
    "b200pdm_silu_f32_to_bf16": [c_p, c_p, i64, c_p],
    "b200pdm_silu_bwd": [c_p, c_p, c_p, i64, c_p],
    "b200pdm_upsample2x_fwd": [c_p, i64, c_p, i64, i32, i32, i32, i32, c_p],
    "b200pdm_upsample2x_bwd": [c_p, i64, c_p, i64, i32, i32, i32, i32, c_p],
    "b200pdm_zero_insert2x": [c_p, i64, c_p, i64, i32, i32, i32, i32, c_p],
    "b200pdm_nchw_f32_to_nhwc_bf16": [c_p, c_p, i64, i32, i32, i32, c_p],
    "b200pdm_nhwc_bf16_to_nchw_f32": [c_p, i64, c_p, i32, i32, i32, c_p],
    "b200pdm_timestep_embedding": [c_p, c_p, i64, i32, i32, c_p],
    "b200pdm_kd_loss_workspace": [i32],
    "b200pdm_kd_loss_fused": [c_p, c_p, c_p, c_p, c_p, c_p, f32, i32, c_p, i32, i64, f32, f32, C.POINTER(FeaturePair), i32,
                              f32, c_p, c_p, C.c_size_t, c_p],
    "b200pdm_adamw_step": [c_p, c_p, c_p, c_p, c_p, i64, f32, f32, f32, f32, f32, i64, f32, i32, c_p],
    "b200pdm_cfg_ddim_step": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i64, i32, i32, f32, c_p],
    "b200pdm_cfg_pndm_step": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i64, i32, i32, f32, c_p],
    "b200pdm_adamw_step_dyn": [c_p, c_p, c_p, c_p, c_p, i64, c_p, f32, f32, f32, f32, f32, i32, c_p],
    "b200pdm_refresh_shadow": [c_p, c_p, i64, c_p],
    "b200pdm_refresh_shadow_zero": [c_p, c_p, c_p, i64, c_p],
    "b200pdm_diffusion_prep": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, i32, i64, c_p],
}
_RESTYPES = {"b200pdm_last_error": C.c_char_p, "b200pdm_launch_count": C.c_uint64}
_RESTYPES.update({n: C.c_size_t for n in _SIGS if n.endswith(("_workspace", "_floats"))})

EXPORTED_SYMBOLS = tuple(_SIGS)


class B200PdmError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles on a CPU-only box)."""
    out = subprocess.run(["make", "-C", str(_PKG / "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0 or not _SO.exists():
        raise B200PdmError("building libb200pdm.so failed (nvcc -gencode arch=compute_100a,code=sm_100a)")
    return _SO


_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the ctypes handle.  Raises loudly when the native library is absent."""
    global _lib
    if _lib is None:
        if not _SO.exists():
            if os.environ.get("B200PDM_AUTOBUILD", "1") == "1":
                build()
            else:
                raise B200PdmError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        h = C.CDLL(str(_SO))
        for name, args in _SIGS.items():
            fn = getattr(h, name)  # AttributeError if the .so does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = h
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b200pdm_last_error()
        raise B200PdmError(f"b200pdm call {what} failed with status {rc}: {msg.decode() if msg else ''}")


def gemm_plan(n, n_groups=1, b_mn=False, tiles_m=1, Z=1, kblocks=1, can_split=False, split_needs_finalize=True) -> dict:
    """Tile plan the GEMM core would choose (host-only query; works without a GPU)."""
    out = (C.c_int * 7)()
    check(lib().b200pdm_gemm_plan(n, n_groups, int(b_mn), tiles_m, Z, kblocks, int(can_split), int(split_needs_finalize), out))
    return dict(zip(("block_n", "splits", "pair", "m_sub", "stages", "tiles", "slots"), list(out)))


def launch_count() -> int:
    return int(lib().b200pdm_launch_count())
