"""ctypes binding of libreed_sm100.so (C-ABI declared in include/reed_b200.h).

The product has no CPU or PyTorch fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# REED_LIB: profiling knob - load an experimental build of the same library (profiles/*.py variants)
LIB_PATH = os.environ.get("REED_LIB") or os.path.join(_HERE, "libreed_sm100.so")

P, I, L, F, D = c_void_p, c_int, c_int64, c_float, c_double

# name -> argtypes (every function returns int; 0 = ok)
_SIGNATURES = {
    "reed_device_check": [c_char_p, I],
    "reed_gemm_reserve_sms": [I],
    "reed_gemm_tcgen05_launches": [P],
    "reed_gemm": [I, P, L, I, P, L, I, P, L, I, I, I, I, I, P, P, L, P, L, I, P, L, I, I, P],
    "reed_attn_fwd": [I, P, P, P, I, I, I, I, I, P],
    "reed_attn_bwd": [I, P, P, P, P, P, P, I, I, I, I, I, P],
    "reed_qk_norm_fwd": [P, I, P, P, P, P, P, P, L, I, I, F, P],
    "reed_qk_norm_bwd": [P, I, P, P, P, P, P, P, P, P, P, L, I, I, P],
    "reed_ln_modulate_fwd": [P, P, P, L, I, P, L, I, P, P, I, I, F, P],
    "reed_gemm_wgrad_bias": [P, L, P, L, P, L, P, I, I, I, I, P],
    "reed_gemm_grouped": [I, P, L, P, L, I, I, P, L, I, I, I, P, I, P],
    "reed_outer_wgrad": [P, I, L, P, I, L, P, L, P, I, I, I, I, P],
    "reed_ln_modulate_bwd": [P, I, P, P, P, P, L, L, I, P, P, P, P, I, I, P],
    "reed_ln_modulate_gate_bwd": [P, I, P, P, P, P, L, L, I, P, P, P, P, P, P, P, P, P, I, I, P],
    "reed_gate_bwd": [P, P, I, P, L, L, I, P, P, P, I, I, P],
    "reed_colsum": [P, I, L, P, I, I, P],
    "reed_unary": [P, I, P, I, I, L, P],
    "reed_act_bwd": [P, I, P, I, P, I, L, P],
    "reed_group_mean_fwd": [P, P, I, I, I, I, P],
    "reed_group_mean_bwd": [P, P, I, I, I, I, P],
    "reed_add_f32": [P, P, P, L, P],
    "reed_siloss_interp": [P, P, P, P, I, I, I, P],
    "reed_siloss_mse_fwd": [P, P, P, P, P, I, I, I, P],
    "reed_siloss_mse_bwd": [P, P, P, P, P, P, I, I, I, P],
    "reed_siloss_cos_fwd": [P, I, P, I, P, P, I, I, I, P],
    "reed_siloss_cos_bwd": [P, I, P, I, P, P, P, I, I, I, P],
    "reed_sampler_step": [P, P, I, P, P, P, P, P, L, I, I, I, I, D, D, D, P],
    "reed_sampler_cast": [P, P, I, L, I, P],
    "reed_preprocess_image": [P, I, P, I, I, I, I, I, P, P, I, P],
    "reed_sample_posterior": [P, P, P, P, F, F, P, I, I, I, P],
    "reed_grad_sumsq": [P, L, P, P],
    "reed_adamw_ema": [P, P, P, P, P, P, L, P, F, F, F, F, F, F, F, I, F, P, P],
    "reed_ema_update": [P, P, L, F, P],
    "reed_nvls_reduce_scatter_sumsq": [P, P, L, L, P, I, P],
    "reed_adamw_ema_mc": [P, P, P, P, P, P, L, P, F, F, F, F, F, F, F, I, F, P, P],
}

EXPORTS = sorted(list(_SIGNATURES) + ["reed_version", "reed_last_error"])

_lib = None


class ReedLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ReedLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C reed_b200/csrc). reed_b200 has no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.reed_version.restype = c_int
    lib.reed_version.argtypes = []
    lib.reed_last_error.restype = c_char_p
    lib.reed_last_error.argtypes = []
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = args
    _lib = lib
    return lib


def call(name: str, *args):
    lib = _lib if _lib is not None else load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise ReedLibraryError(f"{name} failed: {lib.reed_last_error().decode(errors='replace')}")


def version() -> int:
    return load().reed_version()
