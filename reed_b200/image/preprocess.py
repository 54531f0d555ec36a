"""Raw-image preprocessing for the frozen target encoders: drop-in for ``preprocess_raw_image`` of the reference
(/root/reference/image/train.py:53-74), one fused kernel (``reed_preprocess_image``) instead of divide / Normalize /
``F.interpolate(mode='bicubic')``.

This is the first half of SURVEY 8(f) row 3 (the encoder forward is ``reed_b200.image.encoders``).  The formula is pinned
on the CPU by ``oracle/preprocess_oracle.py`` against the reference function; the kernel is checked against it on the B200
(tests/test_zz_next_gpu.py).
"""
from __future__ import annotations

import ctypes

import torch

from .. import ops

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)                 # timm.data constants imported at train.py:31
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
CLIP_DEFAULT_MEAN = (0.48145466, 0.4578275, 0.40821073)       # train.py:37
CLIP_DEFAULT_STD = (0.26862954, 0.26130258, 0.27577711)       # train.py:38


def _plan(enc_type: str, resolution: int):
    """(mean, std, out_size, resize_first) of train.py:55-72, or None for encoder types the reference leaves untouched."""
    resized = 224 * (resolution // 256)
    if "clip" in enc_type:
        return CLIP_DEFAULT_MEAN, CLIP_DEFAULT_STD, resized, 1
    if "mocov3" in enc_type or "mae" in enc_type:
        return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, resolution, 0
    if "dinov2" in enc_type:
        return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, resized, 0
    if "dinov1" in enc_type:                                      # train.py:66
        return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, resolution, 0
    if "jepa" in enc_type:
        return IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, resized, 0
    return None


def preprocess_raw_image(x: torch.Tensor, enc_type: str, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """x: [B, 3, R, R] uint8 (as the data loader delivers it) or float, values 0..255, on the GPU."""
    plan = _plan(enc_type, x.shape[-1])
    if plan is None:
        return x
    if not x.is_cuda:
        raise RuntimeError("reed_b200 preprocess_raw_image runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if x.dim() != 4 or x.shape[-1] != x.shape[-2] or x.shape[1] != 3:
        raise ValueError("expected a [B, 3, R, R] image batch (the normalisation constants are per RGB channel)")
    mean, std, out_size, resize_first = plan
    if out_size < 1:
        raise ValueError(f"resolution {x.shape[-1]} is below 256: the reference resizes to 224 * (resolution // 256) = 0")
    if x.dtype != torch.uint8:
        x = x.float()
    x = x.contiguous()
    B, C = x.shape[0], x.shape[1]
    out = torch.empty((B, C, out_size, out_size), device=x.device, dtype=out_dtype)
    arr = ctypes.c_float * C
    ops._launch("reed_preprocess_image", ops._p(x), 2 if x.dtype == torch.uint8 else 0, ops._p(out), ops._code(out_dtype), B, C,
                x.shape[-1], out_size, ctypes.cast(arr(*mean[:C]), ctypes.c_void_p), ctypes.cast(arr(*std[:C]), ctypes.c_void_p),
                resize_first, ops._stream())
    return out
