"""CPU restatement of the raw-image preprocessing in front of the frozen target encoders.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/image/train.py:53-74 (``preprocess_raw_image``): ``x / 255``, torchvision ``Normalize`` with the
ImageNet (timm ``IMAGENET_DEFAULT_MEAN/STD``, train.py:31) or CLIP (train.py:37-38) statistics, and for the dinov2 / jepa /
clip encoders ``F.interpolate(x, 224 * (resolution // 256), mode='bicubic')`` - normalise-then-resize for dinov2 / jepa,
resize-then-normalise for clip.  The resize is written out tap by tap (ATen ``upsample_bicubic2d``: align_corners=False,
A = -0.75, source index ``scale * (dst + 0.5) - 0.5`` unclamped, taps clamped to the image) so that the CUDA kernel has a
formula-level oracle; it is pinned against the reference function itself in tests/golden/preprocess.pt.
"""
from __future__ import annotations

import torch

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)       # timm.data.constants (published values)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
CLIP_DEFAULT_MEAN = (0.48145466, 0.4578275, 0.40821073)      # train.py:37
CLIP_DEFAULT_STD = (0.26862954, 0.26130258, 0.27577711)      # train.py:38

_A = -0.75


def _cubic_weights(t: torch.Tensor):
    """ATen get_cubic_upsample_coefficients: taps at offsets -1, 0, +1, +2 around floor(src)."""
    def near(x):      # |x| <= 1
        return ((_A + 2) * x - (_A + 3)) * x * x + 1

    def far(x):       # 1 < |x| < 2
        return ((_A * x - 5 * _A) * x + 8 * _A) * x - 4 * _A

    return [far(t + 1), near(t), near(1 - t), far(2 - t)]


def bicubic_resize(x: torch.Tensor, out_size: int) -> torch.Tensor:
    """x [B, C, H, W] float32 -> [B, C, out, out], F.interpolate(mode='bicubic', align_corners=False, antialias=False)."""
    B, C, H, W = x.shape

    def axis(n_in):
        scale = n_in / out_size
        src = scale * (torch.arange(out_size, dtype=torch.float32) + 0.5) - 0.5
        base = torch.floor(src)
        t = src - base
        idx = torch.stack([(base.long() + k).clamp(0, n_in - 1) for k in (-1, 0, 1, 2)])      # [4, out]
        return idx, torch.stack(_cubic_weights(t))                                            # [4, out]

    iy, wy = axis(H)
    ix, wx = axis(W)
    rows = x[:, :, iy, :]                                  # [B, C, 4, out_y, W]
    taps = rows[..., ix]                                   # [B, C, 4, out_y, 4, out_x]
    horiz = (taps * wx.view(1, 1, 1, 1, 4, out_size)).sum(dim=4)          # [B, C, 4, out_y, out_x]
    return (horiz * wy.view(1, 1, 4, out_size, 1)).sum(dim=2)


def _normalize(x, mean, std):
    m = torch.tensor(mean, dtype=x.dtype).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=x.dtype).view(1, -1, 1, 1)
    return (x - m) / s


def preprocess_raw_image(x: torch.Tensor, enc_type: str) -> torch.Tensor:
    """train.py:53-74 for a [B, 3, R, R] uint8 (or float) image batch with values 0..255."""
    resolution = x.shape[-1]
    out = 224 * (resolution // 256)
    if "clip" in enc_type:
        return _normalize(bicubic_resize(x / 255., out), CLIP_DEFAULT_MEAN, CLIP_DEFAULT_STD)
    if "mocov3" in enc_type or "mae" in enc_type:
        return _normalize(x / 255., IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)
    if "dinov2" in enc_type:
        return bicubic_resize(_normalize(x / 255., IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD), out)
    if "dinov1" == enc_type:
        return _normalize(x / 255., IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)
    if "jepa" in enc_type:
        return bicubic_resize(_normalize(x / 255., IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD), out)
    return x                                              # any other encoder type: untouched (train.py:74)
