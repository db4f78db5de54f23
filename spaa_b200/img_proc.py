"""Tensor helpers used inside the hot loops -- /root/reference/src/python/img_proc.py:110-132."""
import torch
import torch.nn.functional as F


def expand_4d(x):
    """:110-114."""
    while x.ndim < 4:
        x = x.unsqueeze(0)
    return x


def resize(x, size):
    """:117-123: area interpolation (adaptive average pooling)."""
    if x.ndim == 3:
        return F.interpolate(x[None], size, mode="area")[0]
    return F.interpolate(x, size, mode="area")


def center_crop(x, size):
    """:126-132: offsets int(round((h - th) / 2)) (Python rounding)."""
    h, w = x.shape[-2:]
    th, tw = size
    i = int(round((h - th) / 2.0))
    j = int(round((w - tw) / 2.0))
    return x[..., i:i + th, j:j + tw]
