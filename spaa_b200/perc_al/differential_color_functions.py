"""Differentiable sRGB->Lab and the reference's CIEDE2000 variant -- API of
/root/reference/src/python/perc_al/differential_color_functions.py:12-190 on the sm_100a colour kernels
(one launch per function and per backward instead of ~150 elementwise kernels)."""
from __future__ import annotations

import torch

from .. import ops
from ..img_proc import expand_4d


class _Rgb2LabFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb):
        ctx.save_for_backward(rgb)
        return ops.rgb2lab(rgb)

    @staticmethod
    def backward(ctx, dlab):
        rgb, = ctx.saved_tensors
        return ops.rgb2lab_bwd(rgb, dlab)


class _DE2000Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lab1, lab2):
        ctx.save_for_backward(lab1, lab2)
        return ops.de2000(lab1, lab2)

    @staticmethod
    def backward(ctx, cot):
        lab1, lab2 = ctx.saved_tensors
        d1, d2 = ops.de2000_bwd(lab1, lab2, cot, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return d1, d2


def rgb2lab_diff(rgb_image, device=None):
    """:39-64.  Bx3xHxW sRGB in [0,1] -> Lab."""
    ops._need_cuda(rgb_image)
    return _Rgb2LabFn.apply(ops._f32c(rgb_image))


def ciede2000_diff(lab1, lab2, device=None):
    """:109-180.  Bx3xHxW x2 -> BxHxW."""
    ops._need_cuda(lab1, lab2)
    return _DE2000Fn.apply(ops._f32c(lab1), ops._f32c(lab2))


def deltaE(x, y):
    """:183-190: mean dE over batch and pixels, no grad, python float."""
    with torch.no_grad():
        x, y = expand_4d(x), expand_4d(y)
        if not x.is_cuda:
            x, y = x.cuda(), y.cuda()
        B, _, H, W = x.shape
        stats, _ = ops.color_loss(x, y, ops.rgb2lab(y), cam_is_lab2=False, de_weighting=False, c_de=0.0, c_l2=0.0, want_grad=False)
        return (stats[:, 0].sum() / (B * H * W)).item()
