"""SSIM -- API of /root/reference/src/python/pytorch_ssim/__init__.py:9-107 on the fused sm_100a kernel
(spaa_ssim_l1_fwd_bwd): replicate padding, 11x11 sigma=1.5 Gaussian moments, C1=1e-4, C2=9e-4, forward and backward
in one launch each."""
from __future__ import annotations

from math import exp

import torch

from .. import ops


def gaussian(window_size, sigma):
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):
    g = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


class _SsimMapFn(torch.autograd.Function):
    """ssim_map(img1, img2); gradient flows to img1 (the prediction) only, as in training."""

    @staticmethod
    def forward(ctx, img1, img2):
        ctx.save_for_backward(img1, img2)
        _, _, smap = ops.ssim_l1(img1, img2, 0.0, 0.0, 0.0, want_grad=False, want_map=True)
        return smap

    @staticmethod
    def backward(ctx, dmap):
        img1, img2 = ctx.saved_tensors
        g1 = g2 = None
        if ctx.needs_input_grad[0]:
            _, g1, _ = ops.ssim_l1(img1, img2, 0.0, 0.0, 0.0, cot_map=dmap.contiguous())
        if ctx.needs_input_grad[1]:
            _, g2, _ = ops.ssim_l1(img2, img1, 0.0, 0.0, 0.0, cot_map=dmap.contiguous())      # SSIM is symmetric
        return g1, g2


class _SsimMeanFn(torch.autograd.Function):
    """mean(ssim_map) with forward value and d/d(img1) produced by ONE kernel launch."""

    @staticmethod
    def forward(ctx, img1, img2):
        need = img1.requires_grad
        sums, grad, _ = ops.ssim_l1(img1, img2, 0.0, 0.0, -1.0, want_grad=need)     # w_ssim=-1 -> grad = +d(mean)/d(img1)
        ctx.grad = grad
        ctx.save_for_backward(img1, img2)
        return sums[2] / img1.numel()

    @staticmethod
    def backward(ctx, d):
        g1 = g2 = None
        if ctx.needs_input_grad[0]:
            g1 = ctx.grad * d
        if ctx.needs_input_grad[1]:
            img1, img2 = ctx.saved_tensors
            _, g2, _ = ops.ssim_l1(img2, img1, 0.0, 0.0, -1.0)
            g2 = g2 * d
        return g1, g2


def _ssim(img1, img2, window, window_size, channel, size_average=True, mask=None, weights=None):
    if window_size != 11:
        raise NotImplementedError("the fused SSIM kernel implements the reference's 11x11 window")
    ops._need_cuda(img1, img2)
    img1, img2 = ops._f32c(img1), ops._f32c(img2)
    if size_average and mask is None and weights is None:
        return _SsimMeanFn.apply(img1, img2)
    ssim_map = _SsimMapFn.apply(img1, img2)
    if weights is not None:
        ssim_map = ssim_map * weights.expand_as(ssim_map)
    if size_average:
        return ssim_map[mask].mean()
    if mask is not None:
        return (ssim_map * mask).mean(1).mean(1).mean(1)
    return ssim_map.mean(1).mean(1).mean(1)


class SSIM(torch.nn.Module):
    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1
        self.register_buffer("window", create_window(window_size, self.channel))

    def forward(self, img1, img2, mask=None, weights=None):
        return _ssim(img1, img2, self.window, self.window_size, img1.shape[1], self.size_average, mask=mask, weights=weights)


def ssim(img1, img2, window_size=11, size_average=True):
    return _ssim(img1, img2, None, window_size, img1.shape[1], size_average)
