"""Image metrics used by validation -- /root/reference/src/python/utils.py:420-491 -- on the fused kernels:
one spaa_ssim_l1_fwd_bwd launch yields MSE and SSIM, one spaa_color_loss_fwd_bwd launch yields L2 and dE."""
from __future__ import annotations

import math
import os

import torch

from . import ops
from .img_proc import expand_4d


def _dev(x, y):
    x, y = expand_4d(x), expand_4d(y)
    if not x.is_cuda:
        x = x.cuda()
    if not y.is_cuda or y.device != x.device:
        y = y.to(x.device)
    return ops._f32c(x), ops._f32c(y)


def calc_img_dists(x, y):
    """utils.py:420-423: (PSNR, RMSE, SSIM, L2, Linf, dE) as python floats."""
    with torch.no_grad():
        x, y = _dev(x, y)
        n = x.numel()
        sums, _, _ = ops.ssim_l1(x, y, 0.0, 0.0, 0.0, want_grad=False)
        B, _, H, W = x.shape
        stats, _ = ops.color_loss(x, y, ops.rgb2lab(y), cam_is_lab2=False, de_weighting=False, c_de=0.0, c_l2=0.0, want_grad=False)
        linf = (x - y).abs().amax(dim=1).mean() * 255
        vals = torch.stack((sums[1] / n, sums[2] / n, stats[:, 1].sum() / (B * H * W) * 255, linf, stats[:, 0].sum() / (B * H * W))).tolist()
    mse, ssim_v, l2, linf_v, de = vals
    return 10 * math.log10(1 / mse), math.sqrt(mse * 3), ssim_v, l2, linf_v, de


def psnr(x, y):
    return calc_img_dists(x, y)[0]


def rmse(x, y):
    return calc_img_dists(x, y)[1]


def ssim(x, y):
    return calc_img_dists(x, y)[2]


def l2_norm(x, y):
    return calc_img_dists(x, y)[3]


def linf_norm(x, y):
    return calc_img_dists(x, y)[4]


def opt_to_string(opt):
    """utils.py:674-675: checkpoint / log title."""
    return (f'{opt["setup_name"]}_{opt["model_name"]}_{opt["loss"]}_{opt["num_train"]}_{opt["batch_size"]}_{opt["max_iters"]}_'
            f'{opt["lr"]}_{opt["lr_drop_ratio"]}_{opt["lr_drop_rate"]}_{opt["l2_reg"]}')


def save_checkpoint(checkpoint_dir, model, title):
    """utils.py:717-721: state_dict only, same file naming."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    fn = os.path.abspath(os.path.join(checkpoint_dir, title + ".pth"))
    torch.save(model.state_dict(), fn)
    print(f"Checkpoint saved to {fn}\n")
    return fn
