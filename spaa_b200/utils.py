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


# ------------------------------------------------------------------------------------------------------------
# on-disk compatibility with the reference: the same PNG layout / naming, so that the reference's
# summarize_* scripts and project_capture_real_attack consume these outputs unchanged
# ------------------------------------------------------------------------------------------------------------

def reset_rng_seeds(seed):
    """utils.py:70-76."""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def torch_imread(filename):
    """utils.py:116-117: RGB float tensor [3,H,W] in [0,1]."""
    import cv2 as cv
    im = cv.imread(filename)
    if im is None:
        raise FileNotFoundError(filename)
    return torch.from_numpy(cv.cvtColor(im, cv.COLOR_BGR2RGB).transpose((2, 0, 1)).copy()).float() / 255


def torch_imread_mt(img_dir, size=None, index=None, gray_scale=False, normalize=False):
    """utils.py:120-142: every image of a directory (sorted), optionally a subset `index`, resized to size=(h,w) -> [N,C,H,W] in [0,1]."""
    import cv2 as cv
    names = sorted(f for f in os.listdir(img_dir) if f.lower().endswith((".png", ".jpg", ".jpeg", ".bmp")))
    if index is not None:
        names = [names[i] for i in index]
    ims = []
    for n in names:
        im = cv.cvtColor(cv.imread(os.path.join(img_dir, n)), cv.COLOR_BGR2RGB)
        if size is not None:
            im = cv.resize(im, (size[1], size[0]))
        ims.append(torch.from_numpy(im))
    imgs = torch.stack(ims).permute(0, 3, 1, 2).float().div(255)
    if gray_scale:
        imgs = (0.2989 * imgs[:, 0] + 0.5870 * imgs[:, 1] + 0.1140 * imgs[:, 2])[:, None]
    if normalize:
        imgs = (imgs - 0.5) / 0.5
    return imgs


def save_imgs(im_4d, path, idx=0):
    """utils.py:146-167: [N,C,H,W] float (x255, truncated to uint8 like np.uint8) or uint8 [N,H,W,C] -> path/img_%04d.png, numbered from idx + 1."""
    import cv2 as cv
    import numpy as np
    os.makedirs(path, exist_ok=True)
    if torch.is_tensor(im_4d):
        imgs = im_4d.detach().cpu().numpy().transpose(0, 2, 3, 1)
    else:
        imgs = im_4d
    if imgs.dtype == np.float32:
        imgs = np.uint8(imgs[:, :, :, ::-1] * 255)
    else:
        imgs = imgs[:, :, :, ::-1]
    for i in range(imgs.shape[0]):
        cv.imwrite(os.path.join(path, "img_{:04d}.png".format(i + 1 + idx)), np.ascontiguousarray(imgs[i]))


def l2_norm_to_mse(x, num_chan):
    """utils.py:489-491: x = per-pixel L2 norm over channels [B,H,W] -> MSE."""
    return (x ** 2).mean() / num_chan


def idx_to_label(imgnet_labels, idx):
    """utils.py:744-746."""
    vals = list(imgnet_labels.values())
    return [vals[x] for x in idx]
