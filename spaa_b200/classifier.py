"""Classifier wrapper with the reference's call convention (/root/reference/src/python/classifier.py:12-75).

The torchvision networks stay cuDNN modules (external operands, BASELINE.json north_star).  `classify` keeps the
reference's return convention `(raw_score, p_sorted numpy, idx numpy)`; the fused attack loops in
`projector_based_attack` / `perc_al` use `logits()` instead, which never leaves the device.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from .img_proc import center_crop as cc, expand_4d, resize

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_URLS = {"vgg16": "vgg16-397923af.pth", "resnet18": "resnet18-5c106cde.pth", "inception_v3": "inception_v3_google-0cc3c7bd.pth"}


def preprocess(im, crop_sz, input_sz):
    """classifier.py:55-59: uint8 -> float/255, centre crop, area resize, ImageNet normalise (one broadcasted op
    instead of the reference's per-sample python loop)."""
    if im.dtype == torch.uint8:
        im = im.type(torch.float32) / 255
    x = resize(cc(expand_4d(im), crop_sz), input_sz)
    mean, std = _norm_consts(x.dtype, x.device)
    return (x - mean) / std


class _PreprocessFn(torch.autograd.Function):
    """Fused crop + area resize + normalise (ops.clf_preprocess) with its adjoint as the backward."""

    @staticmethod
    def forward(ctx, im, crop, out_hw, channels_last, s2d):
        ctx.crop, ctx.hw, ctx.s2d = crop, (im.shape[2], im.shape[3]), s2d
        return ops.clf_preprocess(im, crop, out_hw, IMAGENET_MEAN, IMAGENET_STD, channels_last, s2d)

    @staticmethod
    def backward(ctx, dout):
        return ops.clf_preprocess_bwd(dout, ctx.hw, ctx.crop, IMAGENET_MEAN, IMAGENET_STD, ctx.s2d), None, None, None, None


def preprocess_fused(im, crop_sz, input_sz, channels_last: bool = False, s2d: bool = False):
    """`preprocess` as one kernel (+ one for the backward) for fp32 CUDA image batches [B,3,H,W]; other inputs use `preprocess`.
    s2d: write the 2x2 space-to-depth fold an `S2DStem` reads ([B,16,h/2+3,w/2+3], channels_last) instead of the image; inputs the kernel
    does not cover come back as the plain image (an S2DStem folds those itself)."""
    h, w = im.shape[-2:]
    th, tw = int(crop_sz[0]), int(crop_sz[1])
    ok = (torch.is_tensor(im) and im.is_cuda and im.dim() == 4 and im.shape[1] == 3 and im.dtype == torch.float32 and th <= h and tw <= w
          and th <= 3 * input_sz[0] and tw <= 3 * input_sz[1] and input_sz[0] <= 2 * th and input_sz[1] <= 2 * tw)
    if not ok:
        x = preprocess(im, crop_sz, input_sz)
        return x.contiguous(memory_format=torch.channels_last) if channels_last else x
    top, left = int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))          # img_proc.py:126-132
    s2d = bool(s2d) and input_sz[0] % 2 == 0 and input_sz[1] % 2 == 0
    return _PreprocessFn.apply(im, (top, left, th, tw), (int(input_sz[0]), int(input_sz[1])), bool(channels_last), s2d)


_NORM_CACHE = {}


def _norm_consts(dtype, device):
    """ImageNet mean / std as device constants, created once per (dtype, device): no host->device copy inside the attack
    iteration (which is replayed as a CUDA graph)."""
    key = (dtype, str(device))
    if key not in _NORM_CACHE:
        _NORM_CACHE[key] = (torch.tensor(IMAGENET_MEAN, dtype=dtype, device=device).view(1, 3, 1, 1),
                            torch.tensor(IMAGENET_STD, dtype=dtype, device=device).view(1, 3, 1, 1))
    return _NORM_CACHE[key]


class Classifier(object):
    def __init__(self, model_name, device, device_ids, fix_params=True, sort_results=True, weights_dir=None, seed=0, allow_random_init=None):
        from torchvision import models
        self.name = model_name
        self.fix_params = fix_params
        self.device = torch.device(device)
        self.sort_results = sort_results
        if self.name in ("vgg16", "resnet18"):
            self.input_sz = (224, 224)
            ctor = lambda: getattr(models, self.name)(weights=None)
        elif self.name == "inception_v3":
            self.input_sz = (299, 299)
            ctor = lambda: models.inception_v3(weights=None, init_weights=False, transform_input=True, aux_logits=True)
        else:
            raise ValueError(f"unknown classifier {model_name}")
        # The reference downloads the ImageNet weights and insists on exactly those (classifier.py:36).  Offline: load them from `weights_dir`
        # (or $SPAA_WEIGHTS_DIR / the torch hub cache).  When the checkpoint is missing this raises FileNotFoundError -- attack results against a
        # random network are meaningless -- unless the caller opts in with allow_random_init=True (or $SPAA_ALLOW_RANDOM_INIT=1): benchmarks and
        # tests, where only throughput / parity matter, use a seeded random initialisation; `self.pretrained` records which one was used.
        if allow_random_init is None:
            allow_random_init = os.environ.get("SPAA_ALLOW_RANDOM_INIT", "0") not in ("", "0")
        rng = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.model = ctor()
        torch.random.set_rng_state(rng)
        self.pretrained = False
        for d in (weights_dir, os.environ.get("SPAA_WEIGHTS_DIR"), os.path.join(torch.hub.get_dir(), "checkpoints")):
            if d and os.path.exists(os.path.join(d, _URLS[self.name])):
                self.model.load_state_dict(torch.load(os.path.join(d, _URLS[self.name]), map_location="cpu"))
                self.pretrained = True
                break
        if not self.pretrained:
            if not allow_random_init:
                raise FileNotFoundError(f"pretrained weights {_URLS[self.name]} for {self.name} not found in weights_dir, $SPAA_WEIGHTS_DIR or "
                                        f"{os.path.join(torch.hub.get_dir(), 'checkpoints')}; pass allow_random_init=True to run on a seeded random network")
            import warnings
            warnings.warn(f"Classifier('{self.name}'): pretrained weights not found, using a seeded RANDOM initialisation (allow_random_init)")
        self.model = self.model.to(self.device)
        if len(device_ids) > 1:
            self.model = torch.nn.DataParallel(self.model, device_ids=device_ids)
        if self.fix_params:
            self.model.eval()
            for param in self.model.parameters():
                param.requires_grad = False

    def logits(self, im, crop_sz=(240, 240)):
        """Device-only forward: pre-processing + network, no softmax / host copy."""
        out = self.model(preprocess(im, crop_sz, self.input_sz).to(self.device))
        return out.logits if hasattr(out, "logits") else out

    def classify(self, im, crop_sz=(240, 240)):
        raw_score = self.logits(im, crop_sz)
        p = F.softmax(raw_score, dim=1).detach().cpu()
        if self.sort_results:
            p_sorted, idx = p.sort(descending=True)
        else:
            p_sorted, idx = p, torch.arange(p.shape[1]).repeat(p.shape[0], 1)
        return raw_score, p_sorted.numpy(), idx.numpy()

    def __call__(self, im, crop_sz):
        return self.classify(im, crop_sz)


def device_logits(classifier, im, crop_sz, channels_last: bool = False):
    """Logits of `im` through any classifier object: ours or the reference's `Classifier` (uses .model/.input_sz on
    the device), or an opaque callable with the reference convention (falls back to its own __call__).
    channels_last: hand the network an NHWC input (its weights should have been converted with `use_channels_last`), so
    cuDNN runs its NHWC tensor-core kernels without the per-layer NCHW<->NHWC transposes."""
    model, input_sz = getattr(classifier, "model", None), getattr(classifier, "input_sz", None)
    if model is not None and input_sz is not None:
        x = preprocess_fused(im, crop_sz, input_sz, channels_last, s2d=bool(channels_last and getattr(classifier, "stem_s2d", False)))
        out = model(x)
        return out.logits if hasattr(out, "logits") else out
    return classifier(im, crop_sz)[0]


class _ReluMaxPoolFn(torch.autograd.Function):
    """max_pool2d(relu(x)) (or max_pool2d(x)) and its adjoint as one kernel each (ops.relu_maxpool_nhwc)."""

    @staticmethod
    def forward(ctx, x, k, stride, pad, relu, bias):
        y, idx = ops.relu_maxpool_nhwc(x, k, stride, pad, relu, bias)
        ctx.save_for_backward(idx)
        ctx.geom = (tuple(x.shape[2:]), k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        hw, k, stride, pad = ctx.geom
        return ops.relu_maxpool_nhwc_bwd(dy, idx, hw, k, stride, pad), None, None, None, None, None


class _BiasActFn(torch.autograd.Function):
    """relu?(x + bias[c] + res) as one kernel (ops.bias_act_nhwc); the adjoint is ATen's threshold_backward on the saved output.
    bias is a frozen constant of the private classifier copy (no gradient)."""

    @staticmethod
    def forward(ctx, x, bias, res, relu):
        y = ops.bias_act_nhwc(x, bias, res, relu)
        ctx.relu, ctx.has_res = bool(relu), res is not None
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        g = torch.ops.aten.threshold_backward(dy, ctx.saved_tensors[0], 0) if ctx.relu else dy
        return g, None, (g if ctx.has_res else None), None


def bias_act(x, bias, res=None, relu: bool = True):
    """relu?(x + bias.view(1,-1,1,1) + res): fused kernel for CUDA fp32 channels_last operands, the stock ops otherwise (same order of additions)."""
    if ops.bias_act_supported(x, bias, res):
        return _BiasActFn.apply(x, bias, res, relu)
    y = x if bias is None else x + bias.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return F.relu(y) if relu else y


class FusedReLUMaxPool2d(torch.nn.Module):
    """Stands in for `nn.ReLU -> nn.MaxPool2d(k, stride, pad)` (with_relu) or a lone `nn.MaxPool2d` inside the attack engines'
    private copy of the frozen classifier.  CUDA fp32 channels_last inputs run the fused kernel; anything else runs the stock ops
    the pair was made of, so the private copy computes the same function everywhere."""

    def __init__(self, kernel_size: int, stride: int, padding: int, with_relu: bool, bias=None):
        """bias: the bias of the (frozen) convolution in front, stripped from it and added here (max(x) + b == max(x + b) exactly)."""
        super().__init__()
        self.kernel_size, self.stride, self.padding, self.with_relu = int(kernel_size), int(stride), int(padding), bool(with_relu)
        self.register_buffer("bias", None if bias is None else bias.detach().clone().contiguous())

    def extra_repr(self):
        return (f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}, with_relu={self.with_relu}, "
                f"bias={self.bias is not None}")

    def forward(self, x):
        if ops.relu_maxpool_supported(x, self.kernel_size, self.stride, self.padding) and ops._bias_ok(self.bias, x):
            return _ReluMaxPoolFn.apply(x, self.kernel_size, self.stride, self.padding, self.with_relu, self.bias)
        if self.bias is not None:
            x = x + self.bias.view(1, -1, 1, 1)
        return F.max_pool2d(F.relu(x) if self.with_relu else x, self.kernel_size, self.stride, self.padding)


class ConvBiasAct(torch.nn.Module):
    """A frozen cuDNN convolution whose bias was stripped, followed by relu?(. + bias + residual) as ONE elementwise kernel (bias_act).
    PyTorch adds a convolution's bias with a separate strided kernel; ReLU and the residual add are two more."""

    def __init__(self, conv: torch.nn.Conv2d, relu: bool = True, extra_bias=None):
        super().__init__()
        self.conv, b = _strip_bias(conv)
        if extra_bias is not None:
            b = b + extra_bias
        self.register_buffer("bias", b.detach().clone().contiguous())
        self.relu = bool(relu)

    def forward(self, x, res=None):
        return bias_act(self.conv(x), self.bias, res, self.relu)


class FusedBasicBlock(torch.nn.Module):
    """torchvision BasicBlock (resnet.py: conv1-bn1-relu-conv2-bn2, += identity / downsample(x), relu) after BatchNorm folding: the same
    three cuDNN convolutions, two fused elementwise kernels instead of seven.  The downsample branch's bias is folded into conv2's."""

    def __init__(self, blk):
        super().__init__()
        ds = blk.downsample
        self.dconv, bd = (None, None) if ds is None else _strip_bias(ds[0])
        self.cba1 = ConvBiasAct(blk.conv1, relu=True)
        self.cba2 = ConvBiasAct(blk.conv2, relu=True, extra_bias=bd)

    def forward(self, x):
        out = self.cba1(x)
        return self.cba2(out, x if self.dconv is None else self.dconv(x))


def s2d_fold(x):
    """The 2x2 space-to-depth fold of layout 2 of spaa_clf_preprocess_fwd in torch ops: [B,3,H,W] (H, W even) -> [B,16,H/2+3,W/2+3]; channel
    (dy*2 + dx)*3 + c of cell (I,J) = pixel (2(I-2)+dy, 2(J-2)+dx), channel c; 2 zero cells before, 1 after each axis; channels 12..15 zero."""
    B, C, H, W = x.shape
    lo, hi = ops.S2D_PAD_LO, ops.S2D_PAD_HI
    xp = F.pad(x, (2 * lo, 2 * hi, 2 * lo, 2 * hi))
    Hs, Ws = H // 2 + lo + hi, W // 2 + lo + hi
    y = xp.reshape(B, C, Hs, 2, Ws, 2).permute(0, 3, 5, 1, 2, 4).reshape(B, 4 * C, Hs, Ws)
    return F.pad(y, (0, 0, 0, 0, 0, ops.S2D_C - 4 * C))


class S2DStem(torch.nn.Module):
    """A frozen 7x7 stride-2 pad-3 convolution on 3 channels (torchvision ResNet `conv1`) as the 4x4 stride-1 pad-0 cuDNN convolution over the
    space-to-depth fold of its input: tap k of the 7 is tap (K, d) = ((k + 1) // 2, (k + 1) % 2) of the fold, the 8th (k = -1) has weight 0.
    Same products, still cuDNN -- but a shape with tensor-core NHWC kernels (B=32: 275 vs 532 us forward + input gradient).
    forward() takes the folded input [B,16,h/2+3,w/2+3] (what the fused pre-processing kernel writes) or the plain image (folded here with torch
    ops; odd sizes run the original convolution)."""

    def __init__(self, conv: torch.nn.Conv2d):
        super().__init__()
        self.orig = conv
        w = conv.weight.detach()
        co = w.shape[0]
        w8 = F.pad(w, (1, 0, 1, 0))                                            # taps k' = k + 1 in [0, 8), k' = 0 is the zero tap
        w4 = w8.reshape(co, 3, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4).reshape(co, 12, 4, 4)      # [co, (dy, dx, c), Ky, Kx]
        self.conv = torch.nn.Conv2d(ops.S2D_C, co, 4, 1, 0, bias=conv.bias is not None, device=w.device, dtype=w.dtype)
        with torch.no_grad():
            self.conv.weight.copy_(F.pad(w4, (0, 0, 0, 0, 0, ops.S2D_C - 12)))
            if conv.bias is not None:
                self.conv.bias.copy_(conv.bias)
        for p in self.conv.parameters():
            p.requires_grad = False

    def forward(self, x):
        if x.shape[1] == ops.S2D_C:
            return self.conv(x)
        if x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            return self.conv(s2d_fold(x))
        return self.orig(x)


def fuse_s2d_stem(model) -> int:
    """In place: torchvision ResNet `conv1` (7x7, stride 2, pad 3, 3 input channels) -> S2DStem."""
    from torchvision.models import resnet
    n = 0
    for mod in list(model.modules()):
        c = mod._modules.get("conv1") if isinstance(mod, resnet.ResNet) else None
        if (type(c) is torch.nn.Conv2d and c.in_channels == 3 and tuple(c.kernel_size) == (7, 7) and tuple(c.stride) == (2, 2) and c.padding == (3, 3)
                and tuple(c.dilation) == (1, 1) and c.groups == 1 and c.padding_mode == "zeros"):
            mod._modules["conv1"] = S2DStem(c)
            n += 1
    return n


def _strip_bias(conv: torch.nn.Conv2d):
    """(the same convolution sharing `conv`'s weight but without bias, its bias tensor)."""
    c = torch.nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups, bias=False,
                        padding_mode=conv.padding_mode, device=conv.weight.device, dtype=conv.weight.dtype)
    c.weight = conv.weight
    return c, conv.bias.detach()


def _biased_conv(m) -> bool:
    return type(m) is torch.nn.Conv2d and m.bias is not None and isinstance(m.padding, tuple)


def fuse_bias_act(model) -> int:
    """In place, on a BatchNorm-folded private copy: strip the bias of every cuDNN convolution whose consumer is known and add it in the fused
    kernel that follows -- ResNet stem (conv1 -> relu -> maxpool), BasicBlock, Inception's BasicConv2d, `Conv2d -> ReLU [-> MaxPool2d]` runs of an
    nn.Sequential (VGG `features`).  Returns the number of rewritten sites."""
    from torchvision.models import inception, resnet
    ident = lambda m: type(m) is torch.nn.Identity
    n = 0
    for mod in list(model.modules()):
        kids = mod._modules
        if isinstance(mod, resnet.ResNet) and _biased_conv(kids.get("conv1")) and ident(kids.get("bn1")) and type(kids.get("relu")) is torch.nn.ReLU:
            g = _plain_maxpool(kids.get("maxpool"))
            if g is not None:
                kids["conv1"], b = _strip_bias(kids["conv1"])
                kids["relu"] = torch.nn.Identity()
                kids["maxpool"] = FusedReLUMaxPool2d(*g, with_relu=True, bias=b)
                n += 1
        if isinstance(mod, torch.nn.Sequential):
            names = list(kids.keys())
            for i in range(len(names) - 1):
                conv, act = kids[names[i]], kids[names[i + 1]]
                if not (_biased_conv(conv) and type(act) is torch.nn.ReLU):
                    continue
                g = _plain_maxpool(kids[names[i + 2]]) if i + 2 < len(names) else None
                if g is not None:
                    kids[names[i]], b = _strip_bias(conv)
                    kids[names[i + 1]] = torch.nn.Identity()
                    kids[names[i + 2]] = FusedReLUMaxPool2d(*g, with_relu=True, bias=b)
                else:
                    kids[names[i]] = ConvBiasAct(conv, relu=True)
                    kids[names[i + 1]] = torch.nn.Identity()
                n += 1
        for name, child in list(kids.items()):
            if (type(child) is resnet.BasicBlock and _biased_conv(child.conv1) and _biased_conv(child.conv2) and ident(child.bn1) and ident(child.bn2)
                    and type(child.relu) is torch.nn.ReLU
                    and (child.downsample is None or (isinstance(child.downsample, torch.nn.Sequential) and len(child.downsample) == 2
                                                      and _biased_conv(child.downsample[0]) and ident(child.downsample[1])))):
                kids[name] = FusedBasicBlock(child)
                n += 1
            elif type(child) is inception.BasicConv2d and _biased_conv(child.conv) and ident(child.bn):
                kids[name] = ConvBiasAct(child.conv, relu=True)           # BasicConv2d.forward: conv -> bn -> F.relu
                n += 1
    return n


def _plain_maxpool(m):
    """(k, stride, pad) of an nn.MaxPool2d the fused kernel reproduces exactly, else None."""
    if type(m) is not torch.nn.MaxPool2d or m.ceil_mode or m.return_indices:
        return None

    def one(v):
        v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        return int(v[0]) if len(v) == 2 and v[0] == v[1] else None
    k, st, pd, dl = one(m.kernel_size), one(m.stride if m.stride is not None else m.kernel_size), one(m.padding), one(m.dilation)
    if None in (k, st, pd, dl) or dl != 1 or not (1 <= k <= 15) or 2 * pd > k:
        return None
    return k, st, pd


def fuse_relu_maxpool(model) -> int:
    """In place: replace `ReLU -> MaxPool2d` pairs (torchvision ResNet stem, nn.Sequential neighbours as in VGG's `features`) and lone
    MaxPool2d modules (Inception's maxpool1 / maxpool2) of `model` by FusedReLUMaxPool2d.  Returns the number of replacements."""
    from torchvision.models import resnet
    n = 0
    for mod in list(model.modules()):
        if isinstance(mod, resnet.ResNet) and type(mod._modules.get("relu")) is torch.nn.ReLU:
            g = _plain_maxpool(mod._modules.get("maxpool"))
            if g is not None:                      # ResNet._forward_impl: conv1 -> bn1 -> relu -> maxpool; `relu` is used nowhere else
                mod._modules["relu"] = torch.nn.Identity()
                mod._modules["maxpool"] = FusedReLUMaxPool2d(*g, with_relu=True)
                n += 1
        elif isinstance(mod, torch.nn.Sequential):
            names = list(mod._modules.keys())
            for a, b in zip(names, names[1:]):
                g = _plain_maxpool(mod._modules[b])
                if g is not None and type(mod._modules[a]) is torch.nn.ReLU:
                    mod._modules[a] = torch.nn.Identity()
                    mod._modules[b] = FusedReLUMaxPool2d(*g, with_relu=True)
                    n += 1
    for mod in list(model.modules()):
        for name, child in list(mod._modules.items()):
            g = _plain_maxpool(child)
            if g is not None:
                mod._modules[name] = FusedReLUMaxPool2d(*g, with_relu=False)
                n += 1
    return n


class _FoldedView:
    """What the fused attack loops need of a classifier (`.model`, `.input_sz`), with a private BatchNorm-folded network."""

    def __init__(self, model, input_sz, name):
        self.model, self.input_sz, self.name = model, input_sz, name


def fold_batchnorm(classifier, fuse_pool: Optional[bool] = None, fuse_bias: Optional[bool] = None, fuse_stem: Optional[bool] = None):
    """A view of `classifier` whose network is a PRIVATE copy with every inference-mode BatchNorm2d folded into the cuDNN
    convolution in front of it (torch.nn.utils.fusion.fuse_conv_bn_eval: w' = w * gamma / sqrt(var + eps), b' likewise).
    The classifier is frozen and in eval() (classifier.py:38-42), so its BatchNorm layers are per-channel affine maps: the
    folded network computes the same function (fp32 rounding differences ~1e-6 relative) while its forward AND input-gradient
    lose one elementwise pass per BatchNorm (resnet18, B=32: 20 x bn_fw_inf + 20 x batch_norm_backward + 20 x copy, ~1.1 of
    2.9 ms, ncu launch list profiles/r1_fp16_v3_launches.md).
    fuse_pool (default: $SPAA_FUSE_POOL, on): the copy's `ReLU -> MaxPool2d` pairs and lone MaxPool2d modules become
    FusedReLUMaxPool2d (fuse_relu_maxpool): exact max-pooling, one of our kernels each way instead of ATen's relu / max_pool2d /
    threshold_backward / max_pool2d_backward passes over the largest activation of the iteration.
    fuse_stem (default: $SPAA_FUSE_STEM, on): a ResNet's 7x7 stride-2 stem convolution runs as the equivalent 4x4 stride-1 cuDNN convolution over
    the space-to-depth fold of the input, which the fused pre-processing kernel writes directly (S2DStem).
    fuse_bias (default: $SPAA_FUSE_BIAS, on): the folded biases leave the cuDNN convolutions and are added, together with the residual and the
    ReLU, by one kernel per convolution (fuse_bias_act: ConvBiasAct / FusedBasicBlock) instead of ATen's two or three elementwise passes.
    The user's module is not modified.  Returns `classifier` itself when there is nothing to change (opaque callables,
    training-mode networks)."""
    import copy
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    model, input_sz = getattr(classifier, "model", None), getattr(classifier, "input_sz", None)
    if fuse_pool is None:
        fuse_pool = os.environ.get("SPAA_FUSE_POOL", "1") != "0"
    if fuse_bias is None:
        fuse_bias = os.environ.get("SPAA_FUSE_BIAS", "1") != "0"
    if fuse_stem is None:
        fuse_stem = os.environ.get("SPAA_FUSE_STEM", "1") != "0"
    if not isinstance(model, torch.nn.Module) or input_sz is None or model.training:
        return classifier
    has_bn = any(isinstance(m, torch.nn.BatchNorm2d) for m in model.modules())
    has_pool = fuse_pool and any(_plain_maxpool(m) is not None for m in model.modules())
    has_bias = fuse_bias and isinstance(model, torch.nn.Module) and any(isinstance(m, torch.nn.Sequential) and any(_biased_conv(c) for c in m.children())
                                                                        for m in model.modules())
    if not has_bn and not has_pool and not has_bias:
        return classifier
    folded = copy.deepcopy(model)
    n = 0
    for mod in folded.modules():
        names = list(mod._modules.keys())
        # a Conv2d registered immediately before a matching BatchNorm2d and applied back to back: torchvision's
        # ResNet stem / BasicBlock / Bottleneck / downsample Sequential and Inception's BasicConv2d
        for a, b in zip(names, names[1:]):
            conv, bn = mod._modules[a], mod._modules[b]
            if (type(conv) is torch.nn.Conv2d and type(bn) is torch.nn.BatchNorm2d and bn.num_features == conv.out_channels
                    and bn.track_running_stats and _applied_back_to_back(mod, a, b)):
                mod._modules[a] = fuse_conv_bn_eval(conv, bn)
                mod._modules[b] = torch.nn.Identity()
                n += 1
    if fuse_bias:
        n += fuse_bias_act(folded)
    if fuse_pool:
        n += fuse_relu_maxpool(folded)
    n_stem = fuse_s2d_stem(folded) if fuse_stem else 0
    n += n_stem
    if n == 0:
        return classifier
    for p in folded.parameters():
        p.requires_grad = False
    if next(model.parameters()).is_contiguous(memory_format=torch.channels_last) or any(
            p.dim() == 4 and p.shape[1] > 1 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous() for p in model.parameters()):
        folded.to(memory_format=torch.channels_last)
    view = _FoldedView(folded.eval(), input_sz, getattr(classifier, "name", type(model).__name__))
    view.stem_s2d = n_stem > 0            # device_logits() then asks the pre-processing kernel for the folded input
    return view


def _applied_back_to_back(mod, conv_name: str, bn_name: str) -> bool:
    """Only module types whose forward is known to compute bn(conv(x)) for this attribute pair."""
    from torchvision.models import inception, resnet
    if isinstance(mod, torch.nn.Sequential):
        return True
    if isinstance(mod, (resnet.ResNet, resnet.BasicBlock, resnet.Bottleneck)):
        return conv_name.startswith("conv") and bn_name == "bn" + conv_name[4:]
    if isinstance(mod, inception.BasicConv2d):
        return conv_name == "conv" and bn_name == "bn"
    return False


def use_channels_last(classifier) -> bool:
    """Convert the external network's weights to channels_last strides in place (same values, same cuDNN module).
    Returns whether `device_logits(..., channels_last=True)` should be used."""
    model = getattr(classifier, "model", None)
    # cuDNN's exact-fp32 NHWC kernels are 2x slower than its NCHW ones (measured on B200: resnet18 B=32 fwd+bwd 14.2 vs 7.2 ms);
    # with TF32 allowed (torch's default, the reference's setting) NHWC is the faster layout (3.28 vs 3.51 ms).
    if not torch.backends.cudnn.allow_tf32:
        return False
    if isinstance(model, torch.nn.Module) and getattr(classifier, "input_sz", None) is not None:
        model.to(memory_format=torch.channels_last)
        return True
    return False
