"""Micro-benchmark of the classifier glue kernels (relu_maxpool_fwd/bwd, bias_act) at the shapes of resnet18 / vgg16, B=32: CUDA events, L2 flushed
by a 256 MB read pass between launches, median of 10; GB/s = algorithmic bytes / time, against MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spaa_b200 import ops

dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6554.0
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timed(fn, n=10):
    ts = []
    for _ in range(n + 2):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


print("| kernel | shape (N,C,H,W) | us | GB/s (algorithmic) | frac HBM peak |\n|---|---|---:|---:|---:|")
for name, (N, C, H, W), (k, s, p) in (("resnet18 stem", (32, 64, 112, 112), (3, 2, 1)), ("vgg16 pool1", (32, 64, 224, 224), (2, 2, 0)),
                                      ("vgg16 pool3", (32, 256, 56, 56), (2, 2, 0))):
    x = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(C, device=dev)
    y, idx = ops.relu_maxpool_nhwc(x, k, s, p, True, b)
    dy = torch.randn_like(y)
    nin, nout = x.numel(), y.numel()
    for label, fn, nbytes in ((f"relu_maxpool_fwd {k}x{k} s{s}", lambda: ops.relu_maxpool_nhwc(x, k, s, p, True, b), 4 * (nin + nout) + nout),
                              (f"relu_maxpool_bwd {k}x{k} s{s}", lambda: ops.relu_maxpool_nhwc_bwd(dy, idx, (H, W), k, s, p), 4 * (nin + nout) + nout)):
        us = timed(fn)
        print(f"| {label} ({name}) | {N},{C},{H},{W} | {us:.1f} | {nbytes / us / 1e3:.0f} | {nbytes / us / 1e3 / peak:.3f} |")
for name, (N, C, H, W), with_res in (("resnet18 layer1 conv1", (32, 64, 56, 56), False), ("resnet18 layer1 conv2 + identity", (32, 64, 56, 56), True),
                                     ("resnet18 layer4", (32, 512, 7, 7), True), ("vgg16 conv1_1", (32, 64, 224, 224), False)):
    x = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(C, device=dev)
    r = torch.randn_like(x) if with_res else None
    us = timed(lambda: ops.bias_act_nhwc(x, b, r, True))
    nbytes = 4 * x.numel() * (3 if with_res else 2)
    print(f"| bias_act ({name}) | {N},{C},{H},{W} | {us:.1f} | {nbytes / us / 1e3:.0f} | {nbytes / us / 1e3 / peak:.3f} |")
