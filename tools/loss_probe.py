"""Time (and, under ncu, profile) the fused loss kernels alone at training / attack sizes.  usage: python tools/loss_probe.py [ssim|color]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spaa_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "ssim"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B = 24 if which == "ssim" else 32
a = torch.rand(B, 3, 240, 320, device=dev, generator=g)
b = (a + 0.1 * torch.randn(B, 3, 240, 320, device=dev, generator=g)).clamp(0, 1)
if which == "ssim":
    fn = lambda: ops.ssim_l1(a, b, 1.0, 0.0, 1.0)
else:
    scene = b[:1].contiguous()
    lab = ops.rgb2lab(scene)
    stats, grad = torch.empty(B, 4, device=dev), torch.empty_like(a)
    fn = lambda: ops.color_loss(a, scene, lab, cam_is_lab2=False, de_weighting=False, c_de=1.0, c_l2=1.0, stats=stats, grad=grad)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record(); torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) * 100:.1f} us per launch, B={B}")
