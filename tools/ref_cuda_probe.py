"""Times the UNMODIFIED reference (baseline/_ref, via tests/golden/ref_harness.py) on the GPU: spaa B=32 resnet18, one 50-iteration call
(10 warm-up iterations absorb cuDNN autotune, utils.py:79-81 cudnn.benchmark=True), TF32 default and allow_tf32=False."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import torch
import synth
import ref_harness as rh

CAM_HW, PRJ_HW = (240, 320), (256, 256)
dev = torch.device("cuda:0")
R = rh.load_reference()
torch.backends.cudnn.benchmark = True
P = synth.pcnet_params(100, CAM_HW)
m = torch.nn.DataParallel(rh.ref_pcnet(R, P, CAM_HW, dev).eval(), device_ids=[0])
for p in m.parameters():
    p.requires_grad = False
from torchvision import models
torch.manual_seed(0)
net = models.resnet18(weights=None).to(dev).eval()
for p in net.parameters():
    p.requires_grad = False
clf = rh.ref_classifier(R, torch.nn.DataParallel(net, device_ids=[0]), (224, 224), dev)
scene = synth.textured(0, "bench.scene", (1, 3, *CAM_HW)).to(dev)
setup = {"classifier_crop_sz": (240, 240), "prj_brightness": 0.5, "prj_im_sz": PRJ_HW}
targets = [synth.SPAA_TARGETS10[i % 10] for i in range(32)]
out = {}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    sec, n = rh.time_reference_spaa(R, m, clf, targets, scene, 5.0, "camdE_caml2", dev, setup, warmup=10, steps=40)
    out["tf32" if tf32 else "fp32"] = {"it_per_s": 1 / sec, "ms": sec * 1e3, "n": n}
print(json.dumps(out))
