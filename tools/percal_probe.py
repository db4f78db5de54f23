"""PerC-AL + CompenNet++ attack (BASELINE configs[2]) throughput: iterations/s of PerC_AL.adversary_projector at B=32 on a 240x320 scene with
random-init torchvision classifiers, CUDA-graph replay on / off.  usage: python tools/percal_probe.py [vgg16|inception_v3|resnet18] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import synth
from spaa_b200 import perc_al
from spaa_b200.classifier import Classifier
name = sys.argv[1] if len(sys.argv) > 1 else "vgg16"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
B = 32
scene = synth.textured(0, "pp.scene", (1, 3, 240, 320)).to(dev)
clf = Classifier(name, dev, [0], allow_random_init=True)
targets = torch.tensor([(7 * i) % 1000 for i in range(B)], device=dev)
for graph in (False, True):
    atk = perc_al.PerC_AL(device=dev, max_iterations=iters, alpha_l_init=1, alpha_c_init=0.5, confidence=0)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = atk.adversary_projector(clf, scene.expand(B, -1, -1, -1), targets, None, 11, True, (240, 240), graph=graph)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name} B={B} graph={graph}: {iters / dt:.1f} it/s ({dt / iters * 1e3:.2f} ms per iteration), changed pixels {(out != scene).float().mean().item():.3f}")
