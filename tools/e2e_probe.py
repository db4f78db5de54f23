import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench
from spaa_b200 import models
from spaa_b200.projector_based_attack import spaa, SpaaAttack
dev = torch.device("cuda:0")
scene, P, targets = bench.synthetic_inputs(0)
pcnet = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=bench.CAM_HW)), torch.nn.DataParallel(models.ShadingNetSPAA()))
pcnet.load_state_dict(P, strict=True)
pcnet = models.set_precision(pcnet.to(dev).eval(), "fp16")
for p in pcnet.parameters():
    p.requires_grad = False
clf = bench.make_classifier(dev)
for graph in (False, True, False, True):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        A = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), bench.D_THR, bench.STEALTH, dev, bench.SETUP, graph=graph)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ts = []
        for i in range(6):
            A.step(); torch.cuda.synchronize(); ts.append(time.perf_counter())
        print(f"graph={graph} rep{rep}: init {1e3*(t1-t0):.1f} ms; steps " + " ".join(f"{1e3*(b-a):.1f}" for a, b in zip([t1] + ts, ts)), flush=True)
        del A
