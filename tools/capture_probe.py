"""Where the time of an attack engine's CUDA-graph capture goes: capture_begin / recording the iteration / capture_end (instantiate) / first replay.
usage: python tools/capture_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench
from spaa_b200 import models, ops
from spaa_b200.projector_based_attack import SpaaAttack
dev = torch.device("cuda:0")
scene, P, targets = bench.synthetic_inputs(0)
pcnet = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=bench.CAM_HW)), torch.nn.DataParallel(models.ShadingNetSPAA()))
pcnet.load_state_dict(P, strict=True)
pcnet = models.set_precision(pcnet.to(dev).eval(), "fp16")
for p in pcnet.parameters():
    p.requires_grad = False
clf = bench.make_classifier(dev)
for rep in range(6):
    A = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), bench.D_THR, bench.STEALTH, dev, bench.SETUP, graph=False)
    for _ in range(3):
        A.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    pool = ops.graph_pool(dev)
    t = [time.perf_counter()]
    with torch.cuda.stream(side):
        g.capture_begin(pool=pool) if pool is not None else g.capture_begin()
        t.append(time.perf_counter())
        A._step_eager()
        t.append(time.perf_counter())
        g.capture_end()
        t.append(time.perf_counter())
    torch.cuda.current_stream(dev).wait_stream(side)
    g.replay(); torch.cuda.synchronize(); t.append(time.perf_counter())
    g.replay(); torch.cuda.synchronize(); t.append(time.perf_counter())
    print(f"rep{rep}: capture_begin {1e3*(t[1]-t[0]):.1f} ms, record {1e3*(t[2]-t[1]):.1f} ms, capture_end {1e3*(t[3]-t[2]):.1f} ms, first replay {1e3*(t[4]-t[3]):.1f} ms, second {1e3*(t[5]-t[4]):.1f} ms; "
          f"reserved {torch.cuda.memory_reserved(dev) / 2**30:.2f} GiB, allocated {torch.cuda.memory_allocated(dev) / 2**30:.2f} GiB", flush=True)
    del g, A
    if os.environ.get("PROBE_GC", "0") != "0":
        import gc
        print("   gc.collect() found", gc.collect(), "objects; allocated", f"{torch.cuda.memory_allocated(dev) / 2**30:.2f} GiB", flush=True)
