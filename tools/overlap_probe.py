"""Attack iterations/s of the bench workload (resnet18, 32 targets) for the stream-level schedules of one iteration:
  stock glue / fused pools / + bias/act   the private classifier copy with ATen's elementwise glue, with the fused ReLU + max-pooling kernels,
            and with every convolution's bias / residual add / ReLU in one kernel; one engine, one stream
  (SCHEDULES=1 adds:)
  overlap   stealth-loss kernels on an auxiliary stream beside the classifier (SpaaAttack(overlap=True))
  pipeN     N engines over target slices on N streams (SpaaAttackPipelined), with / without overlap
Prints it/s (CUDA events around K iterations, all streams joined) and the largest difference of the attacked projector images
against the first schedule after the same number of iterations (the free-running loop is chaotic and the warp's scatter-add uses atomics, so
two runs of the SAME schedule differ just as much: the figure only shows that nothing diverged to NaN / garbage)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench
from spaa_b200 import models
from spaa_b200.projector_based_attack import SpaaAttack



class SpaaAttackPipelined:
    """`pipeline` SpaaAttack engines, each on its own stream with its own captured graph, over disjoint slices of the target batch.
    Samples of an attack are independent (no BatchNorm in PCNet, the classifier is in eval(), every update is normalised per
    sample), so a slice's iteration never waits for another slice: while one slice is inside the external classifier's ~150 short
    cuDNN / ATen launches, the other runs its PCNet convolutions.  Results are those of the single-engine attack up to the summation
    order of cuDNN's batch-size-dependent kernel choices (the per-sample arithmetic of our kernels does not depend on the batch size).
    step() only enqueues work; the slices are joined with the caller's stream in sync_streams() / result().
    EXPERIMENT, kept here and not in the package: measured on B200 (profiles/r1_overlap_probe.md) two slices of 16 run at 271-276 it/s and
    four of 8 at 223 it/s against 296 it/s for one engine of 32 -- our persistent convolution kernels fill every SM, so a second stream
    only interleaves at kernel boundaries while each slice pays the smaller batch's tail and launch costs."""

    def __init__(self, pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, pipeline: int = 2, **kw):
        self.device = torch.device(device)
        self.B = len(target_idx)
        n = max(1, min(int(pipeline), self.B))
        cuts = [self.B * k // n for k in range(n + 1)]
        self.slices = [slice(a, b) for a, b in zip(cuts, cuts[1:])]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.slices]
        self.parts = [SpaaAttack(pcnet, classifier, list(target_idx)[sl], targeted, cam_scene, d_thr, stealth_loss, device, setup_info, **kw)
                      for sl in self.slices]

    def step(self):
        main = torch.cuda.current_stream(self.device)
        for A, st in zip(self.parts, self.streams):
            st.wait_stream(main)                     # whatever the caller queued (reset copies, forced inputs) is visible to the slice
            with torch.cuda.stream(st):
                A.step()

    def sync_streams(self):
        main = torch.cuda.current_stream(self.device)
        for st in self.streams:
            main.wait_stream(st)

    def reset(self, cam_scene, target_idx):
        if len(target_idx) != self.B:
            raise ValueError("reset() needs the batch size the engine was built for")
        self.sync_streams()
        for A, sl in zip(self.parts, self.slices):
            A.reset(cam_scene, list(target_idx)[sl])
        return self

    def result(self):
        self.sync_streams()
        res = [A.result() for A in self.parts]
        return torch.cat([r[0] for r in res]), torch.cat([r[1] for r in res])


K = int(os.environ.get("K", "40"))
dev = torch.device("cuda:0")
scene, P, targets = bench.synthetic_inputs(0)
pcnet = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=bench.CAM_HW)), torch.nn.DataParallel(models.ShadingNetSPAA()))
pcnet.load_state_dict(P, strict=True)
pcnet = models.set_precision(pcnet.to(dev).eval(), "fp16")
for p in pcnet.parameters():
    p.requires_grad = False
clf = bench.make_classifier(dev)
args = (pcnet, clf, targets, True, scene.to(dev), bench.D_THR, bench.STEALTH, dev, bench.SETUP)


def run(name, make):
    A = make()
    for _ in range(4):
        A.step()
    if hasattr(A, "sync_streams"):
        A.sync_streams()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        A.step()
    if hasattr(A, "sync_streams"):
        A.sync_streams()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    prj = torch.cat([a.prj_adv for a in A.parts]) if hasattr(A, "parts") else A.prj_adv
    print(f"{name:24s} {ms:7.3f} ms/it  {1e3 / ms:7.1f} it/s", flush=True)
    return prj.clone()


def with_env(fuse_pool, fuse_bias, fuse_stem, make):
    def f():
        os.environ["SPAA_FUSE_POOL"] = "1" if fuse_pool else "0"          # read when the engine builds its private classifier copy
        os.environ["SPAA_FUSE_BIAS"] = "1" if fuse_bias else "0"
        os.environ["SPAA_FUSE_STEM"] = "1" if fuse_stem else "0"
        return make()
    return f


cfgs = [("stock glue", with_env(False, False, False, lambda: SpaaAttack(*args, overlap=False))),
        ("fused pools", with_env(True, False, False, lambda: SpaaAttack(*args, overlap=False))),
        ("fused pools + bias/act", with_env(True, True, False, lambda: SpaaAttack(*args, overlap=False))),
        ("... + space-to-depth stem", with_env(True, True, True, lambda: SpaaAttack(*args, overlap=False)))]
if os.environ.get("SCHEDULES", "0") != "0":
    cfgs += [("fused + overlap", with_env(True, True, True, lambda: SpaaAttack(*args, overlap=True))),
             ("pipe2", with_env(True, True, True, lambda: SpaaAttackPipelined(*args, pipeline=2, overlap=False))),
             ("pipe2 + overlap", with_env(True, True, True, lambda: SpaaAttackPipelined(*args, pipeline=2, overlap=True))),
             ("pipe4 + overlap", with_env(True, True, True, lambda: SpaaAttackPipelined(*args, pipeline=4, overlap=True)))]
cfgs += [("stock glue (again)", with_env(False, False, False, lambda: SpaaAttack(*args, overlap=False)))]
ref = None
for name, make in cfgs:
    try:
        prj = run(name, make)
    except Exception as e:                       # keep measuring the other schedules
        import traceback
        traceback.print_exc()
        print(f"{name}: FAILED ({type(e).__name__}: {e})", flush=True)
        continue
    if ref is None:
        ref = prj
    else:
        d = (prj - ref).abs()
        print(f"    vs base after {K + 4} iterations: max |d prj_adv| = {d.max().item():.3e}, samples differing by > 1e-3: "
              f"{int((d.flatten(1).max(1).values > 1e-3).sum())}/{prj.shape[0]}", flush=True)
