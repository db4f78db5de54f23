#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SPAA hot path (contract: see README / DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16|bf16|fp32]

Workload (BASELINE.json configs[1]): SPAA attack, resnet18 classifier, batch of 32 target images, synthetic 256x256
projector / 240x320 camera data, random-init PCNet.  One "step" = one iteration of the attack loop for the whole batch
(PCNet forward, classifier forward+backward, fused Lab+dE2000+L2 loss, one PCNet backward, normalised masked update).
Prints ONE JSON line (rank 0).  With N > 1 (torchrun) every rank runs an independent batch of 32 targets (attack jobs
shard with no collective: weak scaling); the time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CAM_HW, PRJ_HW, CROP = (240, 320), (256, 256), (240, 240)
BATCH = 32
SETUP = {"classifier_crop_sz": CROP, "prj_brightness": 0.5, "prj_im_sz": PRJ_HW}
STEALTH, D_THR = "camdE_caml2", 5.0
# algorithmic work of the dominant layer family (conv4 / conv4_s / conv5 and their backward-data passes):
# 2 * Hout * Wout * Cout * Cin * k^2 FLOP per sample
HEAVY_FLOP_PER_SAMPLE = 2 * 60 * 80 * 256 * 128 * 9
# mean dram__bytes_read.sum + dram__bytes_write.sum of those launches at B=32, one `ncu --set full` capture (profiles/r2_halo_ncu_full.md)
NCU_HEAVY_DRAM_BYTES = 168.1e6


def synthetic_inputs(seed: int):
    import synth
    scene = synth.textured(seed, "bench.scene", (1, 3, *CAM_HW))
    P = synth.pcnet_params(100 + seed, CAM_HW)
    targets = [synth.SPAA_TARGETS10[i % 10] for i in range(BATCH)]
    return scene, P, targets


def make_classifier(device, name="resnet18"):
    """Seeded random-init torchvision classifier (no pretrained weights offline) in the reference's convention: .model / .input_sz / .name
    (classifier.py:15-33; inception_v3 with transform_input=True, :31)."""
    from torchvision import models
    rng = torch.random.get_rng_state()
    torch.manual_seed(0)
    if name == "inception_v3":
        net = models.inception_v3(weights=None, init_weights=False, transform_input=True, aux_logits=True)
    else:
        net = getattr(models, name)(weights=None)
    torch.random.set_rng_state(rng)
    net = net.to(device).eval()
    for p in net.parameters():
        p.requires_grad = False

    class C:
        pass
    c = C()
    c.model, c.input_sz, c.name = net, ((299, 299) if name == "inception_v3" else (224, 224)), name
    return c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx.append(int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------------------

def cpu_reference_run(sample_B: int, warmup: int, steps: int, device="cpu"):
    """Times oracle.spaa_attack (the CPU restatement of projector_based_attack.py:212-339 on stock PyTorch ops -- the
    same ops the reference dispatches) for `sample_B` targets.  Returns seconds per iteration of that sample."""
    from oracle import spaa_oracle as O
    scene, P, targets = synthetic_inputs(0)
    dev = torch.device(device)
    P = {k: v.to(dev) for k, v in P.items()}
    scene = scene.to(dev)
    clf = make_classifier(dev)
    times = []

    def classify(im):
        if dev.type == "cuda":      # stock ops the reference dispatches (img_proc.py:117-123: F.interpolate(mode='area')), no python pool matrix
            x = torch.nn.functional.interpolate(O.crop_center(im, CROP), clf.input_sz, mode="area")
            mean = torch.tensor(O.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
            std = torch.tensor(O.IMAGENET_STD, device=dev).view(1, 3, 1, 1)
            logits = clf.model((x - mean) / std)
            p_sorted, idx = torch.softmax(logits, 1).detach().sort(descending=True)
            return logits, p_sorted, idx
        return O.classify(clf.model, im, CROP, clf.input_sz)

    def pc(x, s):
        return O.pcnet(P, x, s, CAM_HW)

    class Timer(list):                          # spaa_attack appends one dict per iteration: timestamp them
        def append(self, item):
            if dev.type == "cuda":
                torch.cuda.synchronize()
            times.append(time.perf_counter())

    t0 = time.perf_counter()
    O.spaa_attack(pc, classify, targets[:sample_B], True, scene, D_THR, STEALTH, prj_hw=PRJ_HW, iters=warmup + steps, trace=Timer())
    stamps = [t0] + times
    per_it = [(stamps[i + 1] - stamps[i]) for i in range(warmup, warmup + steps)]
    return sum(per_it) / len(per_it)


def reference_modules():
    """(R, rh): the UNMODIFIED reference imported through tests/golden/ref_harness.py from /root/reference (build container) or its git-ignored copy
    baseline/_ref (GPU box); None when neither exists (then the oracle port stands in, `kind: "port"`)."""
    import ref_harness as rh
    if not rh.available():
        return None
    with open(os.devnull, "w") as nul:
        import contextlib, warnings
        with contextlib.redirect_stderr(nul), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return rh.load_reference(), rh


def reference_spaa_objects(R, rh, device, clf_name="resnet18"):
    """The reference's own objects for the bench workload: DataParallel-wrapped PCNet (train_network.py:550-552) and its Classifier wrapper
    (classifier.py:12-75, DataParallel-wrapped model :39) around the same seeded random-init networks our arm uses."""
    scene, P, targets = synthetic_inputs(0)
    dev = torch.device(device)
    ids = [dev.index or 0] if dev.type == "cuda" else None
    m = rh.ref_pcnet(R, P, CAM_HW, dev).eval()
    if ids:                                      # (on the host cores the bare module: nn.DataParallel without device_ids would grab the box's GPUs)
        m = torch.nn.DataParallel(m, device_ids=ids)
    for p in m.parameters():                     # projector_based_attack.py:63-67
        p.requires_grad = False
    net = make_classifier(dev, clf_name).model
    insz = (299, 299) if clf_name == "inception_v3" else (224, 224)
    clf = rh.ref_classifier(R, torch.nn.DataParallel(net, device_ids=ids) if ids else net, insz, dev, clf_name)
    return m, clf, scene.to(dev), targets


def reference_cpu_spaa(sample_B: int, warmup: int, steps: int, budget_s: float):
    """The reference's CPU path for the bench workload: (seconds per iteration of `sample_B` targets, iterations timed, kind)."""
    mods = reference_modules()
    if mods is None:
        return cpu_reference_run(sample_B, warmup, steps), steps, "port"
    R, rh = mods
    m, clf, scene, targets = reference_spaa_objects(R, rh, "cpu")
    sec, n = rh.time_reference_spaa(R, m, clf, targets[:sample_B], scene, D_THR, STEALTH, "cpu", SETUP, warmup=warmup, steps=steps, budget_s=budget_s)
    return sec, n, "reference"


def run_reference(args):
    """Reference arm: the reference's own `spaa` (projector_based_attack.py:212-339, unmodified, imported from baseline/_ref) on the box's host cores,
    all threads, on the FULL bench workload (32 targets, resnet18, 256x256 / 240x320).  One CPU iteration of that batch takes ~10-15 s, so the
    run is bounded in time, not in batch: the requested warm-up iterations (at most 5), then as many of the requested `--steps` as fit in ~150 s (at least 2)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warmup = max(1, min(args.warmup, 5))
    sec, n, kind = reference_cpu_spaa(BATCH, warmup, max(2, args.steps), budget_s=float(os.environ.get("SPAA_BENCH_REF_BUDGET_S", "150")))
    its = 1.0 / sec
    sample = (f"full batch of {BATCH} targets x {n} timed iterations (+{warmup} warm-up) of the "
              + ("UNMODIFIED reference spaa() imported from baseline/_ref" if kind == "reference" else "oracle port (baseline/_ref absent)")
              + f"; {args.steps} steps requested, bounded to ~150 s of CPU time")
    line = {"impl": "reference", "metric": "spaa_attack_iters_per_sec", "value": its, "unit": "it/s", "n_gpus": args.gpus, "steps": n, "warmup": warmup,
            "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus), "arm": arm_dict("fp32", cpu=True),
            "cpu_baseline": {"value": its, "unit": "it/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": its, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(n_gpus: int):
    """Names the workload only, identically for both arms (`--impl ours` / `--impl reference`); how each arm runs it is in `arm`."""
    return {"workload": "spaa_attack resnet18 B=32 targets/GPU, prj 256x256, cam 240x320, camdE_caml2 d_thr=5 (BASELINE configs[1])",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
            "l2": "per-iteration working set (~3 GB of activations at B=32) is far larger than the 126 MB L2; no explicit flush",
            "parallelism": f"{n_gpus} independent attack jobs, no collective"}


def arm_dict(precision: str, fold_bn: bool = False, cpu: bool = False):
    if cpu:
        return {"precision": "fp32", "classifier": "torchvision resnet18, seeded random init, the reference's stock module on the host cores"}
    return {"precision": precision,
            "classifier": "torchvision resnet18 (cuDNN, external operand, channels_last, TF32 as torch defaults"
                          + ("; private copy: inference-mode BatchNorm folded into the preceding cuDNN convolutions, stem ReLU + max-pooling on spaa_b200's fused kernels; see side leg stock_classifier)"
                             if fold_bn else ")")}



# ---------------------------------------------------------------------------------------------------------------
# training leg: PCNet training step (train_network.py:235-363), data-parallel, one NCCL all-reduce of the flat gradient bucket
# ---------------------------------------------------------------------------------------------------------------
TRAIN_BATCH, TRAIN_N = 24, 500


def train_leg(dev, rank, world, steps, warmup, precision="bf16", dp_mode="weak", phases=(("l1", 0), ("l1+ssim", 401))):
    """img/s of `train_pcnet` (batch 24 per GPU, weak scaling; 500 synthetic pairs resident in HBM) in the L1 phase (iterations
    <= 400) and the L1+SSIM phase (> 400), timed with CUDA events around exactly `steps` optimizer steps (max over ranks)."""
    import random
    import torch.distributed as dist
    import synth
    from spaa_b200 import models, train_network as tn
    torch.manual_seed(123 + rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    prj = torch.rand(TRAIN_N, 3, *PRJ_HW, device=dev, generator=g)
    cam = torch.rand(TRAIN_N, 3, *CAM_HW, device=dev, generator=g)
    scene = synth.textured(rank, "bench.train.scene", (1, 3, *CAM_HW)).to(dev)
    P = synth.pcnet_params(300, CAM_HW)
    model = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=CAM_HW)), torch.nn.DataParallel(models.ShadingNetSPAA()))
    model.load_state_dict(P, strict=True)
    model = models.set_precision(model.to(dev), precision)
    out = {}
    gb = TRAIN_BATCH * world if dp_mode == "weak" else TRAIN_BATCH
    for phase, offset in phases:
        # one train_pcnet call of W + K steps; the first W (>= 5: three eager steps fill the caches, the fourth records the CUDA graph of the
        # step) are untimed, CUDA events bracket exactly the last K
        W = max(5, warmup)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def on_step(i, W=W, e0=e0):
            if i == W:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record()
        cfg = tn.AttrDict(device=str(dev), data_root=None, model_name="PCNet", num_train=TRAIN_N, batch_size=TRAIN_BATCH, max_iters=W + steps, lr=1e-3,
                          lr_drop_ratio=0.2, lr_drop_rate=800, l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, dp_mode=dp_mode, iter_offset=offset,
                          save_checkpoint=False, on_step=on_step)
        random.seed(123)
        tn.train_pcnet(model, dict(cam_scene=scene, cam_train=cam, prj_train=prj, mask=P["mask"]), None, cfg, verbose=False)
        e1.record()
        torch.cuda.synchronize()
        times = [0.0, e0.elapsed_time(e1)]
        t = torch.tensor([times[1]], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item() / steps
        out[phase] = {"img_per_s": gb / (ms / 1e3), "ms_per_step": ms}
    return {"metric": "pcnet_train_img_per_sec", "value": out["l1+ssim"]["img_per_s"], "unit": "img/s", "phases": out, "steps": steps,
            "batch_per_gpu": gb / world, "global_batch": gb, "scaling": "weak" if dp_mode == "weak" else "strong", "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16", "bf16x3": "bf16x3 (fp32-accurate)"}[precision],
            "note": "train_pcnet (3 Adam groups in one flat fp32 bucket, one NCCL all-reduce per step when N>1); " +
                    (("fp16 forward activations / bf16 gradients" if precision == "fp16" else "bf16 activations / gradients") +
                     " on tcgen05 (forward, backward-data, backward-weight), fp32 master weights and accumulation; "
                     if precision not in ("fp32", "bf16x3") else ("exact fp32 CUDA-core convolutions; " if precision == "fp32" else
                                                                  "split-precision (three bf16 parts per value) fp32-accurate convolutions on tcgen05; ")) + "value = L1+SSIM phase (1600 of the reference's 2000 steps)"}


def torch_cuda_train_step_ms(dev, steps=5, batch=TRAIN_BATCH, warm=3):
    """The reference's training step on stock PyTorch ops (oracle port: F.conv2d/grid_sample/SSIM via autograd, torch.optim.Adam) on `dev`:
    the GPU (the torch-CUDA side leg) or the host cores (`train.cpu_baseline`, with a bounded `batch`)."""
    import synth
    from oracle import spaa_oracle as O
    dev = torch.device(dev)
    P = {k: v.to(dev).requires_grad_(v.dtype.is_floating_point and k not in ("mask", "warping_net.ctrl_pts")) for k, v in synth.pcnet_params(300, CAM_HW).items()}
    params = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4)
    g = torch.Generator(device=dev).manual_seed(7)
    prj = torch.rand(batch, 3, *PRJ_HW, device=dev, generator=g)
    cam = torch.rand(batch, 3, *CAM_HW, device=dev, generator=g)
    scene = synth.textured(0, "bench.train.scene", (1, 3, *CAM_HW)).to(dev).expand(batch, -1, -1, -1)

    def step():
        opt.zero_grad()
        loss, _ = O.training_loss(O.pcnet(P, prj, scene, CAM_HW), cam, "l1+ssim")
        loss.backward()
        opt.step()
    for _ in range(warm):
        step()
    if dev.type != "cuda":
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

# ---------------------------------------------------------------------------------------------------------------
# the UNMODIFIED reference on the same B200 through stock PyTorch-CUDA (the denominator of the north-star's ">= 10x" target), with the
# reference's own settings: cudnn.benchmark = True / deterministic = False (utils.py:79-81 set_torch_reproducibility(False)), three
# device->host syncs per iteration, the refinement net on B copies of the grid, two backward passes
# ---------------------------------------------------------------------------------------------------------------

class _RefSettings:
    def __init__(self, exact_fp32: bool):
        self.exact = exact_fp32

    def __enter__(self):
        b = torch.backends
        self.saved = (b.cudnn.benchmark, b.cudnn.deterministic, b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32)
        b.cudnn.benchmark, b.cudnn.deterministic = True, False
        if self.exact:
            b.cudnn.allow_tf32 = b.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        b = torch.backends
        b.cudnn.benchmark, b.cudnn.deterministic, b.cudnn.allow_tf32, b.cuda.matmul.allow_tf32 = self.saved


def torch_cuda_reference_spaa(dev, exact_fp32: bool):
    mods = reference_modules()
    if mods is None:                              # baseline/_ref absent: the oracle port of the same algorithm on the same stock ops
        with _RefSettings(exact_fp32):
            sec = cpu_reference_run(BATCH, 2, 5, device=str(dev))
        return {"value": 1.0 / sec, "unit": "it/s", "steps": 5, "kind": "port", "note": "oracle port on stock PyTorch-CUDA ops (baseline/_ref absent)"}
    R, rh = mods
    with _RefSettings(exact_fp32):
        m, clf, scene, targets = reference_spaa_objects(R, rh, dev)
        sec, n = rh.time_reference_spaa(R, m, clf, targets, scene, D_THR, STEALTH, dev, SETUP, warmup=10, steps=40)
    del m, clf
    torch.cuda.empty_cache()
    return {"value": 1.0 / sec, "unit": "it/s", "steps": n, "kind": "reference",
            "note": "UNMODIFIED reference spaa() (baseline/_ref, projector_based_attack.py:212-339) on stock PyTorch-CUDA, B=32, one call of 50 iterations: 10 untimed "
                    "(cuDNN autotune, cudnn.benchmark=True as utils.py:79-81) + 40 timed; wall clock with synchronize per iteration; "
                    + ("allow_tf32 = False (exact fp32 cuDNN / cuBLAS)" if exact_fp32 else "TF32 as torch defaults (cudnn.allow_tf32=True)")}


def torch_cuda_reference_train(dev):
    """The reference's train_pcnet step (train_network.py:293-320: batch 24 gathered from CPU-resident data, forward, L1+SSIM, three Adam optimisers)."""
    import tempfile
    import synth
    mods = reference_modules()
    if mods is None:
        tms = torch_cuda_train_step_ms(dev)
        return {"img_per_s": TRAIN_BATCH / (tms / 1e3), "ms_per_step": tms, "kind": "port", "note": "oracle port on stock PyTorch-CUDA ops (baseline/_ref absent)"}
    R, rh = mods
    dev = torch.device(dev)
    P = synth.pcnet_params(300, CAM_HW)
    with _RefSettings(False):
        model = torch.nn.DataParallel(rh.ref_pcnet(R, P, CAM_HW, dev), device_ids=[dev.index or 0])
        g = torch.Generator().manual_seed(7)
        n = 96                                   # CPU-resident like the reference's (train_network.py:296-297 copies each batch H2D); 96 of the 500 pairs bound the host RAM / time
        td = dict(cam_scene=synth.textured(0, "bench.train.scene", (1, 3, *CAM_HW)), cam_train=torch.rand(n, 3, *CAM_HW, generator=g),
                  prj_train=torch.rand(n, 3, *PRJ_HW, generator=g), mask=P["mask"])
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "data"), exist_ok=True)
        cfg = R.DictConfig(dict(device=str(dev), data_root=os.path.join(tmp, "data"), setup_name="synth", model_name="PCNet", num_train=n, batch_size=TRAIN_BATCH,
                                max_iters=0, lr=1e-3, lr_drop_ratio=0.2, lr_drop_rate=800, l2_reg=1e-4, plot_on=False, train_plot_rate=50, valid_rate=200, loss="l1+ssim"))
        import random
        random.seed(123)
        sec = rh.time_reference_train_step(R, model, td, cfg, warmup=4, steps=10)
    del model
    torch.cuda.empty_cache()
    return {"img_per_s": TRAIN_BATCH / sec, "ms_per_step": sec * 1e3, "kind": "reference",
            "note": "UNMODIFIED reference train_pcnet (baseline/_ref, train_network.py:235-363) on stock PyTorch-CUDA, batch 24, L1+SSIM phase (iteration counter started at 401), "
                    "cudnn.benchmark=True, TF32 default, batches copied from CPU-resident data every step as the reference does; 4 untimed + 9 timed steps"}


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: PerC-AL + CompenNet++ with vgg16 and inception_v3 (projector_based_attack.py:342-359)
# ---------------------------------------------------------------------------------------------------------------
PERCAL_D_THR = 11.0                               # projector_based_attack.py:186-187


def percal_leg(dev, name, precision, with_reference=True):
    import synth
    from spaa_b200 import models
    from spaa_b200.projector_based_attack import perc_al_compennet_pp
    scene, _, targets = synthetic_inputs(0)
    scene = scene.to(dev)
    C = synth.compennet_pp_params(200)
    cm = models.CompenNetPlusplus(torch.nn.DataParallel(models.WarpingNet(out_size=PRJ_HW)), torch.nn.DataParallel(models.CompenNet()))
    cm.load_state_dict(C, strict=True)
    cm = models.set_precision(cm.to(dev).eval(), precision)
    for p in cm.parameters():
        p.requires_grad = False
    clf = make_classifier(dev, name)
    perc_al_compennet_pp(cm, clf, None, targets, True, scene, PERCAL_D_THR, dev, SETUP, iters=8)          # warm-up call (cuDNN plans, packed weights)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cam_best, prj_best = perc_al_compennet_pp(cm, clf, None, targets, True, scene, PERCAL_D_THR, dev, SETUP)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    out = {"value": 50 / sec, "unit": "it/s", "iters": 50, "ms_per_iter": sec / 50 * 1e3,
           "note": "one perc_al_compennet_pp() call = 50 PerC-AL iterations (classifier forward+backward, fused dE2000 forward+backward, second classifier forward) + one "
                   "CompenNet++ forward; wall clock of the whole call, B=32, d_thr=11"}
    del cm
    mods = reference_modules() if with_reference else None
    if mods is not None:
        R, rh = mods
        labels = {i: f"class{i}" for i in range(1000)}
        with _RefSettings(False):
            rcm = torch.nn.DataParallel(rh.ref_compennet_pp(R, C, PRJ_HW, dev).eval(), device_ids=[dev.index or 0])
            for p in rcm.parameters():
                p.requires_grad = False
            rclf = rh.ref_classifier(R, torch.nn.DataParallel(clf.model, device_ids=[dev.index or 0]), clf.input_sz, dev, name)
            import contextlib, io
            warm = rh.patched(R.pba, "perc_al_compennet_pp", "max_iterations=50", "max_iterations=8")
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                warm(rcm, rclf, labels, targets, True, scene, PERCAL_D_THR, dev, SETUP)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                R.pba.perc_al_compennet_pp(rcm, rclf, labels, targets, True, scene, PERCAL_D_THR, dev, SETUP)
                torch.cuda.synchronize()
                rsec = time.perf_counter() - t0
        out["torch_cuda_reference"] = {"value": 50 / rsec, "unit": "it/s", "kind": "reference",
                                       "note": "UNMODIFIED reference perc_al_compennet_pp() on stock PyTorch-CUDA, same inputs, cudnn.benchmark=True, TF32 default; wall clock of the whole call"}
        del rcm, rclf
    del clf
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: batched attack sweep, 3 classifiers x setups x 100 targets as independent jobs sharded over ranks
# (the job loop of run_projector_based_attack, projector_based_attack.py:83-141), no collective
# ---------------------------------------------------------------------------------------------------------------

def sweep_leg(dev, rank, world, precision, pcnet):
    import torch.distributed as dist
    import synth
    from spaa_b200.projector_based_attack import run_attack_sweep, clear_engines
    n_setups = {1: 1, 2: 3, 4: 5}.get(world, 10)             # 30 jobs at 8 GPUs (the BASELINE sweep); a bounded share of it on fewer GPUs
    names = ("resnet18", "vgg16", "inception_v3")
    clfs = {n: make_classifier(dev, n) for n in names}
    targets = list(range(0, 1000, 10))                        # 100 targets
    jobs = []
    for su in range(n_setups):
        scene = synth.textured(su, "bench.scene", (1, 3, *CAM_HW))
        for n in names:
            jobs.append(dict(model=pcnet, classifier=clfs[n], cam_scene=scene, target_idx=targets, stealth_loss=STEALTH, d_thr=D_THR, setup_info=SETUP,
                             classifier_name=n))
    jobs.sort(key=lambda j: j["classifier_name"])             # same-classifier jobs adjacent per rank: engines (buffers + CUDA graph) are reused
    clear_engines()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # static cost model (ms per iteration of a B=100 batch, measured on B200: PCNet side 9 + classifier leg): balanced LPT partition, no collective
    cost = {"resnet18": 1.0, "inception_v3": 2.2, "vgg16": 2.8}
    res = run_attack_sweep(jobs, dev, iters=50, precision=precision, save=False, rank=rank, world=world, costs=[cost[j["classifier_name"]] for j in jobs])
    torch.cuda.synchronize()
    mine = time.perf_counter() - t0
    t = torch.tensor([mine], device=dev)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    clear_engines()
    torch.cuda.empty_cache()
    return {"jobs": len(jobs), "jobs_per_s": len(jobs) / t.item(), "seconds": t.item(), "targets_per_job": len(targets) + 1, "iters": 50,
            "attack_iterations_per_s": len(jobs) * 2 * 50 / t.item(), "idle_fraction_of_fastest_rank": 1.0 - tmin.item() / t.item(),
            "jobs_this_rank": len(res),
            "note": f"{len(names)} classifiers x {n_setups} synthetic setups, each job = the reference's targeted batch of 100 targets + the untargeted attack of the scene's "
                    "top-1 (projector_based_attack.py:104-125), 50 iterations each, results copied to the host; jobs partitioned over ranks by a static cost model (longest first), no collective; "
                    "wall clock of the slowest rank incl. engine construction and CUDA-graph capture per (classifier, batch size)"}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------

def parity_check(A, P, scene, clf, dev, precision):
    """Numerical self-check of the EXACT configuration that was just timed (same engine, same captured CUDA graph, the projector images the timed
    iterations left behind): one more replayed iteration against the fp32 oracle (oracle/spaa_oracle.py, exact fp32 cuDNN) on the same inputs."""
    from oracle import spaa_oracle as O
    prj0 = A.prj_adv.clone()
    A.step()
    cam, logits = A.cam.clone(), A.logits.clone()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Pd = {k: v.to(dev) for k, v in P.items()}
        with torch.no_grad():
            ref = O.pcnet(Pd, prj0.clamp(0, 1), scene.expand(BATCH, -1, -1, -1), CAM_HW)
            lo_ref = clf.model(O.classifier_preprocess(ref, CROP, clf.input_sz))
            lo_ours = clf.model(O.classifier_preprocess(cam, CROP, clf.input_sz))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    err = (cam - ref).abs()
    tol = 1e-5 if precision == "fp32" else 2e-3
    return {"what": "PCNet output of one replayed iteration of the timed engine vs the fp32 oracle on the same projector images; top-1 of the stock classifier "
                    "(exact fp32) on both camera images, and of the engine's own logits (TF32 cuDNN, folded BatchNorm)",
            "cam_max_abs_err": err.max().item(), "cam_mean_abs_err": err.mean().item(), "tolerance": tol, "within_tolerance": bool(err.max().item() <= tol),
            "top1_agree": int((lo_ref.argmax(1) == lo_ours.argmax(1)).sum()), "top1_agree_engine_logits": int((lo_ref.argmax(1) == logits.argmax(1)).sum()),
            "of": BATCH, "distinct_projector_images": int(torch.unique(prj0.flatten(1)[:, ::4099], dim=0).shape[0])}


def run_ours(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import spaa_b200
    from spaa_b200 import models, ops
    from spaa_b200.projector_based_attack import SpaaAttack, attack_engine, spaa

    scene, P, targets = synthetic_inputs(rank)
    wn = models.WarpingNet(out_size=CAM_HW)
    sn = models.ShadingNetSPAA(use_rough=True)
    pcnet = models.PCNet(P["mask"], torch.nn.DataParallel(wn), torch.nn.DataParallel(sn))
    pcnet.load_state_dict(P, strict=True)
    pcnet = pcnet.to(dev).eval()
    models.set_precision(pcnet, args.precision)
    for p in pcnet.parameters():
        p.requires_grad = False
    clf = make_classifier(dev)

    fold_bn = not args.no_fold_bn and args.precision != "fp32"
    # ---- e2e_cold: the FIRST spaa() call of this process -- the reference's unit of work, one 50-iteration call per (setup, classifier, loss, d_thr),
    # projector_based_attack.py:107-125 -- with host buffers: engine construction, weight packing, cuDNN plan selection and the CUDA-graph capture
    # are all inside the timed region
    scene_host = scene.clone().pin_memory()
    out_cam = torch.empty(BATCH, 3, *CAM_HW).pin_memory()
    out_prj = torch.empty(BATCH, 3, *PRJ_HW).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cold_iters = 0 if args.skip_cold else 50
    if cold_iters:
        cam_best, prj_best = spaa(pcnet, clf, None, targets, True, scene_host.to(dev, non_blocking=True), D_THR, STEALTH, dev, SETUP, iters=cold_iters,
                                  graph=not args.no_graph, fold_bn=fold_bn, deterministic=args.deterministic)
        out_cam.copy_(cam_best, non_blocking=True)
        out_prj.copy_(prj_best, non_blocking=True)
    torch.cuda.synchronize()
    cold_s = time.perf_counter() - t0
    # the engine spaa() built for this job (kept warm across calls of a sweep: buffers + captured CUDA graph)
    A = attack_engine(pcnet, clf, targets, True, scene.to(dev), D_THR, STEALTH, dev, SETUP, graph=not args.no_graph, fold_bn=fold_bn,
                      deterministic=args.deterministic)
    for _ in range(args.warmup):                            # 2 eager iterations, then the CUDA graph is captured and replayed
        A.step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K iterations, CUDA events on the launching stream -----------------------------------
    clocks = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    barrier()
    e0.record()
    for _ in range(args.steps):
        A.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - n0
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    its = args.steps / (ms / 1e3) * world                   # whole-job iterations/s (each rank runs its own 32-target batch)

    parity = parity_check(A, P, scene.to(dev), clf, dev, args.precision)

    # ---- roofline kernel: the same K iterations re-run WITHOUT graph replay so that every launch of the dominant kernel family
    # (conv4 / conv4_s / conv5 forward + backward-data) can be bracketed by a CUDA-event pair on the launching stream --------------
    Ap = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), D_THR, STEALTH, dev, SETUP, graph=False, fold_bn=fold_bn)
    for _ in range(3):
        Ap.step()
    probe = ops.set_probe(lambda kind, spec: spec.k == 3 and {spec.cin, spec.cout} == {128, 256})
    for _ in range(args.steps):
        # host-launched steps are bound by the host (~50 us of Python / ctypes / tensor-map encoding per launch), and an event pair around a launch the
        # GPU is already waiting for would time that host work too: park the GPU on a ~30 ms spin first, so that the whole step is queued behind it and
        # its kernels run back to back
        torch.cuda._sleep(60_000_000)
        Ap.step()
    torch.cuda.synchronize()
    ops.set_probe(None)
    kern_ms = [a.elapsed_time(b) for a, b in probe["events"]]
    del Ap

    # ---- classifier share (external operand, reported separately) --------------------------------------------------
    from spaa_b200.projector_based_attack import _adv_grad
    # timed as it runs inside the iteration: replayed from a CUDA graph (host-launched it is bound by the launch rate of its ~150 short kernels)
    def clf_leg():
        return _adv_grad(A.classifier, A.cam, CROP, A.target, True, A.clf_cl)
    clf_leg()
    clf_replay, clf_how = clf_leg, "host-launched"
    if not args.no_graph:
        try:
            cg, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                cg.capture_begin()
                try:
                    _keep = clf_leg()
                finally:
                    cg.capture_end()
            torch.cuda.current_stream(dev).wait_stream(side)
            clf_replay, clf_how = cg.replay, "CUDA-graph replay"
        except Exception as e:                                              # keep the host-launched number
            print(f"bench: classifier leg not captured ({type(e).__name__}: {e})", file=sys.stderr)
    clf_replay()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(10):
        clf_replay()
    c1.record()
    torch.cuda.synchronize()
    clf_ms = c0.elapsed_time(c1) / 10

    # ---- e2e: the public spaa() call with HOST buffers, copies inside the timed region -----------------------------
    barrier()
    t0 = time.perf_counter()
    cam_best, prj_best = spaa(pcnet, clf, None, targets, True, scene_host.to(dev, non_blocking=True), D_THR, STEALTH, dev, SETUP, iters=args.steps,
                              graph=not args.no_graph, fold_bn=fold_bn, deterministic=args.deterministic)
    out_cam.copy_(cam_best, non_blocking=True)
    out_prj.copy_(prj_best, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_its = args.steps / t.item() * world
    h2d = scene_host.numel() * 4 + len(targets) * 8
    d2h = (out_cam.numel() + out_prj.numel()) * 4

    train = None
    if not args.skip_train:
        from spaa_b200.projector_based_attack import clear_engines
        clear_engines()
        del A
        torch.cuda.empty_cache()
        train = train_leg(dev, rank, world, max(3, min(args.steps, 10)), 2, args.train_precision)
        # BASELINE configs[3] with the reference's own batch: ONE global batch of 24 (train_network.py:293-297) sharded over the ranks (strong scaling);
        # at N = 1 it is the same step as the weak leg
        if world > 1:
            st = train_leg(dev, rank, world, max(3, min(args.steps, 10)), 2, args.train_precision, dp_mode="global", phases=(("l1+ssim", 401),))
            train["strong"] = {"img_per_s": st["value"], "ms_per_step": st["phases"]["l1+ssim"]["ms_per_step"], "global_batch": TRAIN_BATCH,
                               "batch_per_gpu": TRAIN_BATCH / world, "scaling": "strong",
                               "note": "dp_mode='global': every rank draws the same seeded 24-image sample and takes a strided share; one NCCL all-reduce of the flat bucket"}
        else:
            train["strong"] = {"img_per_s": train["value"], "ms_per_step": train["phases"]["l1+ssim"]["ms_per_step"], "global_batch": TRAIN_BATCH,
                               "batch_per_gpu": TRAIN_BATCH, "scaling": "strong", "note": "N = 1: identical to the weak-scaling step"}
        if world == 1 and not args.skip_side_legs and args.train_precision != "fp32":
            t32 = train_leg(dev, rank, world, 3, 1, "fp32")
            train["fp32_mode"] = {"img_per_s": t32["value"], "ms_per_step": t32["phases"]["l1+ssim"]["ms_per_step"]}
            if args.train_precision != "bf16":
                tb = train_leg(dev, rank, world, max(3, min(args.steps, 10)), 2, "bf16", phases=(("l1+ssim", 401),))
                train["bf16_mode"] = {"img_per_s": tb["value"], "ms_per_step": tb["phases"]["l1+ssim"]["ms_per_step"],
                                      "note": "pure bf16 storage: PCNet output 3.3e-3 max-abs from the fp32 oracle (misses the 2e-3 bar the fp16 mode meets)"}
            try:                                      # fp32-accurate training on the tensor cores (split-precision operands, forward + backward-data + backward-weight)
                tx3 = train_leg(dev, rank, world, 5, 1, "bf16x3", phases=(("l1+ssim", 401),))
                train["bf16x3_mode"] = {"img_per_s": tx3["value"], "ms_per_step": tx3["phases"]["l1+ssim"]["ms_per_step"]}
            except Exception as e:
                train["bf16x3_mode"] = {"unavailable": f"{type(e).__name__}: {e}"}
    sweep = None
    if not args.skip_sweep:
        try:
            sweep = sweep_leg(dev, rank, world, args.precision, pcnet)
        except Exception as e:                                  # a reported side leg: never lose the bench line over it
            sweep = {"unavailable": f"{type(e).__name__}: {e}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    k_ms = sum(kern_ms) / max(1, len(kern_ms))
    flop = HEAVY_FLOP_PER_SAMPLE * BATCH
    achieved = flop / (k_ms * 1e-3) / 1e12 if k_ms else 0.0
    # denominator: the BURST cuBLAS figure when the timed region held the maximum SM clock (a 60 ms region does not reach the power cap),
    # the sustained one when the clocks sagged under load (B200_PROFILING.md)
    burst = bool(clk.get("sm_mhz") and clk.get("sm_max_mhz") and clk["sm_mhz"] >= 0.95 * clk["sm_max_mhz"] and "sw_power_cap" not in clk.get("reasons", []))
    peak_tf, peak_name = (pk["bf16_tflops"], "burst") if burst else (pk["bf16_tflops_sustained"], "sustained")
    line = {"metric": "spaa_attack_iters_per_sec", "value": its, "unit": "it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16", "bf16x3": "bf16x3"}[args.precision], "data": "synthetic", "config": config_dict(world), "arm": arm_dict(args.precision, fold_bn),
            "sample_iters_per_sec": its * BATCH, "classifier_ms_per_step": clf_ms, "classifier_timing": clf_how,
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_its, "unit": "it/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "note": "one spaa() call of `steps` iterations on a warm engine (as in a sweep): scene H2D from pinned memory + results D2H inside the timed region; bytes are per call / steps"},
            "roofline": {"bound": "tensor", "kernel": ("conv_halo_kernel<{128,256},64> (tcgen05.mma M128 x N{128,256} x K16, halo-tile TMA) on the 128<->256-channel 3x3 layers" if args.precision != "fp32"
                                    else "conv_gather_kernel<128,64,8,4> on the 128<->256-channel 3x3 layers (fp32 CUDA-core path)"),
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "frac_of_sustained_peak": achieved / pk["bf16_tflops_sustained"],
                         "traffic": (NCU_HEAVY_DRAM_BYTES if args.precision != "fp32" else None), "traffic_unit": "bytes per launch (dram read + write, ncu --set full: profiles/r2_halo_ncu_full.md)",
                         "algorithmic_flop_per_launch": flop, "launches_timed": len(kern_ms), "avg_launch_ms": k_ms, "peak_source": pk["source"] + " bf16 " + peak_name,
                         "note": "per-launch CUDA events need host-launched kernels: timed over the same K iterations re-run without graph replay"},
            "cuda_graph": not args.no_graph,
            "e2e_cold": {"value": (cold_iters / cold_s) if cold_iters else None, "unit": "it/s", "seconds": cold_s, "iters": cold_iters,
                         "note": "FIRST spaa(iters=50) call of the process with host buffers: engine construction, weight packing, cuDNN plan selection, two eager "
                                 "iterations and the CUDA-graph capture inside the timed region (rank 0)"},
            "parity_check": parity}
    if train is not None:
        line["train"] = train
    if sweep is not None:
        line["sweep"] = sweep
    if world == 1 and not args.skip_side_legs:
        # side legs (not the headline): the exact fp32 mode of the same engine, and the reference algorithm on stock PyTorch-CUDA ops
        # (cuDNN TF32 convolutions + ~1.3k ATen kernels per iteration = what the reference executes on a GPU), same batch, same box
        if fold_bn:
            As = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), D_THR, STEALTH, dev, SETUP, fold_bn=False)
            ns = max(5, min(args.steps, 20))
            for _ in range(3):
                As.step()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); f0.record()
            for _ in range(ns):
                As.step()
            f1.record(); torch.cuda.synchronize()
            line["stock_classifier"] = {"value": ns / (f0.elapsed_time(f1) / 1e3), "unit": "it/s", "steps": ns,
                                        "note": "same engine and CUDA-graph replay with the classifier run as the stock module (BatchNorm layers unfolded, ATen pooling)"}
            del As
        if args.precision != "fp32":
            models.set_precision(pcnet, "fp32")
            A32 = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), D_THR, STEALTH, dev, SETUP)
            n32 = max(3, min(args.steps, 10))
            for _ in range(3):
                A32.step()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); f0.record()
            for _ in range(n32):
                A32.step()
            f1.record(); torch.cuda.synchronize()
            line["fp32_mode"] = {"value": n32 / (f0.elapsed_time(f1) / 1e3), "unit": "it/s", "steps": n32,
                                 "note": "same engine, exact CUDA-core fp32 convolutions (the 1e-5 parity mode)"}
            del A32
            # the fp32-ACCURATE tensor-core mode: bf16x3 split-precision operands (three bf16 parts per value, six part products per multiply,
            # fp32 accumulation in TMEM) on the same tcgen05 kernel -- parity-tested against the fp32 oracle at 1e-5 (tests/test_gpu_fullsize.py)
            try:
                models.set_precision(pcnet, "bf16x3")
                Ax = SpaaAttack(pcnet, clf, targets, True, scene.to(dev), D_THR, STEALTH, dev, SETUP)
                for _ in range(3):
                    Ax.step()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); f0.record()
                for _ in range(n32):
                    Ax.step()
                f1.record(); torch.cuda.synchronize()
                line["bf16x3_mode"] = {"value": n32 / (f0.elapsed_time(f1) / 1e3), "unit": "it/s", "steps": n32,
                                       "parity_check": parity_check(Ax, P, scene.to(dev), clf, dev, "fp32"),
                                       "note": "same engine, fp32-accurate split-precision convolutions on tcgen05 (precision='bf16x3')"}
                del Ax
            except Exception as e:
                line["bf16x3_mode"] = {"unavailable": f"{type(e).__name__}: {e}"}
            models.set_precision(pcnet, args.precision)
        def leg(fn, *a_):
            try:
                return fn(*a_)
            except Exception as e:               # reported side numbers: never lose the bench line over one
                return {"unavailable": f"{type(e).__name__}: {e}"}
        if train is not None:
            train["torch_cuda_reference"] = leg(torch_cuda_reference_train, dev)
        line["torch_cuda_reference"] = leg(torch_cuda_reference_spaa, dev, False)
        if isinstance(line["torch_cuda_reference"], dict):
            line["torch_cuda_reference"]["exact_fp32"] = leg(torch_cuda_reference_spaa, dev, True)      # SURVEY.md 8d asks for both TF32 settings
        # BASELINE configs[2]: PerC-AL + CompenNet++ with vgg16 and inception_v3, ours and the unmodified reference on the same GPU
        line["percal"] = {n: leg(percal_leg, dev, n, args.precision) for n in ("vgg16", "inception_v3")}
    if world == 1 and not args.skip_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample_B = 8
        try:
            sec, n_it, kind = reference_cpu_spaa(sample_B, 1, 3, budget_s=60.0)
        except Exception as e:                   # never lose the bench line over the baseline leg: fall back to the oracle port
            print(f"bench: reference CPU leg failed ({type(e).__name__}: {e}); timing the oracle port instead", file=sys.stderr)
            sec, n_it, kind = cpu_reference_run(sample_B, 1, 3), 3, "port"
        line["cpu_baseline"] = {"value": 1.0 / (sec * BATCH / sample_B), "unit": "it/s", "cores": cores, "kind": kind,
                                "sample": f"{sample_B} of {BATCH} targets x {n_it} timed iterations (+1 warm-up) of "
                                          + ("the UNMODIFIED reference spaa() (baseline/_ref)" if kind == "reference" else "the oracle port")
                                          + f" on the host cores, scaled x{BATCH // sample_B} (CPU time is linear in the batch); the full batch: `bench.py --impl reference`"}
        if train is not None:
            try:                                 # the training step of the reference on the host cores, bounded: 4 of the 24 images, 2 steps
                tb = 4
                tms = torch_cuda_train_step_ms("cpu", steps=2, batch=tb, warm=1)
                train["cpu_baseline"] = {"value": tb / (tms / 1e3), "unit": "img/s", "cores": cores, "kind": "port",
                                         "sample": f"batch of {tb} (of {TRAIN_BATCH}) x 2 timed steps (+1 warm-up) of the oracle port of train_pcnet's step (L1+SSIM, Adam); "
                                                   "img/s does not depend on the batch size on the CPU"}
            except Exception as e:               # a reported side number: never lose the bench line over it
                train["cpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp32", "bf16", "fp16", "bf16x3"],
                    help="fp16 (default): tcgen05 convolutions, fp16 activations / bf16 gradients, fp32 accumulation -- the 16-bit mode that meets "
                         "BASELINE.json's 2e-3 bar; bf16: pure bf16 storage; fp32: exact CUDA-core convolutions (1e-5 parity mode)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured CUDA graph")
    ap.add_argument("--skip-train", action="store_true", help="omit the PCNet training leg")
    ap.add_argument("--train-precision", default="fp16", choices=["fp32", "bf16", "fp16", "bf16x3"],
                    help="fp16 (default): the same 16-bit mode as the attack (fp16 forward activations -- PCNet output within 2e-3 --, bf16 gradients, every "
                         "convolution incl. backward-weight on tcgen05); bf16: pure bf16 storage (faster, forward 3.3e-3: side leg train.bf16_mode)")
    ap.add_argument("--no-fold-bn", action="store_true", help="run the external classifier as the stock module (BatchNorm layers not folded)")
    ap.add_argument("--skip-side-legs", action="store_true", help="profiling runs only: omit the fp32 and torch-CUDA side measurements")
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only: omit the CPU baseline leg")
    ap.add_argument("--skip-sweep", action="store_true", help="omit the attack-sweep leg (BASELINE configs[4])")
    ap.add_argument("--deterministic", action="store_true", help="warp adjoint as the atomics-free tiled gather (bit-reproducible run to run)")
    ap.add_argument("--skip-cold", action="store_true", help="profiling runs only: omit the cold first spaa(iters=50) call (e2e_cold)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
