"""SPAA and PerC-AL+CompenNet++ attackers -- API of /root/reference/src/python/projector_based_attack.py:212-359.

`spaa` keeps the reference signature and result, but one iteration is a fixed sequence of sm_100a kernels with no
host synchronisation:
  warp (grid_sample of clamp(prj) * mask, x*s fused) -> 14 fused conv kernels -> classifier (cuDNN, external)
  -> fused Lab + dE2000 + L2 loss/gradient -> device-side decision masks -> per-sample cotangent select
  -> ONE backward through PCNet (each sample consumes exactly one of the reference's two gradients)
  -> grid_sample scatter -> per-sample norm -> fused normalised step + best-so-far copy.
Loop constants the reference recomputes every iteration (sampling grid, skipConv1(scene), Lab(scene)) are hoisted.
"""
from __future__ import annotations

import collections
import itertools
import os
import weakref
from typing import List, Optional

import torch

from . import ops
from .classifier import device_logits, fold_batchnorm, use_channels_last
from .img_proc import expand_4d
from .models import PCNet, _Stack, set_precision
from .ops import MASK_OPEN01
from .perc_al import PerC_AL


_SCOPES = itertools.count(1)


def _unwrap(m):
    return m.module if isinstance(m, torch.nn.DataParallel) else m


def _adv_grad(classifier, cam, cp_sz, target, targeted, channels_last: bool = False):
    """logits (detached) and d(-+ sum_b logit[b, target_b]) / d cam through the external classifier."""
    leaf = cam.detach().requires_grad_(True)
    with torch.enable_grad():
        logits = device_logits(classifier, leaf, cp_sz, channels_last)
        sel = logits.gather(1, target.view(-1, 1)).sum()
        loss = -sel if targeted else sel
        g, = torch.autograd.grad(loss, leaf)
    return logits.detach(), g


class SpaaAttack:
    """State + one-iteration stepper of the SPAA loop (projector_based_attack.py:212-339).  `spaa()` below is
    `SpaaAttack(...)`, `iters` x `step()`, `result()`; bench.py drives `step()` directly to time exactly K iterations."""

    def __init__(self, pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, precision=None,
                 graph: Optional[bool] = None, fold_bn: Optional[bool] = None, deterministic: bool = False, overlap: Optional[bool] = None):
        """overlap: run the fused colour loss (and the prjl2 term), which only need the PCNet output, on an auxiliary stream beside the
        external classifier's forward + backward (fork / join, recorded as a parallel branch of the captured graph).  Same kernels
        and arithmetic.  Opt-in (default None: $SPAA_OVERLAP, off): measured on B200 at B=32 it is 2 % SLOWER (290 vs 296 it/s,
        profiles/r1_overlap_probe.md) -- the issue-bound colour kernel takes SM time from the cuDNN kernels it runs beside.
        deterministic: the warp's backward runs as a gather through a per-attack CSR adjoint map instead of a scatter-add with atomics
        (same result to fp32 rounding, bit-identical run to run, ~20 us slower per iteration at B=32).
        graph: replay one captured CUDA graph per iteration (default: on for the fused PCNet path).  An iteration is a fixed
        sequence of ~80 of our launches + ~300 cuDNN/ATen launches of the external classifier with no host decision in
        between, so after two eager iterations (which fill the packed-weight / workspace caches) the third is captured and
        every later step() is a single cudaGraphLaunch.
        fold_bn: run the frozen external classifier through a private copy with its inference-mode BatchNorm layers folded into
        the preceding cuDNN convolutions (classifier.fold_batchnorm).  Default (None): on exactly when cuDNN may use TF32
        (torch's default, the reference's setting) -- i.e. off in the exact-fp32 parity mode, like the channels_last layout."""
        device = torch.device(device)
        if precision is not None:
            set_precision(_unwrap(pcnet), precision)
        if device.type != "cuda":
            raise RuntimeError("spaa_b200.spaa runs on CUDA devices only (no CPU fallback)")
        self.device, self.classifier, self.targeted, self.d_thr = device, classifier, bool(targeted), float(d_thr)
        self.B = B = len(target_idx)
        self.cp_sz = setup_info["classifier_crop_sz"]
        self.prj_hw = prj_hw = tuple(setup_info["prj_im_sz"])
        self.scene = scene = ops._f32c(expand_4d(cam_scene).to(device)).clone()      # own copy: reset() overwrites it in place
        if scene.shape[0] != 1:
            raise ValueError("cam_scene must be a single image (3xHxW or 1x3xHxW)")
        H, W = scene.shape[2:]
        net_ = _unwrap(pcnet)
        if isinstance(net_, PCNet):
            if tuple(net_.warping_net.out_size) != (H, W):
                raise ValueError(f"the PCNet warps to {tuple(net_.warping_net.out_size)} but cam_scene is {(H, W)} (the reference fails in models.py:340-342)")
            if net_.use_mask and net_.mask.numel() != H * W:
                raise ValueError(f"the PCNet mask has {net_.mask.numel()} elements but cam_scene is {H}x{W}")
        self.hw_cam, self.hw_prj = H * W, prj_hw[0] * prj_hw[1]
        self.target = torch.as_tensor(list(target_idx), dtype=torch.int64, device=device)
        self.gray = setup_info["prj_brightness"] * torch.ones(B, 3, *prj_hw, device=device)
        self.prj_adv = self.gray.clone()
        self.adv_lr, self.col_lr = 2.0, 1.0                                    # :243-244
        self.w_prjl2 = 0.1 if "prjl2" in stealth_loss else 0.0                 # :249-251
        self.w_caml2 = 1.0 if "caml2" in stealth_loss else 0.0
        self.w_camde = 1.0 if "camdE" in stealth_loss else 0.0
        self.p_thresh = 0.9                                                    # :255
        self.prj_best = self.prj_adv.clone()
        self.cam_best = scene.repeat(B, 1, 1, 1)
        self.best_col = 1e6 * torch.ones(B, device=device)
        self.use_col, self.succ, self.better = (torch.zeros(B, dtype=torch.uint8, device=device) for _ in range(3))
        self.col_loss = torch.empty(B, device=device)
        self.stats = torch.empty(B, 4, device=device)
        self.sq = torch.empty(B, device=device)
        self.prjl2sum = torch.empty(B, device=device) if self.w_prjl2 else None
        self.step2 = torch.tensor([-self.adv_lr, -self.col_lr], device=device)
        self.ref_lab = None                        # set below, once the colour arithmetic is known
        self.g_col = torch.empty(B, 3, H, W, device=device)
        self.d_pre6 = torch.empty(B, 3, H, W, device=device)
        self.dprj = torch.empty(B, 3, *prj_hw, device=device)
        self.pcnet = pcnet
        self.net = net = _unwrap(pcnet)
        self.fused = isinstance(net, PCNet)
        self.warp_adj = None
        if self.fused:
            with torch.no_grad():
                self.sh = sh = net.shading_net
                self.grid = net.warping_net.planar_grid(prj_hw).detach()
                self.mask = net.flat_mask()
                # measured at B=32: scatter-add (atomics) 52 us + norm 17 us + zero-fill vs 95 us for the gather form -- the gather is the
                # opt-in, bit-reproducible variant
                self.warp_adj = ops.WarpAdjoint(self.grid, prj_hw, self.mask) if deterministic else None
                self.skip_acts = _Stack.skip1(sh, scene)                      # skipConv1(cam_scene): loop constant
                self.xw = torch.empty(B, 3, H, W, device=device)
                self.tc = _Stack.act_dtype(sh) != torch.float32          # tensor-core path: packed 16-channel boundary tensors
                self.split = _Stack.split(sh)                            # 'bf16x3': three bf16 parts per channel (48 physical channels)
                if self.tc:
                    pc = 48 if self.split else 16
                    self.packed = torch.empty((B, pc, H, W), dtype=_Stack.act_dtype(sh), device=device, memory_format=torch.channels_last)
                    self.d_pre6_packed = torch.empty((B, pc, H, W), dtype=_Stack.grad_dtype(sh), device=device, memory_format=torch.channels_last)
                if net.use_rough:
                    self.sfeat = torch.empty(B, 6, H, W, device=device)
                    self.sfeat[:, :3] = scene
                    self.surf_acts = None
                else:
                    self.sfeat = None
                    self.surf_acts = _Stack.surface_branch(sh, scene) if sh.res1_s is None else tuple(
                        t if t.dim() == 4 else t.unsqueeze(0) for t in (sh.res1_s, sh.res2_s, sh.res3_s, sh.res4_s))
        self.scene_b = scene.expand(B, -1, -1, -1)
        self.cam = self.logits = None
        self.clf_cl = use_channels_last(classifier) if device.type == "cuda" else False
        self.fold_bn = bool(torch.backends.cudnn.allow_tf32) if fold_bn is None else bool(fold_bn)
        if self.fold_bn:
            self.classifier = fold_batchnorm(classifier)
        self.use_graph = self.fused if graph is None else (bool(graph) and self.fused)
        self._graph, self._n_eager = None, 0
        self.overlap = (os.environ.get("SPAA_OVERLAP", "0") != "0") if overlap is None else bool(overlap)
        self._aux = torch.cuda.Stream(device=device) if self.overlap else None
        # colour arithmetic: exact in the fp32-level modes ('fp32', 'bf16x3'); the 16-bit modes ('fp16', 'bf16'), whose camera image already carries
        # 4e-4 .. 3e-3 of storage rounding, use the hardware approximations (MUFU-based, ~1e-6 relative: 40 % fewer instructions in an
        # instruction-bound kernel, measured +1.3 % it/s at B=32) -- the reference Lab image is then computed with the same arithmetic so that
        # equal pixels keep dE = 0 exactly.  $SPAA_FAST_COLOR=0 keeps the exact arithmetic everywhere.
        self.fast_color = bool(self.fused and getattr(self, "tc", False) and not getattr(self, "split", False) and os.environ.get("SPAA_FAST_COLOR", "1") != "0")
        self.ref_lab = ops.rgb2lab(scene, fast=self.fast_color)
        self._scope = next(_SCOPES)                 # private kernel workspaces: engines may run concurrently on different streams

    def step(self):
        """One iteration of the loop body (:265-328)."""
        if self._graph is not None:
            self._graph.replay()
            ops._count(self._graph_launches)
            return
        if self.use_graph and self._n_eager >= 2 and not torch.cuda.is_current_stream_capturing():
            self._capture()
            self._graph.replay()
            return
        self._step_eager()
        self._n_eager += 1

    def _capture(self):
        """Record one iteration into a CUDA graph.  Manual capture_begin/capture_end on a side stream instead of the
        `torch.cuda.graph` context: that context empties the caching allocator first, which turns every capture into seconds
        of cudaFree/cudaMalloc of the multi-GB activation cache (measured 47 ms .. 3 s)."""
        # every tensor whose address is baked into the graph must outlive it: the packed 16-bit weights live in a cache
        self._keepalive = [v[2] for v in ops._packed_cache.values()]
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        n0 = ops.launch_count()
        pool = ops.graph_pool(self.device)
        with torch.cuda.stream(side):
            g.capture_begin(pool=pool) if pool is not None else g.capture_begin()
            try:
                self._step_eager()                  # records the launches only; the caller's replay executes this iteration
            finally:
                g.capture_end()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph_launches = ops.launch_count() - n0      # our kernels per replay (the counter ran during capture)
        self._graph = g

    def reset(self, cam_scene, target_idx):
        """Start a new attack job (new scene / targets, same shapes and configuration) on this engine: every buffer whose
        address the captured graph holds is refreshed IN PLACE, so the graph is reused (no re-capture)."""
        scene = ops._f32c(expand_4d(cam_scene).to(self.device))
        if scene.shape != self.scene.shape or len(target_idx) != self.B:
            raise ValueError("reset() needs the scene shape and batch size the engine was built for")
        with torch.no_grad():
            self.scene.copy_(scene)
            self.target.copy_(torch.as_tensor(list(target_idx), dtype=torch.int64))
            self.prj_adv.copy_(self.gray)
            self.prj_best.copy_(self.gray)
            self.cam_best.copy_(self.scene.expand(self.B, -1, -1, -1))
            self.best_col.fill_(1e6)
            for t in (self.use_col, self.succ, self.better):
                t.zero_()
            self.ref_lab.copy_(ops.rgb2lab(self.scene, fast=self.fast_color))
            if self.fused:
                for dst, src in zip(self.skip_acts, _Stack.skip1(self.sh, self.scene)):
                    dst.copy_(src)
                if self.sfeat is not None:
                    self.sfeat[:, :3] = self.scene
                elif self.sh.res1_s is None:
                    for dst, src in zip(self.surf_acts, _Stack.surface_branch(self.sh, self.scene)):
                        dst.copy_(src)
        if self._graph is None:                   # (a captured graph keeps writing the same `cam` / `logits` buffers: they stay valid across jobs)
            self.cam = self.logits = None
        return self

    def __del__(self):
        try:
            ops.release_scope(self._scope)
        except Exception:
            pass

    def _step_eager(self):
        with ops.ws_scope(self._scope):
            self._iteration()

    def _stealth_terms(self, cam):
        ops.color_loss(cam, self.scene, self.ref_lab, cam_is_lab2=False, de_weighting=False, c_de=self.w_camde / self.hw_cam,
                       c_l2=self.w_caml2 / self.hw_cam, stats=self.stats, grad=self.g_col, fast=self.fast_color)
        if self.w_prjl2:
            ops.chan_l2(self.prj_adv, self.gray, self.prjl2sum)

    def _iteration(self):
        net, scene = self.net, self.scene
        # ---- forward ---------------------------------------------------------------------------------
        if self.fused:
            with torch.no_grad():
                if self.tc and self.split:
                    # exact fp32 warp, then ONE pass that writes [x | s | x*s | 0] as three bf16 parts per channel
                    if net.use_rough:
                        ops.grid_sample(self.prj_adv, self.grid, clamp01=True, mask=self.mask, out=self.xw, rough=scene, out2=self.sfeat[:, 3:])
                        ops.pack_nhwc16(self.xw, self.sfeat, self.packed.dtype, split=True, out=self.packed)
                    else:
                        ops.grid_sample(self.prj_adv, self.grid, clamp01=True, mask=self.mask, out=self.xw)
                        ops.pack_nhwc16(self.xw, None, self.packed.dtype, split=True, out=self.packed)
                    cam, S = _Stack.forward(self.sh, None, None, None, surf_acts=self.surf_acts, skip_acts=self.skip_acts, packed=self.packed)
                    if net.use_rough:
                        S["surf_own"] = True
                elif self.tc:
                    ops.grid_sample_packed(self.prj_adv, self.grid, self.packed.dtype, clamp01=True, mask=self.mask, rough=scene, out=self.packed)
                    cam, S = _Stack.forward(self.sh, None, None, None, surf_acts=self.surf_acts, skip_acts=self.skip_acts, packed=self.packed)
                elif net.use_rough:
                    ops.grid_sample(self.prj_adv, self.grid, clamp01=True, mask=self.mask, out=self.xw, rough=scene, out2=self.sfeat[:, 3:])
                    cam, S = _Stack.forward(self.sh, self.xw, self.sfeat, None, skip_acts=self.skip_acts)
                    S["surf_own"] = True
                else:
                    ops.grid_sample(self.prj_adv, self.grid, clamp01=True, mask=self.mask, out=self.xw)
                    cam, S = _Stack.forward(self.sh, self.xw, None, None, surf_acts=self.surf_acts, skip_acts=self.skip_acts)
        else:
            prj_leaf = self.prj_adv.detach().requires_grad_(True)
            with torch.enable_grad():
                cam_g = self.pcnet(torch.clamp(prj_leaf, 0, 1), self.scene_b)
            cam = cam_g.detach()
        # ---- losses, masks ---------------------------------------------------------------------------
        if self._aux is not None:
            main = torch.cuda.current_stream(self.device)
            self._aux.wait_stream(main)                                        # fork: cam is ready
            with torch.cuda.stream(self._aux):
                self._stealth_terms(cam)
            logits, g_adv = _adv_grad(self.classifier, cam, self.cp_sz, self.target, self.targeted, self.clf_cl)
            main.wait_stream(self._aux)                                        # join before the decisions read the statistics
        else:
            logits, g_adv = _adv_grad(self.classifier, cam, self.cp_sz, self.target, self.targeted, self.clf_cl)
            self._stealth_terms(cam)
        ops.attack_masks(logits, self.target, self.targeted, self.stats, self.prjl2sum, self.hw_cam, self.hw_prj, self.w_prjl2, self.w_caml2,
                         self.w_camde, self.d_thr, self.p_thresh, self.use_col, self.succ, self.better, self.col_loss, self.best_col)
        # ---- one backward with the per-sample selected cotangent ---------------------------------------
        if self.fused:
            d_pre6 = d_pk = None
            if self.tc and self.split:
                ops.select_cotangent(g_adv, self.g_col, self.use_col, cam, MASK_OPEN01, self.d_pre6)
                d_pk = ops.pack_nhwc16(self.d_pre6, None, self.d_pre6_packed.dtype, split=True, out=self.d_pre6_packed)
            elif self.tc:
                d_pk = ops.select_cotangent_packed(g_adv, self.g_col, self.use_col, cam, MASK_OPEN01, self.d_pre6_packed)
            else:
                d_pre6 = ops.select_cotangent(g_adv, self.g_col, self.use_col, cam, MASK_OPEN01, self.d_pre6)
            with torch.no_grad():
                if net.use_rough:
                    dxw, dsf, _ = _Stack.backward(self.sh, S, d_pre6, need_dx=True, surf_grad_channels=(3, 6), d_pre6_packed=d_pk)
                else:
                    dxw, dsf, _ = _Stack.backward(self.sh, S, d_pre6, need_dx=True, d_pre6_packed=d_pk)
                if self.warp_adj is not None:
                    # deterministic mode: adjoint of the warp as a gather through the per-attack CSR map (no atomics); without the prjl2 term the
                    # per-sample squared norm of the clamp-masked gradient (:307,315) comes out of the same pass
                    ops.grid_sample_bwd_gather(self.warp_adj, dxw, dout2=dsf, rough=scene if dsf is not None else None, dimg=self.dprj,
                                               sq=None if self.w_prjl2 else self.sq, x_for_clamp=None if self.w_prjl2 else self.prj_adv)
                else:
                    ops.grid_sample_bwd_input(dxw, self.grid, self.prj_hw, mask=self.mask, dout2=dsf, rough=scene if dsf is not None else None,
                                              dimg=self.dprj)
            S = None
        else:
            ops.select_cotangent(g_adv, self.g_col, self.use_col, None, 0, self.d_pre6)
            g, = torch.autograd.grad(cam_g, prj_leaf, grad_outputs=self.d_pre6)      # clamp backward already applied by autograd
            self.dprj.copy_(g)
        # ---- normalised masked step + best-so-far ------------------------------------------------------
        clamp_in_kernel = self.fused
        if self.w_prjl2:
            ops.chan_l2(self.prj_adv, self.gray, self.sq, c=self.w_prjl2 / self.hw_prj, sel=self.use_col, apply_clamp_mask=clamp_in_kernel,
                        grad=self.dprj)
            clamp_in_kernel = False
        if not (self.fused and self.warp_adj is not None and not self.w_prjl2):    # (deterministic mode without prjl2: the gather kernel produced sq)
            ops.row_sqnorm(self.dprj, self.sq, self.prj_adv if clamp_in_kernel else None)
        ops.row_normalized_step(self.prj_adv, self.dprj, self.sq, self.step2, self.use_col, use_clamp_mask=clamp_in_kernel,
                                copy_dst=self.prj_best, copy_sel=self.succ)
        ops.masked_copy_rows(self.cam_best, cam, self.succ)
        self.cam, self.logits = cam, logits

    def result(self):
        return self.cam_best, torch.clamp(self.prj_best, 0, 1)                 # :337-339


_ENGINES: "collections.OrderedDict" = collections.OrderedDict()
_MAX_ENGINES = 2


def _state_version(module) -> int:
    return sum(int(t._version) for t in list(module.parameters()) + list(module.buffers()))


def attack_engine(pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, precision=None,
                  graph: Optional[bool] = None, fold_bn: Optional[bool] = None, deterministic: bool = False, overlap: Optional[bool] = None) -> SpaaAttack:
    """A SpaaAttack for this job.  Sweeps call spaa() many times with the same model, classifier, batch size and loss
    configuration (run_projector_based_attack, projector_based_attack.py:83-129): the engine -- its buffers and its captured
    CUDA graph -- is kept (LRU of 2) and only reset() for the new scene / targets.  A model whose parameters changed in
    between (version counters) gets a new engine."""
    net = _unwrap(pcnet)
    if not isinstance(net, PCNet) or graph is False:
        return SpaaAttack(pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, precision=precision, graph=graph,
                          fold_bn=fold_bn, deterministic=deterministic, overlap=overlap)
    if precision is not None:
        set_precision(net, precision)
    scene_shape = tuple(expand_4d(cam_scene).shape)
    # ops.weights_epoch(): FlatAdam updates parameters with a raw kernel that does not bump their version counters (train_pcnet between two
    # spaa() calls on the same model object); simplify() adds buffers that start at version 0 -- both must miss the cache
    key = (id(net), _state_version(net), ops.weights_epoch(), net.warping_net.fine_grid is None, net.shading_net.res1_s is None,
           id(getattr(classifier, "model", classifier)), len(target_idx), bool(targeted), stealth_loss, float(d_thr),
           tuple(setup_info["classifier_crop_sz"]), tuple(setup_info["prj_im_sz"]), float(setup_info["prj_brightness"]), scene_shape,
           getattr(net.shading_net, "precision", "fp32"), str(torch.device(device)), bool(torch.backends.cudnn.allow_tf32), fold_bn, bool(deterministic), overlap)
    hit = _ENGINES.get(key)
    if hit is not None and hit[0]() is net and hit[1]() is getattr(classifier, "model", classifier):
        _ENGINES.move_to_end(key)
        return hit[2].reset(cam_scene, target_idx)
    A = SpaaAttack(pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, graph=graph, fold_bn=fold_bn,
                   deterministic=deterministic, overlap=overlap)
    try:
        _ENGINES[key] = (weakref.ref(net), weakref.ref(getattr(classifier, "model", classifier)), A)
    except TypeError:               # classifier object without weak-reference support: do not cache
        return A
    evicted = False
    while len(_ENGINES) > _MAX_ENGINES:
        _ENGINES.popitem(last=False)
        evicted = True
    if evicted:
        # an engine sits in reference cycles (autograd graph of the classifier leg <-> its buffers): without a collection its ~2 GB of buffers and its
        # CUDA graph outlive the eviction by an arbitrary number of jobs (tools/capture_probe.py: +1.4 GB reserved per engine until gc ran)
        import gc
        gc.collect()
    return A


def clear_engines() -> None:
    """Drop every cached engine (and its CUDA graph); the shared graph memory pool goes with them (ops.graph_pool)."""
    _ENGINES.clear()
    import gc
    gc.collect()                                    # engines hold reference cycles (closures over themselves): their graphs must be gone before the pool
    ops.release_graph_pools()


def spaa(pcnet, classifier, imagenet_labels, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, *,
         iters: int = 50, verbose: bool = False, trace: Optional[List[dict]] = None, forced_prj: Optional[List[torch.Tensor]] = None,
         precision: Optional[str] = None, graph: Optional[bool] = None, fold_bn: Optional[bool] = None, deterministic: bool = False,
         overlap: Optional[bool] = None):
    """projector_based_attack.py:212-339.  Returns (cam_infer_best, clamp(prj_adv_best, 0, 1)).
    Keyword-only extras (reference defaults): iters=50; precision None (keep the model's), 'fp32', 'fp16' or 'bf16';
    graph None (CUDA-graph replay of the iteration when the fused PCNet path is used), True or False; fold_bn None (fold the
    frozen classifier's inference-mode BatchNorm into its convolutions when cuDNN may use TF32), True or False; deterministic False (True: the
    warp's backward as an atomics-free gather); overlap None / False (True: the stealth-loss kernels on an auxiliary stream beside the classifier; measured slower)."""
    A = attack_engine(pcnet, classifier, target_idx, targeted, cam_scene, d_thr, stealth_loss, device, setup_info, precision=precision, graph=graph,
                      fold_bn=fold_bn, deterministic=deterministic, overlap=overlap)
    for it in range(iters):
        if forced_prj is not None:
            A.prj_adv.copy_(forced_prj[it])
        prj_in = A.prj_adv.clone() if trace is not None else None
        A.step()
        if trace is not None:
            trace.append(dict(prj_in=prj_in, cam=A.cam.clone(), logits=A.logits.clone(), stats=A.stats.clone(), col_b=A.col_loss.clone(),
                              use_col=A.use_col.bool().clone(), succ=A.succ.bool().clone(), better=A.better.bool().clone(),
                              g=A.dprj.clone(), prj_out=A.prj_adv.clone(), best_col=A.best_col.clone(), best_prj=A.prj_best.clone(),
                              best_cam=A.cam_best.clone()))
        if verbose and (it % 30 == 0 or it == iters - 1):
            v = min(7 if targeted else 0, A.B - 1)                             # :240 (guarded: the reference indexes v=7 blindly)
            print(f"adv_loss_sel_logit = {A.logits[v, A.target[v]].item():<9.4f} | col_loss = {A.col_loss[v].item():.4f} | "
                  f"succ = {bool(A.succ[v].item())}")
    cam_best, prj_best = A.result()
    return cam_best.clone(), prj_best                                          # the engine's buffers are reused by the next job


def perc_al_compennet_pp(compennet_pp, classifier, imgnet_labels, target_idx, targeted, cam_scene, d_thr, device, setup_info, *,
                         iters: int = 50, trace: Optional[List[dict]] = None):
    """projector_based_attack.py:342-359: PerC-AL on the camera image, then one CompenNet++ forward."""
    device = torch.device(device)
    num_target = len(target_idx)
    cp_sz = setup_info["classifier_crop_sz"]
    cam_scene_batch = expand_4d(cam_scene).to(device).expand(num_target, -1, -1, -1)
    confidence = 0 if targeted else 40
    attacker = PerC_AL(device=device, max_iterations=iters, alpha_l_init=1, alpha_c_init=0.5, confidence=confidence)
    cam_infer_best = attacker.adversary_projector(classifier, cam_scene_batch, labels=torch.tensor(target_idx).to(device),
                                                  imagenet_labels=imgnet_labels, d_thr=d_thr, targeted=targeted, cp_sz=cp_sz, trace=trace)
    with torch.no_grad():
        prj_adv_best = compennet_pp(cam_infer_best, cam_scene_batch)
    return cam_infer_best, prj_adv_best


# ------------------------------------------------------------------------------------------------------------
# multi-GPU: attack jobs are independent -- shard them over ranks, no data-path collective
# ------------------------------------------------------------------------------------------------------------

def shard_jobs(jobs, rank: Optional[int] = None, world: Optional[int] = None, costs=None):
    """This rank's share of an attack sweep (the (setup, classifier, stealth_loss, d_thr, targets) tuples that
    run_projector_based_attack iterates serially, projector_based_attack.py:83-129).  Returns [(global job index, job), ...].
    Default: round-robin.  costs (one relative cost per job, the same list on every rank): longest-processing-time-first -- jobs in order of
    decreasing cost, each to the least loaded rank (ties: lowest rank; deterministic, so every rank computes the same partition without a
    collective) -- a vgg16 job costs ~3x a resnet18 job, and round-robin leaves the fastest rank idle for a third of the sweep."""
    import torch.distributed as dist
    if rank is None or world is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
    if costs is None:
        return [(i, jobs[i]) for i in range(rank, len(jobs), world)]
    if len(costs) != len(jobs):
        raise ValueError("shard_jobs: one cost per job")
    load = [0.0] * world
    mine = []
    for i in sorted(range(len(jobs)), key=lambda i: (-float(costs[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        load[r] += float(costs[i])
        if r == rank:
            mine.append(i)
    return [(i, jobs[i]) for i in sorted(mine)]


def gather_job_results(local_results, n_jobs: int):
    """local_results: [(global job index, picklable result), ...] from this rank's shard.  Every rank gets the list of all
    n_jobs results in job order (host-side all_gather_object: bookkeeping, not on the data path)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        parts = [local_results]
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local_results)
    out = [None] * n_jobs
    for part in parts:
        for i, r in part:
            if out[i] is not None:
                raise RuntimeError(f"attack job {i} was executed by more than one rank")
            out[i] = r
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError(f"attack jobs {missing} were not executed by any rank")
    return out


def to_attacker_cfg_str(attacker_name: str):
    """projector_based_attack.py:194-210: the folder name the reference derives from the attacker and its model's training configuration."""
    from .train_network import get_model_train_cfg
    assert attacker_name in ("SPAA", "PerC-AL+CompenNet++"), f"{attacker_name} not supported!"
    c = get_model_train_cfg(model_list=["PCNet" if attacker_name == "SPAA" else "CompenNet++"], single=True)
    model_cfg_str = f"{c.model_name}_{c.loss}_{c.num_train}_{c.batch_size}_{c.max_iters}"
    if attacker_name == "SPAA":
        return f"{attacker_name}_{model_cfg_str}", model_cfg_str
    return f"{attacker_name}_{c.loss}_{c.num_train}_{c.batch_size}_{c.max_iters}", model_cfg_str


def run_attack_sweep(jobs, device, *, attacker_name: str = "SPAA", iters: int = 50, precision: Optional[str] = None, save: bool = True,
                     rank: Optional[int] = None, world: Optional[int] = None, costs=None):
    """The job loop of run_projector_based_attack (projector_based_attack.py:83-141) for models that are already trained, sharded over ranks.

    jobs: list of dicts, one per (setup, stealth_loss, d_thr, classifier) cell of the reference's nested loops:
        model        PCNet (SPAA) or CompenNet++ (PerC-AL), frozen, on `device`
        classifier   object with the reference's Classifier convention (.model / .input_sz / __call__)
        cam_scene    [3,H,W] or [1,3,H,W]
        target_idx   list of target class indices (the targeted batch, :121)
        stealth_loss, d_thr, setup_info (classifier_crop_sz, prj_im_sz, prj_brightness)
        setup_path, classifier_name   (only for `save`)  -> <setup_path>/{cam/infer/adv,prj/adv}/<attacker cfg>/<stealth_loss>/<d_thr>/<classifier_name>/
    Every job runs the reference's two attacks -- untargeted against the scene's own top-1 label (:104-111), then the targeted batch (:121-125) -- and
    stores the n targeted results followed by the untargeted one as img_0001.png .. (:137-138).  Jobs are independent: rank r runs jobs
    r, r + world, ... with no data-path collective; attack engines (buffers + captured CUDA graph) are reused across jobs with the same model,
    classifier and shapes.  Returns [(job index, dict(true_idx, n_targets, cam_infer [n+1,3,H,W] cpu, prj_adv [n+1,3,h,w] cpu)), ...] of this rank."""
    from . import utils as ut
    assert attacker_name in ("SPAA", "PerC-AL+CompenNet++"), f"{attacker_name} not supported!"
    cfg_str = to_attacker_cfg_str(attacker_name)[0]
    out = []
    for ji, job in shard_jobs(jobs, rank, world, costs):           # costs: optional relative job costs for a balanced static partition
        scene = expand_4d(job["cam_scene"]).to(device)
        setup, clf = job["setup_info"], job["classifier"]
        cp_sz = setup["classifier_crop_sz"]
        with torch.no_grad():
            true_idx = int(device_logits(clf, scene, cp_sz).argmax(1)[0])          # :99-102 (top-1 of the un-attacked scene)
        res = []
        for targeted, tidx in ((True, list(job["target_idx"])), (False, [true_idx])):
            if attacker_name == "SPAA":
                cam, prj = spaa(job["model"], clf, None, tidx, targeted, scene, job["d_thr"], job["stealth_loss"], device, setup, iters=iters,
                                precision=precision)
            else:
                cam, prj = perc_al_compennet_pp(job["model"], clf, None, tidx, targeted, scene, job["d_thr"], device, setup, iters=iters)
            res.append((cam.detach().clone(), prj.detach().clone()))
        cam_all = torch.cat([r[0] for r in res], 0).cpu()
        prj_all = torch.cat([r[1] for r in res], 0).cpu()
        if save and job.get("setup_path"):
            folder = os.path.join(cfg_str, job["stealth_loss"], str(job["d_thr"]), job.get("classifier_name", getattr(clf, "name", "classifier")))
            ut.save_imgs(cam_all, os.path.join(job["setup_path"], "cam/infer/adv", folder))
            ut.save_imgs(prj_all, os.path.join(job["setup_path"], "prj/adv", folder))
        out.append((ji, dict(true_idx=true_idx, n_targets=len(job["target_idx"]), cam_infer=cam_all, prj_adv=prj_all)))
    return out
