"""PCNet / CompenNet++ training -- API of /root/reference/src/python/train_network.py:130-473 on the sm_100a kernels.

One training step = batch gather -> PCNet forward (warp + 17 fused convs) -> ONE fused L1+MSE+SSIM loss/gradient kernel
-> backward (data + weight kernels) accumulating into a FLAT gradient buffer -> [NCCL all-reduce of that one bucket
when torch.distributed is initialised, one process per GPU] -> ONE fused Adam launch over the flat parameter buffer with
per-group learning rates read from a device-side schedule table.  No host synchronisation inside a step.
"""
from __future__ import annotations

import math
import os
import random
import time
from os.path import join
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops
from . import utils as ut


class AttrDict(dict):
    """Minimal stand-in for omegaconf.DictConfig (attribute + item access), which this image does not ship."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def copy(self):
        return AttrDict(dict.copy(self))


# ------------------------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------------------------

class _FusedLossFn(torch.autograd.Function):
    """(w_l1 * L1 + w_l2 * MSE + w_ssim * (1 - SSIM), MSE) with the gradient produced by the same launch."""

    @staticmethod
    def forward(ctx, pred, target, w_l1, w_l2, w_ssim):
        need = pred.requires_grad
        sums, grad, _ = ops.ssim_l1(pred, target, w_l1, w_l2, w_ssim, want_grad=need)
        n = pred.numel()
        ctx.grad = grad
        l2 = sums[1] / n
        loss = w_l1 * sums[0] / n + w_l2 * l2 + (w_ssim * (1 - sums[2] / n) if w_ssim else 0.0)
        ctx.mark_non_differentiable(l2)
        return loss, l2

    @staticmethod
    def backward(ctx, dloss, _dl2):
        return ctx.grad * dloss, None, None, None, None


def huber(x, y, delta=0.1):
    """train_network.py:29-36 (pseudo-Huber)."""
    return ((1 + ((x - y) / delta) ** 2).clamp(1e-4).sqrt() - 1) * delta


def compute_loss(prj_infer, prj_train, loss_option):
    """train_network.py:367-392: substring-selected terms; the MSE is always returned second."""
    if loss_option == "":
        raise TypeError("Loss type not specified")
    ops._need_cuda(prj_infer, prj_train)
    w_l1 = 1.0 if "l1" in loss_option else 0.0
    w_l2 = 1.0 if "l2" in loss_option else 0.0
    w_ssim = 1.0 if "ssim" in loss_option else 0.0
    train_loss, l2_loss = _FusedLossFn.apply(ops._f32c(prj_infer), ops._f32c(prj_train.detach()), w_l1, w_l2, w_ssim)
    if "huber" in loss_option:
        train_loss = train_loss + huber(prj_infer, prj_train).abs().mean()
    return train_loss, l2_loss


# ------------------------------------------------------------------------------------------------------------
# flat-buffer Adam with per-group schedules
# ------------------------------------------------------------------------------------------------------------

class FlatAdam:
    """All trainable parameters live in one flat fp32 buffer (the modules' .data and .grad become views of it), so
    zero_grad is one memset, the data-parallel all-reduce is one NCCL call on one bucket and the update is one kernel.
    groups: list of (params, lr, weight_decay, milestones, gamma) -- MultiStepLR / StepLR semantics evaluated per step
    into a device table so no host value is read while training."""

    def __init__(self, groups: Sequence[Tuple[List[torch.nn.Parameter], float, float, Sequence[int], float]], max_steps: int,
                 step_lr: Optional[int] = None):
        groups = [g for g in groups if len(g[0])]
        dev = groups[0][0][0].device
        n = sum(p.numel() for g in groups for p in g[0])
        self.param = torch.empty(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        off, ends = 0, []
        self.params = []
        for plist, *_ in groups:
            for p in plist:
                k = p.numel()
                self.param[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.param[off:off + k].view_as(p.data)
                p.grad = self.grad[off:off + k].view_as(p.data)
                p._spaa_flat_grad = True          # models._StackFn / _RefineFn accumulate into this view directly (zero_grad() clears it)
                self.params.append(p)
                off += k
            ends.append(off)
        self.seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
        self.seg_wd = torch.tensor([g[2] for g in groups], device=dev)
        table = []
        for s in range(max_steps + 1):
            row = []
            for _, lr, _, milestones, gamma in groups:
                if step_lr is not None:
                    row.append(lr * gamma ** (s // step_lr))
                else:
                    row.append(lr * gamma ** sum(1 for m in milestones if s >= m))
            table.append(row)
        self.lr_table_host = table
        self.beta1, self.beta2 = 0.9, 0.999
        # per-step scalars as ONE device table, row t-1 = [1 - beta1^t, sqrt(1 - beta2^t), lr of every group at scheduler step t-1]:
        # advance() copies the row of the coming step into `dyn` (device to device, no host value read), apply() launches the update
        # with pointers into `dyn` -- so apply() can be recorded in a CUDA graph and replayed with fresh scalars every step
        rows = [[1.0 - self.beta1 ** (t + 1), math.sqrt(1.0 - self.beta2 ** (t + 1))] + table[min(t, len(table) - 1)] for t in range(max_steps + 1)]
        self.dyn_table = torch.tensor(rows, device=dev, dtype=torch.float32)
        self.dyn = self.dyn_table[0].clone()
        self.t = 0

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:                      # autograd may have replaced a .grad view; re-attach
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * 4:
                raise RuntimeError("a parameter's .grad was detached from the flat gradient buffer")

    def advance(self):
        """Host side of a step: move the schedule on and stage that step's scalars on the device (outside any CUDA graph)."""
        self.t += 1
        self.dyn.copy_(self.dyn_table[min(self.t - 1, self.dyn_table.shape[0] - 1)])

    def apply(self, grad_scale: float = 1.0):
        """Device side of a step: one Adam launch over the flat buffers with the staged scalars (CUDA-graph capturable)."""
        ops.adam_step_dev(self.param, self.grad, self.m, self.v, self.seg_end, self.dyn[2:], self.seg_wd, self.dyn[:2], beta1=self.beta1,
                          beta2=self.beta2, grad_scale=grad_scale)
        # the raw kernel updates the parameters without touching their autograd version counters: refresh the packed 16-bit copies of the
        # weights the tensor-core convolutions cache (one multi-tensor launch, recorded with the step's CUDA graph), or the next forward
        # would run on the previous step's weights
        ops.repack_packed_weights()

    def step(self, grad_scale: float = 1.0):
        """One Adam update with the learning rates of scheduler step `self.t` (then advances the schedule)."""
        self.advance()
        self.apply(grad_scale)

    def lr(self, group: int = 0) -> float:
        return self.lr_table_host[min(self.t, len(self.lr_table_host) - 1)][group]


def _strip(name: str) -> str:
    return name[len("module."):] if name.startswith("module.") else name


def pcnet_param_groups(model, l2_reg: float, lr_drop_ratio: float):
    """train_network.py:248-265: (affine+theta | refine net | everything else) with their Adam / MultiStepLR settings."""
    named = [(_strip(n), p) for n, p in model.named_parameters() if p.requires_grad]
    g1 = [p for n, p in named if n in ("warping_net.affine_mat", "warping_net.theta")]
    g2 = [p for n, p in named if "warping_net.grid_refine_net" in n]
    g3 = [p for n, p in named if "warping_net" not in n]
    return [(g1, 1e-2, 0.0, [100], lr_drop_ratio), (g2, 5e-3, 0.0, [1200], lr_drop_ratio), (g3, 1e-3, l2_reg, [1800], lr_drop_ratio)]


# ------------------------------------------------------------------------------------------------------------
# data parallelism: one process per GPU, one NCCL all-reduce of the flat gradient bucket per step
# ------------------------------------------------------------------------------------------------------------

def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(idx: Sequence[int], rank: int, world: int) -> List[int]:
    """Global-batch semantics: every rank draws the SAME seeded sample (train_network.py:295) and takes a strided share."""
    return list(idx[rank::world])


def allreduce_grads(opt: FlatAdam, world: int, skip: Optional[Tuple[int, int]] = None) -> float:
    """Sum the flat gradient bucket over ranks; returns the scale that turns the sum into the global-batch mean.
    skip = (lo, hi): that slice has already been reduced (GradOverlap) -- only the rest is exchanged here."""
    if world > 1:
        if skip is None:
            dist.all_reduce(opt.grad, op=dist.ReduceOp.SUM)
        else:
            lo, hi = skip
            if lo > 0:
                dist.all_reduce(opt.grad[:lo], op=dist.ReduceOp.SUM)
            if hi < opt.grad.numel():
                dist.all_reduce(opt.grad[hi:], op=dist.ReduceOp.SUM)
        return 1.0 / world
    return 1.0


class GradOverlap:
    """Overlap of the gradient exchange with the tail of the backward pass.  The conv stack (ShadingNet / CompenNet: 99 % of the parameters)
    finishes its backward before the warp's adjoint, the refinement net and the grid generator run theirs (~0.35 ms at batch 24); its backward
    accumulates straight into the flat bucket, so the moment it returns its slice of the bucket is final: the all-reduce of that slice starts there, on
    a second stream, and the small remainder is exchanged when the backward is done.  Works inside the CUDA graph of the step (fork / join)."""

    def __init__(self, model, opt: FlatAdam, device):
        from .models import _ConvStackNet
        self.opt, self.range, self.stack = opt, None, None
        stacks = [m for m in model.modules() if isinstance(m, _ConvStackNet)]
        if len(stacks) != 1 or torch.device(device).type != "cuda" or os.environ.get("SPAA_NO_GRAD_OVERLAP"):
            return
        base, esz = opt.grad.data_ptr(), opt.grad.element_size()
        spans = []
        for p in stacks[0].parameters():
            if not getattr(p, "_spaa_flat_grad", False) or p.grad is None:
                return
            o = (p.grad.data_ptr() - base) // esz
            spans.append((o, o + p.numel()))
        lo, hi = min(a for a, _ in spans), max(b for _, b in spans)
        if sum(b - a for a, b in spans) != hi - lo:
            return                                   # the stack's gradients are not one contiguous slice of the bucket
        self.range, self.stack = (lo, hi), stacks[0]
        self.side = torch.cuda.Stream(device=device)
        self.fired = False

    def arm(self, world: int):
        """Call before backward(): installs the hook the conv stack's backward fires when its parameter gradients are complete."""
        self.fired = False
        if self.range is None or world <= 1:
            return

        def hook():
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                lo, hi = self.range
                dist.all_reduce(self.opt.grad[lo:hi], op=dist.ReduceOp.SUM)
            self.fired = True
        self.stack._grads_ready_hook = hook

    def finish(self, world: int) -> float:
        """Call after backward(): exchanges what the hook did not and joins the second stream."""
        if self.stack is not None:
            self.stack._grads_ready_hook = None
        if self.fired:
            scale = allreduce_grads(self.opt, world, skip=self.range)
            torch.cuda.current_stream().wait_stream(self.side)
            return scale
        return allreduce_grads(self.opt, world)


# ------------------------------------------------------------------------------------------------------------
# training loops
# ------------------------------------------------------------------------------------------------------------

def _resident(t: torch.Tensor, device) -> torch.Tensor:
    return t if t.device == device else t.to(device)


_GRAPH_WARMUP = 3


def _train_loop(model, inputs, targets, scene, cfg, opt: FlatAdam, loss_of_iter, valid_data, title, verbose):
    device = torch.device(cfg["device"])
    rank, world = _world()
    dp_mode = cfg.get("dp_mode", "global")            # 'global': shard one global batch; 'weak': batch_size per rank
    B = cfg["batch_size"]
    local_B = B if dp_mode == "weak" else len(shard_indices(list(range(B)), rank, world))
    scene_batch = scene.expand(local_B, -1, -1, -1)
    start = time.time()
    history = torch.zeros(cfg["max_iters"], 2, device=device)        # (loss, mse) per step, read only when printing
    valid_psnr = valid_rmse = valid_ssim = 0.0
    print_rate = cfg.get("print_rate", 1 if verbose else 0)
    iters = 0
    # One training step = a fixed sequence of launches with no host decision inside (batch gather, forward, fused loss, zero_grad,
    # backward, gradient all-reduce, Adam): after `_GRAPH_WARMUP` eager steps per loss phase (they fill the packed-weight / workspace
    # caches) the step is recorded in a CUDA graph and replayed; per step the host only stages the batch indices and the scheduler row.
    use_graph = bool(cfg.get("graph", device.type == "cuda"))
    it_static = torch.zeros(local_B, dtype=torch.int64, device=device)
    hist_slot = torch.zeros(2, device=device)
    graphs, eager_steps, keep = {}, {}, []
    overlap = GradOverlap(model, opt, device)

    def device_step(loss_option):
        x_batch, y_batch = inputs.index_select(0, it_static), targets.index_select(0, it_static)
        model.train()
        infer = model(x_batch, scene_batch)
        loss, l2 = compute_loss(infer, y_batch, loss_option)
        opt.zero_grad()
        overlap.arm(world)
        loss.backward()
        scale = overlap.finish(world)
        opt.apply(grad_scale=scale)
        hist_slot[0], hist_slot[1] = loss.detach(), l2

    on_step = cfg.get("on_step")                                     # optional host callback(iteration) before each step (progress / timing hooks)
    while iters < cfg["max_iters"]:
        if on_step is not None:
            on_step(iters)
        idx = random.sample(range(cfg["num_train"]), B)             # :295 (python RNG, same draw on every rank)
        if dp_mode != "weak":
            idx = shard_indices(idx, rank, world)
        it_static.copy_(torch.as_tensor(idx), non_blocking=True)
        cfg["loss"] = loss_of_iter(iters + cfg.get("iter_offset", 0))      # iter_offset: resume / benchmark a later phase of the schedule
        opt.advance()
        g = graphs.get(cfg["loss"])
        if g is not None:
            g.replay()
            ops.bump_weights_epoch()                                # (the replay re-packed the 16-bit weight copies; attack engines built before it are stale)
        elif use_graph and eager_steps.get(cfg["loss"], 0) >= _GRAPH_WARMUP and "huber" not in cfg["loss"]:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    device_step(cfg["loss"])                        # records the launches; the replay below executes this step
                finally:
                    g.capture_end()
            torch.cuda.current_stream(device).wait_stream(side)
            keep.append(side)
            graphs[cfg["loss"]] = g
            g.replay()
            ops.bump_weights_epoch()
        else:
            device_step(cfg["loss"])
            eager_steps[cfg["loss"]] = eager_steps.get(cfg["loss"], 0) + 1
        history[iters].copy_(hist_slot)
        if valid_data is not None and (iters % cfg["valid_rate"] == 0 or iters == cfg["max_iters"] - 1):
            valid_psnr, valid_rmse, valid_ssim, _ = evaluate_model(model, valid_data)
        if print_rate and (iters % print_rate == 0 or iters == cfg["max_iters"] - 1) and rank == 0:
            lv, mv = history[iters].tolist()
            lapse = time.strftime("%H:%M:%S", time.gmtime(time.time() - start))
            print(f"Iter:{iters:5d} | Time: {lapse} | Train Loss: {lv:.4f} | Train RMSE: {math.sqrt(mv * 3):.4f} "
                  f"| Valid PSNR: {f'{valid_psnr:>2.4f}' if valid_psnr else '':7s}  | Valid RMSE: {f'{valid_rmse:.4f}' if valid_rmse else '':6s}  "
                  f"| Valid SSIM: {f'{valid_ssim:.4f}' if valid_ssim else '':6s}  | Learn Rate: {opt.lr(0):.5f} |")
        iters += 1
    cfg["loss_history"] = history
    if cfg.get("data_root") and rank == 0 and cfg.get("save_checkpoint", True):
        ut.save_checkpoint(join(cfg["data_root"], "../checkpoint"), model, title)
    return model, valid_psnr, valid_rmse, valid_ssim


def train_pcnet(model, train_data, valid_data, cfg, verbose: bool = True):
    """train_network.py:235-363.  `model` may be the bare PCNet or wrapped (DataParallel with one device / DDP-style
    wrappers are unwrapped for parameter naming only)."""
    device = torch.device(cfg["device"])
    scene = _resident(train_data["cam_scene"], device)
    cam_train = _resident(train_data["cam_train"], device)         # kept resident in HBM (855 MB for 500 pairs)
    prj_train = _resident(train_data["prj_train"], device)
    if "model_name" not in cfg:
        cfg["model_name"] = model.name if hasattr(model, "name") else model.module.name
    opt = FlatAdam(pcnet_param_groups(model, cfg["l2_reg"], cfg["lr_drop_ratio"]), cfg["max_iters"])
    cfg["loss"] = "l1+ssim"
    title = ut.opt_to_string(cfg) if "setup_name" in cfg else "pcnet"
    loss_of_iter = lambda it: "l1" if it <= 400 else "l1+ssim"     # :300-303
    return _train_loop(model, prj_train, cam_train, scene, cfg, opt, loss_of_iter, valid_data, title, verbose)


def train_compennet_pp(model, train_data, valid_data, cfg, verbose: bool = True):
    """train_network.py:130-232: one Adam (lr, l2_reg), StepLR(lr_drop_rate, lr_drop_ratio), fixed cfg.loss,
    input = camera image, target = projector image."""
    device = torch.device(cfg["device"])
    scene = _resident(train_data["cam_scene"], device)
    cam_train = _resident(train_data["cam_train"], device)
    prj_train = _resident(train_data["prj_train"], device)
    if "model_name" not in cfg:
        cfg["model_name"] = model.name if hasattr(model, "name") else model.module.name
    params = [p for p in model.parameters() if p.requires_grad]
    opt = FlatAdam([(params, cfg["lr"], cfg["l2_reg"], [], cfg["lr_drop_ratio"])], cfg["max_iters"], step_lr=cfg["lr_drop_rate"])
    fixed = cfg["loss"]
    title = ut.opt_to_string(cfg) if "setup_name" in cfg else "compennet_pp"
    return _train_loop(model, cam_train, prj_train, scene, cfg, opt, lambda it: fixed, valid_data, title, verbose)


def evaluate_model(model, valid_data, chunk_sz=10):
    """train_network.py:395-441."""
    cam_scene, cam_valid, prj_valid = valid_data["cam_scene"], valid_data["cam_valid"], valid_data["prj_valid"]
    device = next(model.parameters()).device
    name = model.name if hasattr(model, "name") else model.module.name
    with torch.no_grad():
        model.eval()
        valid_psnr = valid_rmse = valid_ssim = 0.0
        model_infer = torch.zeros(cam_valid.shape if "PCNet" in name else prj_valid.shape)
        num_valid = cam_valid.shape[0]
        for idx in torch.chunk(torch.arange(num_valid), chunk_sz):
            bs = len(idx)
            scene_b = cam_scene[idx].to(device) if cam_scene.shape[0] == num_valid else cam_scene.to(device).expand(bs, -1, -1, -1)
            cam_b, prj_b = cam_valid[idx].to(device), prj_valid[idx].to(device)
            inp, gt = (prj_b, cam_b) if "PCNet" in name else (cam_b, prj_b)
            out = model(inp, scene_b)
            model_infer[idx] = out.detach().cpu()
            m = ut.calc_img_dists(out, gt)
            valid_psnr += m[0] * bs / num_valid
            valid_rmse += m[1] * bs / num_valid
            valid_ssim += m[2] * bs / num_valid
    return valid_psnr, valid_rmse, valid_ssim, model_infer


def get_model_train_cfg(model_list, data_root=None, setup_list=None, device_ids=[0], center_crop=False, load_pretrained=False, plot_on=True,
                        single=False):
    """train_network.py:444-473."""
    cfg = AttrDict()
    cfg.data_root, cfg.setup_list, cfg.device, cfg.device_ids = data_root, setup_list, "cuda", device_ids
    cfg.load_pretrained, cfg.max_iters, cfg.batch_size, cfg.lr = load_pretrained, 2000, 24, 1e-3
    cfg.lr_drop_ratio, cfg.lr_drop_rate, cfg.l2_reg = 0.2, 800, 1e-4
    cfg.train_plot_rate, cfg.valid_rate, cfg.plot_on, cfg.center_crop = 50, 200, plot_on, center_crop
    if single:
        cfg.model_name, cfg.num_train, cfg.loss = model_list[0], 500, "l1+ssim"
    else:
        cfg.model_list, cfg.num_train_list, cfg.loss_list = model_list, [500], ["l1+ssim"]
    return cfg
