"""PerC-AL attacker -- API of /root/reference/src/python/perc_al/__init__.py:15-256 on the sm_100a kernels.

Per iteration: classifier forward+backward (cuDNN, external) -> per-sample norm + masked step -> fused Lab + dE2000
gradient of ||dE map||_2 -> norm + masked step -> fused clamp / quantise / L2 projection -> classifier forward on
the quantised image -> device-side decision masks and best-so-far copy.  No host synchronisation inside the loop.
"""
from __future__ import annotations

from math import cos, pi
from typing import List, Optional

import time

import torch
import torch.nn as nn

from .. import ops
from ..classifier import device_logits, fold_batchnorm, use_channels_last
from .differential_color_functions import ciede2000_diff, deltaE, rgb2lab_diff  # noqa: F401  (re-exported like the reference)


def quantization(x):
    """:15-18."""
    return torch.round(x * 255) / 255


class PerC_AL:
    def __init__(self, max_iterations: int = 1000, alpha_l_init: float = 1., alpha_c_init: float = 0.5, confidence: float = 0,
                 device: torch.device = torch.device("cuda")) -> None:
        self.max_iterations = max_iterations
        self.alpha_l_init = alpha_l_init
        self.alpha_c_init = alpha_c_init
        self.confidence = confidence
        self.device = torch.device(device)

    # ---------------------------------------------------------------------------------------------------------
    def _run(self, logits_fn, inputs, labels, targeted, d_thr, p_thresh, projector_rules, trace, graph=True):
        if inputs.min() < 0 or inputs.max() > 1:
            raise ValueError("Input values should be in the [0, 1] range.")
        ops._need_cuda(inputs)
        if targeted and self.confidence != 0:
            print("Only support setting confidence in untargeted case!")
            return None
        dev = inputs.device
        inputs = ops._f32c(inputs)
        labels = labels.to(dev).long()
        B, _, H, W = inputs.shape
        hw = H * W
        a_l_min, a_c_min = self.alpha_l_init / 100, self.alpha_c_init / 10
        n_it = self.max_iterations
        # cosine-annealed step sizes (:184-185) as a device table: row i = (step for ~mask rows, step for mask rows)
        cosf = [1 + cos(i / n_it * pi) for i in range(n_it)]
        a_l = [a_l_min + 0.5 * (self.alpha_l_init - a_l_min) * c for c in cosf]
        a_c = [a_c_min + 0.5 * (self.alpha_c_init - a_c_min) * c for c in cosf]
        sign = -1.0 if targeted else 1.0
        # delta += a_l * g/||g|| on ~mask rows with g = grad of (sign * CE): table holds +a_l ; colour: -a_c on mask rows
        tab_l = torch.tensor([[a, 0.0] for a in a_l], device=dev)
        tab_c = torch.tensor([[0.0, -a] for a in a_c], device=dev)

        best = inputs.clone()
        ref_lab = ops.rgb2lab(inputs)
        delta = torch.zeros_like(inputs)
        xs = inputs.clone()                       # inputs + delta
        xs2 = torch.empty_like(inputs)
        xq = torch.empty_like(inputs)
        g_c = torch.empty_like(inputs)
        use_col, isadv, better = (torch.zeros(B, dtype=torch.uint8, device=dev) for _ in range(3))
        best_dis = torch.ones(B, device=dev) * 100000
        dis, l2sum, sq = (torch.empty(B, device=dev) for _ in range(3))
        stats = torch.empty(B, 4, device=dev)
        if not targeted and self.confidence != 0:
            mode = 2
        elif targeted:
            mode = 1
        else:
            mode = 0
        if not projector_rules:                   # original `adversary`: no perturbation-size / confidence gates
            d_thr, p_thresh = -1.0, -1.0
        # One iteration is a fixed launch sequence with no host decision (the step sizes of iteration i are staged into two device scalars
        # first): after two eager iterations it is recorded in a CUDA graph and replayed -- the external classifier's two forward passes
        # and one backward pass are ~100-300 small launches per iteration.
        step_l, step_c = torch.empty(2, device=dev), torch.empty(2, device=dev)
        use_graph = bool(graph) and trace is None and inputs.is_cuda
        state = {}

        def body():
            leaf = xs.detach().requires_grad_(True)
            with torch.enable_grad():
                logits = logits_fn(leaf)
                loss = sign * nn.functional.cross_entropy(logits, labels, reduction="sum")
                g_a, = torch.autograd.grad(loss, leaf)
            ops.row_sqnorm(g_a, sq)
            ops.row_normalized_step(delta, g_a, sq, step_l, use_col, base=inputs, sum_out=xs2)
            ops.color_loss(xs2, inputs, ref_lab, cam_is_lab2=True, de_weighting=True, c_de=1.0, c_l2=0.0, stats=stats, grad=g_c)
            ops.row_sqnorm(g_c, sq)
            ops.row_normalized_step(delta, g_c, sq, step_c, use_col)
            ops.percal_project(inputs, delta, xq, xs, l2sum)
            with torch.no_grad():
                logits2 = logits_fn(xq)
            ops.percal_masks(logits2, labels, mode, 40.0, l2sum, hw, d_thr, p_thresh, stats, isadv, use_col, better, dis, best_dis)
            ops.masked_copy_rows(best, xq, isadv if projector_rules else better)      # :244-245 vs :128
            state["g_a"] = g_a

        # Recording costs ~0.1 s for a large classifier, and replay only pays when the eager iteration is bound by the host's launch rate
        # (resnet18 / inception_v3: hundreds of short kernels) -- not when the device is the bottleneck (vgg16 at B=32: 9.1 ms eager vs 12.2 ms per
        # iteration with the capture amortised over 30 iterations).  The second eager iteration is timed both ways to decide.
        g, launch_bound, probe = None, None, None
        for i in range(n_it):
            step_l.copy_(tab_l[i])
            step_c.copy_(tab_c[i])
            if g is not None:
                g.replay()
            elif use_graph and i == 1 and n_it >= 7:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                t0 = time.perf_counter()
                body()
                probe = (time.perf_counter() - t0, e0, e1)
                e1.record()
            elif use_graph and i >= 2 and n_it - i >= 4 and launch_bound is not False and not torch.cuda.is_current_stream_capturing():
                if launch_bound is None and probe is not None:
                    probe[2].synchronize()
                    launch_bound = probe[0] >= 0.7 * probe[1].elapsed_time(probe[2]) * 1e-3
                    if not launch_bound:
                        body()
                        continue
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                pool = ops.graph_pool(dev)
                with torch.cuda.stream(side):
                    g.capture_begin(pool=pool) if pool is not None else g.capture_begin()
                    try:
                        body()                     # records the launches; the replay below executes this iteration
                    finally:
                        g.capture_end()
                torch.cuda.current_stream(dev).wait_stream(side)
                g.replay()
            else:
                body()
            if trace is not None:
                trace.append(dict(g_a=state["g_a"].clone(), g_c=g_c.clone(), delta=delta.clone(), dis=dis.clone(), x_round=xq.clone(),
                                  use_col=use_col.bool().clone(), isadv=isadv.bool().clone(), best=best.clone()))
        return best

    # ---------------------------------------------------------------------------------------------------------
    def adversary(self, model: nn.Module, inputs: torch.Tensor, labels: torch.Tensor, targeted: bool = True) -> torch.Tensor:
        """:53-131: the original digital PerC-AL against a bare model fed (x - 0.5) / 0.5."""
        def fn(x):
            return model((x - 0.5) / 0.5)
        return self._run(fn, inputs, labels, targeted, 0.0, 0.0, False, None)

    def adversary_projector(self, classifier, inputs: torch.Tensor, labels: torch.Tensor, imagenet_labels, d_thr, targeted: bool = True,
                            cp_sz=(240, 240), trace: Optional[List[dict]] = None, graph: bool = True) -> torch.Tensor:
        """:133-256.  The frozen classifier is called as in spaa(): channels_last and BatchNorm-folded (private copy) when cuDNN may
        use TF32 (torch's default), the stock module in the exact-fp32 parity mode."""
        clf_cl = use_channels_last(classifier) if inputs.is_cuda else False
        if torch.backends.cudnn.allow_tf32:
            classifier = fold_batchnorm(classifier)

        def fn(x):
            return device_logits(classifier, x, cp_sz, clf_cl)
        return self._run(fn, inputs, labels, targeted, float(d_thr), 0.9, True, trace, graph)
