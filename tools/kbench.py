#!/usr/bin/env python
"""Per-kernel micro-benchmark at the BASELINE configs[1] sizes (B=32, prj 256x256, cam 240x320): CUDA-event time of each
hot-path kernel launched alone (L2 flushed between launches), achieved algorithmic GB/s or TFLOP/s and the fraction of the
measured peak (MEASURED_PEAKS.json).  usage: python tools/kbench.py [--precision fp16|bf16] [--only substr] [--markdown]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

B, CAM, PRJ = 32, (240, 320), (256, 256)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--markdown", action="store_true")
    ap.add_argument("--profile", action="store_true", help="one forward + backward of the conv stack between cudaProfilerStart/Stop "
                    "(ncu --profile-from-start off), no timing")
    args = ap.parse_args()
    import synth
    from spaa_b200 import models, ops
    from spaa_b200.models import _Stack
    dev = torch.device("cuda:0")
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
        {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros((), dtype=torch.int64, device=dev)
    rows = []

    def bench(name, fn, nbytes=None, flop=None):
        if args.only and args.only not in name:
            return
        for _ in range(2):
            fn()
        ts = []
        for _ in range(args.reps):
            # flush L2 by READING a 256 MB buffer: the lines left behind are clean.  (Writing it, as the first version did, leaves 126 MB of
            # dirty lines whose write-back is then charged to the kernel under test: +19 us at HBM peak for a 100 MB kernel.)
            sink.copy_(flush.view(torch.int64).sum())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        gbs = nbytes / ms / 1e6 if nbytes else None
        tf = flop / ms / 1e9 if flop else None
        rows.append((name, ms * 1e3, gbs, gbs / pk["hbm_gbs"] if gbs else None, tf, tf / pk["bf16_tflops"] if tf else None))

    # ---- model + activations at full size ---------------------------------------------------------------
    P = synth.pcnet_params(100, CAM)
    m = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=CAM)), torch.nn.DataParallel(models.ShadingNetSPAA()))
    m.load_state_dict(P, strict=True)
    m = models.set_precision(m.to(dev).eval(), args.precision)
    sh = m.shading_net
    adt, gdt = _Stack.act_dtype(sh), _Stack.grad_dtype(sh)
    scene = synth.textured(0, "kb.scene", (1, 3, *CAM)).to(dev)
    prj = (0.5 + 0.2 * synth.randn(1, "kb.prj", (B, 3, *PRJ))).to(dev)
    with torch.no_grad():
        grid = m.warping_net.planar_grid(PRJ).detach()
        mask = m.flat_mask()
        skip = _Stack.skip1(sh, scene)
        packed = torch.empty((B, 16, *CAM), dtype=adt, device=dev, memory_format=torch.channels_last)
        HW, PHW = CAM[0] * CAM[1], PRJ[0] * PRJ[1]
        bench("grid_sample_fwd_packed", lambda: ops.grid_sample_packed(prj, grid, adt, clamp01=True, mask=mask, rough=scene, out=packed),
              nbytes=B * (3 * PHW * 4 + 16 * HW * 2))
        cam, S = _Stack.forward(sh, None, None, None, skip_acts=skip, packed=packed)
        if args.profile:
            ops.grid_sample_packed(prj, grid, adt, clamp01=True, mask=mask, rough=scene, out=packed)
            d_pk = torch.empty((B, 16, *CAM), dtype=gdt, device=dev, memory_format=torch.channels_last).zero_()
            d_pk[:, :3].normal_()
            for rep in range(3):
                if rep == 2:
                    torch.cuda.synchronize(); flush.zero_(); torch.cuda.synchronize()
                    torch.cuda.profiler.start()
                cam, S = _Stack.forward(sh, None, None, None, skip_acts=skip, packed=packed)
                _Stack.backward(sh, S, None, need_dx=True, surf_grad_channels=(3, 6), d_pre6_packed=d_pk)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            return
        sp = sh._specs
        W = lambda n: (getattr(sh, n).weight, getattr(sh, n).bias)
        esz = 2

        def conv_case(name, spec_name, x, add=None, epi=ops.EPI_RELU, out_dtype=None, cin_offset=0, real_cin=None):
            spec = sp[spec_name]
            w, b = W(spec_name)
            y = ops.conv_forward(spec, x, w, b, add=add, epi=epi, out_dtype=out_dtype, cin_offset=cin_offset)
            ho, wo = y.shape[2:]
            flop = 2.0 * B * ho * wo * spec.cout * spec.cin * spec.k * spec.k / (spec.stride ** 2 if spec.kind == "convT" else 1)
            nb = x.numel() * x.element_size() + y.numel() * y.element_size() + (add.numel() * add.element_size() if add is not None else 0)
            bench(f"fwd {name}", lambda: ops.conv_forward(spec, x, w, b, add=add, epi=epi, out=y, cin_offset=cin_offset), nbytes=nb, flop=flop)
            return y

        conv_case("conv1_s 16(6)->32 s2", "conv1_s", packed, cin_offset=3)
        conv_case("conv2_s 32->64 s2", "conv2_s", S["r1s"])
        conv_case("conv3_s 64->128", "conv3_s", S["r2s"])
        conv_case("conv4_s 128->256", "conv4_s", S["r3s"])
        conv_case("conv1 16(3)->32 s2 +add", "conv1", packed, add=S["r1s"])
        res2 = conv_case("skipConv2 1x1 32->64", "skipConv2", S["x1"], epi=0)
        conv_case("conv2 32->64 s2 +add", "conv2", S["x1"], add=S["r2s"])
        res3 = conv_case("skipConv3 64->128", "skipConv3", S["x2"], epi=0)
        conv_case("conv3 64->128 +add", "conv3", S["x2"], add=S["r3s"])
        conv_case("conv4 128->256 +add", "conv4", S["x3"], add=S["r4s"])
        conv_case("conv5 256->128 +add", "conv5", S["x4"], add=res3)
        conv_case("transConv1 128->64 k3 up2 +add", "transConv1", S["x5"], add=res2)
        conv_case("transConv2 64->32 k2 up2", "transConv2", S["x6"])
        conv_case("conv6 32->3 fp32 planar out +add", "conv6", S["x7"], add=S["res1"], epi=ops.EPI_RELU | ops.EPI_CLAMP_MAX1, out_dtype=torch.float32)

        # ---- loss / update kernels -------------------------------------------------------------------------
        ref_lab = ops.rgb2lab(scene)
        stats = torch.empty(B, 4, device=dev)
        g_col = torch.empty(B, 3, *CAM, device=dev)
        camf = cam.float().contiguous()
        bench("color_loss fwd+bwd", lambda: ops.color_loss(camf, scene, ref_lab, cam_is_lab2=False, de_weighting=False, c_de=1.0 / HW, c_l2=1.0 / HW,
                                                          stats=stats, grad=g_col), nbytes=B * 3 * HW * 4 * 2 + 2 * 3 * HW * 4)
        g_adv = torch.randn(B, 3, *CAM, device=dev)
        use_col = (torch.arange(B, device=dev) % 2).to(torch.uint8)
        d_pk = torch.empty((B, 16, *CAM), dtype=gdt, device=dev, memory_format=torch.channels_last)
        bench("select_cotangent_packed", lambda: ops.select_cotangent_packed(g_adv, g_col, use_col, camf, ops.MASK_OPEN01, d_pk),
              nbytes=B * (2 * 3 * HW * 4 + 16 * HW * 2))
        # ---- backward chain: time each bwd-data launch through the probe -----------------------------------
        names = []
        probe = ops.set_probe(lambda kind, spec: kind.startswith("bwd_data"))
        dxw, dsf, _ = _Stack.backward(sh, S, None, need_dx=True, surf_grad_channels=(3, 6), d_pre6_packed=d_pk)
        ops.set_probe(None)
        n_bwd = len(probe["events"])
        for rep in range(3):
            probe = ops.set_probe(lambda kind, spec: kind.startswith("bwd_data"))
            _Stack.backward(sh, S, None, need_dx=True, surf_grad_channels=(3, 6), d_pre6_packed=d_pk)
            ops.set_probe(None)
        torch.cuda.synchronize()
        order = ["conv6", "transConv2", "transConv1", "conv5", "conv4", "conv3", "skipConv3", "conv2", "skipConv2", "conv1", "conv4_s", "conv3_s",
                 "conv2_s", "conv1_s"]
        if not args.only or "bwd" in args.only:
            for i, (a, b) in enumerate(probe["events"]):
                nm = order[i] if i < len(order) and n_bwd == len(order) else f"#{i}"
                spec = sp.get(nm)
                flop = None
                if spec is not None:
                    ih, iw = {"conv6": CAM, "transConv2": (120, 160), "transConv1": (60, 80)}.get(nm, None) or (None, None)
                rows.append((f"bwd_data {nm} (in-chain, warm L2)", a.elapsed_time(b) * 1e3, None, None, None, None))
        dprj = torch.empty(B, 3, *PRJ, device=dev)
        bench("grid_sample_bwd_input", lambda: ops.grid_sample_bwd_input(dxw, grid, PRJ, mask=mask, dout2=dsf, rough=scene, dimg=dprj),
              nbytes=B * (2 * 3 * HW * 4 + 2 * 3 * PHW * 4))
        sq = torch.empty(B, device=dev)
        adj = ops.WarpAdjoint(grid, PRJ, mask)
        bench("grid_sample_bwd_gather(+sqnorm)", lambda: ops.grid_sample_bwd_gather(adj, dxw, dout2=dsf, rough=scene, dimg=dprj, sq=sq, x_for_clamp=prj),
              nbytes=B * (2 * 3 * HW * 4 + 2 * 3 * PHW * 4))
        bench("row_sqnorm", lambda: ops.row_sqnorm(dprj, sq, prj), nbytes=B * 2 * 3 * PHW * 4)
        step2 = torch.tensor([-2.0, -1.0], device=dev)
        best = prj.clone()
        succ = use_col.clone()
        bench("row_normalized_step(+best copy)", lambda: ops.row_normalized_step(prj, dprj, sq, step2, use_col, use_clamp_mask=True, copy_dst=best, copy_sel=succ),
              nbytes=B * 3 * PHW * 4 * 3 + (B // 2) * 3 * PHW * 4)
    hdr = ("kernel", "us", "GB/s (algorithmic)", "frac HBM peak", "TFLOP/s", "frac bf16 burst peak")
    if args.markdown:
        print(f"# per-kernel micro-benchmark, B={B}, precision {args.precision} (CUDA events, L2 flushed, median of {args.reps})\n")
        print("| " + " | ".join(hdr) + " |\n|---|---:|---:|---:|---:|---:|")
        for r in rows:
            print("| " + " | ".join([r[0], f"{r[1]:.1f}"] + [("" if v is None else (f"{v:.0f}" if i in (0, 2) else f"{v:.3f}")) for i, v in enumerate(r[2:])]) + " |")
    else:
        for r in rows:
            print(f"{r[0]:48s} {r[1]:9.1f} us " + (f"{r[2]:8.0f} GB/s ({r[3]:.3f})" if r[2] else " " * 24) + (f"  {r[4]:7.1f} TF/s ({r[5]:.3f})" if r[4] else ""))


if __name__ == "__main__":
    main()
