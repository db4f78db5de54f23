#!/usr/bin/env python
"""Per-launch table from `ncu -i rep.ncu-rep --page raw --csv` of the conv stack capture (tools/kbench.py --profile): duration, DRAM traffic, HBM GB/s,
tensor-pipe and issue utilisation.  usage: tools/summarize_ncu_full.py raw.csv [layer names file] > profiles/<name>.md"""
import csv
import sys

LAYERS_F = ["conv1_s f", "conv2_s f", "conv3_s f", "conv4_s f", "conv1 f", "skipConv2 f", "conv2 f", "skipConv3 f", "conv3 f", "conv4 f", "conv5 f",
            "transConv1 f", "transConv2 f", "conv6 f"]
LAYERS_B = ["conv6 b", "transConv2 b", "transConv1 b", "conv5 b", "conv4 b", "conv3 b", "skipConv3 b", "conv2 b", "skipConv2 b", "conv1 b", "conv4_s b",
            "conv3_s b", "conv2_s b", "conv1_s b"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(name):
        for h in hdr:
            if h.endswith(name):
                return ix[h]
        raise KeyError(name)
    c_dur, c_rd, c_wr = col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    c_tensor = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if any(h.endswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for h in hdr) else None
    c_issue = col("sm__inst_issued.avg.pct_of_peak_sustained_active") if any(h.endswith("sm__inst_issued.avg.pct_of_peak_sustained_active") for h in hdr) else None
    c_inst = col("smsp__inst_executed.sum") if any(h.endswith("smsp__inst_executed.sum") for h in hdr) else None

    def num(r, c, unit_scale=True):
        v = float(r[c].replace(",", ""))
        u = units[c]
        if not unit_scale:
            return v
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    names = LAYERS_F + LAYERS_B
    print("| layer | kernel | CTAs x threads | us | DRAM read MB | DRAM write MB | HBM GB/s | tensor pipe active % | issue active % | warp instructions (M) |")
    print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|")
    tot_us = tot_mb = 0.0
    for i, r in enumerate(body):
        us, rd, wr = num(r, c_dur), num(r, c_rd), num(r, c_wr)
        k = r[ix["Kernel Name"]]
        k = k[k.find("conv_halo_kernel"):k.find("(")] if "conv_halo_kernel" in k else k[:40]
        tp = f"{float(r[c_tensor]):.1f}" if c_tensor is not None else ""
        ia = f"{float(r[c_issue]):.1f}" if c_issue is not None else ""
        wi = f"{float(r[c_inst].replace(',', '')) / 1e6:.1f}" if c_inst is not None else ""
        nm = names[i] if i < len(names) else f"launch {i}"
        print(f"| {nm} | `{k}` | {r[ix['Grid Size']].strip('()').split(',')[0]} x {r[ix['Block Size']].strip('()').split(',')[0]} | {us:.1f} | {rd:.1f} | {wr:.1f} | "
              f"{(rd + wr) / us * 1e3:.0f} | {tp} | {ia} | {wi} |")
        tot_us += us
        tot_mb += rd + wr
    print(f"\nSum over the {len(body)} launches: {tot_us:.0f} us, {tot_mb:.0f} MB of DRAM traffic = {tot_mb / tot_us * 1e3:.0f} GB/s on average.")


if __name__ == "__main__":
    main()
