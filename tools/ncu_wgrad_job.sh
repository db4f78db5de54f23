#!/bin/bash
mkdir -p gpurun_out
python tools/train_probe.py fp16 > gpurun_out/wg_plain.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc --launch-skip 36 --launch-count 18 -o gpurun_out/r2_wgrad python tools/train_probe.py fp16 > gpurun_out/wg_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/wg_ncu.log
