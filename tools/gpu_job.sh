#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q -m gpu -k "wgrad or train or backward_weight or weight" > gpurun_out/r2_wg_t1.log 2>&1
tail -4 gpurun_out/r2_wg_t1.log | cut -c1-300
for c in 1 2 3; do
SPAA_WGRAD_CTAS=$c python tools/train_probe.py fp16 > gpurun_out/train_probe_fp16_c$c.log 2>&1
echo "== CTAS $c"; grep "total device\|conv_wgrad_tc_kernel" gpurun_out/train_probe_fp16_c$c.log | cut -c1-220
done
