#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -x -q -m gpu -k "train or warping or training" > gpurun_out/r2_rf_t1.log 2>&1
tail -5 gpurun_out/r2_rf_t1.log | cut -c1-300
python tools/train_probe.py fp16 > gpurun_out/train_probe_rf.log 2>&1
sed -n 1,4p gpurun_out/train_probe_rf.log | cut -c1-200; grep "conv_bwd_weight_kernel\|conv_gather\|scatter\|channel_sum" gpurun_out/train_probe_rf.log | cut -c1-160
