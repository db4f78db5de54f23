#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -s -k "pair_mode or split_precision_conv or forward_and_backward_data" > gpurun_out/r2_t7.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_t7.log | cut -c1-300 | tail -20
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s -k "per_layer or teacher_forced" > gpurun_out/r2_t8.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed\|fullsize" gpurun_out/r2_t8.log | cut -c1-300 | tail -30
python tools/kbench.py --markdown --only conv > gpurun_out/r2_kbench_pair.md 2>&1; grep "conv3\|conv4\|conv5\|skipConv3\|transConv1" gpurun_out/r2_kbench_pair.md
SPAA_TC_PAIR=0 python tools/kbench.py --markdown --only conv > gpurun_out/r2_kbench_nopair.md 2>&1; grep "conv3\|conv4\|conv5\|skipConv3\|transConv1" gpurun_out/r2_kbench_nopair.md
