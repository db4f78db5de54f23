#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -q -m gpu -k "grid_sample or colour_loss_ssim_warp" > gpurun_out/r2_t12.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_t12.log | cut -c1-300 | tail -6
python tools/kbench.py --markdown --only grid_sample > gpurun_out/r2_kbench_gather.md 2>&1; tail -6 gpurun_out/r2_kbench_gather.md
