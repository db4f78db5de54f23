#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s -k "percal_fullsize or training_step_fullsize" > gpurun_out/r2_t6.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed\|fullsize training" gpurun_out/r2_t6.log | cut -c1-300 | tail -30
