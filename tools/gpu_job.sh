#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q -m gpu -k "train or warping or training or grid or tps or compennet" > gpurun_out/r2_cg_t2.log 2>&1
tail -3 gpurun_out/r2_cg_t2.log | cut -c1-300
python tools/train_probe.py fp16 > gpurun_out/train_probe_cg2.log 2>&1
grep "total device\|coarse_grid_bwd\|grid_sample_bwd_grid" gpurun_out/train_probe_cg2.log | cut -c1-160
