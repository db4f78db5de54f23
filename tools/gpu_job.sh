#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "scratch or backward_weight" > gpurun_out/r2_sf_t1.log 2>&1
tail -8 gpurun_out/r2_sf_t1.log | cut -c1-300
