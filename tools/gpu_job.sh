#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "pool or sweep or deterministic or cache" > gpurun_out/r2_pool_t2.log 2>&1
tail -12 gpurun_out/r2_pool_t2.log | cut -c1-300
