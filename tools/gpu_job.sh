#!/bin/bash
# the GPU job of the moment (kept in the tree so that the snapshot gpurun takes when a box frees up runs the CURRENT job)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -s -k "split" > gpurun_out/r2_split.log 2>&1
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_models.py -q -m gpu -s -k "fullsize or tps_full or bf16x3" > gpurun_out/r2_t4.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_b1_ref.json 2> gpurun_out/r2_b1_ref.err
grep -n "^E \|passed\|failed" gpurun_out/r2_split.log | cut -c1-250 | head -20
grep -n "^E \|passed\|failed\|fullsize\[" gpurun_out/r2_t4.log | cut -c1-400 | head -40
tail -5 gpurun_out/r2_b1.err; wc -c gpurun_out/r2_b1.json gpurun_out/r2_b1_ref.json
