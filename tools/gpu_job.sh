#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r2_gpusuite.log 2>&1
tail -5 gpurun_out/r2_gpusuite.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
tail -3 gpurun_out/r2_b2.err; wc -c gpurun_out/r2_b2.json
