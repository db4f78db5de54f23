#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_gpusuite.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_gpusuite.log | cut -c1-300 | tail -12
