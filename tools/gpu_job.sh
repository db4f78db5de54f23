#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s > gpurun_out/r2_gpusuite.log 2>&1
grep -n "^FAILED\|passed\|failed\|fullsize\[" gpurun_out/r2_gpusuite.log | cut -c1-330
