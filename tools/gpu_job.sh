#!/bin/bash
# Full validation on one B200 (run through tools/grun.sh -- 'bash tools/gpu_job.sh'): GPU test-suite, smoke, bench (own arm + reference arm).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/validate_tests.log 2>&1
tail -3 gpurun_out/validate_tests.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/validate_smoke.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/validate_ref.json 2> gpurun_out/validate_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/validate_bench.json
