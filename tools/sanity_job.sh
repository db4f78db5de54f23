#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/sanity_tests.log 2>&1; tail -2 gpurun_out/sanity_tests.log | cut -c1-200
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('attack', d['value'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], 'roofline', d['roofline']['frac'])"
