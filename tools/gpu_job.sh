#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_models.py -q -m gpu -s -k "fullsize_per_layer or bf16x3" > gpurun_out/r2_t5.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed\|fullsize\[" gpurun_out/r2_t5.log | cut -c1-330
python bench.py --steps 20 --warmup 5 --skip-sweep > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; tail -3 gpurun_out/r2_b3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b3.json').read().strip().splitlines()[-1])
print(d['value'], d['train'].get('bf16x3_mode'), d['train'].get('fp32_mode'))
PY
