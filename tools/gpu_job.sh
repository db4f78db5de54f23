#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final5_tests.log 2>&1
tail -3 gpurun_out/r2_final5_tests.log | cut -c1-300
B="--steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-cold"
python bench.py $B > gpurun_out/wg2_bench.json 2> gpurun_out/wg2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/wg2_bench.json').read().strip().splitlines()[-1])
print('attack', d['value'], 'e2e', d['e2e']['value']); t=d['train']; print('train', t['value'], t.get('phases'), 'bf16', t.get('bf16_mode'), 'x3', t.get('bf16x3_mode'), 'fp32', t.get('fp32_mode'), 'ref', str(t.get('torch_cuda_reference'))[:120])
PY
