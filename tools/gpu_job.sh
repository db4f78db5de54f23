#!/bin/bash
mkdir -p gpurun_out
S="--steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold"
for ws in 0 1 0 1; do
SPAA_SBRANCH_STREAM=$ws python bench.py $S > gpurun_out/sb_$ws.json 2> gpurun_out/sb_$ws.err
python - $ws <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/sb_{sys.argv[1]}.json').read().strip().splitlines()[-1]); print('SBRANCH_STREAM', sys.argv[1], 'attack', d['value'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], 'parity', d['parity_check']['cam_max_abs_err'], d['parity_check']['top1_agree'])
except Exception as e:
    print('failed', e); print(open(f'gpurun_out/sb_{sys.argv[1]}.err').read()[-800:])
PY
done
SPAA_SBRANCH_STREAM=1 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/sb_t.log 2>&1; tail -3 gpurun_out/sb_t.log | cut -c1-200
