#!/bin/bash
mkdir -p gpurun_out
S="--steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold"
for ws in 0 1 0 1; do
SPAA_WGRAD_STREAM=$ws python bench.py $S > gpurun_out/ws_$ws.json 2> gpurun_out/ws_$ws.err
python - $ws <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/ws_{sys.argv[1]}.json').read().strip().splitlines()[-1]); print('WGRAD_STREAM', sys.argv[1], 'train', d['train']['value'], d['train']['phases'])
except Exception as e:
    print('failed', e); print(open(f'gpurun_out/ws_{sys.argv[1]}.err').read()[-800:])
PY
done
SPAA_WGRAD_STREAM=1 timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -x -q -m gpu -k "train or training" > gpurun_out/ws_t.log 2>&1; tail -3 gpurun_out/ws_t.log | cut -c1-200
