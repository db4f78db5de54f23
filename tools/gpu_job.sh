#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --skip-sweep --skip-cpu-baseline --precision bf16x3 --train-precision bf16x3 --skip-side-legs > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; tail -2 gpurun_out/r2_b7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b7.json').read().strip().splitlines()[-1])
print(d['value'], d['parity_check']['cam_max_abs_err'], d['train']['value'], d['train']['phases'])
PY
