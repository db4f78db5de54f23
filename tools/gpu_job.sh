#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-train > gpurun_out/r2_b8_a$i.json 2> gpurun_out/r2_b8.err
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-train --deterministic > gpurun_out/r2_b8_d$i.json 2>> gpurun_out/r2_b8.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b8_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['e2e_cold']['value'])
PY
