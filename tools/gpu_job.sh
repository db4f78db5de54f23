#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; tail -2 gpurun_out/r2_b4.err
SPAA_TC_PAIR=0 python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b4_nopair.json 2> gpurun_out/r2_b4_nopair.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_b4.json','gpurun_out/r2_b4_nopair.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['train']['value'], d['parity_check']['cam_max_abs_err'])
PY
