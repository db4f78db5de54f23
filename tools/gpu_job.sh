#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py -q -m gpu -s -k "train_pcnet" > gpurun_out/r2_t13.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed\|^losses" gpurun_out/r2_t13.log | cut -c1-300 | tail -12
for tp in fp16 bf16; do
python bench.py --steps 10 --warmup 3 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold --train-precision $tp > gpurun_out/r2_b9_$tp.json 2> gpurun_out/r2_b9.err; tail -1 gpurun_out/r2_b9.err
done
python - <<'PY'
import json
for tp in ('fp16','bf16'):
    d=json.loads(open(f'gpurun_out/r2_b9_{tp}.json').read().strip().splitlines()[-1]); print(tp, d['train']['value'], d['train']['phases'])
PY
