#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/r2_cs_t1.log 2>&1
tail -4 gpurun_out/r2_cs_t1.log | cut -c1-300
python tools/train_probe.py fp16 > gpurun_out/train_probe_fp16_cs.log 2>&1
grep "total device\|channel_sum\|conv_wgrad_tc_kernel" gpurun_out/train_probe_fp16_cs.log | cut -c1-200
B="--steps 10 --warmup 3 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold"
python bench.py $B > gpurun_out/ab_cs.json 2> gpurun_out/ab_cs.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ab_cs.json').read().strip().splitlines()[-1]); print('attack %.1f it/s'%d['value'], 'train', d.get('train',{}).get('value'), d.get('train',{}).get('phases'))
PY
