#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu > gpurun_out/r2_c2_t1.log 2>&1
tail -5 gpurun_out/r2_c2_t1.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "train" > gpurun_out/r2_c2_t2.log 2>&1
tail -5 gpurun_out/r2_c2_t2.log | cut -c1-300
B="--steps 10 --warmup 3 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold"
run() { # name dir env...
  name=$1; dir=$2; shift 2
  ( cd $dir && env "$@" python bench.py $B > /root/repo/gpurun_out/ab_$name.json 2> /root/repo/gpurun_out/ab_$name.err )
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1]); print('$name', 'attack %.1f it/s'%d['value'], 'train', d.get('train',{}).get('value'), d.get('train',{}).get('phases'), d.get('train',{}).get('bf16_mode'))
except Exception as e: print('$name failed', e); print(open('gpurun_out/ab_$name.err').read()[-800:])
PY
}
run old1 _ab_old X=1
run new1 . X=1
run new2 . X=1
