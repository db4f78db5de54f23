#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu > gpurun_out/r2_lean_t6.log 2>&1
tail -3 gpurun_out/r2_lean_t6.log | cut -c1-300
B="--steps 10 --warmup 3 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold"
run() { # name dir env...
  name=$1; dir=$2; shift 2
  ( cd $dir && env "$@" python bench.py $B > /root/repo/gpurun_out/ab_$name.json 2> /root/repo/gpurun_out/ab_$name.err )
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1]); print('$name', 'attack %.1f it/s'%d['value'], 'train', d.get('train',{}).get('value'), d.get('parity_check',{}).get('cam_max_abs_err'), d.get('parity_check',{}).get('top1_agree'))
except Exception as e: print('$name failed', e); print(open('gpurun_out/ab_$name.err').read()[-800:])
PY
}
run old1 _ab_old X=1
run new1 . X=1
run new2 . X=1
python tools/kbench.py --reps 12 > gpurun_out/ab_kb_new.log 2>&1
grep -i conv gpurun_out/ab_kb_new.log | cut -c1-62
