#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu > gpurun_out/r2_c3_t1.log 2>&1
tail -3 gpurun_out/r2_c3_t1.log | cut -c1-300
for i in 1 2; do
for v in _ab_old .; do
( cd $v && python tools/kbench.py --reps 12 > /root/repo/gpurun_out/bis_$(basename $v | tr -d ._)_$i.log 2>&1 )
done
done
for i in 1 2; do
paste <(grep -i "conv3\|transConv1\|conv4\|conv5" gpurun_out/bis_abold_$i.log | cut -c1-62) <(grep -i "conv3\|transConv1\|conv4\|conv5" gpurun_out/bis__$i.log | cut -c48-62)
done
B="--steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold --skip-train"
python bench.py $B > gpurun_out/ab2_new.json 2> gpurun_out/ab2_new.err
( cd _ab_old && python bench.py $B > /root/repo/gpurun_out/ab2_old.json 2> /root/repo/gpurun_out/ab2_old.err )
python - <<'PY'
import json
for n in ('old','new'):
    d=json.loads(open(f'gpurun_out/ab2_{n}.json').read().strip().splitlines()[-1]); print(n, d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
PY
