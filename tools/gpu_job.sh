#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final4_tests.log 2>&1
tail -3 gpurun_out/r2_final4_tests.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/r2_final4_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final4_smoke.log | cut -c1-300
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final4_bench.json 2> gpurun_out/r2_final4_bench.err ) 2> gpurun_out/r2_final4_bench.time; echo "bench rc=$?"
tail -2 gpurun_out/r2_final4_bench.err | cut -c1-300
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final4_ref.json 2> gpurun_out/r2_final4_ref.err ) 2> gpurun_out/r2_final4_ref.time; echo "ref rc=$?"
python tools/kbench.py --markdown > gpurun_out/r2_kbench_v2.md 2> gpurun_out/r2_kbench_v2.err; echo "kbench rc=$?"
L="--steps 2 --warmup 3 --no-graph --skip-train --skip-side-legs --skip-cpu-baseline --skip-sweep --skip-cold"
python bench.py $L > gpurun_out/r2_l2_plain.json 2> gpurun_out/r2_l2_plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py $L > gpurun_out/r2_l2_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final4_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype') if k in d}); print('e2e',d['e2e']['value']); print('roofline',d['roofline']['achieved'], d['roofline']['frac'], d['roofline'].get('frac_of_sustained_peak')); print('train',d['train']['value'],d['train'].get('phases'), d['train'].get('bf16_mode'), d['train'].get('bf16x3_mode'), d['train'].get('fp32_mode'))
print('clocks',d.get('clocks'))
for k in ('percal','bf16x3_mode','fp32_mode','stock_classifier','e2e_cold','torch_cuda_reference','sweep'):
    if k in d: print(k, str(d[k])[:260])
print('train ref', str(d['train'].get('torch_cuda_reference'))[:300])
r=json.loads(open('gpurun_out/r2_final4_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
