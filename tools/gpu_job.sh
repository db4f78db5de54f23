#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpusuite.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_gpusuite.log | cut -c1-300 | tail -6
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final.json 2> gpurun_out/r2_final.err; tail -2 gpurun_out/r2_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final.json').read().strip().splitlines()[-1])
print('attack', d['value'], 'e2e', d['e2e']['value'], 'cold', d['e2e_cold']['value'], 'roofline', d['roofline']['frac'], d['roofline']['frac_of_sustained_peak'])
print('parity', d['parity_check']['cam_max_abs_err'], d['parity_check']['top1_agree'])
print('stock', d['stock_classifier']['value'], 'fp32', d['fp32_mode']['value'], 'x3', d['bf16x3_mode']['value'], d['bf16x3_mode']['parity_check']['cam_max_abs_err'])
print('ref cuda', d['torch_cuda_reference']['value'], d['torch_cuda_reference']['exact_fp32']['value'])
t=d['train']; print('train', t['value'], t['phases']['l1']['img_per_s'], 'fp32', t['fp32_mode'], 'x3', t['bf16x3_mode'], 'ref', t['torch_cuda_reference']['img_per_s'])
print('percal', {k:(v['value'], v['torch_cuda_reference']['value']) for k,v in d['percal'].items()})
print('sweep', d['sweep']['seconds'], d['sweep']['jobs'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
r=json.loads(open('gpurun_out/r2_final_ref.json').read().strip().splitlines()[-1]); print('ref arm', r['value'], r['steps'], r['warmup'])
PY
