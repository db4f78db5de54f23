#!/bin/bash
# Full validation on one B200: GPU test-suite, smoke, bench (own arm + reference arm), training-step profile.
mkdir -p gpurun_out
S=""
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final9_tests.log 2>&1
tail -3 gpurun_out/r2_final9_tests.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/r2_final9_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final9_smoke.log | cut -c1-300
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final9_bench.json 2> gpurun_out/r2_final9_bench.err ) 2> gpurun_out/r2_final9_bench.time; echo "bench rc=$?"
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final9_ref.json 2> gpurun_out/r2_final9_ref.err ) 2> gpurun_out/r2_final9_ref.time; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final9_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype') if k in d}); print('e2e',d['e2e']['value']); print('roofline',d['roofline']['achieved'], d['roofline']['frac'], d['roofline'].get('frac_of_sustained_peak')); print('train',d['train']['value'],d['train'].get('phases'), d['train'].get('bf16_mode'), d['train'].get('bf16x3_mode'), d['train'].get('fp32_mode'))
print('clocks',d.get('clocks'))
for k in ('bf16x3_mode','fp32_mode','stock_classifier','e2e_cold','torch_cuda_reference','sweep'):
    if k in d: print(k, str(d[k])[:160])
print('percal', d['percal']['vgg16']['value'], d['percal']['vgg16']['torch_cuda_reference']['value'], d['percal']['inception_v3']['value'], d['percal']['inception_v3']['torch_cuda_reference']['value'])
print('train ref', str(d['train'].get('torch_cuda_reference'))[:100], 'exact', d['torch_cuda_reference'].get('exact_fp32',{}).get('value'), 'parity', d['parity_check']['cam_max_abs_err'], d['parity_check']['top1_agree'])
r=json.loads(open('gpurun_out/r2_final9_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
