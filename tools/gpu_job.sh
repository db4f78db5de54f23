#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py -q -m gpu -k "pair_mode or forward_and_backward_data or per_layer or teacher_forced" > gpurun_out/r2_t10.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_t10.log | cut -c1-300 | tail -10
timeout 600 python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b6.json 2> gpurun_out/r2_b6.err; tail -2 gpurun_out/r2_b6.err
SPAA_PDL=0 SPAA_FAST_COLOR=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b6_nopdl.json 2> gpurun_out/r2_b6_nopdl.err
SPAA_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-train > gpurun_out/r2_b6_nopdl_fast.json 2> gpurun_out/r2_b6_nopdl_fast.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_b6.json','gpurun_out/r2_b6_nopdl.json','gpurun_out/r2_b6_nopdl_fast.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d.get('train',{}).get('value'), d['parity_check']['cam_max_abs_err'], d['parity_check']['top1_agree'])
    except Exception as e: print(f, 'ERR', e)
PY
