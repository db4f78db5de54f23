#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_fullsize.py -q -m gpu -k "pair_mode or forward_and_backward_data or per_layer or teacher_forced" > gpurun_out/r2_t9.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed" gpurun_out/r2_t9.log | cut -c1-300 | tail -10
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; tail -2 gpurun_out/r2_b5.err
SPAA_TC_NSPLIT=0 python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs > gpurun_out/r2_b5_nosplit.json 2> gpurun_out/r2_b5_nosplit.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_b5.json','gpurun_out/r2_b5_nosplit.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['train']['value'], d['parity_check']['cam_max_abs_err'])
PY
