#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu -s > gpurun_out/r2_multi.log 2>&1
grep -n "^FAILED\|^E  \|passed\|failed\|skipped" gpurun_out/r2_multi.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b_2gpu.json 2> gpurun_out/r2_b_2gpu.err
tail -4 gpurun_out/r2_b_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['e2e']['value'], d['train']['value'], d['train']['strong'], d['sweep'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_b_2gpu_ref.json 2> gpurun_out/r2_b_2gpu_ref.err
wc -l gpurun_out/r2_b_2gpu_ref.json; tail -2 gpurun_out/r2_b_2gpu_ref.err | cut -c1-200
