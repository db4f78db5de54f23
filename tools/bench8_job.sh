#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 --skip-cpu-baseline --skip-side-legs --skip-cold > gpurun_out/r2_final9_8gpu.json 2> gpurun_out/r2_final9_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final9_8gpu.json').read().strip().splitlines()[-1])
print('8gpu', d['value'], d['e2e']['value'], d['train']['value'], d['train']['phases'], d['train'].get('strong',{}).get('img_per_s'), str(d.get('sweep'))[:200])
PY
