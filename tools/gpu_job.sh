#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_final4_multi.log 2>&1
tail -3 gpurun_out/r2_final4_multi.log | cut -c1-300
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final4_2gpu.json 2> gpurun_out/r2_final4_2gpu.err ) 2> gpurun_out/r2_final4_2gpu.time; echo "bench2 rc=$?"
tail -2 gpurun_out/r2_final4_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final4_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling') if k in d}); print('e2e',d['e2e']['value']); print('train',d['train']['value'], d['train'].get('strong')); print('sweep', str(d.get('sweep'))[:300])
PY
