#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_final9_multi.log 2>&1
tail -2 gpurun_out/r2_final9_multi.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --skip-sweep --skip-cpu-baseline --skip-side-legs --skip-cold > gpurun_out/r2_final9_1gpu.json 2> gpurun_out/r2_final9_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_final9_2gpu.json 2> gpurun_out/r2_final9_2gpu.err; echo "bench2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_final9_2gpu_ref.json 2> gpurun_out/r2_final9_2gpu_ref.err; echo "ref2 rc=$?"; tail -c 300 gpurun_out/r2_final9_2gpu_ref.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2_final9_1gpu.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r2_final9_2gpu.json').read().strip().splitlines()[-1])
print('1gpu', a['value'], a['train']['value'])
print('2gpu', d['value'], d['e2e']['value'], d['train']['value'], d['train']['phases'], d['train'].get('strong',{}).get('img_per_s'))
print('eff attack', d['value']/2/a['value'], 'train', d['train']['value']/2/a['train']['value'])
print('sweep', str(d.get('sweep'))[:160])
PY
