#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/c15_bench8.json 2> gpurun_out/c15_bench8.err
echo "bench8 rc=$?"; tail -c 600 gpurun_out/c15_bench8.err | tail -5
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c15_bench8.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac']); print(d['parity']); print(d.get('strong')); print(d['e2e'])
except Exception as e: print('parse error', e)
PY
