#!/bin/bash
# two GPUs, lean: weak scaling with two SMs left to the exchange kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 2 --steps 200 --warmup 5 --e2e-steps 0 --no-strong > gpurun_out/c28_bench2.json 2> gpurun_out/c28_bench2.err
echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c28_bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['config']['halo'], d['parity']['ok'], d['parity']['prec_err'])
except Exception as e: print('parse error', e)
PY
