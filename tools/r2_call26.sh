#!/bin/bash
# eight GPUs, lean: weak scaling with the peer-to-peer exchange (7 neighbours per subdomain over cudaIpc windows), oracle parity on every rank
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --steps 100 --warmup 5 --e2e-steps 0 --no-strong > gpurun_out/c26_bench8.json 2> gpurun_out/c26_bench8.err
echo "bench8 rc=$?"; tail -c 300 gpurun_out/c26_bench8.err | tail -2
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c26_bench8.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac'], d['config']['halo'], d['config']['setup_s']); print(d['parity'])
except Exception as e: print('parse error', e)
PY
