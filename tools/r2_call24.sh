#!/bin/bash
# two GPUs: the peer-to-peer exchange between rank processes (cudaIpc windows), the driver with two ranks, bench N=2
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/c24_multirank.log 2>&1; echo "multirank pytest rc=$?"; tail -3 gpurun_out/c24_multirank.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c24_bench2.json 2> gpurun_out/c24_bench2.err
echo "bench2 rc=$?"; tail -c 400 gpurun_out/c24_bench2.err | tail -3
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c24_bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac'], d['config']['halo']); print(d['parity']['ok'], d['parity']['prec_err']); s=d.get('strong'); print({k:s[k] for k in ('ms_per_step','one_gpu_ms_per_step','parallel_efficiency','halo')}, s['parity']['ok'])
except Exception as e: print('parse error', e)
PY
