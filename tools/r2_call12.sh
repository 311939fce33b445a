#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/c12_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c12_bench2.json 2> gpurun_out/c12_bench2.err
echo "bench2 rc=$?"; tail -c 800 gpurun_out/c12_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c12_bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['roofline']['frac']); print(d['parity']); print(d.get('strong')); print(d['e2e'])
except Exception as e: print('parse error', e)
PY
for r in 0 4 8 16; do
  MFB_HALO_RESERVE_CTAS=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 200 --warmup 10 --e2e-steps 0 --no-parity --no-strong > gpurun_out/c12_res$r.json 2>/dev/null
  python -c "import json;d=json.loads(open('gpurun_out/c12_res$r.json').read().strip().splitlines()[-1]);print('reserve $r', d['ms_per_step'])"
done
