#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/c11_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c11_smoke.log
timeout 1200 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/c11_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c11_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['parity'], d['e2e']['value'] if d['e2e'] else None)
    print(json.dumps(d.get('configs'))[:3000])
    print(d.get('cpu_baseline'), d.get('cpu_baseline_optimized'))
except Exception as e: print('parse error', e)
PY
