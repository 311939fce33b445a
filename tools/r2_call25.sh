#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/c25_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c25_smoke.log
timeout 1200 python bench.py > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c25_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['parity']['ok'], d['parity']['values_err'], d['parity']['prec_err'])
    for k,v in d['configs'].items(): print(k, {p:round(x['ms_per_step'],4) for p,x in v.get('gpu',{}).items()})
except Exception as e: print('parse error', e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 6 -c 1 -o gpurun_out/r2_ring_final3_ela_full \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --configs none --e2e-steps 0 --no-other-paths --no-parity > gpurun_out/c25_ncu_full.log 2>&1
echo "ncu ring rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --configs none --e2e-steps 2 > gpurun_out/c25_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
