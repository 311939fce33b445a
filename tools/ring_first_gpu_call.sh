#!/bin/bash
# First GPU call for the RING path (run under gpurun from the repository root):
#   gpurun --timeout 1500 -- 'bash tools/ring_first_gpu_call.sh'
# 1. parity of the RING kernel against the oracle (isolated process), 2. racecheck / memcheck on a
# small case, 3. TILED vs RING timing on the EIB mesh, 4. ncu: launch list and one full capture of
# the RING kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tests/ring_gpu_worker.py > gpurun_out/ring_parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/ring_parity.log
tail -5 gpurun_out/ring_parity.log
cat > gpurun_out/ring_sanitize_case.py <<'PY'
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import minifem_b200 as mfb
mesh = mfb.Mesh.generate(9, 8, 7, seed=3)
for op in ("ela", "lap"):
    ctx = mfb.Context(mfb.Setup(mesh, op), path="ring")
    ctx.iteration(); ctx.iteration(); ctx.download(); ctx.close()
print("RING_SANITIZE_DONE")
PY
for tool in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool $tool python gpurun_out/ring_sanitize_case.py > gpurun_out/ring_sanitize_$tool.log 2>&1
    echo "$tool rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/ring_sanitize_$tool.log)"; tail -2 gpurun_out/ring_sanitize_$tool.log
done
timeout 600 python tools/quick_bench.py --paths tiled,ring --steps 20 > gpurun_out/ring_quick_bench.log 2>&1
tail -4 gpurun_out/ring_quick_bench.log
# the tile cut as TILED cuts it (runs of the Morton curve) instead of the bisection
MFB_RING_CUT=morton timeout 300 python tools/quick_bench.py --paths ring --steps 20 > gpurun_out/ring_quick_bench_morton.log 2>&1
tail -1 gpurun_out/ring_quick_bench_morton.log
for caps in "48 820" "64 1100" "24 410"; do
    set -- $caps
    timeout 300 python tools/quick_bench.py --paths ring --steps 20 --tile-rows $1 --tile-elems $2 > gpurun_out/ring_quick_bench_$1.log 2>&1
    tail -1 gpurun_out/ring_quick_bench_$1.log
done
# two CTAs of 384 threads per SM with larger tiles (fewer edges cut by tile borders)
timeout 300 python tools/quick_bench.py --paths ring --steps 20 --threads 384 --tile-rows 54 --tile-elems 960 > gpurun_out/ring_quick_bench_384.log 2>&1
tail -1 gpurun_out/ring_quick_bench_384.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ring_launches.csv \
    python bench.py --path ring --steps 3 --warmup 3 --no-cpu-baseline --no-other-paths > gpurun_out/ring_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/ring_ela_full \
    python bench.py --path ring --steps 3 --warmup 3 --no-cpu-baseline --no-other-paths > gpurun_out/ring_ncu_full.log 2>&1
timeout 900 python bench.py --path ring --no-cpu-baseline > gpurun_out/ring_bench.json 2> gpurun_out/ring_bench.err
tail -1 gpurun_out/ring_bench.json | cut -c1-400
