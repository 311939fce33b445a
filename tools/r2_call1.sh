#!/bin/bash
# Round-2 first GPU call: ncu of the RING kernel + knob timings + parity.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/r2_ring_ela_full \
    python tools/quick_bench.py --paths ring --steps 4 > gpurun_out/c1_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 600 python tools/quick_bench.py --paths tiled,ring --steps 30 > gpurun_out/c1_qb_default.log 2>&1
tail -3 gpurun_out/c1_qb_default.log
MFB_RING_CUT=morton timeout 300 python tools/quick_bench.py --paths ring --steps 30 > gpurun_out/c1_qb_morton.log 2>&1
tail -1 gpurun_out/c1_qb_morton.log
for caps in "48 820" "64 1100" "24 410"; do
    set -- $caps
    timeout 300 python tools/quick_bench.py --paths ring --steps 30 --tile-rows $1 --tile-elems $2 > gpurun_out/c1_qb_$1.log 2>&1
    tail -1 gpurun_out/c1_qb_$1.log
done
timeout 300 python tools/quick_bench.py --paths ring --steps 30 --threads 384 --tile-rows 54 --tile-elems 960 > gpurun_out/c1_qb_384.log 2>&1
tail -1 gpurun_out/c1_qb_384.log
timeout 300 python tools/quick_bench.py --paths ring --steps 30 --op lap > gpurun_out/c1_qb_lap.log 2>&1
tail -1 gpurun_out/c1_qb_lap.log
timeout 600 python tests/ring_gpu_worker.py > gpurun_out/c1_ring_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/c1_ring_parity.log
