#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c6_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c6_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c6_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d256 MFB_X=1 $QB
run t384_54 MFB_X=1 $QB --threads 384 --tile-rows 54 --tile-elems 960
run lap MFB_X=1 $QB --op lap
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/r2_ringv1b_ela_full \
    python tools/quick_bench.py --paths ring --steps 4 > gpurun_out/c6_ncu_full.log 2>&1
echo "ncu rc=$?"
timeout 900 python tests/ring_gpu_worker.py > gpurun_out/c6_ring_parity.log 2>&1
echo "parity rc=$?"; tail -2 gpurun_out/c6_ring_parity.log
