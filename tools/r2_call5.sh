#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c5_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c5_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c5_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d256 MFB_X=1 $QB --threads 256
run lap384 MFB_X=1 $QB --op lap
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/r2_ringv2_ela_full \
    python tools/quick_bench.py --paths ring --steps 4 > gpurun_out/c5_ncu_full.log 2>&1
echo "ncu rc=$?"
