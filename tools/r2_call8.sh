#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tests/ring_gpu_worker.py > gpurun_out/c8_ring_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/c8_ring_parity.log
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c8_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c8_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c8_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d768 MFB_X=1 $QB --threads 768
run t768_48 MFB_X=1 $QB --threads 768 --tile-rows 48 --tile-elems 820
run t384_28 MFB_X=1 $QB --tile-rows 28 --tile-elems 480
run lap384 MFB_X=1 $QB --op lap
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/r2_ringws_ela_full \
    python tools/quick_bench.py --paths ring --steps 4 > gpurun_out/c8_ncu_full.log 2>&1
echo "ncu rc=$?"
