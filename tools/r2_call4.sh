#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tests/ring_gpu_worker.py > gpurun_out/c4_ring_parity.log 2>&1
echo "parity rc=$?"; grep -v "^ring.*e-1[5-9]" gpurun_out/c4_ring_parity.log | tail -12
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c4_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c4_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c4_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d256 MFB_X=1 $QB --threads 256
run t384_36 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 36 --tile-elems 580
run t256_23 MFB_RING_MAXJOBS=256 $QB --threads 256 --tile-rows 23 --tile-elems 370
run lap384 MFB_X=1 $QB --op lap
run lap256 MFB_X=1 $QB --op lap --threads 256
