#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tests/ring_gpu_worker.py > gpurun_out/c10_ring_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/c10_ring_parity.log
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c10_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c10_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c10_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d768 MFB_X=1 $QB --threads 768
run lap384 MFB_X=1 $QB --op lap
run lap768 MFB_X=1 $QB --op lap --threads 768
run t384_mj MFB_RING_MAXJOBS=320 $QB
