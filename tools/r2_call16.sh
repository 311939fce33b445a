#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c16_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c16_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d768 MFB_X=1 $QB --threads 768
run lap384 MFB_X=1 $QB --op lap
timeout 400 python tests/ring_gpu_worker.py > gpurun_out/c16_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/c16_parity.log
