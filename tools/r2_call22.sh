#!/bin/bash
# one GPU: coalesced row write-out against the old shape, per CTA shape; parity of the RING kernel; store probe
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/microbench/store_probe > gpurun_out/c22_store_probe.txt 2>&1; tail -4 gpurun_out/c22_store_probe.txt
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c22_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c22_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
V=$PWD/mini-fem_b200/variants
run c768 MFB_X=1 $QB --threads 768
run c896 MFB_X=1 $QB --threads 896
run c1024 MFB_X=1 $QB --threads 1024
run c640 MFB_X=1 $QB --threads 640
run c384 MFB_X=1 $QB --threads 384
run rows3_768 MFB_LIBRARY=$V/libminifem_b200_rows3.so $QB --threads 768
run ldg768 MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 768
run ldg1024 MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 1024
timeout 400 python tests/ring_gpu_worker.py > gpurun_out/c22_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/c22_parity.log
