#!/bin/bash
# one GPU: P2P exchange between contexts of one process, RING at 1024 threads, timing of the kernel variants
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/c20_p2p.log 2>&1; echo "p2p pytest rc=$?"; tail -4 gpurun_out/c20_p2p.log
timeout 400 python tests/ring_gpu_worker.py > gpurun_out/c20_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/c20_parity.log; grep "1024" gpurun_out/c20_parity.log
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c20_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c20_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
V=$PWD/mini-fem_b200/variants
run base768 MFB_X=1 $QB --threads 768
run base1024 MFB_X=1 $QB --threads 1024
run ldg768 MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 768
run ldg1024 MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 1024
run j20o12 MFB_LIBRARY=$V/libminifem_b200_j20o12.so $QB --threads 1024
run j16r80 MFB_LIBRARY=$V/libminifem_b200_j16r80.so $QB --threads 1024
run ldg_j20o12 MFB_LIBRARY=$V/libminifem_b200_ldg_j20o12.so $QB --threads 1024
run lap768 MFB_X=1 $QB --op lap --threads 768
run lap1024 MFB_X=1 $QB --op lap --threads 1024
run lapldg1024 MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --op lap --threads 1024
