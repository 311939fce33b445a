#!/bin/bash
# one GPU: store-bandwidth probe, P2P with 8 subdomains in one process, CTA shapes 640 / 896, poll sleep
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/microbench/store_probe > gpurun_out/c21_store_probe.txt 2>&1; cat gpurun_out/c21_store_probe.txt
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "p2p" > gpurun_out/c21_p2p.log 2>&1; echo "p2p pytest rc=$?"; tail -4 gpurun_out/c21_p2p.log
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c21_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c21_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
V=$PWD/mini-fem_b200/variants
run t896 MFB_X=1 $QB --threads 896
run t640 MFB_X=1 $QB --threads 640
run t768 MFB_X=1 $QB --threads 768
run poll100 MFB_RING_POLL_NS=100 $QB --threads 768
run poll400 MFB_RING_POLL_NS=400 $QB --threads 768
run ldg768w MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 768
run ldg896w MFB_LIBRARY=$V/libminifem_b200_ldg.so $QB --threads 896
run lap896 MFB_X=1 $QB --op lap --threads 896
run lap1024 MFB_X=1 $QB --op lap --threads 1024
