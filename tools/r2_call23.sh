#!/bin/bash
# one GPU: write-out loop split at the diagonal, 768 threads with 96 / 64 registers per role
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c23_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c23_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
V=$PWD/mini-fem_b200/variants
run base768 MFB_X=1 $QB --threads 768
run splitdiag768 MFB_LIBRARY=$V/libminifem_b200_splitdiag.so $QB --threads 768
run regs768 MFB_LIBRARY=$V/libminifem_b200_regs768.so $QB --threads 768
