#!/bin/bash
# one GPU: suspend-time hint of the barrier waits (does a sleeping warp wake late?)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 100 env "$@" > gpurun_out/c27_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c27_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run hint20000 MFB_X=1 $QB
run hint2000 MFB_RING_WAIT_HINT_NS=2000 $QB
run hint200 MFB_RING_WAIT_HINT_NS=200 $QB
run hint1 MFB_RING_WAIT_HINT_NS=1 $QB
