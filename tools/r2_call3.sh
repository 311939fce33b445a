#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c3_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c3_$name.log | tail -1 | cut -c1-150)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run t384_34 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 34 --tile-elems 640
run t384_33 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 33 --tile-elems 640
run t384_36 MFB_X=1 $QB --threads 384 --tile-rows 36 --tile-elems 640
run t384_54 MFB_X=1 $QB --threads 384 --tile-rows 54 --tile-elems 960
run t384_68 MFB_RING_MAXJOBS=768 $QB --threads 384 --tile-rows 68 --tile-elems 1200
