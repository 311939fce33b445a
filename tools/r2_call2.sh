#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # name env... -- args
    name=$1; shift
    timeout 300 env "$@" > gpurun_out/c2_$name.log 2>&1
    echo "$name: $(grep '^ring' gpurun_out/c2_$name.log | tail -1 | cut -c1-150)"
}
QB="python tools/quick_bench.py --paths ring --steps 30"
run base256 MFB_X=1 $QB
run t384_34 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 34 --tile-elems 640
run t384_33 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 33 --tile-elems 640
run t384_32 MFB_RING_MAXJOBS=384 $QB --threads 384 --tile-rows 32 --tile-elems 640
run t384_36 MFB_X=1 $QB --threads 384 --tile-rows 36 --tile-elems 640
run t256_22 MFB_RING_MAXJOBS=256 $QB --tile-rows 22 --tile-elems 420
run t256_46 MFB_RING_MAXJOBS=512 $QB --tile-rows 46 --tile-elems 820
