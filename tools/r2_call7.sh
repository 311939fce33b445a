#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c7_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c7_$name.log | tail -1 | cut -c1-110)"; }
QB="python tools/quick_bench.py --paths ring --steps 20"
run ctas148 MFB_X=1 $QB --ctas 148
run ctas296 MFB_X=1 $QB --ctas 296
run ctas444 MFB_X=1 $QB --ctas 444
run lap148 MFB_X=1 $QB --ctas 148 --op lap
run lap296 MFB_X=1 $QB --ctas 296 --op lap
run lap444 MFB_X=1 $QB --ctas 444 --op lap
