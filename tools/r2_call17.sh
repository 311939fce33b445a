#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 200 env "$@" > gpurun_out/c17_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c17_$name.log | tail -1 | cut -c1-105)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
run d384 MFB_X=1 $QB
run d768 MFB_X=1 $QB --threads 768
for j in 13 15 16; do
  L=$PWD/mini-fem_b200/libminifem_b200_j$j.so
  run j${j}_384 MFB_LIBRARY=$L $QB
  run j${j}_768 MFB_LIBRARY=$L $QB --threads 768
done
run d768_56 MFB_X=1 $QB --threads 768 --tile-rows 56 --tile-elems 960
run d768_48 MFB_X=1 $QB --threads 768 --tile-rows 48 --tile-elems 820
run lap384 MFB_X=1 $QB --op lap
timeout 400 python tests/ring_gpu_worker.py > gpurun_out/c17_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/c17_parity.log
