#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/c9_$name.log 2>&1; echo "$name: $(grep '^ring' gpurun_out/c9_$name.log | tail -1 | cut -c1-110) $(grep -o "smem_bytes': [0-9]*" gpurun_out/c9_$name.log | tail -1)"; }
QB="python tools/quick_bench.py --paths ring --steps 30"
for j in 8 7 6; do
  L=$PWD/mini-fem_b200/libminifem_b200_j$j.so
  run j${j}_384 MFB_LIBRARY=$L $QB
  run j${j}_768 MFB_LIBRARY=$L $QB --threads 768
done
L=$PWD/mini-fem_b200/libminifem_b200_j6.so
run j6_lap MFB_LIBRARY=$L $QB --op lap
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 4 -c 1 -o gpurun_out/r2_ringws6_ela_full \
    env MFB_LIBRARY=$L python tools/quick_bench.py --paths ring --steps 4 > gpurun_out/c9_ncu_full.log 2>&1
echo "ncu rc=$?"
