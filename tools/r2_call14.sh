#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_integration_binding.py tests/test_gpu_parity.py -m gpu -x -q -k "binding or driver or unstructured_random or fixtures or interval" > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/c14_pytest.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --configs none --e2e-steps 2 > gpurun_out/c14_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# full capture of the RING kernel as bench.py runs it
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ring_assembly -s 6 -c 1 -o gpurun_out/r2_ring_final_ela_full \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --configs none --e2e-steps 0 --no-other-paths --no-parity > gpurun_out/c14_ncu_full.log 2>&1
echo "ncu ring rc=$?"
# DRAM throughput of the two scatter kernels (colour vs atomic), EIB lap and ela
for op in lap ela; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum \
     --clock-control none -k regex:scatter_elements -s 40 -c 34 --csv --log-file gpurun_out/r2_scatter_$op.csv \
     python tools/quick_bench.py --paths atomic,color --op $op --steps 1 > gpurun_out/c14_scatter_$op.log 2>&1
  echo "scatter $op rc=$?"
done
