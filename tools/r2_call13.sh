#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 150 env "$@" > gpurun_out/c13_$name.json 2> gpurun_out/c13_$name.err; echo "$name rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/c13_$name.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d.get('parity'))" 2>&1 | tail -1)"; }
run eager_r8 MFB_X=1 $B --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 0 --no-strong
run eager_r0 MFB_HALO_RESERVE_CTAS=0 $B --master-port 29562 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 0 --no-strong --no-parity
run eager_r16 MFB_HALO_RESERVE_CTAS=16 $B --master-port 29563 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 0 --no-strong --no-parity
run graph_r8 MFB_MULTI_GPU_GRAPH=1 $B --master-port 29564 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 0 --no-strong --no-parity
tail -3 gpurun_out/c13_graph_r8.err
run tiled MFB_X=1 $B --master-port 29565 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 0 --no-strong --no-parity --path tiled
