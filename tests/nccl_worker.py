"""One rank of the NCCL halo test (launched by torch.distributed.run): every path's fused
iteration over a 2-block partition, checked against the oracle's multi-subdomain result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import minifem_b200 as mfb
from minifem_b200 import dist as mdist
from helpers import RTOL, block_scaled_error, row_scaled_error
from oracle_lib import Oracle

rank, world = mdist.init_from_env("nccl")
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
grid = (12, 10, 8)
blocks = mfb.choose_blocks(*grid, world)
oracle = Oracle()
meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=4) for r in range(world)]
for op in ("lap", "ela"):
    dim = 1 if op == "lap" else 9
    setups = [mfb.Setup(m, op) for m in meshes]
    results = [oracle.fem_iteration(s) for s in setups]
    precs = [np.ascontiguousarray(res[1]) for res in results]
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes],
                         [m.neighborsList for m in meshes], dim)
    s = setups[rank]
    want_p = oracle.prec_inversion(precs[rank], s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID)
    for path in ("ring", "tiled", "atomic"):
        ctx = mfb.Context(s, path=path, device=int(os.environ.get("LOCAL_RANK", rank)), nbBlocks=world, rank=rank,
                          tile_rows=32 if path != "ring" else 16, tile_elems=400 if path != "ring" else 300)
        mdist.comm_init(ctx)
        modes = ["fused", "staged"]
        if path == "ring":
            # the same fused iteration over the peer-to-peer windows (cudaIpc between the rank processes), twice
            # (both receive-buffer parities), then NCCL again
            active, why = mdist.p2p_connect(ctx)
            assert active or os.environ.get("MFB_HALO", "") == "nccl", why
            if active:
                modes = ["fused", "fused", "fused", "staged", "nccl"]
        for mode in modes:
            if mode == "nccl":
                ctx.p2p_enable(False)
                ctx.iteration()
            elif mode == "fused":
                ctx.iteration()
            else:
                ctx.stages()
            v, p = ctx.download()
            assert row_scaled_error(v, results[rank][0], s.row, dim) <= RTOL, (op, path, mode)
            assert block_scaled_error(p, want_p, dim) <= RTOL, (op, path, mode)
        ctx.close()
mdist.barrier()
print("NCCL_WORKER_OK", rank, flush=True)
