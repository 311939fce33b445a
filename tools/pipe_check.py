"""Development aid: correctness of the pipelined TILED kernel (threads=0) against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import minifem_b200 as mfb
from helpers import row_scaled_error, block_scaled_error
from oracle_lib import Oracle
orc = Oracle()
for grid in ((3, 3, 3), (12, 10, 9), (30, 30, 30)):
    mesh = mfb.Mesh.generate(*grid, seed=7)
    for op in ("ela", "lap"):
        setup = mfb.Setup(mesh, op)
        want_v, _, want_p = orc.fem_iteration(setup)
        for tr, te in ((0, 0), (16, 200)):
            ctx = mfb.Context(setup, path="tiled", threads=0, tile_rows=tr, tile_elems=te)
            for it in range(3):
                ctx.iteration()
            v, p = ctx.download()
            print(grid, op, tr, te, ctx.plan_stats()["tiles"], "err", row_scaled_error(v, want_v, setup.row, setup.operatorDim),
                  block_scaled_error(p, want_p, setup.operatorDim), flush=True)
            ctx.close()
print("PIPE_CHECK_DONE")
