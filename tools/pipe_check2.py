import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import minifem_b200 as mfb
from helpers import row_scaled_error, block_scaled_error
from oracle_lib import Oracle
orc = Oracle()
for grid in ((5, 4, 3), (1, 1, 1), (16, 9, 12)):
    mesh = mfb.Mesh.generate(*grid, seed=2)
    for op in ("lap", "ela"):
        setup = mfb.Setup(mesh, op)
        want_v, _, want_p = orc.fem_iteration(setup)
        for mode in ("staged", "fused"):
            ctx = mfb.Context(setup, path="tiled")
            if mode == "staged": ctx.assembly()
            else: ctx.iteration()
            v, p = ctx.download()
            err = np.abs(v - want_v)
            bad = np.flatnonzero(err > 1e-9 * np.abs(want_v).max())
            rows = np.searchsorted(setup.row, bad, side="right") - 1
            print(grid, op, mode, ctx.plan_stats()["tiles"], "err", row_scaled_error(v, want_v, setup.row, setup.operatorDim), "bad entries", bad[:10], "rows", rows[:10], flush=True)
            ctx.close()
