"""Small all-paths run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import minifem_b200 as mfb
from helpers import row_scaled_error, block_scaled_error
from oracle_lib import Oracle
orc = Oracle()
mesh = mfb.Mesh.generate(9, 8, 7, seed=3)
for op in ("ela", "lap"):
    for path, kw in (("tiled", {}), ("tiled", {"threads": 768}), ("atomic", {}), ("color", {})):
        setup = mfb.Setup(mesh, op, coloring=(path == "color"))
        want_v, _, want_p = orc.fem_iteration(setup)
        ctx = mfb.Context(setup, path=path, **kw)
        ctx.iteration(); ctx.stages(); ctx.iteration()
        v, p = ctx.download()
        print(op, path, kw, row_scaled_error(v, want_v, setup.row, setup.operatorDim), block_scaled_error(p, want_p, setup.operatorDim), flush=True)
        ctx.close()
# the layout builders on the GPU and the device-side norms
e2n, nbNodes = mesh.elemToNode, mesh.nbNodes
row, col = mfb.device_create_nodeToNode(e2n, nbNodes)
ref_row, ref_col = mfb.create_nodeToNode(e2n, nbNodes)
assert np.array_equal(row, ref_row) and np.array_equal(col, ref_col)
assert np.array_equal(mfb.device_create_elemToEdge(row, col, e2n), mfb.create_elemToEdge(row, col, e2n))
got, want = mfb.device_coloring_creation(e2n, nbNodes), mfb.coloring_creation(e2n, nbNodes)
assert all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3]))
ctx = mfb.Context(mfb.Setup(mesh, "ela"), path="tiled")
ctx.iteration()
print("builders ok; norms", ctx.norms(), flush=True)
ctx.close()
print("SANITIZE_CASE_DONE")
