"""Scratch timing of the three assembly paths on one GPU (development aid, not bench.py)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
import numpy as np
import minifem_b200 as mfb

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, nargs=3, default=[100, 100, 100])
ap.add_argument("--op", default="ela")
ap.add_argument("--paths", default="tiled,atomic,color", help="comma-separated: tiled, ring, atomic, color")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--tile-rows", type=int, default=0)
ap.add_argument("--tile-elems", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0)
ap.add_argument("--no-bank-aware", action="store_true")
ap.add_argument("--file", default=None, help="read the mesh from an input file (src/IO.cc format, e.g. tools/delaunay_mesh.py) instead of generating a Kuhn grid")
ap.add_argument("--shuffle", action="store_true", help="random node and element numbering (unstructured-like input order)")
a = ap.parse_args()

t0 = time.time()
mesh = mfb.Mesh.read(a.file) if a.file else mfb.Mesh.generate(*a.grid, seed=1)
print(f"mesh {a.file or a.grid}: E={mesh.nbElem} N={mesh.nbNodes} Z={mesh.nbEdges}  ({time.time()-t0:.1f}s)", flush=True)
if a.shuffle:
    rng = np.random.default_rng(5)
    nperm = rng.permutation(mesh.nbNodes)                  # old node -> new node
    eperm = rng.permutation(mesh.nbElem)
    coord = np.empty_like(mesh.coord.reshape(-1, 3)); coord[nperm] = mesh.coord.reshape(-1, 3)
    codes = np.empty_like(mesh.boundNodesCode); codes[nperm] = mesh.boundNodesCode
    e2n = (nperm[mesh.elemToNode.reshape(-1, 4) - 1] + 1)[eperm]
    mesh.coord, mesh.boundNodesCode, mesh.elemToNode = coord.ravel(), codes, np.ascontiguousarray(e2n, np.int32).ravel()
E, N, Z = mesh.nbElem, mesh.nbNodes, mesh.nbEdges
alg = (16 * E + 76 * Z + 112 * N) if a.op == "ela" else (16 * E + 12 * Z + 36 * N)
ref = None
for path in a.paths.split(","):
    t0 = time.time()
    setup = mfb.Setup(mesh, a.op, coloring=(path == "color"))
    t1 = time.time()
    ctx = mfb.Context(setup, path=path, tile_rows=a.tile_rows, tile_elems=a.tile_elems, threads=a.threads,
                      use_graph=(path in ("color", "blockcolor")), ctas=a.ctas, bank_aware=not a.no_bank_aware)
    t2 = time.time()
    for _ in range(3):
        ctx.iteration()
    ctx.sync()
    ms = []
    for _ in range(a.steps):
        ctx.iteration()
        ms.append(ctx.stage_ms()[4])
    ms = np.array(ms)
    line = f"{path:7s} setup {t1-t0:.1f}s ctx {t2-t1:.1f}s  iter ms: med {np.median(ms):.4f} min {ms.min():.4f}  -> {E/np.median(ms)/1e6:.2f} Gelem/s, alg {alg/np.median(ms)/1e6:.1f} GB/s"
    if path in ("tiled", "ring"):
        line += f"  plan {ctx.plan_stats()} bytes {ctx.device_bytes()}"
    print(line, flush=True)
    v, p = ctx.download()
    if path not in ("color",):
        if ref is None:
            ref = (v, p)
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from helpers import row_scaled_error, block_scaled_error
            print("   vs first path (row / block scaled): values", row_scaled_error(v, ref[0], setup.row, setup.operatorDim),
                  "prec", block_scaled_error(p, ref[1], setup.operatorDim))
    ctx.close()
