"""Unstructured test input: Delaunay tetrahedralisation of random points in the unit cube, written
in the reference's per-rank input format (src/IO.cc:61-96), single subdomain.
usage: python tools/delaunay_mesh.py NPOINTS OUTFILE [seed]
Slivers thinner than 1e-9 of a typical cell volume are dropped (elem_coef_seq divides by the
volume).  Boundary codes: 52 on x < 0.02, 54 on z < 0.02, 10 on y > 0.98."""
import os, sys
import numpy as np
from scipy.spatial import Delaunay


def delaunay_arrays(npoints, seed=1):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.0, 1.0, size=(npoints, 3))
    tets = Delaunay(pts).simplices.astype(np.int64)
    a, b, c, d = (pts[tets[:, k]] for k in range(4))
    vol = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0
    keep = np.abs(vol) > 1e-9 / npoints
    tets = tets[keep]
    used = np.zeros(npoints, bool)
    used[tets.ravel()] = True                      # every Delaunay vertex is used; kept for safety
    remap = np.cumsum(used) - 1
    pts, tets = pts[used], remap[tets]
    codes = np.zeros(len(pts), np.int32)
    codes[pts[:, 0] < 0.02] = 52
    codes[pts[:, 2] < 0.02] = 54
    codes[pts[:, 1] > 0.98] = 10
    return np.ascontiguousarray(pts.ravel()), np.ascontiguousarray((tets + 1).astype(np.int32).ravel()), codes


def write_input(path, coord, e2n, codes, nbEdges):
    nbNodes, nbElem = coord.size // 3, e2n.size // 4
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        np.array([nbElem, nbNodes, nbEdges, 0, 0, int(np.count_nonzero(codes))], np.int32).tofile(f)
        coord.astype(np.float64).tofile(f)
        e2n.astype(np.int32).tofile(f)
        np.zeros(3, np.int32).tofile(f)             # neighborsList: max(nbIntf, 1) * 3
        np.zeros(1, np.int32).tofile(f)             # intfIndex: nbIntf + 1
        codes.astype(np.int32).tofile(f)            # (intfNodes is empty)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mini-fem_b200", "python"))
    import minifem_b200 as mfb
    n, out = int(sys.argv[1]), sys.argv[2]
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    coord, e2n, codes = delaunay_arrays(n, seed)
    row, col = mfb.create_nodeToNode(e2n, coord.size // 3)
    write_input(out, coord, e2n, codes, int(row[-1]))
    deg = np.diff(row)
    print(f"{out}: {coord.size // 3} nodes, {e2n.size // 4} tets, {int(row[-1])} CSR entries; row length mean {deg.mean():.1f} max {deg.max()}")
