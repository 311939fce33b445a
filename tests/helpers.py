"""Shared helpers of the parity tests."""
import numpy as np

# north_star: assembled values and preconditioner entries within 1e-12 relative error of the
# reference.  Individual entries can cancel to (nearly) zero — e.g. Laplacian couplings of
# orthogonal gradients — so, as SURVEY.md §7 item 5 prescribes, an entry's error is measured
# against the largest magnitude of its CSR row (values) or of its node block (prec), never
# against the entry itself.
RTOL = 1e-12


def row_scaled_error(got, want, row, dim):
    """max over entries of |got-want| / max|want over the CSR row|."""
    got = np.asarray(got, dtype=np.float64).reshape(-1, dim)
    want = np.asarray(want, dtype=np.float64).reshape(-1, dim)
    assert got.shape == want.shape
    if want.size == 0:
        return 0.0
    row = np.asarray(row)
    lens = np.diff(row)
    entry_scale = np.abs(want).max(axis=1)
    starts = row[:-1][lens > 0]
    row_scale = np.maximum.reduceat(entry_scale, starts) if starts.size else np.zeros(0)
    scale = np.repeat(row_scale, lens[lens > 0])
    err = np.abs(got - want).max(axis=1)
    ok = scale > 0
    assert np.all(err[~ok] == 0)
    return float((err[ok] / scale[ok]).max()) if ok.any() else 0.0


def block_scaled_error(got, want, dim):
    """max over nodes of |got-want| / max|want over the node's block| (inf/nan must coincide)."""
    got = np.asarray(got, dtype=np.float64).reshape(-1, dim)
    want = np.asarray(want, dtype=np.float64).reshape(-1, dim)
    assert got.shape == want.shape
    if want.size == 0:
        return 0.0
    finite = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), finite), "non-finite entries differ"
    assert np.array_equal(got[~finite], want[~finite], equal_nan=True)
    w = np.where(finite, want, 0.0)
    g = np.where(finite, got, 0.0)
    scale = np.abs(w).max(axis=1)
    err = np.abs(g - w).max(axis=1)
    ok = scale > 0
    assert np.all(err[~ok] == 0)
    return float((err[ok] / scale[ok]).max()) if ok.any() else 0.0


def random_tet_mesh(rng, nbNodes, nbElem):
    """Unstructured stand-in: random 4-subsets of jittered points (valid for layout tests; the
    volumes are arbitrary but non-zero with probability 1)."""
    coord = rng.uniform(-1, 1, size=(nbNodes, 3))
    elems = np.stack([rng.choice(nbNodes, size=4, replace=False) for _ in range(nbElem)]) + 1
    return coord.ravel(), elems.astype(np.int32).ravel()


class ArrayMesh:
    """Minimal mesh object for Setup/Oracle when the arrays do not come from the generator."""

    def __init__(self, coord, elemToNode, nbNodes, boundNodesCode=None):
        self.coord = np.ascontiguousarray(coord, dtype=np.float64)
        self.elemToNode = np.ascontiguousarray(elemToNode, dtype=np.int32)
        self.nbNodes = int(nbNodes)
        self.nbElem = self.elemToNode.size // 4
        self.boundNodesCode = (np.zeros(nbNodes, np.int32) if boundNodesCode is None
                               else np.ascontiguousarray(boundNodesCode, dtype=np.int32))
        self.nbBoundNodes = int(np.count_nonzero(self.boundNodesCode))
        self.nbIntf, self.nbIntfNodes = 0, 0
        self.neighborsList = np.zeros(3, np.int32)
        self.intfIndex = np.zeros(1, np.int32)
        self.intfNodes = np.zeros(0, np.int32)
        self.nbEdges = 0
        self.globalNode = None
