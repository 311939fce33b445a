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


# ---------------------------------------------------------------------------------------------
# Extended-precision truth for ill-shaped elements.
#
# north_star asks for 1e-12 relative to the reference.  On FEM-quality meshes every path meets that
# against the double-precision oracle directly.  On RANDOM 4-subsets of points (layout stress tests)
# elements can be arbitrarily flat: the reference's own formula (elem_coef_seq, src/assembly.cc:85-121)
# then loses eps * h^3 / V digits in double precision, i.e. the reference is itself further than 1e-12
# from the exact value of its formula.  An evaluation that bases the same gradients at another node of
# the element (the RING path) loses as many digits, but different ones.  For those meshes the bar is
# therefore stated against the formula evaluated in 80-bit arithmetic (numpy longdouble on x86-64):
#   err(path, truth) <= max(1e-12, SLIVER_FACTOR * err(reference, truth)).
SLIVER_FACTOR = 8.0


def extended_truth(setup):
    """nodeToNodeValue of the reference's formulas (src/assembly.cc:85-121, :386-409, :539-541) in
    numpy longdouble.  Meant for meshes of a few thousand elements."""
    ld = np.longdouble
    m = setup.mesh
    dim = setup.operatorDim
    e2n = np.asarray(setup.elemToNode, np.int64).reshape(-1, 4) - 1
    p = np.asarray(m.coord, np.float64).reshape(-1, 3).astype(ld)[e2n]          # [E, 4, 3]
    a, b, c = p[:, 0] - p[:, 3], p[:, 2] - p[:, 3], p[:, 1] - p[:, 3]           # edge vectors from node 3 to nodes 0, 2, 1
    r0, r1, r2 = np.cross(b, c), np.cross(a, b), np.cross(c, a)
    coef = np.stack([r0, r1, r2, -(r0 + r1 + r2)], axis=1)                      # [E, 4, 3]
    vol = np.einsum("ek,ek->e", a, r0)
    coef = coef / vol[:, None, None]
    row, col = np.asarray(setup.row, np.int64), np.asarray(setup.col, np.int64)
    values = np.zeros((int(row[-1]), dim), ld)
    for j in range(4):
        for k in range(4):
            # the reference's search finds the FIRST column equal to node k + 1 in the row of node j (:419-421)
            idx = np.empty(e2n.shape[0], np.int64)
            for e in range(e2n.shape[0]):
                nj, nk = e2n[e, j], e2n[e, k]
                seg = col[row[nj]:row[nj + 1]]
                idx[e] = row[nj] + int(np.nonzero(seg == nk + 1)[0][0])
            cj, ck = coef[:, j], coef[:, k]
            if dim == 1:
                np.add.at(values[:, 0], idx, np.einsum("ek,ek->e", cj, ck))
            else:
                dot = np.einsum("ek,ek->e", cj, ck)
                blk = ld(1.25) * cj[:, :, None] * ck[:, None, :] + dot[:, None, None] * np.eye(3, dtype=ld)[None]
                np.add.at(values, idx, blk.reshape(-1, 9))
    return values.reshape(-1)


def assert_close_or_conditioned(got, want, truth, row, dim, what=""):
    """1e-12 against the reference's result, or — where the reference itself is further than that from
    the exact value of its formula — within SLIVER_FACTOR times the reference's own error of that value."""
    direct = row_scaled_error(got, want, row, dim)
    if direct <= RTOL:
        return direct
    truth64 = np.asarray(truth, np.longdouble)
    def err(x):
        x = np.asarray(x, np.float64).astype(np.longdouble).reshape(-1, dim)
        t = truth64.reshape(-1, dim)
        lens = np.diff(np.asarray(row))
        scale = np.repeat(np.maximum.reduceat(np.abs(t).max(axis=1), np.asarray(row)[:-1][lens > 0]), lens[lens > 0])
        e = np.abs(x - t).max(axis=1)
        ok = scale > 0
        return float((e[ok] / scale[ok]).max()) if ok.any() else 0.0
    e_ref, e_got = err(want), err(got)
    assert e_ref > RTOL / SLIVER_FACTOR, f"{what}: {direct:.2e} from the reference although the reference is within {e_ref:.2e} of the exact formula"
    assert e_got <= SLIVER_FACTOR * e_ref, f"{what}: {e_got:.2e} from the exact formula, the reference {e_ref:.2e}"
    return direct


def extended_truth_prec(setup, values_truth):
    """prec_init + prec_inversion (src/preconditioner.cc:25-87, src/Fortran/elasclpr.f:19-53) of the
    extended-precision matrix, in longdouble (cofactor inverse); nodes without diagonal entry keep the
    masked zero block / 1/0 like the reference."""
    ld = np.longdouble
    m = setup.mesh
    dim = setup.operatorDim
    row, col = np.asarray(setup.row, np.int64), np.asarray(setup.col, np.int64)
    vals = np.asarray(values_truth, ld).reshape(-1, dim)
    prec = np.zeros((m.nbNodes, dim), ld)
    cb = np.asarray(setup.checkBounds).reshape(3, m.nbNodes)
    for i in range(m.nbNodes):
        seg = col[row[i]:row[i + 1]]
        hit = np.nonzero(seg == i + 1)[0]
        has_diag = hit.size > 0
        if has_diag:
            prec[i] = vals[row[i] + hit[0]]
        if dim == 1:
            with np.errstate(divide="ignore"):
                prec[i, 0] = ld(1.0) / prec[i, 0]
            continue
        b = prec[i].reshape(3, 3).copy()
        for c in range(3):
            if cb[c, i] != 0:
                b[c, :] = 0; b[:, c] = 0; b[c, c] = 1
        if has_diag:
            cof = np.empty((3, 3), ld)
            for r in range(3):
                for s in range(3):
                    cof[r, s] = (b[(r + 1) % 3, (s + 1) % 3] * b[(r + 2) % 3, (s + 2) % 3]
                                 - b[(r + 1) % 3, (s + 2) % 3] * b[(r + 2) % 3, (s + 1) % 3])
            det = (b[0] * cof[0]).sum()
            b = cof.T / det
        prec[i] = b.reshape(9)
    return prec.reshape(-1)


def diag_conditioning(setup, values):
    """rho_i = max |K_ij| over row i / max |K_ii|: how much larger than the diagonal block the largest block of the
    row is.  <= ~3 on FEM-quality meshes; unbounded on random tetrahedra (a flat element gives its blunt node a
    huge gradient, hence huge couplings K_ij next to a moderate K_ii)."""
    dim = setup.operatorDim
    row, col = np.asarray(setup.row, np.int64), np.asarray(setup.col, np.int64)
    blocks = np.abs(np.asarray(values, np.float64).reshape(-1, dim)).max(axis=1)
    rho = np.zeros(setup.mesh.nbNodes)
    for i in range(setup.mesh.nbNodes):
        seg = col[row[i]:row[i + 1]]
        hit = np.nonzero(seg == i + 1)[0]
        if hit.size and blocks[row[i] + hit[0]] > 0:
            rho[i] = blocks[row[i]:row[i + 1]].max() / blocks[row[i] + hit[0]]
    return rho


def assert_prec_close_or_conditioned(got, want, truth, dim, what="", rho=None):
    """1e-12 against the reference's preconditioner block, or — on ill-shaped meshes — against the 80-bit
    truth within max(1e-12, SLIVER_FACTOR x the reference's own error) x max(1, rho_i).  The factor rho_i is
    the RING path's: its diagonal block is minus the sum of the row's off-diagonal blocks (zero row sums of
    the element matrices), so it inherits the absolute error the row's LARGEST block is allowed, rho_i times
    larger relative to the diagonal, where the reference sums |grad_i|^2 terms directly
    (helpers.diag_conditioning: rho <= ~3 on FEM-quality meshes, where every path is held to 1e-12 outright)."""
    direct = block_scaled_error(got, want, dim)
    if direct <= RTOL:
        return direct
    t = np.asarray(truth, np.longdouble).reshape(-1, dim)
    def err(x):
        x = np.asarray(x, np.float64).astype(np.longdouble).reshape(-1, dim)
        fin = np.isfinite(t).all(axis=1) & np.isfinite(x).all(axis=1)
        scale = np.abs(t).max(axis=1)
        e = np.zeros(t.shape[0])
        ok = fin & (scale > 0)
        e[ok] = (np.abs(x[ok] - t[ok]).max(axis=1) / scale[ok]).astype(np.float64)
        return e
    e_ref, e_got = err(want), err(got)
    bound = np.full(t.shape[0], max(RTOL, SLIVER_FACTOR * float(e_ref.max())))
    if rho is not None:
        bound = bound * np.maximum(1.0, np.asarray(rho))
    worst = int(np.argmax(e_got - bound))
    assert np.all(e_got <= bound), (f"{what}: prec of node {worst} is {e_got[worst]:.2e} from the exact formula (reference: {e_ref[worst]:.2e}, "
                                    f"rho {0 if rho is None else rho[worst]:.1e}, bound {bound[worst]:.2e})")
    return direct
