"""TEST INFRASTRUCTURE — ctypes handles on the CPU oracle (oracle/libminifem_oracle.so, the
plain-C restatement) and on the reference's own compiled sources (oracle/_ref/*.so).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libminifem_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """oracle/minifem_oracle.c"""

    def __init__(self):
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_norm.restype = C.c_double
        self.lib.orc_norm.argtypes = [C.c_void_p, C.c_long]

    def node_to_elem(self, e2n, nbNodes):
        e2n = _i32(e2n)
        nbElem = e2n.size // 4
        idx = np.zeros(nbNodes + 1, np.int32)
        val = np.zeros(max(nbElem * 4, 1), np.int32)
        self.lib.orc_node_to_elem(_p(e2n), nbElem, nbNodes, _p(idx), _p(val))
        return idx, val[:nbElem * 4]

    def create_nodeToNode(self, e2n, nbNodes):
        e2n = _i32(e2n)
        nbElem = e2n.size // 4
        n = self.lib.orc_create_nodeToNode(_p(e2n), nbElem, nbNodes, None, None)
        row = np.zeros(nbNodes + 1, np.int32)
        col = np.zeros(max(n, 1), np.int32)
        self.lib.orc_create_nodeToNode(_p(e2n), nbElem, nbNodes, _p(row), _p(col))
        return row, col[:n]

    def create_elemToEdge(self, row, col, e2n):
        e2n = _i32(e2n)
        nbElem = e2n.size // 4
        out = np.zeros(max(nbElem * 16, 1), np.int32)
        self.lib.orc_create_elemToEdge(_p(_i32(row)), _p(_i32(col)), _p(e2n), _p(out), nbElem)
        return out[:nbElem * 16]

    def coloring(self, e2n, nbNodes):
        e2n = _i32(e2n)
        nbElem = e2n.size // 4
        part = np.zeros(max(nbElem, 1), np.int32)
        c2e = np.zeros(129, np.int32)
        perm = np.zeros(max(nbElem, 1), np.int32)
        nb = self.lib.orc_coloring(_p(e2n), nbElem, nbNodes, _p(part), _p(c2e), _p(perm))
        if nb < 0:
            return None
        return part[:nbElem], c2e[:nb + 1].copy(), perm[:nbElem], nb

    def permute_int_2d(self, tab, perm, dim):
        t = _i32(tab).copy()
        self.lib.orc_permute_int_2d(_p(t), _p(_i32(perm)), t.size // dim, dim)
        return t

    def boundary_mask(self, codes):
        codes = _i32(codes)
        out = np.zeros(max(codes.size * 3, 1), np.int32)
        self.lib.orc_boundary_mask(_p(codes), codes.size, _p(out))
        return out[:codes.size * 3]

    def elem_coef(self, coord, e2n, elem):
        out = np.zeros(12)
        self.lib.orc_elem_coef(_p(_f64(coord)), _p(_i32(e2n)), elem, _p(out))
        return out

    def assembly(self, coord, row, col, e2n, operatorID, elemToEdge=None, colorToElem=None):
        e2n = _i32(e2n)
        nbElem = e2n.size // 4
        dim = 1 if operatorID == 0 else 9
        nbEdges = int(row[-1])
        values = np.full(max(nbEdges * dim, 1), np.nan)
        nbColors = 0 if colorToElem is None else len(colorToElem) - 1
        self.lib.orc_assembly(_p(_f64(coord)), _p(values), _p(_i32(row)), _p(_i32(col)), _p(e2n),
                              _p(None if elemToEdge is None else _i32(elemToEdge)), nbElem, nbEdges,
                              operatorID, _p(None if colorToElem is None else _i32(colorToElem)), nbColors)
        return values[:nbEdges * dim]

    def prec_init(self, values, row, col, nbNodes, dim):
        prec = np.full(max(nbNodes * dim, 1), np.nan)
        self.lib.orc_prec_init(_p(prec), _p(_f64(values)), _p(_i32(row)), _p(_i32(col)), nbNodes, dim)
        return prec[:nbNodes * dim]

    def prec_inversion(self, prec, row, col, checkBounds, nbNodes, operatorID):
        prec = _f64(prec).copy()
        with np.errstate(all="ignore"):
            self.lib.orc_prec_inversion(_p(prec), _p(_i32(row)), _p(_i32(col)), _p(_i32(checkBounds)),
                                        nbNodes, operatorID)
        return prec

    def halo_exchange(self, precs, intfIndex, intfNodes, neighborsList, dim):
        """In place on the list of per-rank prec arrays."""
        n = len(precs)
        keep = [[_f64(p) for p in precs], [_i32(a) for a in intfIndex], [_i32(a) for a in intfNodes],
                [_i32(a) for a in neighborsList]]
        for r in range(n):
            assert keep[0][r] is precs[r], "prec arrays must be contiguous float64 (updated in place)"
        arr = lambda lst: (C.c_void_p * n)(*[a.ctypes.data for a in lst])
        nbIntf = _i32([len(a) - 1 for a in intfIndex])
        rc = self.lib.orc_halo_exchange(n, arr(keep[0]), arr(keep[1]), arr(keep[2]), arr(keep[3]), _p(nbIntf), dim)
        assert rc == 0, "interface lists of two facing subdomains disagree"

    def norm(self, a):
        a = _f64(a)
        return self.lib.orc_norm(_p(a), a.size)

    def fem_iteration(self, setup):
        """assembly + prec_init + prec_inversion on one subdomain (no halo); `setup` is a
        minifem_b200.Setup-like object.  Returns (values, precInit, precInverted)."""
        m = setup.mesh
        values = self.assembly(m.coord, setup.row, setup.col, setup.elemToNode, setup.operatorID,
                               setup.elemToEdge, setup.colorToElem)
        prec0 = self.prec_init(values, setup.row, setup.col, m.nbNodes, setup.operatorDim)
        prec = self.prec_inversion(prec0, setup.row, setup.col, setup.checkBounds, m.nbNodes, setup.operatorID)
        return values, prec0, prec


class RankData(C.Structure):
    _fields_ = [("coord", C.c_void_p), ("values", C.c_void_p), ("prec", C.c_void_p),
                ("row", C.c_void_p), ("col", C.c_void_p), ("elemToNode", C.c_void_p),
                ("elemToEdge", C.c_void_p), ("intfIndex", C.c_void_p), ("intfNodes", C.c_void_p),
                ("neighborsList", C.c_void_p), ("checkBounds", C.c_void_p),
                ("nbElem", C.c_int), ("nbNodes", C.c_int), ("nbEdges", C.c_int),
                ("nbIntf", C.c_int), ("nbIntfNodes", C.c_int)]


def ref_available(kind="ref"):
    return os.path.exists(os.path.join(REF_DIR, f"libminifem_ref_{kind}.so"))


class Reference:
    """The reference's own sources compiled by oracle/Makefile (kind: ref, ref_opt,
    coloring, coloring_opt)."""

    def __init__(self, kind="ref"):
        self.kind = kind
        self.optimized = kind.endswith("_opt")
        self.lib = C.CDLL(os.path.join(REF_DIR, f"libminifem_ref_{kind}.so"))
        self.lib.mref_norm.restype = C.c_double
        self.lib.mref_nodeToNode_capacity.restype = C.c_long
        self.lib.mref_fem_loop.argtypes = [C.c_int, C.POINTER(RankData), C.c_int, C.c_int,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int]
        self.lib.mref_main.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]

    def create_nodeToNode(self, e2n, nbNodes):
        e2n = _i32(e2n).copy()
        nbElem = e2n.size // 4
        row = np.zeros(nbNodes + 1, np.int32)
        col = np.zeros(max(self.lib.mref_nodeToNode_capacity(nbElem), 1), np.int32)
        n = self.lib.mref_create_nodeToNode(_p(e2n), nbElem, nbNodes, _p(row), _p(col))
        return row, col[:n].copy()

    def create_elemToEdge(self, row, col, e2n):
        e2n = _i32(e2n).copy()
        nbElem = e2n.size // 4
        out = np.zeros(max(nbElem * 16, 1), np.int32)
        self.lib.mref_create_elemToEdge(_p(_i32(row).copy()), _p(_i32(col).copy()), _p(e2n), _p(out), nbElem)
        return out[:nbElem * 16]

    def coloring(self, e2n, nbNodes):
        """Returns (permuted elemToNode, colorPerm, colorToElem, nbTotalColors)."""
        e2n = _i32(e2n).copy()
        nbElem = e2n.size // 4
        perm = np.zeros(max(nbElem, 1), np.int32)
        c2e = np.zeros(129, np.int32)
        nb = self.lib.mref_coloring(_p(e2n), nbElem, nbNodes, _p(perm), _p(c2e))
        assert nb > 0, "mref_coloring needs a COLORING build"
        return e2n, perm[:nbElem], c2e[:nb + 1].copy(), nb

    def set_colors(self, colorToElem):
        c2e = _i32(colorToElem)
        self.lib.mref_set_colors(_p(c2e), len(c2e) - 1)

    def boundary_mask(self, codes):
        codes = _i32(codes).copy()
        out = np.zeros(max(codes.size * 3, 1), np.int32)
        nb = int(np.count_nonzero(codes))
        self.lib.mref_boundary_mask(_p(codes), codes.size, nb, _p(out))
        return out[:codes.size * 3]

    def assembly(self, coord, row, col, e2n, operatorID, elemToEdge=None):
        e2n = _i32(e2n).copy()
        nbElem = e2n.size // 4
        dim = 1 if operatorID == 0 else 9
        nbEdges = int(row[-1])
        values = np.full(max(nbEdges * dim, 1), np.nan)
        if self.optimized:
            assert elemToEdge is not None
        self.lib.mref_assembly(_p(_f64(coord).copy()), _p(values), _p(_i32(row).copy()), _p(_i32(col).copy()),
                               _p(e2n), _p(None if elemToEdge is None else _i32(elemToEdge).copy()),
                               nbElem, nbEdges, dim, operatorID)
        return values[:nbEdges * dim]

    def prec_init(self, values, row, col, nbNodes, dim):
        prec = np.full(max(nbNodes * dim, 1), np.nan)
        self.lib.mref_prec_init(_p(prec), _p(_f64(values).copy()), _p(_i32(row).copy()), _p(_i32(col).copy()),
                                nbNodes, dim)
        return prec[:nbNodes * dim]

    def prec_inversion(self, prec, row, col, checkBounds, nbNodes, operatorID):
        prec = _f64(prec).copy()
        self.lib.mref_prec_inversion(_p(prec), _p(_i32(row).copy()), _p(_i32(col).copy()),
                                     _p(_i32(checkBounds).copy()), nbNodes, operatorID)
        return prec

    def norm(self, a):
        a = _f64(a)
        return self.lib.mref_norm(_p(a), a.size)

    def fem_loop(self, setups, nbIter, verbose=False):
        """The reference's FEM_loop over len(setups) subdomains (threads).  Returns
        (values per rank, prec per rank, cycles[4], tscHz)."""
        n = len(setups)
        ranks = (RankData * n)()
        keep = []
        for r, s in enumerate(setups):
            m = s.mesh
            arrs = dict(coord=_f64(m.coord).copy(), values=np.zeros(max(s.nbEdges * s.operatorDim, 1)),
                        prec=np.zeros(max(m.nbNodes * s.operatorDim, 1)), row=_i32(s.row).copy(),
                        col=_i32(s.col).copy(), e2n=_i32(s.elemToNode).copy(),
                        e2e=None if s.elemToEdge is None else _i32(s.elemToEdge).copy(),
                        ii=_i32(m.intfIndex).copy(), inn=_i32(m.intfNodes).copy() if m.nbIntfNodes else np.zeros(1, np.int32),
                        nl=_i32(m.neighborsList).copy(), cb=_i32(s.checkBounds).copy())
            if self.optimized:
                assert arrs["e2e"] is not None, "OPTIMIZED build needs elemToEdge"
            keep.append(arrs)
            ranks[r] = RankData(_p(arrs["coord"]), _p(arrs["values"]), _p(arrs["prec"]), _p(arrs["row"]),
                                _p(arrs["col"]), _p(arrs["e2n"]), _p(arrs["e2e"]), _p(arrs["ii"]), _p(arrs["inn"]),
                                _p(arrs["nl"]), _p(arrs["cb"]), m.nbElem, m.nbNodes, s.nbEdges, m.nbIntf, m.nbIntfNodes)
        cycles = (C.c_uint64 * 4)()
        hz = C.c_double(0)
        operatorID = setups[0].operatorID
        self.lib.mref_fem_loop(n, ranks, nbIter, operatorID, cycles, C.byref(hz), int(verbose))
        values = [k["values"][:s.nbEdges * s.operatorDim] for k, s in zip(keep, setups)]
        precs = [k["prec"][:s.mesh.nbNodes * s.operatorDim] for k, s in zip(keep, setups)]
        return values, precs, list(cycles), hz.value

    def main(self, mesh, op, nbIter):
        return self.lib.mref_main(mesh.encode(), op.encode(), str(nbIter).encode())
