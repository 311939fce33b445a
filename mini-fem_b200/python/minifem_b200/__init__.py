"""ctypes binding of libminifem_b200.so (include/minifem_b200.h) for the test and bench
harness.  The product is the C-ABI library; this module only marshals numpy arrays.

There is no fallback: if the library is missing, importing this module raises, and every
GPU entry point raises MfbError when the library reports an error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))            # mini-fem_b200/
LIB_PATH = os.environ.get("MFB_LIBRARY") or os.path.join(PKG_ROOT, "libminifem_b200.so")   # MFB_LIBRARY: an experimental build

PATH_TILED, PATH_ATOMIC, PATH_COLOR, PATH_RING = 0, 1, 2, 3
PATH_BLOCKCOLOR = 4
PATH_NAMES = {"tiled": PATH_TILED, "atomic": PATH_ATOMIC, "color": PATH_COLOR, "ring": PATH_RING, "blockcolor": PATH_BLOCKCOLOR}
COMM_ID_BYTES = 128
P2P_CARD_BYTES = 1024


class MfbError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {PKG_ROOT} lib` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
lib = C.CDLL(LIB_PATH)

_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)


class MeshView(C.Structure):
    _fields_ = [("nbElem", C.c_int), ("nbNodes", C.c_int), ("nbEdges", C.c_int),
                ("nbIntf", C.c_int), ("nbIntfNodes", C.c_int), ("nbBoundNodes", C.c_int),
                ("coord", _f64p), ("elemToNode", _i32p), ("neighborsList", _i32p),
                ("intfIndex", _i32p), ("intfNodes", _i32p), ("boundNodesCode", _i32p),
                ("globalNode", C.POINTER(C.c_int64))]


class Problem(C.Structure):
    _fields_ = [("operatorID", C.c_int), ("nbElem", C.c_int), ("nbNodes", C.c_int),
                ("nbEdges", C.c_int), ("coord", C.c_void_p), ("elemToNode", C.c_void_p),
                ("nodeToNodeRow", C.c_void_p), ("nodeToNodeColumn", C.c_void_p),
                ("elemToEdge", C.c_void_p), ("checkBounds", C.c_void_p),
                ("colorToElem", C.c_void_p), ("nbTotalColors", C.c_int),
                ("nbBlocks", C.c_int), ("rank", C.c_int), ("nbIntf", C.c_int),
                ("nbIntfNodes", C.c_int), ("intfIndex", C.c_void_p),
                ("intfNodes", C.c_void_p), ("neighborsList", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("path", C.c_int), ("device", C.c_int), ("tileRows", C.c_int),
                ("tileElems", C.c_int), ("threads", C.c_int), ("useGraph", C.c_int),
                ("ctas", C.c_int), ("bankAware", C.c_int)]


lib.mfb_last_error.restype = C.c_char_p
lib.mfb_version.restype = C.c_char_p
lib.mfb_count_edges.restype = C.c_int64
lib.mfb_double_norm.restype = C.c_double
lib.mfb_double_norm.argtypes = [C.c_void_p, C.c_int64]
lib.mfb_ctx_launch_count.restype = C.c_int64
lib.mfb_ctx_launch_count.argtypes = [C.c_void_p]
lib.mfb_mesh_generate.argtypes = [C.c_int] * 7 + [C.c_uint64, C.POINTER(C.c_void_p)]
lib.mfb_mesh_read.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
lib.mfb_mesh_write.argtypes = [C.c_void_p, C.c_char_p]
lib.mfb_mesh_get.argtypes = [C.c_void_p, C.POINTER(MeshView)]
lib.mfb_mesh_free.argtypes = [C.c_void_p]
lib.mfb_mesh_free.restype = None
lib.mfb_ctx_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(C.c_void_p)]
lib.mfb_ctx_destroy.argtypes = [C.c_void_p]
lib.mfb_ctx_destroy.restype = None
for _name in ("assembly", "assembly_fused", "prec_init", "halo_exchange", "prec_inversion", "iteration", "sync",
              "zero_values"):
    getattr(lib, "mfb_ctx_" + _name).argtypes = [C.c_void_p]
lib.mfb_ctx_assembly_interval.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.mfb_ctx_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.mfb_ctx_upload_coord.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_iteration_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.mfb_ctx_norms.argtypes = [C.c_void_p, _f64p, _f64p]
lib.mfb_ctx_iteration_norms_host.argtypes = [C.c_void_p, C.c_void_p, _f64p]
lib.mfb_ctx_device_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
lib.mfb_ctx_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
lib.mfb_ctx_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
lib.mfb_ctx_device_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
lib.mfb_ctx_plan_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
lib.mfb_comm_unique_id.argtypes = [C.c_void_p]
lib.mfb_ctx_comm_init.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_p2p_card.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_p2p_connect.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_p2p_enable.argtypes = [C.c_void_p, C.c_int]
lib.mfb_ctx_p2p_active.argtypes = [C.c_void_p]
lib.mfb_block_coloring.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_int)]
lib.mfb_ctx_halo_pack_host.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_halo_add_host.argtypes = [C.c_void_p, C.c_void_p]
lib.mfb_ctx_prec_inversion_interface.argtypes = [C.c_void_p]
lib.mfb_ctx_run_timed.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
lib.mfb_tile_plan_selfcheck.argtypes = [C.POINTER(Problem), C.c_int, C.c_int, C.POINTER(C.c_int64)]
lib.mfb_ring_plan_selfcheck.argtypes = [C.POINTER(Problem), C.c_int, C.c_int, C.POINTER(C.c_int64)]
lib.mfb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
lib.mfb_host_free.argtypes = [C.c_void_p]
lib.mfb_host_free.restype = None
lib.mfb_checking_write.argtypes = [C.c_char_p, C.c_double, C.c_double]
lib.mfb_checking_read.argtypes = [C.c_char_p, _f64p, _f64p]
lib.mfb_choose_blocks.argtypes = [C.c_int] * 4 + [_i32p] * 3
lib.mfb_choose_blocks.restype = None
lib.mfb_device_create_nodeToNode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                             _i32p, C.c_int]
lib.mfb_device_create_elemToEdge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
lib.mfb_device_coloring_creation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             _i32p, C.c_int]

# Every symbol include/minifem_b200.h declares (tests check that the library exports them).
DECLARED_SYMBOLS = [
    "mfb_last_error", "mfb_version", "mfb_node_to_elem", "mfb_count_edges",
    "mfb_create_nodeToNode", "mfb_create_elemToEdge", "mfb_coloring_creation",
    "mfb_permute_int_2d", "mfb_boundary_mask", "mfb_double_norm", "mfb_mesh_generate",
    "mfb_mesh_read", "mfb_mesh_write", "mfb_mesh_get", "mfb_mesh_free", "mfb_choose_blocks",
    "mfb_checking_write", "mfb_checking_read", "mfb_ctx_create", "mfb_ctx_destroy",
    "mfb_ctx_assembly", "mfb_ctx_assembly_interval", "mfb_ctx_zero_values", "mfb_ctx_prec_init",
    "mfb_ctx_halo_exchange", "mfb_ctx_prec_inversion", "mfb_ctx_iteration", "mfb_ctx_sync",
    "mfb_ctx_download", "mfb_ctx_upload_coord", "mfb_ctx_iteration_host", "mfb_ctx_device_ptrs",
    "mfb_ctx_stream", "mfb_ctx_stage_ms", "mfb_ctx_launch_count", "mfb_ctx_device_bytes",
    "mfb_ctx_plan_stats", "mfb_comm_unique_id", "mfb_ctx_comm_init", "mfb_ctx_halo_pack_host",
    "mfb_ctx_halo_add_host", "mfb_ctx_assembly_fused", "mfb_ctx_prec_inversion_interface", "mfb_ctx_run_timed",
    "mfb_tile_plan_selfcheck", "mfb_ring_plan_selfcheck", "mfb_host_alloc",
    "mfb_host_free", "mfb_device_count",
    "mfb_device_create_nodeToNode", "mfb_device_create_elemToEdge", "mfb_device_coloring_creation",
    "mfb_ctx_norms", "mfb_ctx_iteration_norms_host",
    "mfb_ctx_p2p_card", "mfb_ctx_p2p_connect", "mfb_ctx_p2p_enable", "mfb_ctx_p2p_active", "mfb_block_coloring",
]


def _check(rc, what=""):
    if rc != 0:
        raise MfbError(f"{what} failed ({rc}): {lib.mfb_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _as_np(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)


# ---------------------------------------------------------------- host-side builders

def node_to_elem(elemToNode, nbNodes):
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    index = np.zeros(nbNodes + 1, np.int32)
    value = np.zeros(max(nbElem * 4, 1), np.int32)
    _check(lib.mfb_node_to_elem(_ptr(e2n), nbElem, nbNodes, _ptr(index), _ptr(value)), "mfb_node_to_elem")
    return index, value[:nbElem * 4]


def create_nodeToNode(elemToNode, nbNodes):
    """(nodeToNodeRow, nodeToNodeColumn) as create_nodeToNode builds them (matrix.cc:55-91)."""
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    n = lib.mfb_count_edges(_ptr(e2n), nbElem, nbNodes)
    if n < 0:
        _check(int(n), "mfb_count_edges")
    row = np.zeros(nbNodes + 1, np.int32)
    col = np.zeros(max(n, 1), np.int32)
    out = C.c_int(0)
    _check(lib.mfb_create_nodeToNode(_ptr(e2n), nbElem, nbNodes, _ptr(row), _ptr(col), C.byref(out)),
           "mfb_create_nodeToNode")
    return row, col[:out.value]


def create_elemToEdge(row, col, elemToNode):
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    out = np.zeros(max(nbElem * 16, 1), np.int32)
    _check(lib.mfb_create_elemToEdge(_ptr(_i32(row)), _ptr(_i32(col)), _ptr(e2n), _ptr(out), nbElem),
           "mfb_create_elemToEdge")
    return out[:nbElem * 16]


def coloring_creation(elemToNode, nbNodes):
    """(colorPart, colorToElem, colorPerm, nbTotalColors) — coloring.cc:84-109."""
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    part = np.zeros(max(nbElem, 1), np.int32)
    c2e = np.zeros(129, np.int32)
    perm = np.zeros(max(nbElem, 1), np.int32)
    nb = C.c_int(0)
    _check(lib.mfb_coloring_creation(_ptr(e2n), nbElem, nbNodes, _ptr(part), _ptr(c2e), _ptr(perm),
                                     C.byref(nb)), "mfb_coloring_creation")
    return part[:nbElem], c2e[:nb.value + 1].copy(), perm[:nbElem], nb.value


# ---------------------------------------------------------------- the same builders on the GPU

def device_create_nodeToNode(elemToNode, nbNodes, device=0):
    """create_nodeToNode (matrix.cc:55-91) built on the GPU; bit-identical to create_nodeToNode()."""
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    row = np.zeros(nbNodes + 1, np.int32)
    col = np.zeros(max(nbElem * 12 + nbNodes, 1), np.int32)       # upper bound of the entry count
    out = C.c_int(0)
    _check(lib.mfb_device_create_nodeToNode(_ptr(e2n), nbElem, nbNodes, _ptr(row), _ptr(col), col.size,
                                            C.byref(out), device), "mfb_device_create_nodeToNode")
    return row, col[:out.value].copy()


def device_create_elemToEdge(row, col, elemToNode, device=0):
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    row = _i32(row)
    out = np.zeros(max(nbElem * 16, 1), np.int32)
    _check(lib.mfb_device_create_elemToEdge(_ptr(row), _ptr(_i32(col)), _ptr(e2n), _ptr(out), nbElem,
                                            row.size - 1, device), "mfb_device_create_elemToEdge")
    return out[:nbElem * 16]


def device_coloring_creation(elemToNode, nbNodes, device=0):
    """coloring_creation (coloring.cc:84-109) evaluated front by front on the GPU; same colours."""
    e2n = _i32(elemToNode).ravel()
    nbElem = e2n.size // 4
    part = np.zeros(max(nbElem, 1), np.int32)
    c2e = np.zeros(129, np.int32)
    perm = np.zeros(max(nbElem, 1), np.int32)
    nb = C.c_int(0)
    _check(lib.mfb_device_coloring_creation(_ptr(e2n), nbElem, nbNodes, _ptr(part), _ptr(c2e), _ptr(perm),
                                            C.byref(nb), device), "mfb_device_coloring_creation")
    return part[:nbElem], c2e[:nb.value + 1].copy(), perm[:nbElem], nb.value


def block_coloring(elemToNode, nbNodes, coord, block_elems=0):
    """mfb_block_coloring: (elemOrder, launchStart, localIndex, localStart, dict of counts)."""
    e2n = _i32(elemToNode)
    nbElem = e2n.size // 4
    xyz = np.ascontiguousarray(coord, np.float64)
    order = np.zeros(max(nbElem, 1), np.int32)
    launch = np.zeros(65, np.int32)
    index = np.zeros(nbElem + 2, np.int32)
    start = np.zeros(2 * nbElem + 2, np.int32)
    counts = (C.c_int * 4)()
    _check(lib.mfb_block_coloring(_ptr(e2n), nbElem, nbNodes, _ptr(xyz), block_elems, _ptr(order), _ptr(launch), _ptr(index),
                                  _ptr(start), counts), "mfb_block_coloring")
    blocks, colors, max_local, entries = [int(v) for v in counts]
    return (order[:nbElem], launch[:colors + 1], index[:blocks + 1], start[:entries],
            dict(blocks=blocks, block_colors=colors, max_local_colors=max_local))


def permute_int_2d(tab, perm, dim):
    t = _i32(tab).ravel().copy()
    _check(lib.mfb_permute_int_2d(_ptr(t), _ptr(_i32(perm)), t.size // dim, dim), "mfb_permute_int_2d")
    return t


def boundary_mask(boundNodesCode):
    codes = _i32(boundNodesCode)
    out = np.zeros(max(codes.size * 3, 1), np.int32)
    nb = C.c_int(0)
    _check(lib.mfb_boundary_mask(_ptr(codes), codes.size, _ptr(out), C.byref(nb)), "mfb_boundary_mask")
    return out[:codes.size * 3], nb.value


def double_norm(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib.mfb_double_norm(_ptr(a), a.size)


def choose_blocks(nx, ny, nz, max_ranks):
    px, py, pz = C.c_int(1), C.c_int(1), C.c_int(1)
    lib.mfb_choose_blocks(nx, ny, nz, max_ranks, C.byref(px), C.byref(py), C.byref(pz))
    return px.value, py.value, pz.value


class Mesh:
    """One subdomain's input arrays as numpy copies (IO.cc:61-96 layout)."""

    FIELDS = ("coord", "elemToNode", "neighborsList", "intfIndex", "intfNodes", "boundNodesCode")

    def __init__(self, handle):
        v = MeshView()
        _check(lib.mfb_mesh_get(handle, C.byref(v)), "mfb_mesh_get")
        self.nbElem, self.nbNodes, self.nbEdges = v.nbElem, v.nbNodes, v.nbEdges
        self.nbIntf, self.nbIntfNodes, self.nbBoundNodes = v.nbIntf, v.nbIntfNodes, v.nbBoundNodes
        self.coord = _as_np(v.coord, v.nbNodes * 3, np.float64)
        self.elemToNode = _as_np(v.elemToNode, v.nbElem * 4, np.int32)
        self.neighborsList = _as_np(v.neighborsList, max(v.nbIntf, 1) * 3, np.int32)
        self.intfIndex = _as_np(v.intfIndex, v.nbIntf + 1, np.int32)
        self.intfNodes = _as_np(v.intfNodes, v.nbIntfNodes, np.int32)
        self.boundNodesCode = _as_np(v.boundNodesCode, v.nbNodes, np.int32)
        self.globalNode = _as_np(v.globalNode, v.nbNodes, np.int64) if v.globalNode else None

    @classmethod
    def generate(cls, nx, ny, nz, blocks=(1, 1, 1), rank=0, seed=1):
        h = C.c_void_p()
        _check(lib.mfb_mesh_generate(nx, ny, nz, blocks[0], blocks[1], blocks[2], rank, seed, C.byref(h)),
               "mfb_mesh_generate")
        try:
            return cls(h)
        finally:
            lib.mfb_mesh_free(h)

    @classmethod
    def read(cls, path):
        h = C.c_void_p()
        _check(lib.mfb_mesh_read(path.encode(), C.byref(h)), "mfb_mesh_read")
        try:
            return cls(h)
        finally:
            lib.mfb_mesh_free(h)

    @staticmethod
    def generate_to_file(path, nx, ny, nz, blocks=(1, 1, 1), rank=0, seed=1):
        h = C.c_void_p()
        _check(lib.mfb_mesh_generate(nx, ny, nz, blocks[0], blocks[1], blocks[2], rank, seed, C.byref(h)),
               "mfb_mesh_generate")
        try:
            _check(lib.mfb_mesh_write(h, path.encode()), "mfb_mesh_write")
        finally:
            lib.mfb_mesh_free(h)


class Setup:
    """What main.cc builds between read_input_data and FEM_loop (main.cc:209-351):
    optional colouring + permutation, CSR, optional elemToEdge, Dirichlet mask."""

    def __init__(self, mesh, operator="ela", coloring=False, elem_to_edge=False, builder="host", device=0):
        self.mesh = mesh
        self.operatorID = {"lap": 0, "ela": 1}[operator]
        self.operatorDim = 1 if self.operatorID == 0 else 9
        self.elemToNode = mesh.elemToNode.copy()
        self.colorToElem, self.nbTotalColors, self.colorPerm = None, 0, None
        if builder not in ("host", "gpu"):
            raise ValueError("builder must be 'host' or 'gpu'")
        gpu = builder == "gpu"           # same layouts, built by the kernels of csrc/kernels_topology.cu
        if coloring:
            color = (lambda e, n: device_coloring_creation(e, n, device)) if gpu else coloring_creation
            _, self.colorToElem, self.colorPerm, self.nbTotalColors = color(self.elemToNode, mesh.nbNodes)
            self.elemToNode = permute_int_2d(self.elemToNode, self.colorPerm, 4)
        if gpu:
            self.row, self.col = device_create_nodeToNode(self.elemToNode, mesh.nbNodes, device)
        else:
            self.row, self.col = create_nodeToNode(self.elemToNode, mesh.nbNodes)
        self.nbEdges = int(self.row[-1])
        self.elemToEdge = None
        if elem_to_edge:
            self.elemToEdge = (device_create_elemToEdge(self.row, self.col, self.elemToNode, device) if gpu
                               else create_elemToEdge(self.row, self.col, self.elemToNode))
        self.checkBounds, _ = boundary_mask(mesh.boundNodesCode)


class Context:
    """GPU context over one subdomain (mfb_ctx_*)."""

    def __init__(self, setup, path="tiled", device=0, nbBlocks=1, rank=0, tile_rows=0, tile_elems=0,
                 threads=0, use_graph=False, ctas=0, bank_aware=True):
        m = setup.mesh
        self.setup = setup
        self._keep = dict(coord=np.ascontiguousarray(m.coord), e2n=_i32(setup.elemToNode), row=_i32(setup.row),
                          col=_i32(setup.col), e2e=None if setup.elemToEdge is None else _i32(setup.elemToEdge),
                          cb=_i32(setup.checkBounds),
                          c2e=None if setup.colorToElem is None else _i32(setup.colorToElem),
                          ii=_i32(m.intfIndex), inn=_i32(m.intfNodes), nl=_i32(m.neighborsList))
        k = self._keep
        p = Problem(setup.operatorID, m.nbElem, m.nbNodes, setup.nbEdges, _ptr(k["coord"]), _ptr(k["e2n"]),
                    _ptr(k["row"]), _ptr(k["col"]), _ptr(k["e2e"]), _ptr(k["cb"]), _ptr(k["c2e"]),
                    setup.nbTotalColors, nbBlocks, rank, m.nbIntf, m.nbIntfNodes, _ptr(k["ii"]),
                    _ptr(k["inn"]), _ptr(k["nl"]))
        o = Options(PATH_NAMES[path] if isinstance(path, str) else path, device, tile_rows, tile_elems,
                    threads, int(use_graph), ctas, 1 if bank_aware else -1)
        self.handle = C.c_void_p()
        self.path = o.path
        _check(lib.mfb_ctx_create(C.byref(p), C.byref(o), C.byref(self.handle)), "mfb_ctx_create")
        self.nbValues = setup.nbEdges * setup.operatorDim
        self.nbPrec = m.nbNodes * setup.operatorDim

    def close(self):
        if self.handle:
            lib.mfb_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def assembly(self): _check(lib.mfb_ctx_assembly(self.handle), "mfb_ctx_assembly")
    def assembly_fused(self): _check(lib.mfb_ctx_assembly_fused(self.handle), "mfb_ctx_assembly_fused")
    def zero_values(self): _check(lib.mfb_ctx_zero_values(self.handle), "mfb_ctx_zero_values")
    def assembly_interval(self, first, last): _check(lib.mfb_ctx_assembly_interval(self.handle, first, last), "mfb_ctx_assembly_interval")
    def prec_init(self): _check(lib.mfb_ctx_prec_init(self.handle), "mfb_ctx_prec_init")
    def halo_exchange(self): _check(lib.mfb_ctx_halo_exchange(self.handle), "mfb_ctx_halo_exchange")
    def prec_inversion(self): _check(lib.mfb_ctx_prec_inversion(self.handle), "mfb_ctx_prec_inversion")
    def iteration(self): _check(lib.mfb_ctx_iteration(self.handle), "mfb_ctx_iteration")
    def sync(self): _check(lib.mfb_ctx_sync(self.handle), "mfb_ctx_sync")

    def stages(self):
        """The four stage calls of FEM_loop in order (FEM.cc:183-233)."""
        self.assembly(); self.prec_init(); self.halo_exchange(); self.prec_inversion()

    def download(self, values=True, prec=True):
        v = np.empty(max(self.nbValues, 1), np.float64) if values else None
        p = np.empty(max(self.nbPrec, 1), np.float64) if prec else None
        _check(lib.mfb_ctx_download(self.handle, _ptr(v), _ptr(p)), "mfb_ctx_download")
        return (None if v is None else v[:self.nbValues]), (None if p is None else p[:self.nbPrec])

    def norms(self):
        """(matrix norm, prec norm) as check_results computes them (FEM.cc:68-76), on the device."""
        a, b = C.c_double(0), C.c_double(0)
        _check(lib.mfb_ctx_norms(self.handle, C.byref(a), C.byref(b)), "mfb_ctx_norms")
        return a.value, b.value

    def iteration_norms_host(self, coord_ptr):
        out = (C.c_double * 2)()
        _check(lib.mfb_ctx_iteration_norms_host(self.handle, coord_ptr, out), "mfb_ctx_iteration_norms_host")
        return out[0], out[1]

    def iteration_host(self, coord_ptr, values_ptr, prec_ptr):
        _check(lib.mfb_ctx_iteration_host(self.handle, coord_ptr, values_ptr, prec_ptr), "mfb_ctx_iteration_host")

    def halo_pack_host(self):
        buf = np.zeros(max(self.setup.mesh.nbIntfNodes * self.setup.operatorDim, 1), np.float64)
        _check(lib.mfb_ctx_halo_pack_host(self.handle, _ptr(buf)), "mfb_ctx_halo_pack_host")
        return buf[:self.setup.mesh.nbIntfNodes * self.setup.operatorDim]

    def halo_add_host(self, recvbuf):
        buf = np.ascontiguousarray(recvbuf, dtype=np.float64)
        if buf.size == 0:
            buf = np.zeros(1)
        _check(lib.mfb_ctx_halo_add_host(self.handle, _ptr(buf)), "mfb_ctx_halo_add_host")

    def prec_inversion_interface(self):
        _check(lib.mfb_ctx_prec_inversion_interface(self.handle), "mfb_ctx_prec_inversion_interface")

    def run_timed(self, steps):
        ms = C.c_float(0)
        _check(lib.mfb_ctx_run_timed(self.handle, steps, C.byref(ms)), "mfb_ctx_run_timed")
        return ms.value

    def stage_ms(self):
        ms = (C.c_float * 5)()
        _check(lib.mfb_ctx_stage_ms(self.handle, ms), "mfb_ctx_stage_ms")
        return list(ms)

    def launch_count(self):
        return int(lib.mfb_ctx_launch_count(self.handle))

    def stream(self):
        s = C.c_void_p()
        _check(lib.mfb_ctx_stream(self.handle, C.byref(s)), "mfb_ctx_stream")
        return s.value or 0

    def device_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _check(lib.mfb_ctx_device_bytes(self.handle, C.byref(a), C.byref(b)), "mfb_ctx_device_bytes")
        return a.value, b.value

    def plan_stats(self):
        s = (C.c_int64 * 8)()
        _check(lib.mfb_ctx_plan_stats(self.handle, s), "mfb_ctx_plan_stats")
        if self.path == PATH_BLOCKCOLOR:
            return dict(blocks=s[0], block_colors=s[1], max_local_colors=s[2])
        if self.path == PATH_RING:
            return dict(tiles=s[0], jobs=s[1], ring_steps=s[2], max_rows=s[3], max_nodes=s[4], smem_bytes=s[5],
                        padded_lane_steps=s[6], max_blob_bytes=s[7])
        return dict(tiles=s[0], tile_elems=s[1], contributions=s[2], max_rows=s[3], max_elems=s[4], smem_bytes=s[5],
                    padded_lane_steps=s[6], max_blob_bytes=s[7])

    def comm_init(self, unique_id: bytes):
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(unique_id)
        _check(lib.mfb_ctx_comm_init(self.handle, buf), "mfb_ctx_comm_init")

    # peer-to-peer exchange of the fused RING iteration (include/minifem_b200.h: mfb_ctx_p2p_*)
    def p2p_card(self) -> bytes:
        buf = (C.c_ubyte * P2P_CARD_BYTES)()
        _check(lib.mfb_ctx_p2p_card(self.handle, buf), "mfb_ctx_p2p_card")
        return bytes(buf)

    def p2p_connect(self, cards):
        """cards: the p2p_card() of every subdomain, in rank order."""
        blob = b"".join(cards)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _check(lib.mfb_ctx_p2p_connect(self.handle, buf), "mfb_ctx_p2p_connect")

    def p2p_enable(self, on=True): _check(lib.mfb_ctx_p2p_enable(self.handle, int(bool(on))), "mfb_ctx_p2p_enable")
    def p2p_active(self): return bool(lib.mfb_ctx_p2p_active(self.handle))


def comm_unique_id() -> bytes:
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(lib.mfb_comm_unique_id(buf), "mfb_comm_unique_id")
    return bytes(buf)


def device_count():
    return int(lib.mfb_device_count())


class PinnedArray:
    """numpy view over cudaMallocHost memory (for the *_host calls)."""

    def __init__(self, count, dtype=np.float64):
        self.ptr = C.c_void_p()
        self.bytes = int(count) * np.dtype(dtype).itemsize
        _check(lib.mfb_host_alloc(C.byref(self.ptr), self.bytes), "mfb_host_alloc")
        buf = (C.c_char * max(self.bytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(count))

    def free(self):
        if self.ptr:
            self.array = None
            lib.mfb_host_free(self.ptr)
            self.ptr = C.c_void_p()
