"""Dry run of tests/ring_gpu_worker.py without a GPU: Context is replaced by the host-thread emulation of
the RING kernel (and by the oracle-backed TILED stand-in), big meshes are shrunk.  Catches Python-level
mistakes in the worker, nothing else."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), "tests")); sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), "mini-fem_b200", "python"))
import ring_gpu_worker as w
mfb = w.mfb
klib = C.CDLL(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "libmfb_ringkernel_host.so"))
klib.mfb_ring_kernel_host.argtypes = [C.c_int]*3 + [C.c_void_p]*6 + [C.c_int]*4 + [C.c_void_p]*2 + [C.c_int]
def p(a): return None if a is None else a.ctypes.data_as(C.c_void_p)
class FakeCtx:
    def __init__(self, setup, path="tiled", nbBlocks=1, rank=0, tile_rows=0, tile_elems=0, ctas=0, bank_aware=True, **kw):
        self.s, self.rows, self.entries, self.ctas, self.nb = setup, tile_rows, tile_elems, ctas, nbBlocks
        self.intf = None
        m = setup.mesh
        if nbBlocks > 1:
            self.intf = np.zeros(m.nbNodes, np.uint8); self.intf[m.intfNodes - 1] = 1
        self.v = self.p = None; self.n = 0
    def _run(self, fuse):
        s, m = self.s, self.s.mesh; dim = s.operatorDim
        v = np.full(s.nbEdges*dim, np.nan); pr = np.full(m.nbNodes*dim, np.nan)
        keep = [np.ascontiguousarray(s.elemToNode, np.int32), np.ascontiguousarray(s.row, np.int32), np.ascontiguousarray(s.col, np.int32),
                np.ascontiguousarray(m.coord, np.float64), np.ascontiguousarray(s.checkBounds, np.int32)]
        rc = klib.mfb_ring_kernel_host(s.operatorID, m.nbNodes, keep[0].size//4, *[p(k) for k in keep], p(self.intf), self.rows, self.entries,
                                       3 if self.ctas <= 0 else min(self.ctas, 4), fuse, p(v), p(pr), 256)
        assert rc == 0
        self.v, self.n = v, self.n + 1
        if fuse: self.p = pr
    def iteration(self): self._run(1)
    def assembly_fused(self): self._run(1)
    def assembly(self): self._run(0)
    def prec_init(self): self.p = w.oracle_.prec_init(self.v, self.s.row, self.s.col, self.s.mesh.nbNodes, self.s.operatorDim)
    def halo_exchange(self): pass
    def prec_inversion(self): self.p = w.oracle_.prec_inversion(self.p, self.s.row, self.s.col, self.s.checkBounds, self.s.mesh.nbNodes, self.s.operatorID)
    def download(self, values=True, prec=True): return self.v.copy(), self.p.copy()
    def launch_count(self): return self.n
    def run_timed(self, steps): return 1.0
    def close(self): pass
    def sync(self): pass
    def halo_pack_host(self):
        m = self.s.mesh; dim = self.s.operatorDim
        return np.ascontiguousarray(self.p.reshape(-1, dim)[m.intfNodes - 1]).ravel()
    def halo_add_host(self, recv):
        m = self.s.mesh; dim = self.s.operatorDim
        P = self.p.reshape(-1, dim); R = recv.reshape(-1, dim)
        for j, n in enumerate(m.intfNodes - 1): P[n] += R[j]
    def prec_inversion_interface(self):
        m = self.s.mesh; dim = self.s.operatorDim
        full = w.oracle_.prec_inversion(self.p.copy(), self.s.row, self.s.col, self.s.checkBounds, m.nbNodes, self.s.operatorID).reshape(-1, dim)
        idx = np.unique(m.intfNodes - 1)
        self.p.reshape(-1, dim)[idx] = full[idx]
w.oracle_ = w.Oracle()
w.mfb.Context = FakeCtx
w.mfb.device_count = lambda: 1
gen = w.mfb.Mesh.generate
def small(*grid, **kw):
    if grid[0] * grid[1] * grid[2] > 3000: grid = (7, 6, 5)
    return gen(*grid, **kw)
w.mfb.Mesh.generate = staticmethod(small)
w.main()
