"""Dry run of bench.py main() (--path auto) on a CPU box: CUDA contexts replaced by the host-thread RING
emulation, pinned memory by numpy, torch.cuda calls by no-ops; the probe subprocess is replaced by an
in-process verdict.  Only checks that the Python of the measured path and of the JSON line holds together."""
import sys, json, types, argparse
import os
__file__ = os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])), "ring_gpu_worker_dryrun.py")
src = open(__file__).read().replace("w.main()", "")
exec(src)
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
import bench
class Ctx(FakeCtx):
    def __init__(self, setup, path="tiled", device=0, nbBlocks=1, rank=0, tile_rows=0, tile_elems=0, use_graph=False, **kw):
        super().__init__(setup, path=path, nbBlocks=nbBlocks, rank=rank, tile_rows=tile_rows, tile_elems=tile_elems)
        self.path = path
        self.nbValues = setup.nbEdges * setup.operatorDim; self.nbPrec = setup.mesh.nbNodes * setup.operatorDim
    def run_timed(self, steps): self.n += steps; return 0.5 * steps
    def iteration_host(self, a, b, c): self.iteration()
    def iteration_norms_host(self, a): self.iteration(); return (1.0, 2.0)
    def plan_stats(self): return {"tiles": 1}
    def device_bytes(self): return (1, 2)
class Pin:
    def __init__(self, count): self.array = np.zeros(int(count)); self.bytes = self.array.nbytes; self.ptr = 0
    def free(self): pass
mfb.Context = Ctx; mfb.PinnedArray = Pin; mfb.device_count = lambda: 1
torch.cuda.set_device = lambda d: None; torch.cuda.synchronize = lambda: None
bench.choose_path = lambda args, rank, local: {"probe": {"ok": True, "ring_ms": 0.3, "tiled_ms": 0.6}, "chosen": "ring"}
sys.argv = ["bench.py", "--grid", "5", "4", "3", "--steps", "3", "--warmup", "3", "--e2e-steps", "1", "--no-cpu-baseline", "--no-other-paths"]
bench.main()
