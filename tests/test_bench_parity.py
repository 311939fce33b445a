"""Host logic of bench.py without a GPU: the `parity` record compares what a context holds with the CPU
oracle on the rank's own subdomain and, with several ranks, moves the oracle's interface values between the
ranks in the message pattern of halo.cc — checked here with two gloo ranks and a stand-in context that
returns the reference's multi-subdomain result (and a corrupted copy of it)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import bench
import minifem_b200 as mfb
from oracle_lib import Oracle, Reference, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import minifem_b200 as mfb
from minifem_b200 import dist as mdist
from oracle_lib import Oracle
rank, world = mdist.init_from_env("gloo")
grid, blocks = (8, 6, 5), mfb.choose_blocks(8, 6, 5, world)
meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=3) for r in range(world)]
oracle = Oracle()
for op in ("ela", "lap"):
    dim = 9 if op == "ela" else 1
    setups = [mfb.Setup(m, op) for m in meshes]
    res = [oracle.fem_iteration(s) for s in setups]
    precs = [np.array(r[1], copy=True) for r in res]              # halo_exchange works in place
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes], [m.neighborsList for m in meshes], dim)
    s = setups[rank]
    want_p = oracle.prec_inversion(precs[rank], s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID)
    class Ctx:
        def __init__(self, v, p): self.v, self.p = v, p
        def download(self): return self.v, self.p
    good = bench.oracle_parity(Ctx(res[rank][0], want_p), s, meshes[rank], world)
    assert good["ok"] and good["values_err"] == 0.0 and good["prec_err"] == 0.0 and good["ranks"] == world, good
    bad_p = want_p.copy()
    if rank == 1:                                   # one interface block of one rank is off by 1e-9
        node = int(meshes[rank].intfNodes[0]) - 1
        bad_p[node * dim] *= 1.0 + 1e-9
    bad = bench.oracle_parity(Ctx(res[rank][0], bad_p), s, meshes[rank], world)
    assert not bad["ok"] and bad["prec_err"] > 1e-12, bad   # seen by every rank: the record is the max over ranks
    # without the exchange the interface blocks would differ: the harness really moves them
    alone = oracle.prec_inversion(np.ascontiguousarray(res[rank][1]), s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID)
    assert not bench.oracle_parity(Ctx(res[rank][0], alone), s, meshes[rank], world)["ok"]
# bench.warm_up: a peer-to-peer exchange that times out on ONE rank is switched off on every rank and the warm-up repeated
class TimeoutCtx:
    def __init__(self, fails): self.fails, self.p2p, self.iterations = fails, True, 0
    def iteration(self): self.iterations += 1
    def sync(self):
        if self.fails and self.p2p: raise mfb.MfbError("peer-to-peer halo exchange timed out waiting for a neighbour's flag")
    def p2p_enable(self, on): self.p2p = on
tc, halo = TimeoutCtx(rank == 1), {"transport": "p2p"}
bench.warm_up(mfb, mdist, tc, 3, halo)
assert halo["transport"] == "nccl" and tc.p2p is False and tc.iterations == 6 and "timed out" in halo["why"], (rank, halo)
tc, halo = TimeoutCtx(False), {"transport": "p2p"}
bench.warm_up(mfb, mdist, tc, 3, halo)
assert halo == {"transport": "p2p"} and tc.iterations == 3
print("PARITY_WORKER_OK", rank, flush=True)
'''


def test_parity_record_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = 29600 + os.getpid() % 300
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.count("PARITY_WORKER_OK") == 2, res.stdout[-3000:]


def test_parity_record_single_rank():
    mesh = mfb.Mesh.generate(6, 5, 4, seed=2)
    setup = mfb.Setup(mesh, "ela")
    v, _, p = Oracle().fem_iteration(setup)

    class Ctx:
        def download(self):
            return v, p
    rec = bench.oracle_parity(Ctx(), setup, mesh, 1)
    assert rec["ok"] and rec["values_err"] == 0.0 and rec["rtol"] == 1e-12


@pytest.mark.skipif(not ref_available("ref"), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_arm_builds_its_own_layouts():
    """The reference arm's CSR, elemToEdge and Dirichlet mask come from the reference's compiled functions and
    equal the ones this repository builds (so both arms time the same matrix)."""
    mesh = mfb.Mesh.generate(7, 6, 5, seed=4)
    ours = mfb.Setup(mesh, "ela", elem_to_edge=True)
    theirs = bench.RefSetup(Reference("ref"), mesh, "ela", elem_to_edge=True)
    assert np.array_equal(ours.row, theirs.row) and np.array_equal(ours.col, theirs.col)
    assert np.array_equal(ours.elemToEdge, theirs.elemToEdge) and np.array_equal(ours.checkBounds, theirs.checkBounds)


def test_algorithmic_bytes_match_survey():
    E, N, Z = 6000000, 1030301, 15210901
    assert bench.algorithmic_bytes("ela", E, Z, N) == 16 * E + 76 * Z + 112 * N == 1367422188
    assert bench.algorithmic_bytes("lap", E, Z, N) == 16 * E + 12 * Z + 36 * N


def test_time_paths_records_every_path_and_survives_a_failing_one():
    """bench.time_paths (the `other_paths` / `configs` side measurements) with stand-in contexts: the blocked colouring
    reports its launch structure, and a path that fails is recorded instead of taking the bench line with it."""
    class Setup:
        def __init__(self, mesh, op, coloring=False): self.nbEdges, self.nbTotalColors = 100, 7
    class Ctx:
        def __init__(self, setup, path, device=0, use_graph=False):
            if path == "tiled": raise mfb.MfbError("tile plan needs more than 227 KB")
            if path == "atomic": raise KeyError("anything else")
            self.path, self.graph = path, use_graph
        def iteration(self): pass
        def sync(self): pass
        def run_timed(self, steps): return 2.0 * steps
        def plan_stats(self): return dict(blocks=5, block_colors=3, max_local_colors=9)
        def close(self): pass
    class Fake:
        MfbError = mfb.MfbError
    Fake.Setup, Fake.Context = Setup, Ctx
    class Mesh: nbElem, nbNodes = 1000, 300
    out = bench.time_paths(Fake, Mesh, "ela", ["ring", "tiled", "atomic", "color", "blockcolor"], 0, 5, peak=1000.0)
    assert out["ring"]["ms_per_step"] == 2.0 and out["ring"]["value"] == 1000 / 2e-3 and "colors" not in out["ring"]
    assert "227 KB" in out["tiled"]["error"] and "KeyError" in out["atomic"]["error"]
    assert out["color"]["colors"] == 7
    assert out["blockcolor"]["block_colors"] == 3 and out["blockcolor"]["blocks"] == 5 and out["blockcolor"]["ms_per_step"] == 2.0
