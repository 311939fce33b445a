"""Several subdomains = several contexts.  On one GPU the interface values travel through
the host halves of the exchange (mfb_ctx_halo_pack_host / _add_host); with >= 2 GPUs the
NCCL path is run under torch.distributed.run (tests/nccl_worker.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import RTOL, block_scaled_error, row_scaled_error
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_exchange(ctxs, meshes, dim):
    send = [c.halo_pack_host() for c in ctxs]
    for r, (c, m) in enumerate(zip(ctxs, meshes)):
        recv = np.zeros_like(send[r])
        for i in range(m.nbIntf):
            s = int(m.neighborsList[i]) - 1
            o = meshes[s]
            q = [k for k in range(o.nbIntf) if o.neighborsList[k] - 1 == r][0]
            recv[m.intfIndex[i] * dim:m.intfIndex[i + 1] * dim] = send[s][o.intfIndex[q] * dim:o.intfIndex[q + 1] * dim]
        c.halo_add_host(recv)


@pytest.mark.parametrize("path", ["tiled", "atomic"])
@pytest.mark.parametrize("op", ["lap", "ela"])
def test_blocks_fixture(path, op):
    g = np.load(os.path.join(GOLDEN, "blocks_2x2x1_of_4x4x3.npz"))
    grid, blocks = tuple(g["grid"]), tuple(g["blocks"])
    n = int(np.prod(blocks))
    dim = 1 if op == "lap" else 9
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=11) for r in range(n)]
    setups = [mfb.Setup(m, op) for m in meshes]
    ctxs = [mfb.Context(s, path=path, nbBlocks=n, rank=r) for r, s in enumerate(setups)]
    for c in ctxs:
        c.assembly(); c.prec_init()
    host_exchange(ctxs, meshes, dim)
    for r, c in enumerate(ctxs):
        c.prec_inversion()
        v, p = c.download()
        assert row_scaled_error(v, g[f"r{r}_{op}_values"], setups[r].row, dim) <= RTOL
        assert block_scaled_error(p, g[f"r{r}_{op}_prec"], dim) <= RTOL
        with pytest.raises(mfb.MfbError, match="comm_init"):
            c.halo_exchange()                      # NCCL path refuses to run without a communicator
        c.close()


def test_fused_interface_split_matches_oracle():
    """TILED fused kernel with nbBlocks > 1: interface rows keep the raw diagonal block until
    the halo sum, every other row is inverted in the kernel."""
    grid, blocks = (9, 8, 7), (2, 2, 1)
    oracle = Oracle()
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=5) for r in range(4)]
    setups = [mfb.Setup(m, "ela") for m in meshes]
    precs = [np.ascontiguousarray(oracle.fem_iteration(s)[1]) for s in setups]
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes],
                         [m.neighborsList for m in meshes], 9)
    want = [oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, 1) for r, s in enumerate(setups)]
    ctxs = [mfb.Context(s, path="tiled", nbBlocks=4, rank=r, tile_rows=16, tile_elems=300) for r, s in enumerate(setups)]
    for c in ctxs:
        assert c.plan_stats()["tiles"] > 4
        c.assembly_fused()                         # one launch: values + prec (interior inverted)
    for r, c in enumerate(ctxs):                   # before the exchange: interface rows are raw
        _, p = c.download(values=False)
        intf = np.unique(meshes[r].intfNodes - 1)
        raw = oracle.fem_iteration(setups[r])[1].reshape(-1, 9)
        assert block_scaled_error(p.reshape(-1, 9)[intf], raw[intf], 9) <= RTOL
    host_exchange(ctxs, meshes, 9)
    for r, c in enumerate(ctxs):
        c.prec_inversion_interface()
        v, p = c.download()
        assert block_scaled_error(p, want[r], 9) <= RTOL
        assert row_scaled_error(v, oracle.fem_iteration(setups[r])[0], setups[r].row, 9) <= RTOL
        c.close()
    atomic = mfb.Context(setups[0], path="atomic", nbBlocks=4, rank=0)
    with pytest.raises(mfb.MfbError, match="TILED"):
        atomic.assembly_fused()
    atomic.close()


@pytest.mark.parametrize("nranks", [2])
def test_nccl_halo_under_torchrun(nranks, tmp_path):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.count("NCCL_WORKER_OK") == nranks, res.stdout[-3000:]
