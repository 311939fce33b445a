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


@pytest.mark.parametrize("op,grid,blocks,threads", [("ela", (12, 10, 8), (2, 1, 1), 0), ("lap", (9, 8, 7), (2, 2, 1), 0),
                                                     ("ela", (8, 8, 8), (2, 2, 1), 384), ("ela", (9, 8, 7), (2, 2, 1), 1024)])
def test_p2p_exchange_between_contexts_of_one_process(op, grid, blocks, threads):
    """The peer-to-peer halo (csrc/kernels_halo_p2p.cu) with every subdomain's context in this process, on one GPU:
    windows are connected through plain device pointers, every iteration is two launches per subdomain (exchange
    kernel + RING assembly kernel over all tiles) and the kernels of the subdomains find each other through the
    epoch flags.  Several iterations in a row exercise both receive-buffer parities; results against the oracle's
    multi-subdomain FEM iteration (faces, edges and corners shared by up to 8 subdomains)."""
    n = int(np.prod(blocks))
    dim = 1 if op == "lap" else 9
    oracle = Oracle()
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=5) for r in range(n)]
    setups = [mfb.Setup(m, op) for m in meshes]
    results = [oracle.fem_iteration(s) for s in setups]
    precs = [np.ascontiguousarray(res[1]) for res in results]
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes],
                         [m.neighborsList for m in meshes], dim)
    want = [oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID) for r, s in enumerate(setups)]
    ctxs = [mfb.Context(s, path="ring", nbBlocks=n, rank=r, tile_rows=12, tile_elems=240, threads=threads) for r, s in enumerate(setups)]
    with pytest.raises(mfb.MfbError, match="comm_init or mfb_ctx_p2p_connect"):
        ctxs[0].iteration()
    cards = [c.p2p_card() for c in ctxs]
    for c in ctxs:
        c.p2p_connect(cards)
        assert c.p2p_active()
    for it in range(3):
        before = [c.launch_count() for c in ctxs]
        for c in ctxs:
            c.iteration()                              # asynchronous: the kernels of all subdomains are in flight together
        for r, c in enumerate(ctxs):
            c.sync()
            assert c.launch_count() - before[r] == 2
            v, p = c.download()
            assert row_scaled_error(v, results[r][0], setups[r].row, dim) <= RTOL, (it, r)
            assert block_scaled_error(p, want[r], dim) <= RTOL, (it, r)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("blocks,op", [((2, 2, 2), "ela"), ((3, 1, 1), "lap")])
def test_p2p_all_subdomains_of_a_partition_in_one_process(blocks, op):
    """2 x 2 x 2: every subdomain has 7 neighbours (faces, edges and the corner node shared by all 8); 3 x 1 x 1: the
    middle subdomain has two.  Also the failure path: a subdomain that skips an exchange makes its neighbours report
    a timeout through mfb_ctx_sync instead of hanging (tests/p2p_worker.py)."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "p2p_worker.py")] + [str(b) for b in blocks] + [op]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "P2P_WORKER_OK" in res.stdout, res.stdout[-3000:]


def test_p2p_card_checks():
    meshes = [mfb.Mesh.generate(6, 5, 4, blocks=(2, 1, 1), rank=r, seed=5) for r in range(2)]
    setups = [mfb.Setup(m, "ela") for m in meshes]
    tiled = mfb.Context(setups[0], path="tiled", nbBlocks=2, rank=0)
    with pytest.raises(mfb.MfbError, match="RING"):
        tiled.p2p_card()
    tiled.close()
    ring = mfb.Context(setups[0], path="ring", nbBlocks=2, rank=0)
    with pytest.raises(mfb.MfbError, match="p2p_card first"):
        ring.p2p_connect([bytes(mfb.P2P_CARD_BYTES)] * 2)
    card = ring.p2p_card()
    with pytest.raises(mfb.MfbError, match="published no card"):
        ring.p2p_connect([card, bytes(mfb.P2P_CARD_BYTES)])
    assert not ring.p2p_active()
    ring.close()


@pytest.mark.parametrize("nranks", [2])
def test_nccl_halo_under_torchrun(nranks, tmp_path):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_worker.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.count("NCCL_WORKER_OK") == nranks, res.stdout[-3000:]


def test_driver_two_ranks(tmp_path):
    """minifem_b200 as two processes (RANK / WORLD_SIZE like mpirun -np 2): per-rank input files,
    NCCL id and timer reduction through the rendezvous directory, numerical check per rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    grid, blocks = (12, 8, 8), (2, 1, 1)
    data, rdv = str(tmp_path / "data"), str(tmp_path / "rdv")
    os.makedirs(rdv)
    oracle = Oracle()
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=6) for r in range(2)]
    for op in ("ela",):
        setups = [mfb.Setup(m, op) for m in meshes]
        results = [oracle.fem_iteration(s) for s in setups]
        precs = [np.ascontiguousarray(res[1]) for res in results]
        oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes],
                             [m.neighborsList for m in meshes], 9)
        for r, s in enumerate(setups):
            mfb.Mesh.generate_to_file(os.path.join(data, "EIB", "inputs", f"{op}_2_{r}"), *grid, blocks=blocks, rank=r, seed=6)
            p = oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, 1)
            path = os.path.join(data, "EIB", "checkings", f"{op}_2_{r}").encode()
            assert mfb.lib.mfb_checking_write(path, oracle.norm(results[r][0]), oracle.norm(p)) == 0
        exe = os.path.join(ROOT, "mini-fem_b200", "minifem_b200")
        procs = []
        for r in range(2):
            env = dict(os.environ, MINIFEM_DATA_PATH=data, MINIFEM_RENDEZVOUS=rdv, RANK=str(r), WORLD_SIZE="2",
                       LOCAL_RANK=str(r), MINIFEM_FUSED=str(r % 2 * 0 + 1))
            procs.append(subprocess.Popen([exe, "EIB", op, "3"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=300)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        assert "Average cycles" in outs[0] and "Average cycles" not in outs[1]          # rank 0 prints (FEM.cc:125)
        for r in range(2):
            report = open(tmp_path / f"numerical_results_{r}").read()
            diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
            assert len(diffs) == 2 and max(diffs) < 1e-13, report
