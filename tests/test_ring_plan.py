"""RING path without a GPU: the plan's structural self-check through the C ABI, and a host replay of
the plan (tools/ring_replay.cc — lane by lane, in the kernel's order, with the arithmetic header
the kernel compiles) against the oracle.  This pins the plan format, the edge-ring formulation
(gradients rebuilt from coordinates, one node per element), the transposed block of interior
edges and the row-sum diagonal to the reference's numbers; the kernel itself is covered by the
`-m gpu` tests (test_gpu_parity.py, path "ring")."""
import ctypes as C
import os

import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import (RTOL, ArrayMesh, assert_close_or_conditioned, assert_prec_close_or_conditioned, block_scaled_error,
                     diag_conditioning, extended_truth, extended_truth_prec, random_tet_mesh, row_scaled_error)
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The RING path evaluates the same gradients as elem_coef_seq with an algebraically identical
# formula based at another node of the element (ring_math.h).  On FEM-quality meshes (Kuhn,
# Delaunay) the two agree to ~1e-15 and are held to 1e-12.  On RANDOM 4-subsets of points the elements
# can be arbitrarily flat: both formulas lose eps * (h^3 / volume) digits, different ones, and the
# reference itself is up to 1.7e-12 away from the exact value of its own formula.  Those layout-stress
# meshes are judged against an 80-bit evaluation of the reference's formula (helpers.extended_truth):
# within SLIVER_FACTOR times the reference's own error of it, wherever 1e-12 against the reference fails.
SLIVER = "sliver"


@pytest.fixture(scope="module")
def ringlib():
    lib = C.CDLL(os.path.join(ROOT, "tools", "libmfb_ringcheck.so"))
    lib.mfb_ring_replay_error.restype = C.c_char_p
    lib.mfb_ring_replay.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p] * 3
    return lib


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def replay(lib, setup, rows=0, entries=0, bank_aware=1, interface=None):
    m = setup.mesh
    dim = setup.operatorDim
    values = np.full(setup.nbEdges * dim, np.nan)
    prec = np.full(m.nbNodes * dim, np.nan)
    stats = np.zeros(16, np.int64)
    keep = [np.ascontiguousarray(setup.elemToNode, np.int32), np.ascontiguousarray(setup.row, np.int32),
            np.ascontiguousarray(setup.col, np.int32), np.ascontiguousarray(m.coord, np.float64),
            np.ascontiguousarray(setup.checkBounds, np.int32)]
    rc = lib.mfb_ring_replay(setup.operatorID, m.nbNodes, keep[0].size // 4, _p(keep[0]), _p(keep[1]), _p(keep[2]),
                             _p(keep[3]), _p(keep[4]), _p(interface), rows, entries, bank_aware,
                             _p(values), _p(prec), _p(stats))
    assert rc == 0, lib.mfb_ring_replay_error().decode()
    return values, prec, stats


def check_against_oracle(oracle, setup, values, prec, interface=None, rtol=RTOL):
    want_v, want_p0, want_p = oracle.fem_iteration(setup)
    dim = setup.operatorDim
    if rtol == SLIVER:
        truth = extended_truth(setup)
        assert_close_or_conditioned(values, want_v, truth, setup.row, dim)
        assert interface is None
        assert_prec_close_or_conditioned(prec, want_p, extended_truth_prec(setup, truth), dim, rho=diag_conditioning(setup, want_v))
        return
    assert row_scaled_error(values, want_v, setup.row, dim) <= rtol
    if interface is None:
        assert block_scaled_error(prec, want_p, dim) <= rtol
    else:          # interface rows keep the raw diagonal block for the halo sum, the others are inverted
        intf = interface.astype(bool)
        got, raw, inv = prec.reshape(-1, dim), want_p0.reshape(-1, dim), want_p.reshape(-1, dim)
        assert block_scaled_error(got[intf], raw[intf], dim) <= RTOL
        assert block_scaled_error(got[~intf], inv[~intf], dim) <= RTOL


@pytest.mark.parametrize("op", ["ela", "lap"])
@pytest.mark.parametrize("grid,rows,entries", [((12, 10, 9), 0, 0), ((7, 6, 5), 16, 200), ((9, 9, 9), 64, 960), ((3, 2, 2), 1, 32)])
def test_ring_replay_matches_oracle(ringlib, oracle, op, grid, rows, entries):
    mesh = mfb.Mesh.generate(*grid, seed=5)
    setup = mfb.Setup(mesh, op)
    values, prec, stats = replay(ringlib, setup, rows, entries)
    check_against_oracle(oracle, setup, values, prec)
    # every off-diagonal CSR entry comes from one job, as its own or its transposed block
    nb_offdiag = setup.nbEdges - mesh.nbNodes
    assert stats[1] + stats[2] == nb_offdiag
    assert stats[5] == 0                                   # a conforming mesh: one chain per edge
    if rows:
        assert stats[11] <= rows and stats[13] <= entries + 7 * rows      # slab slots: up to 7 idle ones per row


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_replay_plain_order_and_shuffled_numbering(ringlib, oracle, op):
    mesh = mfb.Mesh.generate(8, 7, 6, seed=3)
    rng = np.random.default_rng(1)
    nperm, eperm = rng.permutation(mesh.nbNodes), rng.permutation(mesh.nbElem)
    coord = np.empty_like(mesh.coord).reshape(-1, 3)
    coord[nperm] = mesh.coord.reshape(-1, 3)
    e2n = np.empty_like(mesh.elemToNode).reshape(-1, 4)
    e2n[eperm] = nperm[mesh.elemToNode.reshape(-1, 4) - 1] + 1
    codes = np.empty_like(mesh.boundNodesCode)
    codes[nperm] = mesh.boundNodesCode
    shuffled = mfb.Setup(ArrayMesh(coord.ravel(), e2n.ravel(), mesh.nbNodes, codes), op)
    for setup, bank in ((mfb.Setup(mesh, op), 0), (shuffled, 1)):
        values, prec, _ = replay(ringlib, setup, bank_aware=bank)
        check_against_oracle(oracle, setup, values, prec)


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_replay_random_tets(ringlib, oracle, op):
    """Random 4-subsets: edge links are arbitrary graphs (several chains per edge, breaks), rows are
    long, nodes can be isolated."""
    rng = np.random.default_rng(12)
    coord, e2n = random_tet_mesh(rng, 60, 150)
    codes = rng.choice([0, 0, 0, 52, 53, 54, 10], size=60).astype(np.int32)
    setup = mfb.Setup(ArrayMesh(coord, e2n, 60, codes), op)
    values, prec, stats = replay(ringlib, setup, rows=8, entries=400)
    check_against_oracle(oracle, setup, values, prec, rtol=SLIVER)
    assert 6 * 150 <= stats[3] <= 12 * 150                 # an (element, edge) pair is visited once, or once per tile when the edge crosses tiles
    coord, e2n = random_tet_mesh(rng, 25, 400)             # dense: many elements around every edge, chains with breaks
    setup = mfb.Setup(ArrayMesh(coord, e2n, 25), op)
    values, prec, stats = replay(ringlib, setup, rows=25, entries=640)
    check_against_oracle(oracle, setup, values, prec, rtol=SLIVER)
    assert stats[0] == 1 and stats[3] == 6 * 400 and stats[5] > 0


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_replay_delaunay(ringlib, oracle, op):
    """A genuinely unstructured mesh: Delaunay tetrahedralisation of random points, slivers included
    (tools/delaunay_mesh.py).  Measured against an 80-bit evaluation of the reference's formula on
    20,000 points, the oracle itself is 7.7e-14 away (row-scaled) and the RING replay 1.7e-13: the
    1e-12 bar holds with the margin the slivers leave to any double-precision evaluation."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from delaunay_mesh import delaunay_arrays
    coord, e2n, codes = delaunay_arrays(1500, seed=4)
    setup = mfb.Setup(ArrayMesh(coord, e2n, coord.size // 3, codes), op)
    values, prec, stats = replay(ringlib, setup)
    check_against_oracle(oracle, setup, values, prec)
    assert stats[5] == 0                                   # Delaunay links are closed polygons or open fans


def test_ring_replay_interface_rows_keep_raw_blocks(ringlib, oracle):
    mesh = mfb.Mesh.generate(6, 6, 6, blocks=(2, 1, 1), rank=1, seed=2)
    setup = mfb.Setup(mesh, "ela")
    interface = np.zeros(mesh.nbNodes, np.uint8)
    interface[mesh.intfNodes - 1] = 1
    values, prec, _ = replay(ringlib, setup, interface=interface)
    check_against_oracle(oracle, setup, values, prec, interface)


def test_ring_plan_selfcheck_cabi():
    """The structural check of the plan through the product library (no GPU)."""
    mesh = mfb.Mesh.generate(6, 6, 6, blocks=(2, 1, 1), rank=0, seed=2)
    s = mfb.Setup(mesh, "ela")
    keep = [np.ascontiguousarray(mesh.coord), s.elemToNode, s.row, s.col, mesh.intfIndex, mesh.intfNodes, mesh.neighborsList]
    p = mfb.Problem(1, mesh.nbElem, mesh.nbNodes, s.nbEdges, *[k.ctypes.data for k in keep[:4]], None, None, None, 0,
                    2, 0, mesh.nbIntf, mesh.nbIntfNodes, *[k.ctypes.data for k in keep[4:]])
    stats = (C.c_int64 * 12)()
    assert mfb.lib.mfb_ring_plan_selfcheck(C.byref(p), 0, 0, stats) == 0, mfb.lib.mfb_last_error()
    assert stats[1] + stats[2] == s.nbEdges - mesh.nbNodes
    assert 0 < stats[11] <= stats[0]                       # tiles that own interface nodes come first
    assert stats[6] >= stats[7] > 0 and stats[8] >= stats[9] > 0
    # caps that a single node cannot meet are reported, not silently exceeded
    assert mfb.lib.mfb_ring_plan_selfcheck(C.byref(p), 4, 3, stats) != 0
    assert b"exceeds the tile caps" in mfb.lib.mfb_last_error()


def test_ring_plan_rejects_degenerate_elements():
    coord = np.arange(15, dtype=np.float64) ** 1.5
    e2n = np.array([1, 2, 3, 4, 2, 3, 3, 5], np.int32)      # the second element names node 3 twice
    mesh = ArrayMesh(coord, e2n, 5)
    s = mfb.Setup(mesh, "ela")
    keep = [np.ascontiguousarray(coord), s.elemToNode, s.row, s.col]
    p = mfb.Problem(1, 2, 5, s.nbEdges, *[k.ctypes.data for k in keep], None, None, None, 0, 1, 0, 0, 0, None, None, None)
    stats = (C.c_int64 * 12)()
    assert mfb.lib.mfb_ring_plan_selfcheck(C.byref(p), 0, 0, stats) != 0
    assert b"names a node twice" in mfb.lib.mfb_last_error()


def test_ring_plan_property(ringlib, oracle):
    """Random unstructured connectivity and caps, both operators: a plan that builds verifies and its
    replay reproduces the oracle; caps are met or reported."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=30, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), nbNodes=st.integers(4, 90), density=st.floats(0.5, 5.0),
           rows=st.integers(1, 40), entries=st.integers(16, 700), operatorID=st.integers(0, 1))
    def check(seed, nbNodes, density, rows, entries, operatorID):
        rng = np.random.default_rng(seed)
        nbElem = max(1, int(nbNodes * density))
        coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
        setup = mfb.Setup(ArrayMesh(coord, e2n, nbNodes), "ela" if operatorID else "lap")
        m = setup.mesh
        stats = np.zeros(16, np.int64)
        values = np.full(setup.nbEdges * setup.operatorDim, np.nan)
        prec = np.full(nbNodes * setup.operatorDim, np.nan)
        keep = [setup.elemToNode, setup.row, setup.col, np.ascontiguousarray(m.coord), setup.checkBounds]
        rc = ringlib.mfb_ring_replay(operatorID, nbNodes, nbElem, *[_p(np.ascontiguousarray(k)) for k in keep], None,
                                     rows, entries, 1, _p(values), _p(prec), _p(stats))
        if rc != 0:
            msg = ringlib.mfb_ring_replay_error()
            assert rc == -1 and (b"exceeds the tile caps" in msg or b"more than 254 nodes" in msg), msg
            return
        assert stats[11] <= rows and stats[13] <= entries + 7 * rows and stats[12] <= 254
        want_v, _, want_p = oracle.fem_iteration(setup)
        # random tetrahedra can be arbitrarily flat: compare where the reference's own numbers are finite
        if np.all(np.isfinite(want_v)) and np.all(np.isfinite(want_p)) and np.abs(want_v).max() < 1e12:
            assert row_scaled_error(values, want_v, setup.row, setup.operatorDim) <= 1e-9
    check()


# ------------------------------------------------------------------------------------------------
# The kernel's SOURCE on host threads (tools/ring_kernel_host.cc + tools/cuda_cta_emulation.h): the
# same csrc/kernels_ring.cu that nvcc compiles for sm_100a, built with g++, one std::thread per CUDA
# thread, real barriers, emulated mbarriers / TMA / cp.async / shuffles.  Covers what the replay
# cannot: the prefetch protocol across tiles, the mbarrier phase arithmetic, buffer offsets, the
# persistent-grid stride, the lane mappings of the write-out.
# ------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def kernel_host():
    # MFB_RINGKERNEL_HOST_LIB: another build of the same test aid (e.g. -DMFB_RING_STAGE_LDG=1, the other staging variant)
    lib = C.CDLL(os.environ.get("MFB_RINGKERNEL_HOST_LIB") or os.path.join(ROOT, "tools", "libmfb_ringkernel_host.so"))
    lib.mfb_ring_kernel_host_error.restype = C.c_char_p
    lib.mfb_ring_kernel_host.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_int]
    return lib


def run_kernel_on_host(lib, setup, rows=0, entries=0, ctas=3, fuse=1, interface=None, threads=384):
    m = setup.mesh
    dim = setup.operatorDim
    values = np.full(setup.nbEdges * dim, np.nan)
    prec = np.full(m.nbNodes * dim, np.nan)
    keep = [np.ascontiguousarray(setup.elemToNode, np.int32), np.ascontiguousarray(setup.row, np.int32),
            np.ascontiguousarray(setup.col, np.int32), np.ascontiguousarray(m.coord, np.float64),
            np.ascontiguousarray(setup.checkBounds, np.int32)]
    rc = lib.mfb_ring_kernel_host(setup.operatorID, m.nbNodes, keep[0].size // 4, *[_p(k) for k in keep], _p(interface),
                                  rows, entries, ctas, fuse, _p(values), _p(prec), threads)
    assert rc == 0, lib.mfb_ring_kernel_host_error().decode()
    return values, prec


@pytest.mark.parametrize("op", ["ela", "lap"])
@pytest.mark.parametrize("grid,rows,entries,ctas", [((7, 6, 5), 0, 0, 3), ((7, 6, 5), 0, 0, 1), ((6, 5, 4), 5, 100, 2),
                                                    ((8, 8, 6), 64, 960, 2), ((2, 2, 1), 0, 0, 4)])
def test_ring_kernel_source_on_host_threads(kernel_host, oracle, op, grid, rows, entries, ctas):
    mesh = mfb.Mesh.generate(*grid, seed=6)
    setup = mfb.Setup(mesh, op)
    values, prec = run_kernel_on_host(kernel_host, setup, rows, entries, ctas)
    check_against_oracle(oracle, setup, values, prec)


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_kernel_source_equals_replay_bit_for_bit(ringlib, kernel_host, op):
    """Two independent readings of the plan — the sequential replay and the kernel's own source on 384
    host threads per CTA (8 job warps, 4 write-out warps) — give identical bits (both test aids are built without FMA contraction)."""
    mesh = mfb.Mesh.generate(9, 8, 7, seed=8)
    setup = mfb.Setup(mesh, op)
    v_replay, p_replay, _ = replay(ringlib, setup)
    v_kernel, p_kernel = run_kernel_on_host(kernel_host, setup, ctas=2)
    assert np.array_equal(v_replay, v_kernel) and np.array_equal(p_replay, p_kernel, equal_nan=True)


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_kernel_source_768_threads(kernel_host, oracle, op):
    """The one-CTA-per-SM instantiation (16 job warps + 8 write-out warps, larger tiles)."""
    mesh = mfb.Mesh.generate(9, 8, 8, seed=9)
    setup = mfb.Setup(mesh, op)
    values, prec = run_kernel_on_host(kernel_host, setup, rows=54, entries=960, ctas=2, threads=768)
    check_against_oracle(oracle, setup, values, prec)


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_kernel_source_1024_threads(kernel_host, oracle, op):
    """The 1024-thread instantiation (warpgroups with their own register counts on the device: 16 job + 16 write-out warps
    for elasticity, 24 + 8 for the Laplacian; the write-out warps that take a second row group change from tile to tile)."""
    mesh = mfb.Mesh.generate(9, 8, 8, seed=10)
    setup = mfb.Setup(mesh, op)
    values, prec = run_kernel_on_host(kernel_host, setup, rows=64, entries=1100, ctas=2, threads=1024)
    check_against_oracle(oracle, setup, values, prec)


@pytest.mark.parametrize("threads", [640, 896])
def test_ring_kernel_source_other_cta_shapes(kernel_host, oracle, threads):
    """640 threads (11 job + 9 write-out warps at the registers of the launch) and 896 threads (16 + 12, per-warpgroup
    register counts on the device)."""
    mesh = mfb.Mesh.generate(8, 7, 6, seed=11)
    setup = mfb.Setup(mesh, "ela")
    values, prec = run_kernel_on_host(kernel_host, setup, rows=40, entries=700, ctas=2, threads=threads)
    check_against_oracle(oracle, setup, values, prec)


def test_ring_kernel_source_unfused_interface_and_random_tets(kernel_host, oracle):
    mesh = mfb.Mesh.generate(6, 6, 6, blocks=(2, 1, 1), rank=1, seed=2)
    setup = mfb.Setup(mesh, "ela")
    interface = np.zeros(mesh.nbNodes, np.uint8)
    interface[mesh.intfNodes - 1] = 1
    values, prec = run_kernel_on_host(kernel_host, setup, interface=interface)
    check_against_oracle(oracle, setup, values, prec, interface)
    values, prec = run_kernel_on_host(kernel_host, setup, fuse=0)          # values only: prec stays untouched
    want_v, _, _ = oracle.fem_iteration(setup)
    assert row_scaled_error(values, want_v, setup.row, 9) <= RTOL and np.all(np.isnan(prec))
    rng = np.random.default_rng(3)
    coord, e2n = random_tet_mesh(rng, 40, 200)                             # chains with breaks, long rows
    setup = mfb.Setup(ArrayMesh(coord, e2n, 40), "ela")
    values, prec = run_kernel_on_host(kernel_host, setup, rows=10, entries=400, ctas=2)
    check_against_oracle(oracle, setup, values, prec, rtol=SLIVER)


def test_ring_kernel_protocol_under_thread_sanitizer():
    """The same host-thread run inside a ThreadSanitizer build: the asynchronous copies are performed
    at issue (the earliest the hardware could touch their destination), so a missing barrier or a
    wrong mbarrier phase between the tiles of a CTA is reported as a data race.  (Removing the block
    barrier after the job phase makes this test report 11 races.)"""
    import subprocess
    pkg = os.path.join(ROOT, "mini-fem_b200")
    build = subprocess.run(["make", "-C", pkg, "ringkernel-tsan"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if build.returncode != 0:
        pytest.skip("no ThreadSanitizer build here: " + build.stdout[-300:])
    # one plan-builder thread: libgomp is not instrumented, its barriers would show up as races
    env = dict(os.environ, OMP_NUM_THREADS="1", MFB_PLAN_THREADS="1", TSAN_OPTIONS="halt_on_error=0")
    for cfg in (["6", "8", "140", "2"], ["7", "0", "0", "1"]):
        res = subprocess.run([os.path.join(ROOT, "tools", "ring_kernel_tsan")] + cfg, env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=600)
        assert "TSAN_RUN_DONE" in res.stdout, res.stdout[-2000:]
        assert "ThreadSanitizer" not in res.stdout, res.stdout[-4000:]
        assert "NaNs left 0" in res.stdout


@pytest.mark.parametrize("op", ["ela", "lap"])
def test_ring_mesh_without_elements(ringlib, kernel_host, oracle, op):
    """Nodes but no element: empty rows, tiles without jobs (a zero-byte tail), prec from an all-zero
    diagonal — what the reference computes from its zero-filled arrays."""
    codes = np.array([0, 52, 10, 0, 54], np.int32)
    mesh = ArrayMesh(np.arange(15, dtype=np.float64), np.zeros(0, np.int32), 5, codes)
    setup = mfb.Setup(mesh, op)
    assert setup.nbEdges == 0
    _, want_p0, want_p = oracle.fem_iteration(setup)
    for values, prec in (replay(ringlib, setup)[:2], run_kernel_on_host(kernel_host, setup, ctas=2)):
        assert values.size == 0
        assert np.array_equal(prec, want_p, equal_nan=True)


def test_ring_kernel_source_randomized_against_replay(ringlib, kernel_host):
    """Random meshes (Kuhn, random tetrahedra with chain breaks, partitioned blocks with interface rows),
    tile caps, grid sizes and both CTA shapes: the kernel's source on host threads equals the sequential
    replay bit for bit (60 such trials were run once; the suite keeps 16)."""
    rng = np.random.default_rng(2024)
    for _ in range(16):
        kind = int(rng.integers(0, 3))
        op = "ela" if rng.integers(0, 2) else "lap"
        if kind == 0:
            mesh = mfb.Mesh.generate(*(int(x) for x in rng.integers(1, 8, size=3)), seed=int(rng.integers(1, 99)))
        elif kind == 1:
            n = int(rng.integers(5, 60))
            coord, e2n = random_tet_mesh(rng, n, max(int(n * rng.uniform(0.5, 5)), 1))
            mesh = ArrayMesh(coord, e2n, n)
        else:
            mesh = mfb.Mesh.generate(*(int(x) for x in rng.integers(2, 6, size=3)), blocks=(2, 1, 1),
                                     rank=int(rng.integers(0, 2)), seed=3)
        setup = mfb.Setup(mesh, op)
        rows = int(rng.choice([0, 0, 3, 9, 20, 54]))
        entries = 0 if rows == 0 else int(rows * rng.integers(20, 40))
        interface = None
        if kind == 2 and mesh.nbIntfNodes > 0:
            interface = np.zeros(mesh.nbNodes, np.uint8)
            interface[mesh.intfNodes - 1] = 1
        try:
            v_replay, p_replay, _ = replay(ringlib, setup, rows, entries, 1, interface)
        except AssertionError as refused:                   # a row that cannot fit the caps: reported, not hidden
            assert "exceeds the tile caps" in str(refused) or "254 nodes" in str(refused)
            continue
        v_kernel, p_kernel = run_kernel_on_host(kernel_host, setup, rows, entries, int(rng.integers(1, 5)), 1, interface,
                                                int(rng.choice([384, 384, 768])))
        assert np.array_equal(v_replay, v_kernel, equal_nan=True) and np.array_equal(p_replay, p_kernel, equal_nan=True)
