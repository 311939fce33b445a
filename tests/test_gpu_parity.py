"""Parity of the CUDA path (through the C ABI) with the CPU oracle, on the same seeded
inputs, with the fixtures the reference produced, and — at BASELINE.json's EIB size — through
size-independent properties.  Tolerance: helpers.RTOL (1e-12, north_star), measured per entry
against the row / block scale; CSR structure is bit-exact by construction (same arrays)."""
import os
import subprocess

import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import (RTOL, ArrayMesh, assert_close_or_conditioned, assert_prec_close_or_conditioned, block_scaled_error,
                     diag_conditioning, extended_truth, extended_truth_prec, random_tet_mesh, row_scaled_error)
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATHS = ["ring", "tiled", "atomic", "color"]      # ring = the default write-once path (the one bench.py measures)


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def check_against_oracle(oracle, setup, ctx, fused, slivers=False):
    """slivers: random 4-subsets of points — where 1e-12 against the reference fails, the result is held to the
    80-bit evaluation of the reference's formula (helpers.extended_truth, SLIVER_FACTOR)."""
    want_v, want_p0, want_p = oracle.fem_iteration(setup)
    dim = setup.operatorDim
    truth = extended_truth(setup) if slivers else None
    if fused:
        ctx.iteration()
    else:
        ctx.assembly()
        ctx.prec_init()
        _, p0 = ctx.download(values=False)
        if not slivers:                         # (with slivers p0, a copy of diagonal blocks, is covered by the values)
            assert block_scaled_error(p0, want_p0, dim) <= RTOL
        ctx.halo_exchange()                     # nbBlocks == 1: no-op (halo.cc:44)
        ctx.prec_inversion()
    v, p = ctx.download()
    if slivers:
        assert_close_or_conditioned(v, want_v, truth, setup.row, dim)
        assert_prec_close_or_conditioned(p, want_p, extended_truth_prec(setup, truth), dim, rho=diag_conditioning(setup, want_v))
    else:
        assert row_scaled_error(v, want_v, setup.row, dim) <= RTOL
        assert block_scaled_error(p, want_p, dim) <= RTOL
    return v, p


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("op", ["lap", "ela"])
@pytest.mark.parametrize("grid,seed", [((1, 1, 1), 1), ((5, 4, 3), 2), ((16, 9, 12), 3), ((25, 25, 40), 4)])
@pytest.mark.parametrize("fused", [False, True])
def test_structured_meshes(oracle, path, op, grid, seed, fused):
    """(25,25,40) is the LM6-like case of BASELINE.json (27,716 nodes, 150,000 tets)."""
    mesh = mfb.Mesh.generate(*grid, seed=seed)
    setup = mfb.Setup(mesh, op, coloring=(path == "color"))
    ctx = mfb.Context(setup, path=path, use_graph=(path == "color" and fused))
    check_against_oracle(oracle, setup, ctx, fused)
    if fused:                                   # a second iteration rewrites the same result
        v1, p1 = ctx.download()
        ctx.iteration()
        v2, p2 = ctx.download()
        if path != "atomic":                    # atomics sum in arrival order
            assert np.array_equal(v1, v2) and np.array_equal(p1, p2, equal_nan=True)
        else:
            assert row_scaled_error(v2, v1, setup.row, setup.operatorDim) <= RTOL
    assert ctx.launch_count() > 0
    ctx.close()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("op", ["lap", "ela"])
@pytest.mark.parametrize("name", ["kuhn_5x4x3", "random_40n_90e"])
def test_reference_fixtures(path, op, name):
    """Against what the reference's own sources computed (tests/golden)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    build = "col" if path == "color" else "ref"
    dim = 1 if op == "lap" else 9
    mesh = ArrayMesh(g["coord"], g["elemToNode"], int(g["nbNodes"]), g["boundNodesCode"])
    setup = mfb.Setup(mesh, op, coloring=(path == "color"))
    assert np.array_equal(setup.row, g[f"{build}_row"]) and np.array_equal(setup.col, g[f"{build}_col"])
    ctx = mfb.Context(setup, path=path)
    truth = extended_truth(setup) if name.startswith("random") else None      # random tetrahedra: see helpers.extended_truth
    truth_p = extended_truth_prec(setup, truth) if truth is not None else None
    for run in (ctx.stages, ctx.iteration):
        run()
        v, p = ctx.download()
        if truth is None:
            assert row_scaled_error(v, g[f"{build}_{op}_values"], setup.row, dim) <= RTOL
            assert block_scaled_error(p, g[f"{build}_{op}_prec"], dim) <= RTOL
        else:
            assert_close_or_conditioned(v, g[f"{build}_{op}_values"], truth, setup.row, dim)
            block_scaled_error(p, g[f"{build}_{op}_prec"], dim)               # incl. the isolated node: inf / masked block coincide
            assert_prec_close_or_conditioned(p, g[f"{build}_{op}_prec"], truth_p, dim, rho=diag_conditioning(setup, g[f"{build}_{op}_values"]))
    ctx.close()


@pytest.mark.parametrize("op", ["lap", "ela"])
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_unstructured_random(oracle, op, seed):
    rng = np.random.default_rng(seed)
    nbNodes, nbElem = int(rng.integers(30, 400)), int(rng.integers(20, 900))
    coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
    codes = rng.choice([0, 0, 52, 53, 54, 10, 50, 200], size=nbNodes).astype(np.int32)
    mesh = ArrayMesh(coord, e2n, nbNodes, codes)
    for path in PATHS:
        setup = mfb.Setup(mesh, op, coloring=(path == "color"))
        ctx = mfb.Context(setup, path=path, tile_rows=8, tile_elems=1024)
        check_against_oracle(oracle, setup, ctx, fused=True, slivers=True)
        ctx.close()


@pytest.mark.parametrize("op", ["lap", "ela"])
def test_arbitrary_input_numbering(oracle, op):
    """The same mesh with random node and element numbering (what an unstructured mesh generator
    hands over): the CSR layout follows the new ids, the tile plan only looks at coordinates."""
    mesh = mfb.Mesh.generate(11, 9, 10, seed=9)
    rng = np.random.default_rng(17)
    nperm, eperm = rng.permutation(mesh.nbNodes), rng.permutation(mesh.nbElem)
    coord = np.empty((mesh.nbNodes, 3)); coord[nperm] = mesh.coord.reshape(-1, 3)
    codes = np.empty_like(mesh.boundNodesCode); codes[nperm] = mesh.boundNodesCode
    e2n = (nperm[mesh.elemToNode.reshape(-1, 4) - 1] + 1)[eperm]
    shuffled = ArrayMesh(coord.ravel(), e2n.ravel(), mesh.nbNodes, codes)
    for path in PATHS:
        setup = mfb.Setup(shuffled, op, coloring=(path == "color"))
        ctx = mfb.Context(setup, path=path)
        check_against_oracle(oracle, setup, ctx, fused=True)
        ctx.close()


@pytest.mark.parametrize("rows,elems,threads", [(1, 64, 32), (7, 120, 64), (32, 400, 128), (64, 640, 256), (128, 1200, 256)])
def test_tile_shapes(oracle, rows, elems, threads):
    """Ragged tiles: one row per tile, tiles far below a warp batch, tiles near the caps."""
    mesh = mfb.Mesh.generate(9, 7, 8, seed=5)
    for op in ("lap", "ela"):
        setup = mfb.Setup(mesh, op)
        ctx = mfb.Context(setup, path="tiled", tile_rows=rows, tile_elems=elems, threads=threads)
        stats = ctx.plan_stats()
        assert stats["contributions"] == 16 * mesh.nbElem and stats["max_rows"] <= rows
        check_against_oracle(oracle, setup, ctx, fused=True)
        check_against_oracle(oracle, setup, ctx, fused=False)
        ctx.close()


@pytest.mark.parametrize("path", PATHS)
def test_mesh_without_elements(oracle, path):
    """Nodes but no element: empty CSR rows, lap prec = 1/0 = inf, ela prec = masked zero block."""
    mesh = ArrayMesh(np.arange(15, dtype=np.float64), np.zeros(0, np.int32), 5, np.array([0, 52, 10, 0, 54], np.int32))
    for op in ("lap", "ela"):
        setup = mfb.Setup(mesh, op, coloring=(path == "color"))
        assert setup.nbEdges == 0
        ctx = mfb.Context(setup, path=path)
        check_against_oracle(oracle, setup, ctx, fused=True)
        check_against_oracle(oracle, setup, ctx, fused=False)
        ctx.close()


@pytest.mark.parametrize("variant", ["plain", "pipeline"])
@pytest.mark.parametrize("op", ["lap", "ela"])
def test_tiled_kernel_variants(oracle, op, variant, monkeypatch):
    """The two non-default TILED kernels stay parity-green: the same phases without prefetch
    (MFB_TILED_VARIANT=plain) and the warp-specialised pipelined kernel (768-thread CTAs)."""
    if variant == "plain":
        monkeypatch.setenv("MFB_TILED_VARIANT", "plain")
    for grid, seed in (((3, 2, 2), 1), ((14, 11, 9), 2)):
        mesh = mfb.Mesh.generate(*grid, seed=seed)
        setup = mfb.Setup(mesh, op)
        ctx = mfb.Context(setup, path="tiled", threads=768 if variant == "pipeline" else 0)
        check_against_oracle(oracle, setup, ctx, fused=True)
        check_against_oracle(oracle, setup, ctx, fused=False)
        v1, p1 = ctx.download()
        ctx.iteration()
        v2, p2 = ctx.download()
        assert np.array_equal(v1, v2) and np.array_equal(p1, p2, equal_nan=True)
        ctx.close()


def test_element_interval_callback(oracle):
    """assembly_{lap,ela}_seq(userArgs, first, last) — inclusive interval, no zeroing."""
    mesh = mfb.Mesh.generate(6, 5, 7, seed=6)
    for op in ("lap", "ela"):
        setup = mfb.Setup(mesh, op)
        want, _, _ = oracle.fem_iteration(setup)
        ctx = mfb.Context(setup, path="atomic")
        ctx.zero_values()
        cut = mesh.nbElem // 3
        ctx.assembly_interval(0, cut)
        ctx.assembly_interval(cut + 1, mesh.nbElem - 1)
        v, _ = ctx.download(prec=False)
        assert row_scaled_error(v, want, setup.row, setup.operatorDim) <= RTOL
        ctx.assembly_interval(0, mesh.nbElem - 1)            # adds on top: exactly twice the matrix
        v2, _ = ctx.download(prec=False)
        assert row_scaled_error(v2, 2 * want, setup.row, setup.operatorDim) <= RTOL
        with pytest.raises(mfb.MfbError):
            ctx.assembly_interval(0, mesh.nbElem)
        ctx.close()
        tiled = mfb.Context(setup, path="tiled")
        with pytest.raises(mfb.MfbError, match="ATOMIC / COLOR"):
            tiled.assembly_interval(0, 1)
        tiled.close()


def test_moving_coordinates(oracle):
    """The plan depends on connectivity only: new coordinates need no rebuild."""
    mesh = mfb.Mesh.generate(7, 7, 7, seed=8)
    setup = mfb.Setup(mesh, "ela")
    ctx = mfb.Context(setup, path="tiled")
    ctx.iteration()
    rng = np.random.default_rng(3)
    mesh.coord = mesh.coord + rng.uniform(-0.05, 0.05, mesh.coord.shape)
    pinned = [mfb.PinnedArray(mesh.coord.size), mfb.PinnedArray(ctx.nbValues), mfb.PinnedArray(ctx.nbPrec)]
    pinned[0].array[:] = mesh.coord
    ctx.iteration_host(pinned[0].ptr, pinned[1].ptr, pinned[2].ptr)
    want_v, _, want_p = oracle.fem_iteration(setup)
    assert row_scaled_error(pinned[1].array, want_v, setup.row, 9) <= RTOL
    assert block_scaled_error(pinned[2].array, want_p, 9) <= RTOL
    ctx.close()
    for p in pinned:
        p.free()


@pytest.mark.parametrize("op", ["lap", "ela"])
def test_device_norms(oracle, op):
    """mfb_ctx_norms = compute_double_norm (FEM.cc:48-56) of values and prec, reduced on the device."""
    mesh = mfb.Mesh.generate(16, 9, 12, seed=3)
    setup = mfb.Setup(mesh, op)
    ctx = mfb.Context(setup, path="tiled")
    ctx.iteration()
    v, p = ctx.download()
    got = ctx.norms()
    assert got == ctx.norms()                                   # reproducible
    want = (mfb.double_norm(v), mfb.double_norm(p))
    assert abs(got[0] - want[0]) <= RTOL * want[0] and abs(got[1] - want[1]) <= RTOL * want[1]
    want_v, _, want_p = oracle.fem_iteration(setup)
    assert abs(got[0] - oracle.norm(want_v)) <= RTOL * got[0] and abs(got[1] - oracle.norm(want_p)) <= RTOL * got[1]
    pinned = mfb.PinnedArray(mesh.coord.size)
    pinned.array[:] = mesh.coord
    assert ctx.iteration_norms_host(pinned.ptr) == got
    pinned.free()
    ctx.close()


@pytest.mark.parametrize("op", ["lap", "ela"])
def test_eib_size_against_oracle(oracle, op):
    """The configuration every number is quoted on — BASELINE.json's EIB-like mesh (100^3 cubes: 1,030,301
    nodes, 6,000,000 tets) — entry by entry against the CPU oracle (about 2 s per operator on one core), for
    both write-once paths, fused and staged."""
    mesh = mfb.Mesh.generate(100, 100, 100, seed=1)
    setup = mfb.Setup(mesh, op)
    dim = setup.operatorDim
    want_v, _, want_p = oracle.fem_iteration(setup)
    for path in ("ring", "tiled"):
        ctx = mfb.Context(setup, path=path)
        for mode in ("fused", "staged"):
            ctx.iteration() if mode == "fused" else ctx.stages()
            v, p = ctx.download()
            ev, ep = row_scaled_error(v, want_v, setup.row, dim), block_scaled_error(p, want_p, dim)
            print(f"EIB {op} {path} {mode}: values {ev:.2e} prec {ep:.2e}")
            assert ev <= RTOL and ep <= RTOL, (path, mode)
        assert ctx.launch_count() > 0
        ctx.close()


@pytest.mark.parametrize("op", ["lap", "ela"])
def test_eib_size_properties(op):
    """BASELINE.json's EIB-like size (100^3 cubes: 1,030,301 nodes, 6,000,000 tets), where the
    oracle would take minutes: size-independent properties of the assembled operator.
      * every row sums to zero (the four gradient rows of a tetrahedron sum to zero,
        assembly.cc:115-117, so sum_k K_jk = 0 element by element);
      * K_ji = K_ij^T (the CSR structure is symmetric);
      * prec * diagonal block = I on nodes without Dirichlet component;
      * the write-once path (RING) agrees with the atomic path entry by entry;
      * two runs of the write-once path are bit-identical."""
    mesh = mfb.Mesh.generate(100, 100, 100, seed=1)
    dim = 1 if op == "lap" else 9
    setup = mfb.Setup(mesh, op, elem_to_edge=True)
    ctx = mfb.Context(setup, path="ring")
    ctx.iteration()
    v, p = ctx.download()
    ctx.iteration()
    v2, p2 = ctx.download()
    assert np.array_equal(v, v2) and np.array_equal(p, p2)
    ctx.close()
    row, lens = setup.row, np.diff(setup.row)
    blocks = v.reshape(-1, dim)
    scale = np.maximum.reduceat(np.abs(blocks).max(axis=1), row[:-1])
    sums = np.add.reduceat(blocks, row[:-1], axis=0)
    assert (np.abs(sums).max(axis=1) / scale).max() < 1e-11
    # symmetry through elemToEdge: entry (j,k) of an element vs entry (k,j)
    e2e = setup.elemToEdge.reshape(-1, 4, 4)[::97]
    a, b = e2e.reshape(-1, 16), e2e.transpose(0, 2, 1).reshape(-1, 16)
    if dim == 1:      # (i,j) and (j,i) sum the same products, possibly in a different order
        assert np.abs(v[a] - v[b]).max() <= 1e-13 * np.abs(v).max()
    else:
        va, vb = v.reshape(-1, 3, 3)[a], v.reshape(-1, 3, 3)[b].transpose(0, 1, 3, 2)
        assert np.abs(va - vb).max() <= 1e-13 * np.abs(va).max()
    # preconditioner = inverse of the diagonal block on free nodes
    pick = np.flatnonzero(mesh.boundNodesCode == 0)[::509][:2000]
    diag_idx = np.array([row[i] + list(setup.col[row[i]:row[i + 1]]).index(i + 1) for i in pick])
    d = v.reshape(-1, dim)[diag_idx]
    if dim == 1:
        assert np.abs(d[:, 0] * p[pick] - 1).max() < 1e-13
    else:
        prod = np.einsum("nij,njk->nik", p.reshape(-1, 3, 3)[pick], d.reshape(-1, 3, 3))
        assert np.abs(prod - np.eye(3)).max() < 1e-12
    atomic = mfb.Context(setup, path="atomic")
    atomic.iteration()
    va, pa = atomic.download()
    atomic.close()
    assert row_scaled_error(va, v, row, dim) <= RTOL
    assert block_scaled_error(pa, p, dim) <= RTOL


def test_driver_cli(tmp_path, oracle):
    """minifem_b200 $USE_CASE $OPERATOR $NB_ITERATIONS: the reference's driver contract
    (main.cc:57-94, FEM.cc:125-135, FEM.cc:68-97)."""
    from test_io_format import write_case
    data = str(tmp_path / "data")
    exe = os.path.join(ROOT, "mini-fem_b200", "minifem_b200")
    for op, path, fused in (("ela", "tiled", "1"), ("lap", "color", "0"), ("ela", "atomic", "0")):
        write_case(data, "LM6", op, (10, 8, 6), 3)
        env = dict(os.environ, MINIFEM_DATA_PATH=data, MINIFEM_PATH=path, MINIFEM_FUSED=fused)
        res = subprocess.run([exe, "LM6", op, "4"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=300)
        assert res.returncode == 0, res.stdout
        out = res.stdout
        for needle in ("* Mini-FEM *", 'Test case              : "LM6"', "Creating CSR matrix...", "Main FEM loop",
                       "3. Matrix assembly...                done", "Average cycles", "Preconditioner inversion      :",
                       "Numerical stability of rank 0"):
            assert needle in out, out
        report = open(tmp_path / "numerical_results_0").read()
        diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
        assert len(diffs) == 2 and max(diffs) < 1e-13, report
    # the same case with the layouts built on the device and the two norms reduced there
    env = dict(os.environ, MINIFEM_DATA_PATH=data, MINIFEM_PATH="color", MINIFEM_GPU_SETUP="1", MINIFEM_DEVICE_NORMS="1")
    res = subprocess.run([exe, "LM6", "ela", "3"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout
    report = open(tmp_path / "numerical_results_0").read()
    diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
    assert len(diffs) == 2 and max(diffs) < 1e-13, report
    # MINIFEM_STORE_CHECKINGS=1 writes the checkings file (store_ref_assembly_, IO.cc:42-58) ...
    os.remove(os.path.join(data, "LM6", "checkings", "ela_1_0"))
    env = dict(os.environ, MINIFEM_DATA_PATH=data, MINIFEM_STORE_CHECKINGS="1")
    res = subprocess.run([exe, "LM6", "ela", "2"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and "Stored reference checking" in res.stdout, res.stdout
    # ... and the next run compares against it
    env.pop("MINIFEM_STORE_CHECKINGS")
    res = subprocess.run([exe, "LM6", "ela", "2"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and "difference : 0.0e+00" in res.stdout, res.stdout
    res = subprocess.run([exe, "LM6", "foo", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and 'Incorrect argument "foo"' in res.stdout
    res = subprocess.run([exe, "EIB", "ela"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and "Please specify" in res.stdout
    res = subprocess.run([exe, "EIB", "ela", "2"], cwd=str(tmp_path), env=dict(os.environ, MINIFEM_DATA_PATH=data),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and "cannot read input data" in res.stdout
