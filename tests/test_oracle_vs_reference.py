"""The oracle against the reference's own sources (oracle/_ref) on fresh seeded inputs.
Runs where oracle/_ref was built (the build container, and the GPU box, which receives the
prebuilt libraries); skipped if they are absent."""
import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import ArrayMesh, random_tet_mesh
from oracle_lib import Oracle, Reference, ref_available

pytestmark = pytest.mark.skipif(not ref_available("ref"), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("grid,seed", [((3, 3, 3), 1), ((9, 4, 6), 2), ((1, 1, 1), 3), ((13, 11, 7), 4)])
def test_structured(oracle, grid, seed):
    mesh = mfb.Mesh.generate(*grid, seed=seed)
    ref, refc = Reference("ref"), Reference("coloring")
    row, col = oracle.create_nodeToNode(mesh.elemToNode, mesh.nbNodes)
    rrow, rcol = ref.create_nodeToNode(mesh.elemToNode, mesh.nbNodes)
    assert np.array_equal(row, rrow) and np.array_equal(col, rcol) and len(col) == mesh.nbEdges
    assert np.array_equal(oracle.create_elemToEdge(row, col, mesh.elemToNode),
                          ref.create_elemToEdge(rrow, rcol, mesh.elemToNode))
    part, c2e, perm, nb = oracle.coloring(mesh.elemToNode, mesh.nbNodes)
    re2n, rperm, rc2e, rnb = refc.coloring(mesh.elemToNode, mesh.nbNodes)
    assert nb == rnb and np.array_equal(perm, rperm) and np.array_equal(c2e, rc2e)
    assert np.array_equal(oracle.permute_int_2d(mesh.elemToNode, perm, 4), re2n)
    assert np.array_equal(oracle.boundary_mask(mesh.boundNodesCode), ref.boundary_mask(mesh.boundNodesCode))
    for op in ("lap", "ela"):
        s = mfb.Setup(mesh, op, elem_to_edge=True)
        v, _, p = oracle.fem_iteration(s)
        rv, rp, cycles, hz = ref.fem_loop([s], 2)
        assert np.array_equal(v, rv[0]) and np.array_equal(p, rp[0], equal_nan=True)
        assert hz > 1e8 and len(cycles) == 4


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_unstructured_random(oracle, seed):
    rng = np.random.default_rng(seed)
    nbNodes, nbElem = int(rng.integers(8, 60)), int(rng.integers(1, 150))
    coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
    codes = rng.choice([0, 0, 52, 53, 54, 10, 50, 200], size=nbNodes).astype(np.int32)
    mesh = ArrayMesh(coord, e2n, nbNodes, codes)
    ref = Reference("ref")
    for op in ("lap", "ela"):
        s = mfb.Setup(mesh, op, elem_to_edge=False)
        v, _, p = oracle.fem_iteration(s)
        rv, rp, _, _ = ref.fem_loop([s], 1)        # nbIter == 1 is timed too (FEM.cc:182)
        assert np.array_equal(v, rv[0]) and np.array_equal(p, rp[0], equal_nan=True)


def test_multi_rank_halo(oracle):
    grid, blocks = (5, 4, 4), (2, 1, 2)
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=9) for r in range(4)]
    ref = Reference("ref")
    for op in ("lap", "ela"):
        setups = [mfb.Setup(m, op) for m in meshes]
        rv, rp, _, _ = ref.fem_loop(setups, 3)
        precs = []
        for s in setups:
            v, p0, _ = oracle.fem_iteration(s)
            precs.append(np.ascontiguousarray(p0))
        oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes],
                             [m.neighborsList for m in meshes], setups[0].operatorDim)
        for r, s in enumerate(setups):
            out = oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID)
            assert np.array_equal(out, rp[r])


def test_coloring_build_with_installed_colours(oracle):
    """bench.py times the COLORING build with colours and permutation computed by this repository's
    host code and installed through mref_set_colors: same result as the reference colouring itself."""
    mesh = mfb.Mesh.generate(6, 5, 4, seed=8)
    refc = Reference("coloring")
    for op in ("lap", "ela"):
        s = mfb.Setup(mesh, op, coloring=True)
        refc.set_colors(s.colorToElem)
        v1, p1, _, _ = refc.fem_loop([s], 2)
        e2n, perm, c2e, nb = refc.coloring(mesh.elemToNode, mesh.nbNodes)       # the reference's own colouring
        assert np.array_equal(e2n, s.elemToNode) and np.array_equal(c2e, s.colorToElem)
        v2, p2, _, _ = refc.fem_loop([s], 2)
        assert np.array_equal(v1[0], v2[0]) and np.array_equal(p1[0], p2[0])
        want_v, _, want_p = oracle.fem_iteration(s)
        assert np.array_equal(v1[0], want_v) and np.array_equal(p1[0], want_p)


def test_modified_coloring_build_matches_the_shipped_one(oracle):
    """oracle/_ref/libminifem_ref_coloring_omp.so restores the per-colour `#pragma omp parallel for`
    the reference ships commented out (src/assembly.cc:362,516); bench.py reports it as the
    "modified" CPU baseline.  Elements of one colour share no node, so the threaded loop must give
    the same bits as the shipped serial one."""
    mesh = mfb.Mesh.generate(9, 7, 6, seed=12)
    shipped, modified = Reference("coloring"), Reference("coloring_omp")
    for op in ("lap", "ela"):
        s = mfb.Setup(mesh, op, coloring=True)
        shipped.set_colors(s.colorToElem)
        modified.set_colors(s.colorToElem)
        v1, p1, _, _ = shipped.fem_loop([s], 2)
        v2, p2, _, _ = modified.fem_loop([s], 2)
        assert np.array_equal(v1[0], v2[0]) and np.array_equal(p1[0], p2[0])
        want_v, _, want_p = oracle.fem_iteration(s)
        assert np.array_equal(v2[0], want_v) and np.array_equal(p2[0], want_p)
