"""MFB_PATH_BLOCKCOLOR on the GPU against the CPU oracle: one launch per block colour, one CTA per block, the block's local
colours with a barrier in between (csrc/kernels_scatter.cu: scatter_blocks_kernel; layout: tests/test_block_coloring.py)."""
import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import RTOL, ArrayMesh, assert_close_or_conditioned, block_scaled_error, extended_truth, random_tet_mesh, row_scaled_error
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("op", ["lap", "ela"])
@pytest.mark.parametrize("grid,block_elems", [((1, 1, 1), 0), ((5, 4, 3), 16), ((16, 9, 12), 256), ((25, 25, 40), 0)])
def test_structured_meshes(oracle, op, grid, block_elems):
    mesh = mfb.Mesh.generate(*grid, seed=3)
    setup = mfb.Setup(mesh, op)
    want_v, want_p0, want_p = oracle.fem_iteration(setup)
    dim = setup.operatorDim
    ctx = mfb.Context(setup, path="blockcolor", tile_elems=block_elems, use_graph=True)
    st = ctx.plan_stats()
    assert st["blocks"] >= 1 and 1 <= st["block_colors"] <= 64 and st["max_local_colors"] >= 1
    before = ctx.launch_count()
    ctx.assembly()
    assert ctx.launch_count() - before == st["block_colors"]            # one launch per block colour
    ctx.prec_init(); ctx.halo_exchange(); ctx.prec_inversion()
    v, p = ctx.download()
    assert row_scaled_error(v, want_v, setup.row, dim) <= RTOL and block_scaled_error(p, want_p, dim) <= RTOL
    ctx.iteration()                                                     # fused entry point (captured as a CUDA graph)
    v1, p1 = ctx.download()
    ctx.iteration()
    v2, p2 = ctx.download()
    assert row_scaled_error(v1, want_v, setup.row, dim) <= RTOL and block_scaled_error(p1, want_p, dim) <= RTOL
    assert np.array_equal(v1, v2) and np.array_equal(p1, p2, equal_nan=True)      # no atomics: the order is fixed
    with pytest.raises(mfb.MfbError, match="own element order"):
        ctx.assembly_interval(0, mesh.nbElem - 1)
    ctx.close()


def test_random_tetrahedra(oracle):
    rng = np.random.default_rng(31)
    for nbNodes, nbElem, block in ((60, 150, 32), (25, 400, 64), (300, 900, 0)):
        coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
        codes = rng.choice([0, 0, 52, 53, 54, 10], size=nbNodes).astype(np.int32)
        setup = mfb.Setup(ArrayMesh(coord, e2n, nbNodes, codes), "ela")
        want_v = oracle.fem_iteration(setup)[0]
        ctx = mfb.Context(setup, path="blockcolor", tile_elems=block)
        ctx.iteration()
        v, _ = ctx.download()
        assert_close_or_conditioned(v, want_v, extended_truth(setup), setup.row, 9)
        ctx.close()
