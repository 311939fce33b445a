"""The C-ABI library: it loads without a GPU, exports every symbol include/minifem_b200.h
declares, and its GPU entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import minifem_b200 as mfb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_in_header():
    text = open(os.path.join(ROOT, "include", "minifem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = declared_in_header()
    assert len(names) >= 40
    for name in names:
        assert hasattr(mfb.lib, name), f"{name} is declared in the header but not exported"
    assert sorted(set(mfb.DECLARED_SYMBOLS) - set(names)) == []


def test_header_compiles_as_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "minifem_b200.h"\nint main(void){mfb_problem p; mfb_options o; (void)p; (void)o; return MFB_OK;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", os.path.join(d, "t.o")])


def test_version_and_error_string():
    assert b"sm_100a" in mfb.lib.mfb_version()
    with pytest.raises(mfb.MfbError, match="cannot read input data"):
        mfb.Mesh.read("/nonexistent/file")


@pytest.mark.skipif(mfb.device_count() > 0, reason="checks the behaviour on a box without GPU")
def test_no_cpu_fallback():
    mesh = mfb.Mesh.generate(2, 2, 2)
    setup = mfb.Setup(mesh, "ela")
    for path in ("tiled", "atomic"):
        with pytest.raises(mfb.MfbError, match="no CUDA device"):
            mfb.Context(setup, path=path)
    # the GPU layout builders do not quietly run on the host either
    with pytest.raises(mfb.MfbError, match="no such CUDA device"):
        mfb.device_create_nodeToNode(mesh.elemToNode, mesh.nbNodes)
    with pytest.raises(mfb.MfbError, match="no such CUDA device"):
        mfb.device_create_elemToEdge(setup.row, setup.col, mesh.elemToNode)
    with pytest.raises(mfb.MfbError, match="no such CUDA device"):
        mfb.device_coloring_creation(mesh.elemToNode, mesh.nbNodes)
    with pytest.raises(mfb.MfbError, match="no such CUDA device"):
        mfb.Setup(mesh, "ela", builder="gpu")


def test_argument_validation_happens_before_cuda():
    mesh = mfb.Mesh.generate(2, 2, 2)
    setup = mfb.Setup(mesh, "ela")
    setup.nbEdges += 1
    with pytest.raises(mfb.MfbError, match="nbEdges"):
        mfb.Context(setup)
    setup.nbEdges -= 1
    with pytest.raises(mfb.MfbError, match="COLOR"):
        mfb.Context(setup, path="color")


@pytest.mark.parametrize("operatorID", [0, 1])
@pytest.mark.parametrize("grid,rows,elems", [((6, 5, 4), 0, 0), ((6, 5, 4), 16, 200), ((9, 9, 9), 64, 704), ((3, 2, 2), 1, 64)])
def test_tile_plan_selfcheck(grid, rows, elems, operatorID):
    """Host replay of the TILED plan against the reference's (element, j, k) loop (the Laplacian
    plan stores shared-memory slots of dot products, the elasticity plan (element, a, b) codes)."""
    mesh = mfb.Mesh.generate(*grid, seed=4)
    s = mfb.Setup(mesh, "ela" if operatorID else "lap")
    keep = [np.ascontiguousarray(mesh.coord), s.elemToNode, s.row, s.col]
    p = mfb.Problem(operatorID, mesh.nbElem, mesh.nbNodes, s.nbEdges, *[k.ctypes.data for k in keep], None, None, None, 0,
                    1, 0, 0, 0, None, None, None)
    stats = (C.c_int64 * 6)()
    rc = mfb.lib.mfb_tile_plan_selfcheck(C.byref(p), rows, elems, stats)
    assert rc == 0, mfb.lib.mfb_last_error()
    assert stats[2] == 16 * mesh.nbElem                 # every contribution exactly once
    assert stats[1] >= mesh.nbElem and stats[0] >= 1
    if rows:
        assert stats[3] <= rows and stats[4] <= elems


def test_tile_plan_selfcheck_unstructured_and_multirank():
    from helpers import random_tet_mesh, ArrayMesh
    rng = np.random.default_rng(8)
    coord, e2n = random_tet_mesh(rng, 50, 160)
    s = mfb.Setup(ArrayMesh(coord, e2n, 50), "lap")
    keep = [np.ascontiguousarray(coord), s.elemToNode, s.row, s.col]
    p = mfb.Problem(0, 160, 50, s.nbEdges, *[k.ctypes.data for k in keep], None, None, None, 0, 1, 0, 0, 0, None, None, None)
    stats = (C.c_int64 * 6)()
    assert mfb.lib.mfb_tile_plan_selfcheck(C.byref(p), 8, 400, stats) == 0, mfb.lib.mfb_last_error()
    assert stats[2] == 16 * 160
    # caps that a single node cannot meet are reported, not silently exceeded
    assert mfb.lib.mfb_tile_plan_selfcheck(C.byref(p), 8, 2, stats) != 0
    assert b"exceeds the tile caps" in mfb.lib.mfb_last_error()
    mesh = mfb.Mesh.generate(6, 6, 6, blocks=(2, 1, 1), rank=1, seed=2)
    s = mfb.Setup(mesh, "ela")
    keep = [np.ascontiguousarray(mesh.coord), s.elemToNode, s.row, s.col, mesh.intfIndex, mesh.intfNodes, mesh.neighborsList]
    p = mfb.Problem(1, mesh.nbElem, mesh.nbNodes, s.nbEdges, *[k.ctypes.data for k in keep[:4]], None, None, None, 0,
                    2, 1, mesh.nbIntf, mesh.nbIntfNodes, *[k.ctypes.data for k in keep[4:]])
    assert mfb.lib.mfb_tile_plan_selfcheck(C.byref(p), 16, 256, stats) == 0, mfb.lib.mfb_last_error()


def test_tile_plan_selfcheck_property():
    """Random unstructured connectivity, random caps, both operators: whenever a plan is built
    it replays the reference's (element, j, k) loop exactly; caps are either met or reported."""
    from hypothesis import given, settings, strategies as st
    from helpers import random_tet_mesh, ArrayMesh

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), nbNodes=st.integers(4, 120), density=st.floats(0.5, 6.0),
           rows=st.integers(1, 40), elems=st.integers(8, 400), operatorID=st.integers(0, 1))
    def check(seed, nbNodes, density, rows, elems, operatorID):
        rng = np.random.default_rng(seed)
        nbElem = max(1, int(nbNodes * density))
        coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
        s = mfb.Setup(ArrayMesh(coord, e2n, nbNodes), "ela" if operatorID else "lap")
        keep = [np.ascontiguousarray(coord), s.elemToNode, s.row, s.col]
        p = mfb.Problem(operatorID, nbElem, nbNodes, s.nbEdges, *[k.ctypes.data for k in keep], None, None, None, 0,
                        1, 0, 0, 0, None, None, None)
        stats = (C.c_int64 * 6)()
        rc = mfb.lib.mfb_tile_plan_selfcheck(C.byref(p), rows, elems, stats)
        if rc != 0:
            assert b"exceeds the tile caps" in mfb.lib.mfb_last_error(), mfb.lib.mfb_last_error()
            return
        assert stats[2] == 16 * nbElem
        assert stats[3] <= rows and stats[4] <= elems + 3      # ids include the coset numbering's holes

    check()
