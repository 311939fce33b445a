"""The layout builders on the GPU (mfb_device_create_nodeToNode / _elemToEdge /
_coloring_creation) against the host builders and against the
fixtures the reference's own create_nodeToNode / create_elemToEdge / coloring_creation produced
(tests/golden/).  Everything here is integer structure: the bar is bit-exact."""
import os
import time

import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import ArrayMesh, random_tet_mesh

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shuffled(e2n, rng):
    """Same elements, random element order and random local node order."""
    e = np.asarray(e2n).reshape(-1, 4).copy()
    e = e[rng.permutation(e.shape[0])]
    for k in range(e.shape[0]):
        e[k] = e[k][rng.permutation(4)]
    return np.ascontiguousarray(e.ravel(), dtype=np.int32)


def meshes():
    rng = np.random.default_rng(11)
    out = []
    for grid in [(1, 1, 1), (5, 4, 3), (16, 9, 12), (25, 25, 40)]:
        m = mfb.Mesh.generate(*grid, seed=3)
        out.append((f"kuhn{grid}", m.elemToNode.copy(), m.nbNodes))
        out.append((f"kuhn{grid}-shuffled", shuffled(m.elemToNode, rng), m.nbNodes))
    for nbNodes, nbElem in [(40, 90), (500, 3000), (3000, 2000)]:     # the last one leaves isolated nodes
        _, e2n = random_tet_mesh(rng, nbNodes, nbElem)
        out.append((f"random{nbNodes}n{nbElem}e", e2n, nbNodes))
    return out


MESHES = meshes()


@pytest.mark.parametrize("name,e2n,nbNodes", MESHES, ids=[m[0] for m in MESHES])
def test_csr_and_elem_to_edge_match_host(name, e2n, nbNodes):
    row_h, col_h = mfb.create_nodeToNode(e2n, nbNodes)
    row_d, col_d = mfb.device_create_nodeToNode(e2n, nbNodes)
    assert np.array_equal(row_d, row_h)
    assert np.array_equal(col_d, col_h)          # first-seen order, not merely the same sets
    e2e_h = mfb.create_elemToEdge(row_h, col_h, e2n)
    e2e_d = mfb.device_create_elemToEdge(row_d, col_d, e2n)
    assert np.array_equal(e2e_d, e2e_h)


@pytest.mark.parametrize("name,e2n,nbNodes", MESHES, ids=[m[0] for m in MESHES])
def test_coloring_matches_host(name, e2n, nbNodes):
    part_h, c2e_h, perm_h, nb_h = mfb.coloring_creation(e2n, nbNodes)
    part_d, c2e_d, perm_d, nb_d = mfb.device_coloring_creation(e2n, nbNodes)
    assert nb_d == nb_h
    assert np.array_equal(part_d, part_h)        # the sequential first-fit's colours
    assert np.array_equal(c2e_d, c2e_h)
    assert np.array_equal(perm_d, perm_h)        # stable permutation (coloring.cc:107)


@pytest.mark.parametrize("name", ["kuhn_5x4x3", "random_40n_90e"])
def test_reference_fixtures(name):
    """Outputs of the reference's own builders (oracle/_ref, see tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nbNodes = int(g["nbNodes"])
    row, col = mfb.device_create_nodeToNode(g["elemToNode"], nbNodes)
    assert np.array_equal(row, g["ref_row"]) and np.array_equal(col, g["ref_col"])
    assert np.array_equal(mfb.device_create_elemToEdge(row, col, g["elemToNode"]), g["ref_elemToEdge"])
    _, c2e, perm, nb = mfb.device_coloring_creation(g["elemToNode"], nbNodes)
    assert np.array_equal(perm, g["col_perm"])
    assert np.array_equal(c2e, g["col_colorToElem"][:nb + 1])
    sorted_e2n = mfb.permute_int_2d(g["elemToNode"], perm, 4)
    assert np.array_equal(sorted_e2n, g["col_elemToNode"])
    row, col = mfb.device_create_nodeToNode(sorted_e2n, nbNodes)
    assert np.array_equal(row, g["col_row"]) and np.array_equal(col, g["col_col"])


def test_setup_with_gpu_builder_runs_the_colour_path():
    mesh = mfb.Mesh.generate(9, 8, 7, seed=5)
    host = mfb.Setup(mesh, "ela", coloring=True, elem_to_edge=True)
    gpu = mfb.Setup(mesh, "ela", coloring=True, elem_to_edge=True, builder="gpu")
    for key in ("elemToNode", "row", "col", "elemToEdge", "colorToElem", "colorPerm"):
        assert np.array_equal(getattr(gpu, key), getattr(host, key)), key
    a, b = mfb.Context(host, path="color"), mfb.Context(gpu, path="color")
    a.iteration(); b.iteration()
    va, pa = a.download(); vb, pb = b.download()
    assert np.array_equal(va, vb) and np.array_equal(pa, pb)
    a.close(); b.close()


def test_empty_and_bad_input():
    row, col = mfb.device_create_nodeToNode(np.zeros(0, np.int32), 5)
    assert np.array_equal(row, np.zeros(6, np.int32)) and col.size == 0
    part, c2e, perm, nb = mfb.device_coloring_creation(np.zeros(0, np.int32), 5)
    assert nb == 0 and part.size == 0
    with pytest.raises(mfb.MfbError):
        mfb.device_create_nodeToNode(np.array([1, 2, 3, 9], np.int32), 4)        # id out of range
    with pytest.raises(mfb.MfbError):
        mfb.device_coloring_creation(np.array([1, 2, 2, 3], np.int32), 4)         # node named twice


def test_more_than_128_colours_is_reported():
    """129 elements around one node need 129 colours: coloring.cc:66-69 aborts; so do both builders."""
    n = 129
    e2n = np.zeros((n, 4), np.int32)
    e2n[:, 0] = 1
    e2n[:, 1:] = 2 + 3 * np.arange(n)[:, None] + np.arange(3)[None, :]
    nbNodes = int(e2n.max())
    with pytest.raises(mfb.MfbError):
        mfb.coloring_creation(e2n.ravel(), nbNodes)
    with pytest.raises(mfb.MfbError):
        mfb.device_coloring_creation(e2n.ravel(), nbNodes)
    ok = e2n[:128]
    part_h, *_ = mfb.coloring_creation(ok.ravel(), nbNodes)
    part_d, *_ = mfb.device_coloring_creation(ok.ravel(), nbNodes)
    assert np.array_equal(part_d, part_h) and part_d.max() == 127


def test_eib_size(capsys):
    """BASELINE.json's EIB counts (6 M tets): identical layouts, and the setup times side by side."""
    mesh = mfb.Mesh.generate(100, 100, 100, seed=1)
    e2n, nbNodes = mesh.elemToNode, mesh.nbNodes
    t = [time.perf_counter()]
    row_h, col_h = mfb.create_nodeToNode(e2n, nbNodes); t.append(time.perf_counter())
    row_d, col_d = mfb.device_create_nodeToNode(e2n, nbNodes); t.append(time.perf_counter())
    assert np.array_equal(row_d, row_h) and np.array_equal(col_d, col_h)
    t.append(time.perf_counter())
    part_h, c2e_h, perm_h, nb_h = mfb.coloring_creation(e2n, nbNodes); t.append(time.perf_counter())
    part_d, c2e_d, perm_d, nb_d = mfb.device_coloring_creation(e2n, nbNodes); t.append(time.perf_counter())
    assert nb_d == nb_h and np.array_equal(part_d, part_h) and np.array_equal(perm_d, perm_h)
    assert np.array_equal(c2e_d, c2e_h)
    with capsys.disabled():
        print(f"\n  EIB setup: CSR host {t[1]-t[0]:.3f}s gpu {t[2]-t[1]:.3f}s | colouring host {t[4]-t[3]:.3f}s "
              f"gpu {t[5]-t[4]:.3f}s ({nb_d} colours), host<->device copies included")
