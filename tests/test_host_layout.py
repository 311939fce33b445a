"""The product's host-side layout builders (mini-fem_b200/host, through the C ABI) against
the reference fixtures and the oracle: CSR rows / columns, elemToEdge, colours, permutation
and Dirichlet mask must be bit-exact."""
import os

import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import random_tet_mesh
from oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["kuhn_5x4x3", "random_40n_90e"])
def test_against_reference_fixtures(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nbNodes = int(g["nbNodes"])
    row, col = mfb.create_nodeToNode(g["elemToNode"], nbNodes)
    assert np.array_equal(row, g["ref_row"]) and np.array_equal(col, g["ref_col"])
    assert np.array_equal(mfb.create_elemToEdge(row, col, g["elemToNode"]), g["ref_elemToEdge"])
    mask, nbBound = mfb.boundary_mask(g["boundNodesCode"])
    assert np.array_equal(mask, g["ref_checkBounds"]) and nbBound == np.count_nonzero(g["boundNodesCode"])
    part, c2e, perm, nb = mfb.coloring_creation(g["elemToNode"], nbNodes)
    assert np.array_equal(perm, g["col_perm"]) and np.array_equal(c2e, g["col_colorToElem"])
    e2n = mfb.permute_int_2d(g["elemToNode"], perm, 4)
    assert np.array_equal(e2n, g["col_elemToNode"])
    rowc, colc = mfb.create_nodeToNode(e2n, nbNodes)
    assert np.array_equal(rowc, g["col_row"]) and np.array_equal(colc, g["col_col"])
    assert np.array_equal(mfb.create_elemToEdge(rowc, colc, e2n), g["col_elemToEdge"])
    assert mfb.double_norm(g["ref_ela_values"]) == g["ref_ela_norms"][0]


@pytest.mark.parametrize("seed", range(6))
def test_against_oracle_random(seed):
    rng = np.random.default_rng(100 + seed)
    if seed % 2 == 0:
        grid = tuple(int(x) for x in rng.integers(1, 9, size=3))
        mesh = mfb.Mesh.generate(*grid, seed=seed)
        e2n, nbNodes = mesh.elemToNode, mesh.nbNodes
    else:
        nbNodes = int(rng.integers(5, 80))
        _, e2n = random_tet_mesh(rng, nbNodes, int(rng.integers(1, 200)))
    orc = Oracle()
    idx, val = mfb.node_to_elem(e2n, nbNodes)
    oidx, oval = orc.node_to_elem(e2n, nbNodes)
    assert np.array_equal(idx, oidx) and np.array_equal(val, oval)
    row, col = mfb.create_nodeToNode(e2n, nbNodes)
    orow, ocol = orc.create_nodeToNode(e2n, nbNodes)
    assert np.array_equal(row, orow) and np.array_equal(col, ocol)
    assert np.array_equal(mfb.create_elemToEdge(row, col, e2n), orc.create_elemToEdge(row, col, e2n))
    got, want = mfb.coloring_creation(e2n, nbNodes), orc.coloring(e2n, nbNodes)
    assert got[3] == want[3]
    for a, b in zip(got[:3], want[:3]):
        assert np.array_equal(a, b)


def test_empty_mesh():
    row, col = mfb.create_nodeToNode(np.zeros(0, np.int32), 3)
    assert row.tolist() == [0, 0, 0, 0] and col.size == 0
    part, c2e, perm, nb = mfb.coloring_creation(np.zeros(0, np.int32), 3)
    assert nb == 1 and c2e.tolist() == [0, 0]          # coloring.cc:77 counts one colour


def test_more_than_128_colours_is_an_error():
    # 129 tetrahedra around one shared node all conflict: the reference prints
    # "Error: Not enough colors." and exits (coloring.cc:66-69).
    n = 129
    e2n = np.array([[1, 2 + 3 * k, 3 + 3 * k, 4 + 3 * k] for k in range(n)], np.int32).ravel()
    with pytest.raises(mfb.MfbError, match="Not enough colors"):
        mfb.coloring_creation(e2n, 1 + 3 * n)
    assert Oracle().coloring(e2n, 1 + 3 * n) is None
    part, c2e, perm, nb = mfb.coloring_creation(e2n[:4 * 128], 1 + 3 * n)
    assert nb == 128


def test_elem_to_edge_reports_missing_pair():
    e2n = np.array([1, 2, 3, 4], np.int32)
    row, col = mfb.create_nodeToNode(e2n, 4)
    col = col.copy(); col[1] = 4                      # break (1,2)
    with pytest.raises(mfb.MfbError):
        mfb.create_elemToEdge(row, col, e2n)


@pytest.mark.parametrize("grid,blocks", [((4, 4, 4), (2, 2, 2)), ((7, 3, 5), (3, 1, 2)), ((6, 6, 2), (2, 3, 1))])
def test_partition_interfaces_pair_up(grid, blocks):
    """halo.cc:113-116 adds position j of what neighbour s sent to position j of my list: both
    lists must name the same physical nodes in the same order."""
    n = blocks[0] * blocks[1] * blocks[2]
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=3) for r in range(n)]
    whole = mfb.Mesh.generate(*grid, seed=3)
    assert sum(m.nbElem for m in meshes) == whole.nbElem
    for r, m in enumerate(meshes):
        assert np.array_equal(m.coord.reshape(-1, 3), whole.coord.reshape(-1, 3)[m.globalNode])
        assert np.array_equal(m.boundNodesCode, whole.boundNodesCode[m.globalNode])
        for i in range(m.nbIntf):
            s = m.neighborsList[i] - 1
            o = meshes[s]
            back = [q for q in range(o.nbIntf) if o.neighborsList[q] - 1 == r]
            assert len(back) == 1
            mine = m.globalNode[m.intfNodes[m.intfIndex[i]:m.intfIndex[i + 1]] - 1]
            theirs = o.globalNode[o.intfNodes[o.intfIndex[back[0]]:o.intfIndex[back[0] + 1]] - 1]
            assert np.array_equal(mine, theirs)
    # every node shared by two subdomains is listed by that pair
    owners = {}
    for r, m in enumerate(meshes):
        for gnode in m.globalNode:
            owners.setdefault(int(gnode), set()).add(r)
    for r, m in enumerate(meshes):
        listed = {(int(m.neighborsList[i]) - 1, int(gn)) for i in range(m.nbIntf)
                  for gn in m.globalNode[m.intfNodes[m.intfIndex[i]:m.intfIndex[i + 1]] - 1]}
        expect = {(s, gn) for gn, rs in owners.items() if r in rs for s in rs if s != r}
        assert listed == expect


def test_choose_blocks():
    assert sorted(mfb.choose_blocks(100, 100, 100, 8)) == [2, 2, 2]
    assert np.prod(mfb.choose_blocks(100, 100, 100, 7)) == 7
    assert mfb.choose_blocks(4, 1, 1, 16) == (4, 1, 1)
