"""Host builder of the locality-blocked colouring (MFB_PATH_BLOCKCOLOR, host/mesh_topology.h: build_block_coloring) without
a GPU: the layout the kernel walks — launches (block colours) -> blocks (one CTA each) -> local colours (a barrier in
between) -> elements — visits every element exactly once, two blocks of one launch share no node, two elements of one
local colour of a block share no node.  Those three facts are what makes the kernel's plain read-modify-write race-free;
the arithmetic per element is the ATOMIC / COLOR paths' own function."""
import numpy as np
import pytest

import minifem_b200 as mfb
from helpers import random_tet_mesh


def check_layout(e2n, nbNodes, coord, block_elems):
    e2n = np.asarray(e2n, np.int32).reshape(-1, 4)
    nbElem = len(e2n)
    order, launch, index, start, st = mfb.block_coloring(e2n.ravel(), nbNodes, coord, block_elems)
    assert sorted(order.tolist()) == list(range(nbElem))                       # a permutation of the elements
    assert launch[0] == 0 and launch[-1] == st["blocks"] and np.all(np.diff(launch) > 0) if nbElem else True
    assert len(index) == st["blocks"] + 1
    seen = 0
    for c in range(st["block_colors"]):
        nodes_of_launch = {}
        for b in range(launch[c], launch[c + 1]):
            nb_local = index[b + 1] - index[b] - 1
            assert 1 <= nb_local <= st["max_local_colors"]
            block_nodes = set()
            for q in range(nb_local):
                lo, hi = start[index[b] + q], start[index[b] + q + 1]
                assert lo == seen and hi > lo                                   # consecutive, non-empty
                seen = hi
                elems = order[lo:hi]
                assert np.all(np.diff(elems) > 0)                               # increasing element id inside a colour
                touched = e2n[elems].ravel()
                assert len(set(touched.tolist())) == touched.size               # no node twice inside a local colour
                block_nodes.update(touched.tolist())
            assert hi - start[index[b]] <= max(block_elems, 1) if block_elems > 0 else True
            for n in block_nodes:                                               # no node in two blocks of one launch
                assert nodes_of_launch.setdefault(n, b) == b
    assert seen == nbElem
    return st


@pytest.mark.parametrize("grid,block_elems", [((6, 5, 4), 64), ((12, 10, 9), 256), ((9, 9, 9), 1024), ((3, 2, 2), 1), ((7, 7, 7), 0)])
def test_kuhn_meshes(grid, block_elems):
    mesh = mfb.Mesh.generate(*grid, seed=4)
    st = check_layout(mesh.elemToNode, mesh.nbNodes, mesh.coord, block_elems)
    if block_elems == 1:
        assert st["max_local_colors"] == 1 and st["blocks"] == mesh.nbElem
    if grid == (12, 10, 9):
        assert st["block_colors"] <= 16 and st["max_local_colors"] <= 40     # 24 elements meet at an interior node


def test_random_tetrahedra_and_degenerate_sizes():
    rng = np.random.default_rng(8)
    for nbNodes, nbElem, block in ((40, 200, 32), (25, 400, 100), (300, 900, 128)):
        coord, e2n = random_tet_mesh(rng, nbNodes, nbElem)
        check_layout(e2n, nbNodes, coord, block)
    order, launch, index, start, st = mfb.block_coloring(np.zeros(0, np.int32), 5, np.zeros(15), 16)
    assert st == dict(blocks=0, block_colors=0, max_local_colors=0) and len(order) == 0
    with pytest.raises(mfb.MfbError, match="out of range"):
        mfb.block_coloring(np.array([1, 2, 3, 9], np.int32), 5, np.zeros(15), 16)


def test_too_many_colours_are_reported():
    # 70 single-element blocks around one node: more than 64 block colours
    nbElem = 70
    e2n = np.array([[1, 3 * k + 2, 3 * k + 3, 3 * k + 4] for k in range(nbElem)], np.int32)
    nbNodes = int(e2n.max())
    coord = np.random.default_rng(1).random(nbNodes * 3)
    with pytest.raises(mfb.MfbError, match="more than 64 colours"):
        mfb.block_coloring(e2n.ravel(), nbNodes, coord, 1)
    # 130 elements of ONE block around one node: more than 128 local colours
    e2n = np.array([[1, 3 * k + 2, 3 * k + 3, 3 * k + 4] for k in range(130)], np.int32)
    nbNodes = int(e2n.max())
    with pytest.raises(mfb.MfbError, match="more than 128 local colours"):
        mfb.block_coloring(e2n.ravel(), nbNodes, np.random.default_rng(2).random(nbNodes * 3), 1000)
