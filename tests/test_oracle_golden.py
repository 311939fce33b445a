"""The CPU oracle (oracle/minifem_oracle.c) against the fixtures that the reference's own
compiled sources produced (tests/golden/make_golden.py).  Structure is compared bit for
bit; values too, because the oracle follows the reference's expression order and is built
without FMA contraction like the reference's -mavx build."""
import os

import numpy as np
import pytest

from helpers import ArrayMesh
from oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SINGLE = ["kuhn_5x4x3", "random_40n_90e"]


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("name", SINGLE)
def test_layout_matches_reference(oracle, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nbNodes = int(g["nbNodes"])
    row, col = oracle.create_nodeToNode(g["elemToNode"], nbNodes)
    assert np.array_equal(row, g["ref_row"]) and np.array_equal(col, g["ref_col"])
    assert np.array_equal(oracle.create_elemToEdge(row, col, g["elemToNode"]), g["ref_elemToEdge"])
    assert np.array_equal(oracle.boundary_mask(g["boundNodesCode"]), g["ref_checkBounds"])
    part, c2e, perm, nb = oracle.coloring(g["elemToNode"], nbNodes)
    assert np.array_equal(perm, g["col_perm"]) and np.array_equal(c2e, g["col_colorToElem"])
    e2n = oracle.permute_int_2d(g["elemToNode"], perm, 4)
    assert np.array_equal(e2n, g["col_elemToNode"])
    rowc, colc = oracle.create_nodeToNode(e2n, nbNodes)
    assert np.array_equal(rowc, g["col_row"]) and np.array_equal(colc, g["col_col"])
    # colours are conflict-free: no node is touched twice inside a colour
    for c in range(nb):
        nodes = e2n.reshape(-1, 4)[c2e[c]:c2e[c + 1]].ravel()
        assert len(np.unique(nodes)) == nodes.size


@pytest.mark.parametrize("name", SINGLE)
@pytest.mark.parametrize("op", ["lap", "ela"])
@pytest.mark.parametrize("build", ["ref", "col"])
@pytest.mark.parametrize("optimized", [False, True])
def test_values_and_prec_match_reference(oracle, name, op, build, optimized):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nbNodes = int(g["nbNodes"])
    opID, dim = (0, 1) if op == "lap" else (1, 9)
    e2n = g["elemToNode"] if build == "ref" else g["col_elemToNode"]
    row, col = g[f"{build}_row"], g[f"{build}_col"]
    values = oracle.assembly(g["coord"], row, col, e2n, opID,
                             elemToEdge=g[f"{build}_elemToEdge"] if optimized else None,
                             colorToElem=g["col_colorToElem"] if build == "col" else None)
    assert np.array_equal(values, g[f"{build}_{op}_values"])
    prec0 = oracle.prec_init(values, row, col, nbNodes, dim)
    assert np.array_equal(prec0, g[f"{build}_{op}_precInit"])
    prec = oracle.prec_inversion(prec0, row, col, g[f"{build}_checkBounds"], nbNodes, opID)
    assert np.array_equal(prec, g[f"{build}_{op}_prec"], equal_nan=True)
    norms = g[f"{build}_{op}_norms"]
    assert oracle.norm(values) == norms[0]
    assert oracle.norm(prec) == norms[1] or (np.isnan(norms[1]) or np.isinf(norms[1]))


def test_isolated_node_semantics(oracle):
    """random_40n_90e has a node without elements and with boundary code 54: empty CSR row,
    lap prec = 1/0 = inf (preconditioner.cc:40), ela block masked but not inverted
    (elasclpr.f:29-32)."""
    g = np.load(os.path.join(GOLDEN, "random_40n_90e.npz"))
    row = g["ref_row"]
    assert row[17] - row[16] == 0
    assert np.isinf(g["ref_lap_prec"][16])
    blk = g["ref_ela_prec"][16 * 9:17 * 9]
    assert np.array_equal(blk, [0, 0, 0, 0, 0, 0, 0, 0, 1.0])


@pytest.mark.parametrize("op", ["lap", "ela"])
def test_halo_sum_matches_reference(oracle, op):
    g = np.load(os.path.join(GOLDEN, "blocks_2x2x1_of_4x4x3.npz"))
    n = int(np.prod(g["blocks"]))
    opID, dim = (0, 1) if op == "lap" else (1, 9)
    precs, rows, cols, cbs = [], [], [], []
    for r in range(n):
        nbNodes = int(g[f"r{r}_nbNodes"])
        row, col = oracle.create_nodeToNode(g[f"r{r}_elemToNode"], nbNodes)
        values = oracle.assembly(g[f"r{r}_coord"], row, col, g[f"r{r}_elemToNode"], opID)
        assert np.array_equal(values, g[f"r{r}_{op}_values"])
        precs.append(oracle.prec_init(values, row, col, nbNodes, dim))
        rows.append(row); cols.append(col); cbs.append(oracle.boundary_mask(g[f"r{r}_boundNodesCode"]))
    oracle.halo_exchange(precs, [g[f"r{r}_intfIndex"] for r in range(n)], [g[f"r{r}_intfNodes"] for r in range(n)],
                         [g[f"r{r}_neighborsList"] for r in range(n)], dim)
    for r in range(n):
        out = oracle.prec_inversion(precs[r], rows[r], cols[r], cbs[r], int(g[f"r{r}_nbNodes"]), opID)
        assert np.array_equal(out, g[f"r{r}_{op}_prec"])


def test_partitioned_prec_equals_whole_domain_prec(oracle):
    """Size-independent property: after the halo sum every copy of an interface node holds the
    preconditioner block of the unpartitioned mesh."""
    g = np.load(os.path.join(GOLDEN, "blocks_2x2x1_of_4x4x3.npz"))
    whole = g["whole_ela_prec"].reshape(-1, 9)
    for r in range(int(np.prod(g["blocks"]))):
        mine = g[f"r{r}_ela_prec"].reshape(-1, 9)
        want = whole[g[f"r{r}_globalNode"]]
        assert np.abs(mine - want).max() <= 1e-13 * np.abs(want).max()
