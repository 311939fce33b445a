"""Generates tests/golden/*.npz by running the REFERENCE'S OWN SOURCES (oracle/_ref, built
from /root/reference by oracle/Makefile) on small seeded inputs.

    python tests/golden/make_golden.py        (in the build container; needs /root/reference)

Each fixture stores its inputs next to the reference's outputs, so the parity tests need
neither the generator nor the reference at run time.  The reference ships no golden
vectors for this path, which is why these are produced here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import minifem_b200 as mfb                      # mesh generator / file format only
from helpers import ArrayMesh, random_tet_mesh
from oracle_lib import Reference


class RefSetup:
    """main.cc:209-351 done with the reference's own functions."""

    def __init__(self, mesh, op, coloring):
        self.mesh = mesh
        self.operatorID = {"lap": 0, "ela": 1}[op]
        self.operatorDim = 1 if self.operatorID == 0 else 9
        ref = Reference("coloring" if coloring else "ref")
        self.elemToNode = mesh.elemToNode.copy()
        self.colorPerm = self.colorToElem = None
        self.nbTotalColors = 0
        if coloring:
            self.elemToNode, self.colorPerm, self.colorToElem, self.nbTotalColors = ref.coloring(mesh.elemToNode, mesh.nbNodes)
        self.row, self.col = ref.create_nodeToNode(self.elemToNode, mesh.nbNodes)
        self.nbEdges = int(self.row[-1])
        self.elemToEdge = ref.create_elemToEdge(self.row, self.col, self.elemToNode)
        self.checkBounds = ref.boundary_mask(mesh.boundNodesCode)


def single_domain(name, mesh):
    out = dict(coord=mesh.coord, elemToNode=mesh.elemToNode, boundNodesCode=mesh.boundNodesCode,
               nbNodes=np.int32(mesh.nbNodes))
    for coloring in (False, True):
        tag = "col" if coloring else "ref"
        ref = Reference("coloring" if coloring else "ref")
        ref_opt = Reference("coloring_opt" if coloring else "ref_opt")
        for op in ("lap", "ela"):
            s = RefSetup(mesh, op, coloring)
            if op == "lap":
                out[f"{tag}_row"], out[f"{tag}_col"], out[f"{tag}_elemToEdge"] = s.row, s.col, s.elemToEdge
                out[f"{tag}_checkBounds"] = s.checkBounds
                if coloring:
                    out["col_elemToNode"], out["col_perm"], out["col_colorToElem"] = s.elemToNode, s.colorPerm, s.colorToElem
            if coloring:                                          # colorToElem is a global of each library
                ref.coloring(mesh.elemToNode, mesh.nbNodes)
                ref_opt.coloring(mesh.elemToNode, mesh.nbNodes)
            values, precs, _, _ = ref.fem_loop([s], 2)            # the reference's FEM_loop
            values_opt, precs_opt, _, _ = ref_opt.fem_loop([s], 2)
            assert np.array_equal(values[0], values_opt[0]) and np.array_equal(precs[0], precs_opt[0], equal_nan=True)
            out[f"{tag}_{op}_values"], out[f"{tag}_{op}_prec"] = values[0], precs[0]
            out[f"{tag}_{op}_precInit"] = ref.prec_init(values[0], s.row, s.col, mesh.nbNodes, s.operatorDim)
            out[f"{tag}_{op}_norms"] = np.array([ref.norm(values[0]), ref.norm(precs[0])])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k.endswith("values")})


def multi_domain(name, grid, blocks):
    n = blocks[0] * blocks[1] * blocks[2]
    out = dict(grid=np.array(grid, np.int32), blocks=np.array(blocks, np.int32))
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=11) for r in range(n)]
    ref = Reference("ref")
    for op in ("lap", "ela"):
        setups = [RefSetup(m, op, False) for m in meshes]
        values, precs, _, _ = ref.fem_loop(setups, 2)
        for r in range(n):
            if op == "lap":
                m = meshes[r]
                for f in ("coord", "elemToNode", "boundNodesCode", "intfIndex", "intfNodes", "neighborsList", "globalNode"):
                    out[f"r{r}_{f}"] = getattr(m, f)
                out[f"r{r}_nbNodes"] = np.int32(m.nbNodes)
            out[f"r{r}_{op}_values"], out[f"r{r}_{op}_prec"] = values[r], precs[r]
    # the same global mesh as ONE domain: its prec must equal the halo-summed prec
    whole = mfb.Mesh.generate(*grid, seed=11)
    s = RefSetup(whole, "ela", False)
    _, precs, _, _ = ref.fem_loop([s], 2)
    out["whole_ela_prec"] = precs[0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, n, "ranks")


if __name__ == "__main__":
    single_domain("kuhn_5x4x3", mfb.Mesh.generate(5, 4, 3, seed=5))

    rng = np.random.default_rng(2024)
    coord, e2n = random_tet_mesh(rng, 40, 90)
    e2n[e2n == 17] = 18                                  # node 17 (1-based) becomes isolated
    for e in range(90):                                  # keep 4 distinct nodes per element
        while len(set(e2n[4 * e:4 * e + 4])) < 4:
            e2n[4 * e:4 * e + 4] = rng.choice([n for n in range(1, 41) if n != 17], size=4, replace=False)
    codes = rng.choice([0, 0, 0, 52, 53, 54, 10, 200], size=40).astype(np.int32)
    codes[16] = 54                                       # the isolated node carries a code too
    single_domain("random_40n_90e", ArrayMesh(coord, e2n, 40, codes))

    multi_domain("blocks_2x2x1_of_4x4x3", (4, 4, 3), (2, 2, 1))
