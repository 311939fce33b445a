"""On-disk formats of the path (src/IO.cc): the per-rank binary input file and the two-line
"checkings" file.  The strongest check runs the reference's OWN driver (main.cc, compiled
into oracle/_ref) on files this repository wrote."""
import os
import subprocess
import sys

import numpy as np
import pytest

import minifem_b200 as mfb
from oracle_lib import Oracle, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_input_file_layout_follows_IO_cc(tmp_path):
    mesh = mfb.Mesh.generate(4, 3, 5, blocks=(2, 1, 1), rank=1, seed=6)
    path = str(tmp_path / "EIB" / "inputs" / "ela_2_1")
    mfb.Mesh.generate_to_file(path, 4, 3, 5, blocks=(2, 1, 1), rank=1, seed=6)
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, np.int32, 6)                       # IO.cc:75-80
    nbElem, nbNodes, nbEdges, nbIntf, nbIntfNodes, nbBound = hdr
    assert (nbElem, nbNodes, nbEdges) == (mesh.nbElem, mesh.nbNodes, mesh.nbEdges)
    off = 24
    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off)
        off += a.nbytes
        return a
    assert np.array_equal(take(np.float64, nbNodes * 3), mesh.coord)          # IO.cc:89
    assert np.array_equal(take(np.int32, nbElem * 4), mesh.elemToNode)        # :90
    assert np.array_equal(take(np.int32, max(nbIntf, 1) * 3), mesh.neighborsList)   # :91
    assert np.array_equal(take(np.int32, nbIntf + 1), mesh.intfIndex)         # :92
    assert np.array_equal(take(np.int32, nbIntfNodes), mesh.intfNodes)        # :93
    assert np.array_equal(take(np.int32, nbNodes), mesh.boundNodesCode)       # :94
    assert off == len(raw)
    back = mfb.Mesh.read(path)
    for f in mfb.Mesh.FIELDS:
        assert np.array_equal(getattr(back, f), getattr(mesh, f))
    assert nbEdges == len(mfb.create_nodeToNode(mesh.elemToNode, mesh.nbNodes)[1])
    assert nbBound == np.count_nonzero(mesh.boundNodesCode)


def test_checking_file_roundtrip(tmp_path):
    import ctypes as C
    path = str(tmp_path / "LM6" / "checkings" / "lap_1_0").encode()
    a, b = 15977.719137496881, 3.2519381513914287e-05
    assert mfb.lib.mfb_checking_write(path, a, b) == 0
    lines = open(path).read().split()
    assert len(lines) == 2 and float(lines[0]) == a and float(lines[1]) == b   # setprecision(17)
    x, y = C.c_double(), C.c_double()
    assert mfb.lib.mfb_checking_read(path, C.byref(x), C.byref(y)) == 0
    assert (x.value, y.value) == (a, b)
    assert mfb.lib.mfb_checking_read(b"/nonexistent", C.byref(x), C.byref(y)) != 0
    assert b"cannot read reference checking" in mfb.lib.mfb_last_error()


def write_case(data, mesh_name, op, grid, seed):
    """inputs/ + checkings/ of a single-domain case; the norms come from the oracle."""
    mfb.Mesh.generate_to_file(os.path.join(data, mesh_name, "inputs", f"{op}_1_0"), *grid, seed=seed)
    mesh = mfb.Mesh.generate(*grid, seed=seed)
    orc = Oracle()
    v, _, p = orc.fem_iteration(mfb.Setup(mesh, op))
    path = os.path.join(data, mesh_name, "checkings", f"{op}_1_0").encode()
    assert mfb.lib.mfb_checking_write(path, orc.norm(v), orc.norm(p)) == 0


@pytest.mark.skipif(not ref_available("ref"), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,op", [("ref", "ela"), ("ref_opt", "lap"), ("coloring", "ela")])
def test_reference_driver_accepts_our_files(tmp_path, kind, op):
    data = str(tmp_path / "data")
    write_case(data, "LM6", op, (6, 5, 4), 12)
    code = ("import sys; sys.path.insert(0, %r); from oracle_lib import Reference; "
            "sys.exit(Reference(%r).main('LM6', %r, 3))" % (os.path.join(ROOT, "tests"), kind, op))
    env = dict(os.environ, MINIFEM_DATA_PATH=data, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=120)
    assert res.returncode == 0, res.stdout
    assert "Main FEM loop" in res.stdout and "Average cycles" in res.stdout
    report = open(tmp_path / "numerical_results_0").read()
    diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
    assert len(diffs) == 2
    # REF sums in element order like the oracle (difference exactly 0); COLORING sums colour
    # by colour, so its norm moves in the last bits
    assert all(d <= (0.0 if kind.startswith("ref") else 1e-14) for d in diffs), report


def test_meshgen_cli(tmp_path):
    """minifem_meshgen writes the per-rank input files of a block partition (both operators)."""
    exe = os.path.join(ROOT, "mini-fem_b200", "minifem_meshgen")
    data = str(tmp_path / "data")
    res = subprocess.run([exe, data, "EIB", "6", "4", "5", "4", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    blocks = mfb.choose_blocks(6, 4, 5, 4)
    for r in range(4):
        want = mfb.Mesh.generate(6, 4, 5, blocks=blocks, rank=r, seed=3)
        for op in ("lap", "ela"):
            got = mfb.Mesh.read(os.path.join(data, "EIB", "inputs", f"{op}_4_{r}"))
            for f in mfb.Mesh.FIELDS:
                assert np.array_equal(getattr(got, f), getattr(want, f))
    res = subprocess.run([exe, data, "EIB", "6", "4", "5", "7"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode != 0 and "cannot cut" in res.stdout
    assert subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).returncode != 0
