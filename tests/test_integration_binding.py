"""The reference's UNMODIFIED driver (src/main.cc, FEM.cc, IO.cc, matrix.cc, coloring.cc, compiled where they
lie by oracle/Makefile target `b200`) running its four stages on the GPU through integration/b200_backend.cc —
the binding a Mini-FEM maintainer would add, here as code.  The reference's own
check_results (FEM.cc:59-98) compares the two norms with the checkings file and writes numerical_results_0."""
import os
import subprocess

import pytest

import minifem_b200 as mfb
from test_io_format import write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def exe(kind):
    return os.path.join(REF_DIR, f"minifem_b200_{kind}")


needs_binary = pytest.mark.skipif(not os.path.exists(exe("ref")), reason="oracle/_ref/minifem_b200_* not built (no /root/reference here)")


def run(kind, data, cwd, op, iters="4", **env):
    return subprocess.run([exe(kind), "LM6", op, iters], cwd=cwd, env=dict(os.environ, MINIFEM_DATA_PATH=data, OMP_NUM_THREADS="2", **env),
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("kind,op", [("ref", "ela"), ("ref", "lap"), ("coloring", "ela"), ("coloring", "lap")])
def test_reference_driver_on_gpu_stages(tmp_path, kind, op):
    """LM6-like files (25 x 25 x 40 cubes: 27,716 nodes, 150,000 tets): REF build -> RING path, COLORING build ->
    COLOR path with the colours and the permutation the reference's own coloring_creation computes."""
    data = str(tmp_path / "data")
    write_case(data, "LM6", op, (25, 25, 40), 4)
    res = run(kind, data, str(tmp_path), op)
    assert res.returncode == 0, res.stdout
    for needle in ("* Mini-FEM *", "Main FEM loop", "3. Matrix assembly...                done", "Average cycles",
                   "Preconditioner inversion      :", "Numerical stability of rank 0"):
        assert needle in res.stdout, res.stdout
    report = open(tmp_path / "numerical_results_0").read()
    diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
    assert len(diffs) == 2 and max(diffs) < 1e-13, report


@needs_binary
@pytest.mark.gpu
def test_reference_driver_other_paths(tmp_path):
    data = str(tmp_path / "data")
    write_case(data, "LM6", "ela", (10, 8, 6), 3)
    for path in ("tiled", "atomic"):
        res = run("ref", data, str(tmp_path), "ela", MINIFEM_B200_PATH=path)
        assert res.returncode == 0, res.stdout
        report = open(tmp_path / "numerical_results_0").read()
        diffs = [float(l.split(":")[1]) for l in report.splitlines() if "difference" in l]
        assert len(diffs) == 2 and max(diffs) < 1e-13, report


@needs_binary
@pytest.mark.skipif(mfb.device_count() > 0, reason="checks the behaviour on a box without GPU")
def test_reference_driver_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback behind the binding either: the reference's exit(EXIT_FAILURE) convention with the library's message."""
    data = str(tmp_path / "data")
    write_case(data, "LM6", "ela", (4, 3, 3), 3)
    res = run("ref", data, str(tmp_path), "ela", "2")
    assert res.returncode != 0 and "Error: GPU context" in res.stdout and "no CUDA device" in res.stdout, res.stdout
