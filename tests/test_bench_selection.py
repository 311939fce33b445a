"""bench.py --path auto: RING is measured only after an isolated probe has shown it equal to TILED and
faster; anything else — a failed probe, a crash, garbage, a timeout — leaves the TILED path."""
import argparse
import json
import subprocess
import types

import bench


def _args():
    return argparse.Namespace(op="ela", grid=[100, 100, 100])


def _fake_run(stdout, returncode=0, raises=None):
    def run(cmd, **kw):
        assert "--probe-ring" in cmd
        if raises:
            raise raises
        return types.SimpleNamespace(stdout=stdout, returncode=returncode)
    return run


def test_choose_path(monkeypatch):
    good = {"ok": True, "ring_ms": 0.35, "tiled_ms": 0.57, "values_err": 1e-15, "prec_err": 1e-15, "split_err": 1e-15}
    monkeypatch.setattr(subprocess, "run", _fake_run("noise\nPROBE " + json.dumps(good) + "\n"))
    assert bench.choose_path(_args(), 0, 0)["chosen"] == "ring"
    slow = dict(good, ring_ms=0.6)
    monkeypatch.setattr(subprocess, "run", _fake_run("PROBE " + json.dumps(slow)))
    assert bench.choose_path(_args(), 0, 0)["chosen"] == "tiled"
    wrong = dict(good, ok=False, values_err=1e-3)
    monkeypatch.setattr(subprocess, "run", _fake_run("PROBE " + json.dumps(wrong)))
    verdict = bench.choose_path(_args(), 0, 0)
    assert verdict["chosen"] == "tiled" and verdict["probe"]["values_err"] == 1e-3
    monkeypatch.setattr(subprocess, "run", _fake_run("Segmentation fault", returncode=-11))
    verdict = bench.choose_path(_args(), 0, 0)
    assert verdict["chosen"] == "tiled" and "no verdict" in verdict["probe"]["error"]
    monkeypatch.setattr(subprocess, "run", _fake_run("", raises=subprocess.TimeoutExpired("bench.py", 300)))
    assert bench.choose_path(_args(), 0, 0)["chosen"] == "tiled"
    # ranks other than 0 do not probe; without a process group they keep TILED
    monkeypatch.setattr(subprocess, "run", _fake_run("", raises=AssertionError("rank 1 must not probe")))
    assert bench.choose_path(_args(), 1, 1)["chosen"] == "tiled"


def test_probe_without_gpu_reports_instead_of_raising(capsys):
    bench.probe_ring(argparse.Namespace(op="ela", grid=[3, 3, 3], device=0))
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("PROBE ")][-1]
    verdict = json.loads(line[6:])
    import minifem_b200 as mfb
    if mfb.device_count() == 0:
        assert verdict["ok"] is False and "no CUDA device" in verdict["error"]


SELECTION_WORKER = r'''
import argparse, json, os, subprocess, sys, types
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from minifem_b200 import dist as mdist
rank, world = mdist.init_from_env("gloo")
def run(cmd, **kw):
    assert rank == 0, "only rank 0 may probe"
    verdict = {"ok": True, "ring_ms": 0.3, "tiled_ms": 0.6}
    return types.SimpleNamespace(stdout="PROBE " + json.dumps(verdict), returncode=0)
subprocess.run = run
verdict = bench.choose_path(argparse.Namespace(op="ela", grid=[100, 100, 100]), rank, rank)
assert verdict["chosen"] == "ring", verdict                      # every rank measures the path rank 0 chose
assert (verdict["probe"] is not None) == (rank == 0)
mdist.barrier()
print("SELECTION_WORKER_OK", rank, flush=True)
'''


def test_choice_of_rank_0_reaches_every_rank(tmp_path):
    """World size 2 over gloo: rank 0 probes, both ranks end up on the same path."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "selection_worker.py"
    script.write_text(SELECTION_WORKER % root)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29549", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0 and res.stdout.count("SELECTION_WORKER_OK") == 2, res.stdout[-3000:]
