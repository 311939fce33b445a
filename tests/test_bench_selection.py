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
