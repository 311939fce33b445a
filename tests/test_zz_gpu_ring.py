"""GPU parity of the RING path (MFB_PATH_RING).  The kernel was written after this round's GPU
budget had run out: it compiles for sm_100a and its plan + arithmetic are replayed against the
oracle on the host (test_ring_plan.py), but it had not run on hardware when this file was
committed.  Hence: the cases run in a process of their own (a fault cannot poison the CUDA
context of the other tests), last in the suite, and a failure is reported as XFAIL until the
first measured round confirms the kernel; a pass shows up as XPASS."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(reason="RING kernel: first run on hardware pending (round-1 GPU budget was spent)", strict=False)
def test_ring_kernel_matches_oracle():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ring_gpu_worker.py")], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=720)
    print(res.stdout[-6000:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ring_gpu_worker.log"), "w") as f:
        f.write(res.stdout)
    assert res.returncode == 0 and "RING_GPU_OK" in res.stdout, res.stdout[-3000:]
