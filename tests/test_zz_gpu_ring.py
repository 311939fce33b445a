"""GPU parity of the RING path (MFB_PATH_RING), beyond the path matrix of test_gpu_parity.py: tile shapes,
CTA shapes (384 / 768 threads), one CTA, one tile per CTA, plan-order numbering, random tetrahedra with
chain breaks (held to the extended-precision truth, helpers.extended_truth), four subdomains with the
fused interface split, and the EIB size against the TILED path.  The cases run in a process of their own
(tests/ring_gpu_worker.py) so that a kernel fault cannot poison the CUDA context of the other tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ring_kernel_matches_oracle():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ring_gpu_worker.py")], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(res.stdout[-6000:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ring_gpu_worker.log"), "w") as f:
        f.write(res.stdout)
    assert res.returncode == 0 and "RING_GPU_OK" in res.stdout, res.stdout[-3000:]
