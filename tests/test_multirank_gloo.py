"""World-size-2 run of the multi-process plumbing on CPU (gloo): rendezvous, broadcast of the
communicator id bytes, the message pattern of the halo exchange (halo.cc:52-96) and the
timer max-reduce (FEM.cc:113).  The arithmetic around the exchange is the oracle's; the
result must equal the reference fixture."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
ROOT = %r
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import minifem_b200 as mfb
from minifem_b200 import dist as mdist
from oracle_lib import Oracle
rank, world = mdist.init_from_env("gloo")
assert world == 2
payload = bytes(range(128)) if rank == 0 else bytes(128)
assert mdist.broadcast_bytes(payload, 128) == bytes(range(128))
assert mdist.max_over_ranks(10 + rank) == 11 and mdist.sum_over_ranks(1 + rank) == 3
oracle = Oracle()
grid, blocks = (6, 5, 4), (2, 1, 1)
meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=13) for r in range(2)]
for op in ("lap", "ela"):
    dim = 1 if op == "lap" else 9
    setups = [mfb.Setup(m, op) for m in meshes]
    both = [np.ascontiguousarray(oracle.fem_iteration(s)[1]) for s in setups]
    want = [b.copy() for b in both]
    oracle.halo_exchange(want, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes], [m.neighborsList for m in meshes], dim)
    m, prec = meshes[rank], both[rank]
    send = prec.reshape(-1, dim)[m.intfNodes - 1].ravel()                       # halo.cc:77-80
    recv = mdist.exchange_host(send, m.intfIndex, m.neighborsList, dim)          # halo.cc:52-96
    np.add.at(prec.reshape(-1, dim), m.intfNodes - 1, recv.reshape(-1, dim))     # halo.cc:113-116
    assert np.array_equal(prec, want[rank]), op
mdist.barrier()
print("GLOO_WORKER_OK", rank, flush=True)
'''


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0 and res.stdout.count("GLOO_WORKER_OK") == 2, res.stdout[-3000:]
