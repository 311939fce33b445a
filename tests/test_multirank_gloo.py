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
cards = mdist.all_gather_bytes(bytes([rank + 1]) * 1024, 1024)                   # the peer-to-peer cards, in rank order
assert cards == [bytes([1]) * 1024, bytes([2]) * 1024]


class FakeCtx:                               # mfb_ctx_p2p_*: all ranks switch to the peer-to-peer exchange, or none does
    def __init__(self, card_ok, connect_ok): self.card_ok, self.connect_ok, self.enabled, self.seen = card_ok, connect_ok, None, None
    def p2p_card(self):
        if not self.card_ok: raise mfb.MfbError("no window")
        return bytes([7]) * mfb.P2P_CARD_BYTES
    def p2p_connect(self, cards):
        self.seen = cards
        if not self.connect_ok: raise mfb.MfbError("no peer access")
    def p2p_enable(self, on): self.enabled = on


ok = FakeCtx(True, True)
assert mdist.p2p_connect(ok) == (True, "") and len(ok.seen) == 2 and ok.enabled is None
one_bad = FakeCtx(True, rank != 1)           # rank 1 cannot map its neighbour: rank 0 must switch off again
active, why = mdist.p2p_connect(one_bad)
assert not active and (one_bad.enabled is False if rank == 0 else "no peer access" in why), (rank, why, one_bad.enabled)
no_card = FakeCtx(rank != 0, True)           # rank 0 has no window to publish
active, why = mdist.p2p_connect(no_card)
assert not active and ("no window" in why if rank == 0 else True)
os.environ["MFB_HALO"] = "nccl"
assert mdist.p2p_connect(FakeCtx(True, True)) == (False, "MFB_HALO=nccl")
del os.environ["MFB_HALO"]
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
