"""The peer-to-peer halo exchange (csrc/kernels_halo_p2p.cu) with ALL subdomains of a partition in one process on one GPU
(run in a process of its own: CUDA_DEVICE_MAX_CONNECTIONS must be set before CUDA starts so that the 2 streams of each of
up to 8 contexts get hardware queues of their own — the kernels of different subdomains wait for each other).
    python tests/p2p_worker.py px py pz [op] [iterations]
Checks every iteration of every subdomain against the oracle's multi-subdomain FEM iteration, then lets one subdomain
skip an iteration and expects its neighbours to report the timeout instead of hanging.  Prints P2P_WORKER_OK."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np                                                                  # noqa: E402

import minifem_b200 as mfb                                                          # noqa: E402
from helpers import RTOL, block_scaled_error, row_scaled_error                      # noqa: E402
from oracle_lib import Oracle                                                       # noqa: E402


def main():
    blocks = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 2, 2)
    op = sys.argv[4] if len(sys.argv) > 4 else "ela"
    iterations = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    grid = (5 * blocks[0], 4 * blocks[1], 4 * blocks[2])
    n = int(np.prod(blocks))
    dim = 1 if op == "lap" else 9
    oracle = Oracle()
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=7) for r in range(n)]
    setups = [mfb.Setup(m, op) for m in meshes]
    results = [oracle.fem_iteration(s) for s in setups]
    precs = [np.ascontiguousarray(res[1]) for res in results]
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes], [m.neighborsList for m in meshes], dim)
    want = [oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, s.operatorID) for r, s in enumerate(setups)]
    ctxs = [mfb.Context(s, path="ring", nbBlocks=n, rank=r, tile_rows=12, tile_elems=240) for r, s in enumerate(setups)]
    cards = [c.p2p_card() for c in ctxs]
    for c in ctxs:
        c.p2p_connect(cards)
    print(f"{n} subdomains, interfaces per subdomain: {[m.nbIntf for m in meshes]}", flush=True)
    for it in range(iterations):
        for c in ctxs:
            c.iteration()
        worst = 0.0
        for r, c in enumerate(ctxs):
            c.sync()
            v, p = c.download()
            ev, ep = row_scaled_error(v, results[r][0], setups[r].row, dim), block_scaled_error(p, want[r], dim)
            worst = max(worst, ev, ep)
            assert ev <= RTOL and ep <= RTOL, (it, r, ev, ep)
        print(f"iteration {it}: worst error {worst:.2e}", flush=True)
    if n > 1 and os.environ.get("P2P_WORKER_TIMEOUT_CASE", "1") != "0":
        # subdomain 0 stays out of one exchange: its neighbours' bounded waits run out (~3 s) and sync reports it
        for c in ctxs[1:]:
            c.iteration()
        failed = 0
        for c in ctxs[1:]:
            try:
                c.sync()
            except mfb.MfbError as e:
                assert "timed out" in str(e), e
                failed += 1
        assert failed >= 1, "a missing neighbour went unnoticed"
        print(f"missing neighbour: {failed} subdomains reported the timeout", flush=True)
    for c in ctxs:
        c.close()
    print("P2P_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
