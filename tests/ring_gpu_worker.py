"""Parity of the RING kernel with the CPU oracle, run in a process of its own so that a kernel
fault cannot take the rest of the GPU suite down with it (tests/test_zz_gpu_ring.py).
Prints one line per case and RING_GPU_OK at the end; exits non-zero on the first mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import minifem_b200 as mfb                                                          # noqa: E402
from helpers import (RTOL, ArrayMesh, assert_close_or_conditioned, assert_prec_close_or_conditioned, block_scaled_error,   # noqa: E402
                     diag_conditioning, extended_truth, extended_truth_prec, random_tet_mesh, row_scaled_error)
from oracle_lib import Oracle                                                       # noqa: E402


SLIVER = "sliver"        # random 4-subsets of points: arbitrarily flat elements, judged against helpers.extended_truth


def check(oracle, name, setup, fused=True, rtol=RTOL, **ctx_args):
    want_v, want_p0, want_p = oracle.fem_iteration(setup)
    dim = setup.operatorDim
    ctx = mfb.Context(setup, path="ring", **ctx_args)
    if fused:
        ctx.iteration()
    else:
        ctx.assembly()
        ctx.prec_init()
        ctx.halo_exchange()
        ctx.prec_inversion()
    v, p = ctx.download()
    ev, ep = row_scaled_error(v, want_v, setup.row, dim), block_scaled_error(p, want_p, dim)
    if fused:                               # bit-reproducible: the plan fixes the summation order
        ctx.iteration()
        v2, p2 = ctx.download()
        assert np.array_equal(v, v2) and np.array_equal(p, p2, equal_nan=True), f"{name}: second iteration differs"
    launches = ctx.launch_count()
    ctx.close()
    print(f"ring {name}: values {ev:.2e} prec {ep:.2e} launches {launches}", flush=True)
    assert launches > 0
    if rtol == SLIVER:
        truth = extended_truth(setup)
        assert_close_or_conditioned(v, want_v, truth, setup.row, dim, name)
        assert_prec_close_or_conditioned(p, want_p, extended_truth_prec(setup, truth), dim, name, rho=diag_conditioning(setup, want_v))
    else:
        assert ev <= rtol and ep <= rtol, f"{name}: differs from the oracle"


def main():
    if mfb.device_count() < 1:
        sys.exit("ring_gpu_worker: no CUDA device")
    oracle = Oracle()
    rng = np.random.default_rng(21)
    if os.environ.get("RING_WORKER_SHAPES_ONLY"):                      # development aid: the CTA shapes only
        mesh = mfb.Mesh.generate(14, 12, 10, seed=9)
        for op in ("ela", "lap"):
            for threads in (384, 640, 768, 896, 1024):
                check(oracle, f"{op} {threads} threads", mfb.Setup(mesh, op), threads=threads)
        print("RING_GPU_OK")
        return
    for op in ("ela", "lap"):
        for grid, seed in (((1, 1, 1), 1), ((5, 4, 3), 2), ((16, 9, 12), 3), ((25, 25, 40), 4)):
            mesh = mfb.Mesh.generate(*grid, seed=seed)
            for fused in (True, False):
                check(oracle, f"{op} kuhn{grid} fused={fused}", mfb.Setup(mesh, op), fused)
        mesh = mfb.Mesh.generate(14, 12, 10, seed=9)
        check(oracle, f"{op} small tiles", mfb.Setup(mesh, op), tile_rows=7, tile_elems=120)
        check(oracle, f"{op} large tiles", mfb.Setup(mesh, op), tile_rows=64, tile_elems=960)
        check(oracle, f"{op} 384 threads (two CTAs per SM)", mfb.Setup(mesh, op), threads=384)
        check(oracle, f"{op} 384 threads, small caps", mfb.Setup(mesh, op), threads=384, tile_rows=22, tile_elems=352)
        check(oracle, f"{op} 1024 threads (warpgroups with their own register counts)", mfb.Setup(mesh, op), threads=1024)
        check(oracle, f"{op} 1024 threads, small caps, staged", mfb.Setup(mesh, op), False, threads=1024, tile_rows=22, tile_elems=352)
        check(oracle, f"{op} 896 threads (16 job warps at 80 + 12 write-out warps at 56 registers)", mfb.Setup(mesh, op), threads=896)
        check(oracle, f"{op} 640 threads", mfb.Setup(mesh, op), threads=640, tile_rows=40, tile_elems=700)
        check(oracle, f"{op} one CTA", mfb.Setup(mesh, op), ctas=1)
        check(oracle, f"{op} one tile per CTA", mfb.Setup(mesh, op), ctas=-1)
        check(oracle, f"{op} plan order", mfb.Setup(mesh, op), bank_aware=False)
        coord, e2n = random_tet_mesh(rng, 60, 150)
        codes = rng.choice([0, 0, 0, 52, 53, 54, 10], size=60).astype(np.int32)
        check(oracle, f"{op} random tets", mfb.Setup(ArrayMesh(coord, e2n, 60, codes), op), rtol=SLIVER, tile_rows=8, tile_elems=400)
        coord, e2n = random_tet_mesh(rng, 25, 400)
        check(oracle, f"{op} dense random tets (chains with breaks)", mfb.Setup(ArrayMesh(coord, e2n, 25), op), rtol=SLIVER, tile_rows=25, tile_elems=640)
    # four subdomains on one GPU, interface values through the host halves of the exchange: interface rows
    # leave the fused kernel raw, everything else inverted (as test_gpu_multirank.py does for TILED)
    grid, blocks = (9, 8, 7), (2, 2, 1)
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=5) for r in range(4)]
    setups = [mfb.Setup(m, "ela") for m in meshes]
    precs = [np.ascontiguousarray(oracle.fem_iteration(s)[1]) for s in setups]
    oracle.halo_exchange(precs, [m.intfIndex for m in meshes], [m.intfNodes for m in meshes], [m.neighborsList for m in meshes], 9)
    want = [oracle.prec_inversion(precs[r], s.row, s.col, s.checkBounds, s.mesh.nbNodes, 1) for r, s in enumerate(setups)]
    ctxs = [mfb.Context(s, path="ring", nbBlocks=4, rank=r, tile_rows=16, tile_elems=260) for r, s in enumerate(setups)]
    for c in ctxs:
        c.assembly_fused()
    send = [c.halo_pack_host() for c in ctxs]
    for r, (c, m) in enumerate(zip(ctxs, meshes)):
        recv = np.zeros_like(send[r])
        for i in range(m.nbIntf):
            src = int(m.neighborsList[i]) - 1
            o = meshes[src]
            q = [k for k in range(o.nbIntf) if o.neighborsList[k] - 1 == r][0]
            recv[m.intfIndex[i] * 9:m.intfIndex[i + 1] * 9] = send[src][o.intfIndex[q] * 9:o.intfIndex[q + 1] * 9]
        c.halo_add_host(recv)
    for r, c in enumerate(ctxs):
        c.prec_inversion_interface()
        v, p = c.download()
        ev = row_scaled_error(v, oracle.fem_iteration(setups[r])[0], setups[r].row, 9)
        ep = block_scaled_error(p, want[r], 9)
        print(f"ring 4 subdomains, rank {r}: values {ev:.2e} prec {ep:.2e}", flush=True)
        assert ev <= RTOL and ep <= RTOL
        c.close()
    # EIB size through a property: the element matrices have zero row sums, so every block row of the
    # assembled matrix sums to zero, and RING and TILED agree entry by entry
    mesh = mfb.Mesh.generate(100, 100, 100, seed=1)
    setup = mfb.Setup(mesh, "ela")
    ring = mfb.Context(setup, path="ring")
    ring.iteration()
    v_ring, p_ring = ring.download()
    ms = ring.run_timed(20) / 20
    ring.close()
    tiled = mfb.Context(setup, path="tiled")
    tiled.iteration()
    v_tiled, p_tiled = tiled.download()
    ms_tiled = tiled.run_timed(20) / 20
    tiled.close()
    ev, ep = row_scaled_error(v_ring, v_tiled, setup.row, 9), block_scaled_error(p_ring, p_tiled, 9)
    print(f"ring EIB ela vs tiled: values {ev:.2e} prec {ep:.2e}; ring {ms:.3f} ms, tiled {ms_tiled:.3f} ms per iteration", flush=True)
    assert ev <= RTOL and ep <= RTOL
    print("RING_GPU_OK")


if __name__ == "__main__":
    main()
