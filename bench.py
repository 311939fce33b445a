#!/usr/bin/env python
"""Benchmark of the Mini-FEM assembly hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

A "step" is one iteration of FEM_loop (src/FEM.cc:177-257): assembly + preconditioner
initialisation + halo sum + preconditioner inversion over one subdomain per GPU.  The
workload at N = 1 is the EIB-like elasticity case BASELINE.json quotes its metric on
(100^3 cubes x 6 tetrahedra: 1,030,301 nodes, 6,000,000 elements, 15,210,901 CSR blocks of
3x3); for N > 1 every GPU gets one such block of an N-times larger mesh (weak scaling) and
only interface preconditioner values cross NVLink (NCCL), as in the reference's domain
decomposition.  Prints ONE JSON line (rank 0).

The measured path is RING (mini-fem_b200/csrc/kernels_ring.cu), the default write-once path; the line also
carries, outside every timed region: `parity` (values and prec of the measured contexts against the CPU
oracle on the same subdomains, at every N — at N > 1 the oracle's interface values travel through
torch.distributed in the message pattern of halo.cc), `strong` (N > 1: the 100^3 mesh cut into N
subdomains, next to the same mesh on one GPU), `configs` (N = 1: the other configurations of
BASELINE.json — LM6 lap / ela, EIB lap, 200^3 ela — and an unstructured Delaunay mesh, every path),
`other_paths`, and the CPU baselines (the reference's own FEM_loop from oracle/_ref).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mini-fem_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

METRIC = "EIB ela assembly+precond elements/s"      # BASELINE.json's metric (--op lap renames it)
UNIT = "elements/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=[100, 100, 100], help="cubes per GPU block")
    ap.add_argument("--op", default="ela", choices=["ela", "lap"])
    ap.add_argument("--path", default="ring", choices=["ring", "tiled", "atomic", "color"])
    ap.add_argument("--configs", default="all", choices=["all", "fast", "none"],
                    help="N = 1: also time the other BASELINE.json configurations (fast: without the 200^3 and Delaunay meshes)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--tile-elems", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-paths", action="store_true")
    ap.add_argument("--cpu-ranks", type=int, default=0, help="ranks of the CPU reference (0 = host cores)")
    return ap.parse_args()


def algorithmic_bytes(op, E, Z, N):
    """every input read once, every output written once."""
    return 16 * E + 76 * Z + 112 * N if op == "ela" else 16 * E + 12 * Z + 36 * N


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(grid, op, E, N, Z):
    return (f"EIB-like Kuhn mesh, {grid[0]}x{grid[1]}x{grid[2]} cubes x 6 tets per GPU "
            f"({E} elements, {N} nodes, {Z} CSR entries per GPU), operator {op}")


def global_layout(args, world):
    """Grid of cubes of the whole job and its block partition."""
    import minifem_b200 as mfb
    if args.scaling == "strong" or world == 1:
        grid = tuple(args.grid)
        return grid, mfb.choose_blocks(*grid, world)
    px, py, pz = mfb.choose_blocks(world, world, world, world)     # most cubic factorisation of N
    return (args.grid[0] * px, args.grid[1] * py, args.grid[2] * pz), (px, py, pz)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.samples, self.stop_flag, self.window = [], False, [0.0, 0.0]
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.replace(",", "").isdigit() else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                   # pragma: no cover
            self.err = str(e)

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if self.window[0] <= s[0] <= self.window[1]] or self.samples
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown}
        seen = sorted(n for n, bit in names.items() if any(s[2] & bit for s in inside))
        mhz = sorted(s[1] for s in inside)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": seen,
                "samples_in_timed_region": len([s for s in self.samples if self.window[0] <= s[0] <= self.window[1]])}


def load_traffic(key):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------ reference arm (CPU)

class RefSetup:
    """What main.cc builds between read_input_data and FEM_loop, by the reference's own compiled functions
    (oracle/_ref: create_nodeToNode, create_elemToEdge, dqmrd4_ / e_essbcm_) — the reference arm uses this
    repository's library for nothing but generating the synthetic mesh."""

    def __init__(self, ref, mesh, op, elem_to_edge=False):
        self.mesh = mesh
        self.operatorID = {"lap": 0, "ela": 1}[op]
        self.operatorDim = 1 if self.operatorID == 0 else 9
        self.elemToNode = mesh.elemToNode.copy()
        self.colorToElem, self.nbTotalColors, self.colorPerm = None, 0, None
        self.row, self.col = ref.create_nodeToNode(self.elemToNode, mesh.nbNodes)
        self.nbEdges = int(self.row[-1])
        self.elemToEdge = ref.create_elemToEdge(self.row, self.col, self.elemToNode) if elem_to_edge else None
        self.checkBounds = ref.boundary_mask(mesh.boundNodesCode)


def run_reference_cpu(op, grid, nb_iter, ranks, kind="ref"):
    """The reference's own FEM_loop (oracle/_ref, REF build = pure MPI, one rank per
    subdomain, ranks as threads) on the EIB-like mesh cut into `ranks` blocks.  kind = "ref_opt": the same
    with -DOPTIMIZED (elemToEdge instead of the row search, build/CMakeLists.txt:67-69).  Falls back
    to the C oracle port when oracle/_ref was not built."""
    import minifem_b200 as mfb
    from oracle_lib import Oracle, Reference, ref_available
    blocks = mfb.choose_blocks(*grid, ranks)
    n = blocks[0] * blocks[1] * blocks[2]
    meshes = [mfb.Mesh.generate(*grid, blocks=blocks, rank=r, seed=1) for r in range(n)]
    elements = sum(m.nbElem for m in meshes)
    if ref_available(kind):
        ref = Reference(kind)
        setups = [RefSetup(ref, m, op, elem_to_edge=kind.endswith("_opt")) for m in meshes]
        _, _, cycles, hz = ref.fem_loop(setups, nb_iter)
        seconds = sum(cycles) / hz                                   # FEM.cc:132: total of the 4 stage averages
        detail = (f"oracle/_ref libminifem_ref_{kind}.so (src/*.cc, -O2 -mavx, REF build, " +
                  ("-DOPTIMIZED: elemToEdge)" if kind.endswith("_opt") else "search path as shipped)"))
        kind = "reference"
    else:
        oracle = Oracle()
        setups = [mfb.Setup(m, op) for m in meshes]
        t0 = time.perf_counter()
        for s in setups:
            oracle.fem_iteration(s)
        seconds = (time.perf_counter() - t0)
        n, kind, detail = 1, "port", "oracle/minifem_oracle.c, one thread"
    return elements / seconds, n, kind, detail, elements, blocks


def run_reference_coloring(op, grid, nb_iter, kind="coloring"):
    """The reference's COLORING build (MPI + OpenMP, one rank, OMP_NUM_THREADS = host cores) as
    shipped: the per-colour `#pragma omp parallel for` is commented out in the reference
    (src/assembly.cc:362,516), so only the zero-fill and the preconditioner loops are threaded."""
    import minifem_b200 as mfb
    from oracle_lib import Reference, ref_available
    if not ref_available(kind):
        return None
    mesh = mfb.Mesh.generate(*grid, seed=1)
    setup = mfb.Setup(mesh, op, coloring=True)          # colours + permutation: bit-identical to coloring.cc (tested)
    ref = Reference(kind)
    ref.set_colors(setup.colorToElem)
    _, _, cycles, hz = ref.fem_loop([setup], nb_iter)
    return mesh.nbElem / (sum(cycles) / hz), setup.nbTotalColors


def reference_main(args, rank, world):
    if rank != 0:
        return
    cores = args.cpu_ranks or (os.cpu_count() or 1)
    grid = tuple(args.grid)
    one_iter_guess = 6.0e6 * (np.prod(grid) / 1e6) / 4.0e6 / max(min(cores, 64), 1) * 3
    nb_timed = max(1, min(args.steps, int(60.0 / max(one_iter_guess, 1e-3))))
    import minifem_b200 as mfb
    whole = mfb.Mesh.generate(*grid, seed=1)                      # header counts of the undivided mesh
    t0 = time.perf_counter()
    value, n, kind, detail, elements, blocks = run_reference_cpu(args.op, grid, nb_timed + 1, min(cores, 64))
    # the same element count as the GPU arm at N > 1 is not attempted: the reference arm always runs ONE mesh of the
    # --grid size on the host cores (rates are comparable, the work is not) — said in config.arm
    wall = time.perf_counter() - t0
    sample = (f"{grid[0]}x{grid[1]}x{grid[2]}-cube EIB-like mesh ({elements} elements) in {blocks[0]}x{blocks[1]}x{blocks[2]} "
              f"subdomains, {nb_timed} timed iterations after 1 untimed (FEM.cc:182); {detail}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": nb_timed, "warmup": 1, "ms_per_step": 1e3 * elements / value, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(grid, args.op, whole.nbElem, whole.nbNodes, whole.nbEdges),
                       "arm": "the reference's CPU implementation on the host cores (no GPU work); always ONE mesh of the --grid size, "
                              "whatever --gpus says (the GPU arm's weak scaling multiplies the mesh by N)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ checks and side measurements (GPU)

def oracle_parity(ctx, setup, mesh, world):
    """values and prec held by `ctx` (its last iteration) against the CPU oracle on the same subdomain
    (oracle/minifem_oracle.c, bit-equal to the reference's compiled sources: tests/test_oracle_vs_reference.py).
    At N > 1 the oracle's prec_init values of the interface nodes are exchanged between the ranks through
    torch.distributed in the message pattern of MPI_halo_exchange (halo.cc:52-116) and added in its order.
    Outside every timed region.  Returns the max over ranks."""
    from minifem_b200 import dist as mdist
    from helpers import RTOL, block_scaled_error, row_scaled_error
    from oracle_lib import Oracle
    oracle = Oracle()
    dim = setup.operatorDim
    t0 = time.perf_counter()
    want_v = oracle.assembly(mesh.coord, setup.row, setup.col, setup.elemToNode, setup.operatorID, setup.elemToEdge, setup.colorToElem)
    p0 = np.ascontiguousarray(oracle.prec_init(want_v, setup.row, setup.col, mesh.nbNodes, dim))
    if world > 1 and mesh.nbIntfNodes > 0:
        nodes = np.asarray(mesh.intfNodes, np.int64) - 1
        send = np.ascontiguousarray(p0.reshape(-1, dim)[nodes]).ravel()          # halo.cc:77-80, pre-exchange values
        recv = mdist.exchange_host(send, mesh.intfIndex, mesh.neighborsList, dim)
        np.add.at(p0.reshape(-1, dim), nodes, recv.reshape(-1, dim))             # halo.cc:113-116, in list order
    want_p = oracle.prec_inversion(p0, setup.row, setup.col, setup.checkBounds, mesh.nbNodes, setup.operatorID)
    v, p = ctx.download()
    ev = mdist.max_over_ranks(row_scaled_error(v, want_v, setup.row, dim))
    ep = mdist.max_over_ranks(block_scaled_error(p, want_p, dim))
    return {"values_err": ev, "prec_err": ep, "rtol": RTOL, "ok": bool(ev <= RTOL and ep <= RTOL), "ranks": world,
            "against": "oracle/minifem_oracle.c on every rank's own subdomain, entry by entry (row / block scaled, tests/helpers.py)"
                       + ("; interface prec summed across ranks as in halo.cc:52-116" if world > 1 else ""),
            "seconds": round(time.perf_counter() - t0, 1)}


def time_paths(mfb, mesh, op, paths, device, steps=20, peak=None):
    """ms per fused iteration of each path on one GPU for one mesh (side measurements of the `configs` record)."""
    out = {}
    E, N = mesh.nbElem, mesh.nbNodes
    for path in paths:
        try:
            t0 = time.perf_counter()
            setup = mfb.Setup(mesh, op, coloring=(path == "color"))
            ctx = mfb.Context(setup, path=path, device=device, use_graph=(path in ("color", "blockcolor")))
            build_s = time.perf_counter() - t0
            for _ in range(3):
                ctx.iteration()
            ctx.sync()
            ms = ctx.run_timed(steps) / steps
            alg = algorithmic_bytes(op, E, setup.nbEdges, N)
            out[path] = {"ms_per_step": ms, "value": E / (ms * 1e-3), "frac": alg / (ms * 1e-3) / 1e9 / peak if peak else None,
                         "setup_s": round(build_s, 2)}
            if path == "color":
                out[path]["colors"] = setup.nbTotalColors
            if path == "blockcolor":                       # locality-blocked colouring: launches = block colours
                out[path].update({k: int(v) for k, v in ctx.plan_stats().items()})
            ctx.close()
        except Exception as e:                             # noqa: BLE001 (a side measurement must not take the bench line with it)
            out[path] = {"error": repr(e)[:200]}
    return out


def other_configs(mfb, args, device, peak):
    """BASELINE.json's configurations 1, 2, 3, 5 and an unstructured mesh, N = 1, every path, with the reference's
    CPU builds where BASELINE.json names them."""
    from oracle_lib import Reference, ref_available
    out = {}
    paths = ["ring", "tiled", "atomic", "color", "blockcolor"]
    cores = os.cpu_count() or 1
    lm6 = mfb.Mesh.generate(25, 25, 40, seed=1)
    for op, key in (("lap", "1_LM6_lap"), ("ela", "2_LM6_ela")):
        rec = {"workload": workload_text((25, 25, 40), op, lm6.nbElem, lm6.nbNodes, lm6.nbEdges),
               "gpu": time_paths(mfb, lm6, op, paths, device, 50, peak)}
        try:
            if ref_available("ref"):
                ref = Reference("ref")
                _, _, cycles, hz = ref.fem_loop([RefSetup(ref, lm6, op)], 10)
                rec["cpu_ref_np1"] = {"value": lm6.nbElem / (sum(cycles) / hz), "unit": UNIT, "cores": 1,
                                      "sample": "mpirun -np 1 equivalent: one rank thread, 10 iterations (9 timed, FEM.cc:182), oracle/_ref REF build"}
            col = run_reference_coloring(op, (25, 25, 40), 10)
            if col:
                rec["cpu_coloring"] = {"value": col[0], "unit": UNIT, "cores": cores, "colors": col[1],
                                       "sample": "one rank, OMP_NUM_THREADS = host cores, 10 iterations; COLORING build as shipped"}
        except Exception as e:                                   # noqa: BLE001
            rec["cpu_error"] = repr(e)[:200]
        out[key] = rec
    eib = mfb.Mesh.generate(*args.grid, seed=1)
    out["3_EIB_lap"] = {"workload": workload_text(args.grid, "lap", eib.nbElem, eib.nbNodes, eib.nbEdges),
                        "gpu": time_paths(mfb, eib, "lap", paths, device, 20, peak),
                        "note": "colour-by-colour (coloring.cc's colours, one launch per colour in a CUDA graph) against native FP64 atomics; "
                                "ncu DRAM figures of both kernels: profiles/r2_scatter_ncu.txt"}
    del eib
    if args.configs == "all":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from delaunay_mesh import delaunay_arrays
            from helpers import ArrayMesh
            coord, e2n, codes = delaunay_arrays(200000, seed=4)
            dm = ArrayMesh(coord, e2n, coord.size // 3, codes)
            g = time_paths(mfb, dm, "ela", ["ring", "tiled", "atomic"], device, 20, peak)
            out["unstructured_delaunay_200k_ela"] = {"workload": f"Delaunay tetrahedralisation of 200,000 random points ({dm.nbElem} elements), elasticity", "gpu": g}
        except Exception as e:                                   # noqa: BLE001
            out["unstructured_delaunay_200k_ela"] = {"error": repr(e)[:200]}
        try:
            t0 = time.perf_counter()
            big = mfb.Mesh.generate(200, 200, 200, seed=1)
            g = time_paths(mfb, big, "ela", ["ring"], device, 10, peak)
            out["5_8M_ela_1gpu"] = {"workload": workload_text((200, 200, 200), "ela", big.nbElem, big.nbNodes, big.nbEdges), "gpu": g,
                                    "wall_s": round(time.perf_counter() - t0, 1),
                                    "note": "the 1-GPU point of the 48 M-tetrahedra sweep; its 8-GPU point is this bench at --gpus 8 "
                                            "(2 x 2 x 2 blocks of 100^3 cubes = the same 200^3 mesh)"}
        except Exception as e:                                   # noqa: BLE001
            out["5_8M_ela_1gpu"] = {"error": repr(e)[:200]}
    return out


# ------------------------------------------------------------------------ our arm (GPU)

class Watchdog(threading.Thread):
    """A step that makes no progress for `limit` seconds ends the process with a message instead of hanging the
    box (a multi-GPU exchange that never completes would otherwise sit there until the driver's own limit)."""

    def __init__(self, limit=420.0):
        super().__init__(daemon=True)
        self.limit, self.last, self.what = limit, time.monotonic(), "start"

    def stage(self, what):
        self.last, self.what = time.monotonic(), what

    def run(self):
        while True:
            time.sleep(2.0)
            if time.monotonic() - self.last > self.limit:
                print(f"bench.py: no progress for {self.limit:.0f} s in stage '{self.what}' (rank {os.environ.get('RANK', '0')}); giving up",
                      file=sys.stderr, flush=True)
                os._exit(3)


def connect_halo(mdist, ctx, path):
    """RING at N > 1: the peer-to-peer windows (mfb_ctx_p2p_*), all ranks or none; every other path keeps NCCL."""
    if path != "ring":
        return {"transport": "nccl", "why": "path " + path}
    active, why = mdist.p2p_connect(ctx)
    return {"transport": "p2p"} if active else {"transport": "nccl", "why": why}


def warm_up(mfb, mdist, ctx, steps, halo):
    """Untimed iterations.  A peer-to-peer exchange whose bounded waits run out on any rank (MFB_ERR_COMM from sync)
    is switched off on every rank and the warm-up repeated over NCCL."""
    ok = 1.0
    try:
        for _ in range(steps):
            ctx.iteration()
        ctx.sync()
    except mfb.MfbError as e:
        if not (halo and halo.get("transport") == "p2p"):
            raise
        ok, halo["why"] = 0.0, str(e)[:200]
    if halo and halo.get("transport") == "p2p" and -mdist.max_over_ranks(-ok) < 1.0:
        halo["transport"] = "nccl"
        halo.setdefault("why", "the peer-to-peer exchange timed out on another rank")
        ctx.p2p_enable(False)
        for _ in range(steps):
            ctx.iteration()
        ctx.sync()


def main():
    global METRIC
    args = parse_args()
    dog = Watchdog()
    dog.start()
    if args.op == "lap":
        METRIC = "EIB lap assembly+precond elements/s"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_main(args, rank, world)
        return
    if args.gpus != world:
        if args.gpus > 1:
            sys.exit(f"bench.py --gpus {args.gpus}: launch with `python -m torch.distributed.run --nproc-per-node {args.gpus} "
                     f"--master-addr 127.0.0.1 ... bench.py --gpus {args.gpus}` (one process per GPU)")
    import torch
    import minifem_b200 as mfb
    from minifem_b200 import dist as mdist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mfb.device_count() < 1:
        sys.exit("bench.py: no CUDA device; the assembly path has no CPU fallback")
    torch.cuda.set_device(local)
    mdist.init_from_env("nccl")

    grid, blocks = global_layout(args, world)
    dog.stage("setup")
    t0 = time.perf_counter()
    mesh = mfb.Mesh.generate(*grid, blocks=blocks, rank=rank, seed=1)
    setup = mfb.Setup(mesh, args.op, coloring=(args.path == "color"))
    ctx = mfb.Context(setup, path=args.path, device=local, nbBlocks=world, rank=rank, tile_rows=args.tile_rows,
                      tile_elems=args.tile_elems, use_graph=True)
    halo = None
    if world > 1:
        mdist.comm_init(ctx)
        halo = connect_halo(mdist, ctx, args.path)
    setup_s = time.perf_counter() - t0
    E, N, Z = mesh.nbElem, mesh.nbNodes, setup.nbEdges
    total_elements = mdist.sum_over_ranks(E)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler and sampler.ok:
        sampler.start()

    dog.stage("warm-up and timed region")
    warm_up(mfb, mdist, ctx, max(args.warmup, 3), halo)
    launches_before = ctx.launch_count()
    mdist.barrier(); torch.cuda.synchronize()
    t_begin = time.perf_counter()
    ms_total = ctx.run_timed(args.steps)                 # CUDA events on the launching stream
    ctx.sync()
    torch.cuda.synchronize(); mdist.barrier()
    t_end = time.perf_counter()
    if sampler:
        sampler.window = [t_begin, t_end]
    launches = ctx.launch_count() - launches_before
    ms_step = mdist.max_over_ranks(ms_total) / args.steps
    value = total_elements / (ms_step * 1e-3)

    # end to end through the host-buffer entry point: H2D coord, iteration, D2H values + prec
    e2e, e2e_resident = None, None
    dog.stage("end-to-end legs")
    if args.e2e_steps > 0:
        pins = [mfb.PinnedArray(N * 3), mfb.PinnedArray(ctx.nbValues), mfb.PinnedArray(ctx.nbPrec)]
        pins[0].array[:] = mesh.coord
        ctx.iteration_host(pins[0].ptr, pins[1].ptr, pins[2].ptr)          # warm
        mdist.barrier(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.iteration_host(pins[0].ptr, pins[1].ptr, pins[2].ptr)
        torch.cuda.synchronize(); mdist.barrier()
        e2e_s = mdist.max_over_ranks((time.perf_counter() - t1) / args.e2e_steps)
        checksum = float(pins[2].array[:9].sum())
        e2e = {"value": total_elements / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(pins[0].bytes),
               "d2h_bytes_per_step": int(pins[1].bytes + pins[2].bytes), "ms_per_step": 1e3 * e2e_s,
               "steps": args.e2e_steps, "note": "per GPU: pinned coord -> device, fused iteration, values + prec -> pinned host",
               "prec_checksum": checksum}
        # The same step for a caller that keeps the matrix on the device (a solver on the GPU): coord
        # H2D, fused iteration, the two norms check_results looks at reduced on the device, 16 bytes
        # D2H.  Reported beside `e2e`, not instead of it.
        ctx.iteration_norms_host(pins[0].ptr)
        mdist.barrier(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            norms = ctx.iteration_norms_host(pins[0].ptr)
        torch.cuda.synchronize(); mdist.barrier()
        res_s = mdist.max_over_ranks((time.perf_counter() - t1) / args.e2e_steps)
        e2e_resident = {"value": total_elements / res_s, "unit": UNIT, "h2d_bytes_per_step": int(pins[0].bytes),
                        "d2h_bytes_per_step": 16, "ms_per_step": 1e3 * res_s, "steps": args.e2e_steps,
                        "note": "per GPU: pinned coord -> device, fused iteration, matrix and prec norms reduced on the "
                                "device (mfb_ctx_iteration_norms_host); the matrix stays in HBM",
                        "norms": [float(norms[0]), float(norms[1])]}
        for p in pins:
            p.free()
    if sampler and sampler.ok:
        sampler.stop_flag = True
        sampler.join()

    stats = ctx.plan_stats() if args.path in ("tiled", "ring") else None
    mesh_bytes, plan_bytes = ctx.device_bytes()
    parity = None
    dog.stage("parity against the oracle")
    if not args.no_parity:
        ctx.iteration()
        ctx.sync()
        parity = oracle_parity(ctx, setup, mesh, world)
        if halo and halo["transport"] == "p2p" and not parity["ok"]:
            # the peer-to-peer exchange produced a wrong interface sum on this box: report it, fall back to NCCL for
            # the number of record (every rank sees the same all-reduced verdict)
            halo = {"transport": "nccl", "why": "peer-to-peer exchange failed the oracle check: " + json.dumps(parity)[:200]}
            ctx.p2p_enable(False)
            warm_up(mfb, mdist, ctx, 3, halo)
            mdist.barrier(); torch.cuda.synchronize()
            ms_step = mdist.max_over_ranks(ctx.run_timed(args.steps)) / args.steps
            value = total_elements / (ms_step * 1e-3)
            ctx.iteration(); ctx.sync()
            parity = oracle_parity(ctx, setup, mesh, world)
    if halo and halo["transport"] == "p2p":
        # the same iterations with the NCCL exchange, for the record
        ctx.p2p_enable(False)
        for _ in range(3):
            ctx.iteration()
        ctx.sync()
        mdist.barrier(); torch.cuda.synchronize()
        halo["nccl_ms_per_step"] = mdist.max_over_ranks(ctx.run_timed(args.steps)) / args.steps
        ctx.p2p_enable(True)
    ctx.close()

    # strong scaling (N > 1): the N = 1 workload itself — the 100^3 mesh — cut into N subdomains, and, for the
    # single-GPU time of the same box, the undivided mesh on every GPU at once (no communication)
    strong = None
    dog.stage("strong scaling")
    if world > 1 and not args.no_strong and args.scaling == "weak":
        sgrid = tuple(args.grid)
        sblocks = mfb.choose_blocks(*sgrid, world)
        smesh = mfb.Mesh.generate(*sgrid, blocks=sblocks, rank=rank, seed=1)
        ssetup = mfb.Setup(smesh, args.op)
        sctx = mfb.Context(ssetup, path=args.path, device=local, nbBlocks=world, rank=rank, use_graph=True)
        mdist.comm_init(sctx)
        shalo = connect_halo(mdist, sctx, args.path)
        warm_up(mfb, mdist, sctx, 5, shalo)
        mdist.barrier(); torch.cuda.synchronize()
        s_ms = mdist.max_over_ranks(sctx.run_timed(args.steps)) / args.steps
        sctx.iteration(); sctx.sync()
        s_par = None if args.no_parity else oracle_parity(sctx, ssetup, smesh, world)
        if shalo["transport"] == "p2p":
            sctx.p2p_enable(False)
            for _ in range(3):
                sctx.iteration()
            sctx.sync()
            mdist.barrier(); torch.cuda.synchronize()
            s_nccl = mdist.max_over_ranks(sctx.run_timed(args.steps)) / args.steps
            shalo["nccl_ms_per_step"] = s_nccl
            if s_par is not None and not s_par["ok"]:
                shalo = {"transport": "nccl", "why": "peer-to-peer exchange failed the oracle check"}
                s_ms = s_nccl
                sctx.iteration(); sctx.sync()
                s_par = oracle_parity(sctx, ssetup, smesh, world)
        sctx.close()
        whole = mfb.Mesh.generate(*sgrid, seed=1)
        wsetup = mfb.Setup(whole, args.op)
        wctx = mfb.Context(wsetup, path=args.path, device=local)
        for _ in range(5):
            wctx.iteration()
        wctx.sync()
        mdist.barrier(); torch.cuda.synchronize()
        w_ms = mdist.max_over_ranks(wctx.run_timed(args.steps)) / args.steps
        wctx.close()
        strong = {"workload": workload_text(sgrid, args.op, whole.nbElem, whole.nbNodes, whole.nbEdges) + f", cut into {sblocks[0]}x{sblocks[1]}x{sblocks[2]} subdomains",
                  "n_gpus": world, "ms_per_step": s_ms, "value": whole.nbElem / (s_ms * 1e-3), "unit": UNIT,
                  "one_gpu_ms_per_step": w_ms, "speedup": w_ms / s_ms, "parallel_efficiency": w_ms / s_ms / world,
                  "parity": s_par, "halo": shalo,
                  "note": "device time (CUDA events), max over ranks; one_gpu = the undivided mesh on every GPU of this run at once"}

    peak, peak_src = measured_peak()
    dog.stage("other paths and configurations")
    dog.limit = 900.0
    other, configs = {}, None
    if world == 1 and not args.no_other_paths:
        other = time_paths(mfb, mesh, args.op, [q for q in ("ring", "tiled", "atomic", "color", "blockcolor") if q != args.path], local, 10, peak)
    if world == 1 and args.configs != "none" and rank == 0:
        try:
            configs = other_configs(mfb, args, local, peak)
        except Exception as e:                                   # noqa: BLE001 (the bench line must still appear)
            configs = {"error": repr(e)[:300]}

    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()
    if rank != 0:
        return
    alg = algorithmic_bytes(args.op, E, Z, N)
    achieved = alg / (ms_step * 1e-3) / 1e9
    workload = workload_text(args.grid, args.op, E, N, Z)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "path": args.path, "global_grid": list(grid), "blocks": list(blocks),
                       "parallelism": (f"dd{world} (one subdomain per GPU, interface sum over "
                                       + ("NVLink peer windows, csrc/kernels_halo_p2p.cu" if halo and halo["transport"] == "p2p" else "NCCL send / recv") + ")") if world > 1 else "single subdomain",
                       "halo": halo,
                       "l2": "inputs + outputs per step (>= 1.5 GB for ela) exceed the 126 MB L2; no flush needed",
                       "setup_s": round(setup_s, 2), "device_mesh_bytes": mesh_bytes, "device_plan_bytes": plan_bytes,
                       "plan": stats},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(f"{args.op}_{args.path}_{args.grid[0]}"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "note": "one fused kernel launch per step at N=1; duration = CUDA events over the timed region / steps"},
            "e2e": e2e, "e2e_device_resident": e2e_resident, "gpu_launches": int(launches), "clocks": sampler.summary() if sampler else None,
            "parity": parity, "other_paths": other}
    if strong:
        line["strong"] = strong
    if configs:
        line["configs"] = configs
    dog.stage("CPU baselines")
    if world == 1 and not args.no_cpu_baseline:
        cores = args.cpu_ranks or (os.cpu_count() or 1)
        try:
            cv, n, kind, detail, elements, cb = run_reference_cpu(args.op, tuple(args.grid), 11, min(cores, 64))
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": n, "kind": kind,
                                    "sample": f"whole {elements}-element mesh in {cb[0]}x{cb[1]}x{cb[2]} subdomains (one rank thread each), "
                                              f"10 timed iterations after 1 untimed; {detail}"}
            ov, on, okind, odetail, _, _ = run_reference_cpu(args.op, tuple(args.grid), 11, min(cores, 64), kind="ref_opt")
            if okind == "reference":
                line["cpu_baseline_optimized"] = {"value": ov, "unit": UNIT, "cores": on, "kind": okind,
                                                  "sample": f"the same mesh and ranks, 10 timed iterations after 1 untimed; {odetail}"}
            col = run_reference_coloring(args.op, tuple(args.grid), 2)
            if col:
                line["cpu_baseline_coloring"] = {
                    "value": col[0], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                    "sample": f"whole mesh, one rank, OMP_NUM_THREADS = host cores, {col[1]} colours, 1 timed iteration after 1 untimed; "
                              "oracle/_ref libminifem_ref_coloring.so as shipped (per-colour omp pragma disabled in the reference, assembly.cc:362)"}
            mod = run_reference_coloring(args.op, tuple(args.grid), 3, kind="coloring_omp")
            if mod:
                line["cpu_baseline_coloring_modified"] = {
                    "value": mod[0], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference (modified)",
                    "sample": f"whole mesh, one rank, OMP_NUM_THREADS = host cores, {mod[1]} colours, 2 timed iterations after 1 untimed; "
                              "oracle/_ref libminifem_ref_coloring_omp.so = the COLORING build with the per-colour "
                              "`#pragma omp parallel for` of assembly.cc:362,516 restored (same bits as the shipped build, tested)"}
        except Exception as e:                                   # the bench line must still appear
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
