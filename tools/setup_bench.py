"""Setup cost of the layouts, host builders vs the GPU builders (host<->device copies included).
usage: python tools/setup_bench.py [n] [--shuffle]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mini-fem_b200", "python"))
import minifem_b200 as mfb

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
mesh = mfb.Mesh.generate(n, n, n, seed=1)
e2n, nbNodes = mesh.elemToNode.copy(), mesh.nbNodes
if "--shuffle" in sys.argv:
    rng = np.random.default_rng(0)
    e2n = np.ascontiguousarray(e2n.reshape(-1, 4)[rng.permutation(e2n.size // 4)].ravel())

def best(f, reps=3):
    out, ts = None, []
    for _ in range(reps):
        t = time.perf_counter(); out = f(); ts.append(time.perf_counter() - t)
    return out, min(ts)

mfb.device_create_nodeToNode(e2n[:4000], nbNodes)      # CUDA context
(row_h, col_h), t_csr_h = best(lambda: mfb.create_nodeToNode(e2n, nbNodes))
(row_d, col_d), t_csr_d = best(lambda: mfb.device_create_nodeToNode(e2n, nbNodes))
assert np.array_equal(row_h, row_d) and np.array_equal(col_h, col_d)
e2e_h, t_e2e_h = best(lambda: mfb.create_elemToEdge(row_h, col_h, e2n))
e2e_d, t_e2e_d = best(lambda: mfb.device_create_elemToEdge(row_h, col_h, e2n))
assert np.array_equal(e2e_h, e2e_d)
col_h_, t_col_h = best(lambda: mfb.coloring_creation(e2n, nbNodes))
col_d_, t_col_d = best(lambda: mfb.device_coloring_creation(e2n, nbNodes))
assert all(np.array_equal(a, b) for a, b in zip(col_h_[:3], col_d_[:3]))
print(f"{n}^3 cubes, {e2n.size // 4} tets{' (shuffled)' if '--shuffle' in sys.argv else ''}: "
      f"CSR host {t_csr_h*1e3:.1f} ms gpu {t_csr_d*1e3:.1f} ms | elemToEdge host {t_e2e_h*1e3:.1f} ms gpu {t_e2e_d*1e3:.1f} ms | "
      f"colouring ({col_h_[3]} colours) host {t_col_h*1e3:.1f} ms gpu {t_col_d*1e3:.1f} ms")
