// TEST AID: runs the RING kernel's source on host threads (tools/ring_kernel_host.cc) from an executable
// built with -fsanitize=thread, so that ThreadSanitizer checks the kernel's synchronisation protocol.
//   make -C mini-fem_b200 ringkernel-tsan && OMP_NUM_THREADS=1 tools/ring_kernel_tsan [grid rows entries ctas]
// prints TSAN_RUN_DONE; any "WARNING: ThreadSanitizer: data race" above it is a bug in the kernel.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mini-fem_b200/host/mesh_data.h"
#include "../mini-fem_b200/host/mesh_topology.h"
extern "C" int mfb_ring_kernel_host (int, int, int, const int*, const int*, const int*, const double*, const int*, const uint8_t*, int, int, int, int, double*, double*, int);
extern "C" const char *mfb_ring_kernel_host_error (void);
using namespace mfb;
int main (int argc, char **argv)
{
    int gx = argc > 1 ? atoi (argv[1]) : 7, rows = argc > 2 ? atoi (argv[2]) : 0, entries = argc > 3 ? atoi (argv[3]) : 0, ctas = argc > 4 ? atoi (argv[4]) : 2;
    SubMesh m; generate_block (gx, gx - 1, gx - 2, 1, 1, 1, 0, 3, m);
    std::vector<int> row (m.nbNodes + 1), col ((size_t)count_csr_entries (m.elemToNode.data (), m.nbElem, m.nbNodes));
    build_csr (m.elemToNode.data (), m.nbElem, m.nbNodes, row.data (), col.data ());
    std::vector<int> cb ((size_t)m.nbNodes * 3);
    boundary_mask (m.boundNodesCode.data (), m.nbNodes, cb.data ());
    for (int op = 0; op < 2; op++) {
        const int dim = op ? 9 : 1;
        std::vector<double> values ((size_t)row[m.nbNodes] * dim, NAN), prec ((size_t)m.nbNodes * dim, NAN);
        int rc = mfb_ring_kernel_host (op, m.nbNodes, m.nbElem, m.elemToNode.data (), row.data (), col.data (), m.coord.data (), cb.data (), nullptr,
                                       rows, entries, ctas, 1, values.data (), prec.data (), argc > 5 ? atoi (argv[5]) : 256);
        if (rc) { printf ("error: %s\n", mfb_ring_kernel_host_error ()); return 1; }
        size_t nans = 0; for (double v : values) nans += std::isnan (v); for (double v : prec) nans += std::isnan (v);
        printf ("op %d: %d nodes %d elems, NaNs left %zu\n", op, m.nbNodes, m.nbElem, nans);
    }
    printf ("TSAN_RUN_DONE\n");
}
