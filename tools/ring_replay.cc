// TEST AID, not product code: replays a RING plan (mini-fem_b200/host/ring_plan.h) on the host,
// lane by lane and in the kernel's order of operations, with the arithmetic header the kernel
// itself compiles (csrc/ring_math.h, csrc/device_math.cuh).  tests/test_ring_plan.py compares
// the result with the oracle, so the plan format, the edge-ring formulation and the row-sum
// diagonal are checked without a GPU.  Nothing in libminifem_b200.so links or calls this file;
// it is built into tools/libmfb_ringcheck.so by `make -C mini-fem_b200 ringcheck`.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../mini-fem_b200/csrc/device_math.cuh"
#include "../mini-fem_b200/csrc/ring_math.h"
#include "../mini-fem_b200/host/ring_plan.h"

using namespace mfb;

static std::string g_error;

template <int OPDIM>
static int replay (const RingPlan &plan, const double *coord, const int *checkBounds, int nbNodes, int fusePrec,
                   double *values, double *prec)
{
    const double kNaN = std::numeric_limits<double>::quiet_NaN ();
    constexpr int SLAB = OPDIM == 9 ? 10 : 1;          // doubles per slab entry, as in the kernel (ring_slab_stride)
    std::vector<double> X, Y, Z, slab, sDiag;
    for (int t = 0; t < plan.nbTiles; t++) {
        const uint8_t *base = plan.blob.data () + plan.tileOffset[t];
        const RingTileHeader &h = *plan.header (t);
        const RingRow *rows = reinterpret_cast<const RingRow*> (base + sizeof (RingTileHeader));
        const int *tileNodes = reinterpret_cast<const int*> (base + h.offNodes);
        const RingBatch *batches = reinterpret_cast<const RingBatch*> (base + h.headBytes);
        const uint64_t *jobs = reinterpret_cast<const uint64_t*> (base + h.offJobs);
        const uint64_t *codes = reinterpret_cast<const uint64_t*> (base + h.offCodes);
        X.resize (h.nbNodes); Y.resize (h.nbNodes); Z.resize (h.nbNodes);
        for (int n = 0; n < h.nbNodes; n++) {
            X[n] = coord[(size_t)tileNodes[n] * 3]; Y[n] = coord[(size_t)tileNodes[n] * 3 + 1]; Z[n] = coord[(size_t)tileNodes[n] * 3 + 2];
        }
        slab.assign ((size_t)h.nbEntries * SLAB, kNaN);       // a slot nobody writes must show up
        sDiag.assign ((size_t)h.nbRows * OPDIM, kNaN);
        // ---- job phase: one lane per job ---------------------------------------------------
        for (int b = 0; b < h.nbBatches; b++) {
            const RingBatch rb = batches[b];
            for (int lane = 0; lane < 32; lane++) {
                const uint64_t job = jobs[(size_t)b * 32 + lane];
                const int i = (int)(job & 0xFF), j = (int)((job >> 8) & 0xFF);
                const int sIJ = (int)((job >> 16) & 0xFFFF), sJI = (int)((job >> 32) & 0xFFFF);
                const double xi[3] = {X[i], Y[i], Z[i]};
                const double d[3] = {X[j] - xi[0], Y[j] - xi[1], Z[j] - xi[2]};
                const int len = (int)(job >> 48), nbSteps = rb.nbSteps;
                double acc[OPDIM], u[3] = {0, 0, 0};
                for (int k = 0; k < OPDIM; k++) acc[k] = 0.0;
                const uint64_t *cw = codes + rb.codeBase + lane;
                uint64_t word = 0;
                if (rb.flags == 0) {                                 // regular batch: no branch inside the step, padding masked
                    if (nbSteps > 0) {
                        word = cw[0];
                        const int id = (int)(word & 0xFF);
                        u[0] = X[id] - xi[0]; u[1] = Y[id] - xi[1]; u[2] = Z[id] - xi[2];
                    }
                    for (int k = 1; k < nbSteps; k++) {
                        if ((k & 7) == 0) word = cw[(size_t)(k >> 3) * 32]; else word >>= 8;
                        const int id = (int)(word & 0xFF);
                        const double w[3] = {X[id] - xi[0], Y[id] - xi[1], Z[id] - xi[2]};
                        ring_accumulate<OPDIM> (d, u, w, acc, k < len);
                        u[0] = w[0]; u[1] = w[1]; u[2] = w[2];
                    }
                }
                else {                                               // general batch: chains separated by breaks
                    bool have = false;
                    for (int k = 0; k < nbSteps; k++) {
                        if ((k & 7) == 0) word = cw[(size_t)(k >> 3) * 32]; else word >>= 8;
                        const int id = (int)(word & 0xFF);
                        if (k >= len) continue;
                        if (id == kRingBreak) { have = false; continue; }
                        const double w[3] = {X[id] - xi[0], Y[id] - xi[1], Z[id] - xi[2]};
                        if (have) ring_accumulate<OPDIM> (d, u, w, acc);
                        u[0] = w[0]; u[1] = w[1]; u[2] = w[2];
                        have = true;
                    }
                }
                if (sIJ == 0xFFFF) continue;
                if (OPDIM == 1) {
                    slab[sIJ] = acc[0];
                    if (sJI != 0xFFFF) slab[sJI] = acc[0];
                }
                else {
                    double k9[9];
                    ring_block (acc, k9);
                    for (int k = 0; k < 9; k++) {
                        slab[(size_t)sIJ * SLAB + k] = k9[k];
                        if (sJI != 0xFFFF) slab[(size_t)sJI * SLAB + ring_transposed (k)] = k9[k];
                    }
                }
            }
        }
        // ---- write-out: one warp per row; the diagonal entry is minus the sum of the run ----------
        for (int r = 0; r < h.nbRows; r++) {
            const RingRow rr = rows[r];
            const int len = rr.len, diagOff = rr.diagOff;
            double *out = values + (size_t)rr.valueStart * OPDIM;
            const double *src = slab.data () + (size_t)rr.localStart * SLAB;
            if (OPDIM == 1) {                                       // one lane walks the row
                double a = 0.0;
                for (int k = 0; k < len; k++) if (k != diagOff) { a += src[k]; out[k] = src[k]; }
                const double diag = 0.0 - a;
                if (diagOff != 0xFFFF) out[diagOff] = diag;
                sDiag[r] = diag;
            }
            else {
                for (int comp = 0; comp < 9; comp++) {              // one lane per component walks the row
                    double a = 0.0;
                    for (int k = 0; k < len; k++) {
                        if (k == diagOff) continue;
                        const double v = src[(size_t)k * SLAB + comp];
                        a += v;
                        out[(size_t)k * 9 + comp] = v;
                    }
                    const double diag = 0.0 - a;
                    if (diagOff != 0xFFFF) out[(size_t)diagOff * 9 + comp] = diag;
                    sDiag[(size_t)r * 9 + comp] = diag;
                }
            }
        }
        // ---- fused preconditioner: prec_init + prec_inversion of the owned rows (the mask comes from the plan) ----
        if (!fusePrec) continue;
        for (int r = 0; r < h.nbRows; r++) {
            const RingRow rr = rows[r];
            const int node = rr.node & kRingNodeMask;
            const bool isInterface = rr.node < 0, hasDiag = rr.diagOff != 0xFFFF;
            if (checkBounds) {
                for (int c = 0; c < 3; c++) if (((rr.node >> (28 + c)) & 1) != (checkBounds[(size_t)c * nbNodes + node] != 0)) { g_error = "mask bits"; return -1; }
            }
            if (OPDIM == 1) {
                const double dgl = sDiag[r];
                prec[node] = isInterface ? dgl : 1.0 / dgl;
            }
            else {
                double b[9], inv[9];
                for (int q = 0; q < 9; q++) b[q] = sDiag[(size_t)r * 9 + q];
                if (!isInterface) {
                    mask_block (b, (rr.node >> 28) & 1, (rr.node >> 29) & 1, (rr.node >> 30) & 1);
                    if (hasDiag) {
                        if (invert3_adj (b, inv, [] (double x) { return ring_rcp (x); })) { for (int q = 0; q < 9; q++) b[q] = inv[q]; }
                        else invert3_lu (b);
                    }
                }
                for (int q = 0; q < 9; q++) prec[(size_t)node * 9 + q] = b[q];
            }
        }
    }
    return 0;
}

extern "C" const char *mfb_ring_replay_error (void) { return g_error.c_str (); }

// Builds the RING plan, verifies its structure, replays it.  values[nbEdges * dim] and
// prec[nbNodes * dim] may be NULL (plan + statistics only).  stats: [0] tiles [1] jobs
// [2] jobs that also write the transposed block [3] ring steps (element visits) [4] padded
// lane-steps [5] chain breaks [6] modelled gather wavefronts [7] their conflict-free count
// [8] modelled slab-store wavefronts per block component [9] their conflict-free count
// [10] plan bytes [11] max rows [12] max nodes [13] max entries [14] max head bytes
// [15] max tail bytes.
extern "C" int mfb_ring_replay (int operatorID, int nbNodes, int nbElem, const int *elemToNode, const int *row,
                                const int *col, const double *coord, const int *checkBounds, const uint8_t *isInterface,
                                int maxRows, int maxEntries, int bankAware, double *values, double *prec, int64_t stats[16])
{
    RingPlanLimits lim;
    if (maxRows > 0) lim.maxRows = maxRows;
    if (maxEntries > 0) lim.maxEntries = maxEntries;
    lim.bankAware = bankAware != 0;
    RingPlan plan;
    if (build_ring_plan (nbNodes, nbElem, elemToNode, row, col, coord, isInterface, checkBounds, lim, plan, g_error) != 0) return -1;
    if (verify_ring_plan (plan, nbNodes, nbElem, elemToNode, row, col, checkBounds, g_error) != 0) { g_error = "verify_ring_plan: " + g_error; return -2; }
    if (stats) {
        const int64_t s[16] = {plan.nbTiles, plan.nbJobs, plan.nbSymmetricJobs, plan.nbRingSteps, plan.nbPaddedSteps, plan.nbBreaks,
                               plan.gatherWavefronts, plan.gatherIdeal, plan.slabWriteWavefronts, plan.slabWriteIdeal,
                               (int64_t)plan.blob.size (), plan.maxRows, plan.maxNodes, plan.maxEntries, plan.maxHeadBytes, plan.maxTailBytes};
        memcpy (stats, s, sizeof s);
    }
    if (!values || !prec) return 0;
    return operatorID == 0 ? replay<1> (plan, coord, checkBounds, nbNodes, 1, values, prec)
                           : replay<9> (plan, coord, checkBounds, nbNodes, 1, values, prec);
}
