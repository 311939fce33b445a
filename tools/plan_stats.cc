// Development aid: builds the TILED plan for a structured mesh and models the shared-memory
// wavefronts of each kernel phase (64-bit accesses: per half-warp, wavefronts = largest number
// of distinct 8-byte words falling into one of the 16 bank pairs).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>
#include "../mini-fem_b200/host/mesh_data.h"
#include "../mini-fem_b200/host/mesh_topology.h"
#include "../mini-fem_b200/host/tile_plan.h"
using namespace mfb;

static int wavefronts16 (const int *slots, int n)
{
    int cnt[16] = {0}, best = 0;
    for (int i = 0; i < n; i++) {
        bool dup = false;
        for (int j = 0; j < i; j++) dup |= slots[j] == slots[i];
        if (dup || slots[i] < 0) continue;
        best = std::max (best, ++cnt[slots[i] & 15]);
    }
    return std::max (best, n > 0 ? 1 : 0);
}

int main (int argc, char **argv)
{
    int g = argc > 1 ? atoi (argv[1]) : 40;
    TilePlanLimits lim;
    if (argc > 2) lim.maxRows = atoi (argv[2]);
    if (argc > 3) lim.maxElems = atoi (argv[3]);
    lim.maxNodesRef = lim.maxElems; lim.maxEntries = 65535;
    if (argc > 4) lim.bankAware = atoi (argv[4]) != 0;
    SubMesh m;
    if (argc > 1 && strchr (argv[1], '/')) {                  // an input file (src/IO.cc format), e.g. from tools/delaunay_mesh.py
        if (read_input (argv[1], m)) { printf ("cannot read %s\n", argv[1]); return 1; }
    }
    else generate_block (g, g, g, 1, 1, 1, 0, 1, m);
    std::vector<int> row (m.nbNodes + 1), col (m.nbEdges);
    build_csr (m.elemToNode.data (), m.nbElem, m.nbNodes, row.data (), col.data ());
    TilePlan plan; std::string err;
    if (build_tile_plan (m.nbNodes, m.nbElem, m.elemToNode.data (), row.data (), col.data (), m.coord.data (), nullptr, lim, plan, err)) { printf ("%s\n", err.c_str ()); return 1; }
    const int stride = plan.elemStride;
    double wfP4 = 0, instrP4 = 0, wfDiag = 0, instrDiag = 0, wfP2 = 0, instrP2 = 0;
    for (int t = 0; t < plan.nbTiles; t++) {
        const uint8_t *base = plan.blob.data () + plan.tileOffset[t];
        const TileBlobHeader &h = *plan.header (t);
        const TileRow *rows = (const TileRow*)(base + sizeof (TileBlobHeader));
        const uint16_t *elems = (const uint16_t*)(base + h.offElems);
        const TileBatch *batches = (const TileBatch*)(base + h.offBatches);
        const uint16_t *diag = (const uint16_t*)(base + h.offDiag);
        const uint16_t *pair = (const uint16_t*)(base + h.offPair);
        for (int b = 0; b < h.nbBatches; b++) {
            for (int s = 0; s < batches[b].steps; s++) {
                for (int half = 0; half < 2; half++) {
                    int sa[16], sb[16];
                    for (int l = 0; l < 16; l++) {
                        int code = pair[batches[b].codeBase + s * 32 + half * 16 + l];
                        int e = code >> 4, a = (code >> 2) & 3, bb = code & 3;
                        sa[l] = a * stride + e; sb[l] = bb * stride + e;
                    }
                    wfP4 += 3 * (wavefronts16 (sa, 16) + wavefronts16 (sb, 16));
                }
                instrP4 += 6;
            }
        }
        for (int r0 = 0; r0 < h.nbRows; r0 += 8) {
            int steps = 0;
            for (int r = r0; r < std::min (r0 + 8, (int)h.nbRows); r++) steps = std::max (steps, (rows[r + 1].diagCodeBase - rows[r].diagCodeBase + 3) / 4);
            for (int s = 0; s < steps; s++) {
                for (int half = 0; half < 2; half++) {
                    int sl[16];
                    for (int l = 0; l < 16; l++) {
                        int r = r0 + half * 4 + l / 4, k = (r < h.nbRows ? rows[r].diagCodeBase : 0) + s * 4 + (l & 3);
                        sl[l] = (r < h.nbRows && k < rows[r + 1].diagCodeBase) ? (diag[k] & 3) * stride + (diag[k] >> 2) : -1;
                    }
                    wfDiag += 3 * wavefronts16 (sl, 16);
                }
                instrDiag += 3;
            }
        }
        for (int e0 = 0; e0 < h.nbElems; e0 += 32) {
            for (int k = 0; k < 4; k++) {
                for (int half = 0; half < 2; half++) {
                    int sl[16];
                    for (int l = 0; l < 16; l++) { int e = e0 + half * 16 + l; sl[l] = (e < h.nbElems && elems[e * 4 + k] != 0xFFFF) ? elems[e * 4 + k] : -1; }
                    wfP2 += 3 * wavefronts16 (sl, 16);
                }
                instrP2 += 3;
            }
        }
    }
    printf ("%s: %d elements, tiles %d (%.1f rows, %.1f elements each) tileElems %.3f x, padded steps %.3f x, blob max %u\n", argc > 1 ? argv[1] : "40", m.nbElem, plan.nbTiles,
            (double)m.nbNodes / plan.nbTiles, (double)plan.nbTileElems / plan.nbTiles,
            (double)plan.nbTileElems / m.nbElem, (double)plan.nbPaddedSteps * 32 / (12.0 * m.nbElem), plan.maxBlobBytes);
    printf ("per element: P4 %.1f wavefronts (%.2f per LDS.64), diag %.1f (%.2f), P2 reads %.1f (%.2f)\n",
            wfP4 / m.nbElem, wfP4 / instrP4, wfDiag / m.nbElem, wfDiag / instrDiag, wfP2 / m.nbElem, wfP2 / instrP2);
    return 0;
}
