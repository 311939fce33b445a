// Development aid: builds the RING plan (host/ring_plan.h) for a Kuhn mesh or an input file and prints
// its counts and its shared-memory model next to the TILED plan's size.
//   g++ -O2 -std=c++17 -fopenmp -o /tmp/ring_plan_stats tools/ring_plan_stats.cc mini-fem_b200/host/{ring_plan,tile_plan,mesh_data,mesh_topology}.cc
//   /tmp/ring_plan_stats 100 [maxRows [maxEntries]]      |      /tmp/ring_plan_stats FILE [maxRows [maxEntries]]
// env: RING_MORTON=1 (Morton runs instead of bisection), RING_NOBANK=1, RING_PASSES=n (refinement rounds)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../mini-fem_b200/host/mesh_data.h"
#include "../mini-fem_b200/host/mesh_topology.h"
#include "../mini-fem_b200/host/ring_plan.h"
#include "../mini-fem_b200/host/tile_plan.h"
using namespace mfb;

int main (int argc, char **argv)
{
    SubMesh m;
    const char *what = argc > 1 ? argv[1] : "40";
    const int g = atoi (what);
    if (g > 0) generate_block (g, g, g, 1, 1, 1, 0, 1, m);
    else if (read_input (what, m) != 0) { printf ("cannot read %s\n", what); return 1; }
    std::vector<int> row (m.nbNodes + 1), col ((size_t)count_csr_entries (m.elemToNode.data (), m.nbElem, m.nbNodes));
    build_csr (m.elemToNode.data (), m.nbElem, m.nbNodes, row.data (), col.data ());
    RingPlanLimits lim;
    if (argc > 2) lim.maxRows = atoi (argv[2]);
    if (argc > 3) lim.maxEntries = atoi (argv[3]);
    if (getenv ("RING_MORTON")) lim.bisection = false;
    if (getenv ("RING_NOBANK")) lim.bankAware = false;
    if (getenv ("RING_PASSES")) lim.refinePasses = atoi (getenv ("RING_PASSES"));
    if (getenv ("RING_SWEEPS")) lim.rotationSweeps = atoi (getenv ("RING_SWEEPS"));
    if (getenv ("RING_MAXJOBS")) lim.maxJobs = atoi (getenv ("RING_MAXJOBS"));
    const int warps = getenv ("RING_WARPS") ? atoi (getenv ("RING_WARPS")) : 8;      // warps per CTA: rounds of the job phase / write-out
    RingPlan plan;
    std::string err;
    auto t0 = std::chrono::steady_clock::now ();
    if (build_ring_plan (m.nbNodes, m.nbElem, m.elemToNode.data (), row.data (), col.data (), m.coord.data (), nullptr, nullptr, lim, plan, err)) {
        printf ("ring plan: %s\n", err.c_str ());
        return 1;
    }
    const double seconds = std::chrono::duration<double> (std::chrono::steady_clock::now () - t0).count ();
    const double E = m.nbElem, Z = row[m.nbNodes];
    const double gather = plan.gatherWavefronts / E, slabW = 4.5 * plan.slabWriteWavefronts / E;
    const double slabR = 2.0 * (Z - m.nbNodes) / 3.0 / E;             // 3 entries per warp iteration, 2 wavefronts
    const double misc = (plan.nbPaddedSteps / 32.0 / 7.0 * 4.0 + plan.nbTiles * 3.0 * plan.maxNodes * 8.0 / 128.0) / E;
    printf ("%s: E %d N %d Z %.0f | rows<=%d entries<=%d | build %.2f s | tiles %d | plan %.1f MB (%.1f B/elem)\n", what, m.nbElem, m.nbNodes, Z,
            lim.maxRows, lim.maxEntries, seconds, plan.nbTiles, plan.blob.size () / 1e6, plan.blob.size () / E);
    printf ("  jobs/E %.3f (transposed too: %.3f)  ring steps/E %.3f  padded lane-steps/E %.3f  breaks %lld  maxNodes %d  slab slots %d  head %u B tail %u B\n",
            plan.nbJobs / E, plan.nbSymmetricJobs / E, plan.nbRingSteps / E, plan.nbPaddedSteps / E, (long long)plan.nbBreaks,
            plan.maxNodes, plan.maxEntries, plan.maxHeadBytes, plan.maxTailBytes);
    printf ("  modelled shared-memory wavefronts per element: gather %.2f (x%.2f of conflict-free) + slab stores %.2f (x%.2f) + slab reads %.2f + codes/coords %.2f = %.2f  -> %.1f M per iteration\n",
            gather, (double)plan.gatherWavefronts / plan.gatherIdeal, slabW, (double)plan.slabWriteWavefronts / plan.slabWriteIdeal, slabR, misc,
            gather + slabW + slabR + misc, (gather + slabW + slabR + misc) * E / 1e6);
    {   // per tile: warp batches of the job phase, row groups of the write-out -> rounds of a CTA of `warps` warps
        std::vector<long long> histB (64, 0), histR (64, 0);
        long long roundsJob = 0, roundsOut = 0, batches = 0, groups = 0;
        for (int t = 0; t < plan.nbTiles; t++) {
            const RingTileHeader &h = *plan.header (t);
            const int g = (h.nbRows + 2) / 3;
            histB[std::min<int> (h.nbBatches, 63)]++; histR[std::min<int> (h.nbRows, 63)]++;
            batches += h.nbBatches; groups += g;
            roundsJob += (h.nbBatches + warps - 1) / warps; roundsOut += (g + warps - 1) / warps;
        }
        printf ("  %d warps per CTA: job phase %.3f rounds per tile (batches / warps = %.3f: efficiency %.2f), write-out %.3f rounds (row groups / warps = %.3f: efficiency %.2f)\n",
                warps, (double)roundsJob / plan.nbTiles, (double)batches / plan.nbTiles / warps, (double)batches / warps / roundsJob,
                (double)roundsOut / plan.nbTiles, (double)groups / plan.nbTiles / warps, (double)groups / warps / roundsOut);
        printf ("  batches per tile:");
        for (int b = 0; b < 64; b++) if (histB[b]) printf (" %d:%lld", b, histB[b]);
        printf ("\n  rows per tile:");
        for (int b = 0; b < 64; b++) if (histR[b]) printf (" %d:%lld", b, histR[b]);
        printf ("\n");
    }
    {   // hash of the serialized plan: a refactoring of host/ring_plan.cc must leave it unchanged
        uint64_t h = 1469598103934665603ull;
        for (uint8_t b : plan.blob) { h ^= b; h *= 1099511628211ull; }
        for (uint64_t o : plan.tileOffset) { h ^= o; h *= 1099511628211ull; }
        printf ("  plan hash %016llx\n", (unsigned long long)h);
    }
    if (verify_ring_plan (plan, m.nbNodes, m.nbElem, m.elemToNode.data (), row.data (), col.data (), nullptr, err)) { printf ("  VERIFY FAILED: %s\n", err.c_str ()); return 1; }
    printf ("  verify ok\n");
    return 0;
}
