// Development aid: builds the TILED plan for a fixed set of meshes / caps and prints a hash of every
// blob, so that a refactoring of host/tile_plan.cc can be checked to leave the plans byte-identical.
//   g++ -O2 -std=c++17 -fopenmp -o /tmp/plan_hash tools/plan_hash.cc mini-fem_b200/host/{tile_plan,mesh_data,mesh_topology}.cc
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
#include <string>
#include <vector>
#include "../mini-fem_b200/host/mesh_data.h"
#include "../mini-fem_b200/host/mesh_topology.h"
#include "../mini-fem_b200/host/tile_plan.h"
using namespace mfb;
static uint64_t fnv (const void *p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const uint8_t *b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static void run (const char *name, int nbNodes, int nbElem, const std::vector<int> &e2n, const std::vector<double> &coord,
                 const uint8_t *intf, TilePlanLimits lim)
{
    std::vector<int> row (nbNodes + 1), col ((size_t)count_csr_entries (e2n.data (), nbElem, nbNodes));
    build_csr (e2n.data (), nbElem, nbNodes, row.data (), col.data ());
    TilePlan plan; std::string err;
    auto t0 = std::chrono::steady_clock::now ();
    if (build_tile_plan (nbNodes, nbElem, e2n.data (), row.data (), col.data (), coord.data (), intf, lim, plan, err)) { printf ("%s: ERROR %s\n", name, err.c_str ()); return; }
    double s = std::chrono::duration<double> (std::chrono::steady_clock::now () - t0).count ();
    uint64_t h = fnv (plan.blob.data (), plan.blob.size ());
    h = fnv (plan.tileOffset.data (), plan.tileOffset.size () * 8, h);
    std::string verr;
    int v = verify_tile_plan (plan, nbNodes, nbElem, e2n.data (), row.data (), col.data (), verr);
    printf ("%-28s tiles %6d intf %5d bytes %10zu hash %016llx verify %d  (%.2f s)\n", name, plan.nbTiles, plan.nbInterfaceTiles, plan.blob.size (), (unsigned long long)h, v, s);
}
int main ()
{
    for (int lap = 0; lap < 2; lap++) {
        for (int g : {7, 24, 40}) {
            SubMesh m; generate_block (g, g, g - 2, 1, 1, 1, 0, 3, m);
            TilePlanLimits lim; lim.laplacian = lap;
            char name[64]; snprintf (name, sizeof name, "kuhn%d %s", g, lap ? "lap" : "ela");
            run (name, m.nbNodes, m.nbElem, m.elemToNode, m.coord, nullptr, lim);
            if (g == 24) {
                std::vector<uint8_t> intf (m.nbNodes, 0);
                for (int n = 0; n < m.nbNodes; n += 7) intf[n] = 1;
                snprintf (name, sizeof name, "kuhn%d %s intf", g, lap ? "lap" : "ela");
                run (name, m.nbNodes, m.nbElem, m.elemToNode, m.coord, intf.data (), lim);
                TilePlanLimits l2 = lim; l2.maxRows = 64; l2.maxElems = 624; l2.maxNodesRef = 624;
                snprintf (name, sizeof name, "kuhn%d %s 64/624", g, lap ? "lap" : "ela");
                run (name, m.nbNodes, m.nbElem, m.elemToNode, m.coord, nullptr, l2);
                TilePlanLimits l3 = lim; l3.bankAware = false;
                snprintf (name, sizeof name, "kuhn%d %s plain order", g, lap ? "lap" : "ela");
                run (name, m.nbNodes, m.nbElem, m.elemToNode, m.coord, nullptr, l3);
                TilePlanLimits l4 = lim; l4.maxRows = 5; l4.maxElems = 100; l4.maxNodesRef = 100;
                snprintf (name, sizeof name, "kuhn%d %s 5/100", g, lap ? "lap" : "ela");
                run (name, m.nbNodes, m.nbElem, m.elemToNode, m.coord, nullptr, l4);
                // shuffled numbering
                std::mt19937 rng (5);
                std::vector<int> nperm (m.nbNodes), eperm (m.nbElem);
                for (int i = 0; i < m.nbNodes; i++) nperm[i] = i;
                for (int i = 0; i < m.nbElem; i++) eperm[i] = i;
                std::shuffle (nperm.begin (), nperm.end (), rng); std::shuffle (eperm.begin (), eperm.end (), rng);
                std::vector<double> c2 (m.coord.size ()); std::vector<int> e2 (m.elemToNode.size ());
                for (int i = 0; i < m.nbNodes; i++) for (int k = 0; k < 3; k++) c2[(size_t)nperm[i] * 3 + k] = m.coord[(size_t)i * 3 + k];
                for (int e = 0; e < m.nbElem; e++) for (int k = 0; k < 4; k++) e2[(size_t)eperm[e] * 4 + k] = nperm[m.elemToNode[(size_t)e * 4 + k] - 1] + 1;
                snprintf (name, sizeof name, "kuhn%d %s shuffled", g, lap ? "lap" : "ela");
                run (name, m.nbNodes, m.nbElem, e2, c2, nullptr, lim);
            }
        }
        // random tets: high, irregular degrees
        std::mt19937 rng (11);
        int nbNodes = 600, nbElem = 4000;
        std::vector<double> coord (nbNodes * 3); for (auto &c : coord) c = (rng () % 100000) / 50000.0 - 1.0;
        std::vector<int> e2n (nbElem * 4);
        for (int e = 0; e < nbElem; e++) { int ids[4]; for (int k = 0; k < 4; ) { int c = rng () % nbNodes + 1; bool dup = false; for (int j = 0; j < k; j++) dup |= ids[j] == c; if (!dup) ids[k++] = c; } for (int k = 0; k < 4; k++) e2n[e * 4 + k] = ids[k]; }
        TilePlanLimits lim; lim.laplacian = lap;
        run (lap ? "random lap" : "random ela", nbNodes, nbElem, e2n, coord, nullptr, lim);
    }
}
