#include "ring_plan.h"

#include <omp.h>

#include <algorithm>
#include <cstring>

#include "mesh_topology.h"
#include "tile_plan.h"

namespace mfb {

namespace {

inline uint32_t align16 (uint32_t x) { return (x + 15u) & ~15u; }

struct RingJob {
    int i = 0, j = 0;                 // indices into TileWork::nodes until the final numbering
    int slotIJ = 0xFFFF, slotJI = 0xFFFF;
    int codeStart = 0, codeLen = 0;   // bytes in TileWork::codes: node indices, kRingBreak between chains
    bool closedSingle = false;        // one closed chain: every rotation / direction walks the same ring
    bool single = false;              // exactly one chain (may be reversed)
};

// Everything one thread needs to plan one tile; vectors are reused from tile to tile.
struct TileWork {
    std::vector<int> nodes;           // global ids, owned rows first
    std::vector<RingRow> rows;
    std::vector<RingJob> jobs;
    std::vector<uint8_t> codes;       // ring codes of all jobs (node indices < 254)
    std::vector<int> order;           // job index per lane slot after sorting, -1 = idle
    std::vector<RingBatch> batches;
    std::vector<int> newId;           // node index -> tile-local id
    // scratch of the edge-link walk
    std::vector<int> linkP, linkQ;
    std::vector<uint8_t> linkUsed;
    void clear ()
    {
        nodes.clear (); rows.clear (); jobs.clear (); codes.clear (); order.clear (); batches.clear (); newId.clear ();
    }
};

struct TileResult {
    std::vector<uint8_t> blob;
    uint32_t headBytes = 0;
    int nbRows = 0, nbNodes = 0, nbEntries = 0, nbBatches = 0;
    bool hasInterface = false;
    int bad = 0;                      // 1 = CSR lacks a pair, 2 = element names a node twice, 3 = too many nodes
    int64_t jobs = 0, symmetric = 0, ringSteps = 0, paddedSteps = 0, breaks = 0;
    int64_t gatherWf = 0, gatherIdeal = 0, slabWf = 0, slabIdeal = 0;
};

// Splits the link of one mesh edge — the graph whose vertices are the "other" nodes and whose
// edges are the elements (p, q) around the mesh edge — into chains of consecutive elements and
// appends them to `codes`: node, node, ... [kRingBreak node, node, ...].  On a conforming mesh the
// link is one closed polygon (interior edge) or one open fan (boundary edge); anything else
// still works, with one break per extra chain.
void append_chains (TileWork &w, RingJob &job, int &breaks)
{
    const int m = (int)w.linkP.size ();
    job.codeStart = (int)w.codes.size ();
    job.single = false; job.closedSingle = false;
    if (m == 0) { job.codeLen = 0; return; }
    w.linkUsed.assign ((size_t)m, 0);
    int remaining = m, chains = 0;
    bool closed = false;
    while (remaining > 0) {
        // a vertex of odd remaining degree starts an open chain; otherwise any remaining vertex
        int start = -1;
        for (int k = 0; k < m && start < 0; k++) {
            if (w.linkUsed[k]) continue;
            for (int side = 0; side < 2 && start < 0; side++) {
                const int v = side ? w.linkQ[k] : w.linkP[k];
                int deg = 0;
                for (int t = 0; t < m; t++) if (!w.linkUsed[t]) deg += (w.linkP[t] == v) + (w.linkQ[t] == v);
                if (deg & 1) start = v;
            }
        }
        if (start < 0) for (int k = 0; k < m; k++) if (!w.linkUsed[k]) { start = w.linkP[k]; break; }
        if (chains > 0) { w.codes.push_back ((uint8_t)kRingBreak); breaks++; }
        chains++;
        int cur = start;
        w.codes.push_back ((uint8_t)cur);
        for (;;) {
            int next = -1;
            for (int k = 0; k < m; k++) {
                if (w.linkUsed[k]) continue;
                if (w.linkP[k] == cur) next = w.linkQ[k];
                else if (w.linkQ[k] == cur) next = w.linkP[k];
                else continue;
                w.linkUsed[k] = 1;
                break;
            }
            if (next < 0) break;
            w.codes.push_back ((uint8_t)next);
            cur = next;
            remaining--;
        }
        closed = cur == start;
    }
    job.codeLen = (int)w.codes.size () - job.codeStart;
    job.single = chains == 1;
    job.closedSingle = chains == 1 && closed && job.codeLen >= 3;
}

// Bank model of the shared-memory coordinate planes: 8-byte words, 16 per 128-byte line, a
// half-warp's LDS.64 takes as many wavefronts as the busiest bank has DISTINCT words.
struct HalfWarpBanks {
    static constexpr int kMaxSteps = 64;
    uint8_t node[kMaxSteps][16][16];
    uint8_t refs[kMaxSteps][16][16];          // lanes of the half-warp loading that word
    uint8_t count[kMaxSteps][16];
    uint8_t busiest[kMaxSteps];               // max over the banks of count[s][.]
    int steps = 0;
    void reset (int nbSteps)
    {
        steps = std::min (nbSteps, (int)kMaxSteps);
        memset (count, 0, sizeof (count[0]) * (size_t)steps);
        memset (busiest, 0, (size_t)steps);
    }
    int cost (int s, int id) const
    {
        if (s >= steps) return 0;
        const int b = id & 15, n = count[s][b];
        for (int k = 0; k < n; k++) if (node[s][b][k] == id) return 0;   // same word: broadcast
        // the step costs as many wavefronts as its busiest bank holds words: joining a bank that
        // already is the busiest one adds a wavefront, joining a quieter one only makes that likelier
        return n == 0 ? 0 : (n >= busiest[s] ? 16 + n : n);
    }
    void add (int s, int id)
    {
        if (s >= steps) return;
        const int b = id & 15, n = count[s][b];
        for (int k = 0; k < n; k++) if (node[s][b][k] == id) { refs[s][b][k]++; return; }
        if (n < 16) {
            node[s][b][n] = (uint8_t)id; refs[s][b][n] = 1; count[s][b] = (uint8_t)(n + 1);
            busiest[s] = std::max (busiest[s], count[s][b]);
        }
    }
    void remove (int s, int id)
    {
        if (s >= steps) return;
        const int b = id & 15, n = count[s][b];
        for (int k = 0; k < n; k++) {
            if (node[s][b][k] != id) continue;
            if (--refs[s][b][k] == 0) {
                node[s][b][k] = node[s][b][n - 1]; refs[s][b][k] = refs[s][b][n - 1]; count[s][b] = (uint8_t)(n - 1);
                if (n == busiest[s]) { uint8_t m = 0; for (int q = 0; q < 16; q++) m = std::max (m, count[s][q]); busiest[s] = m; }
            }
            return;
        }
    }
    int wavefronts (int s) const { return busiest[s]; }
};

// Plans one tile.  Global -> tile maps are thread-private dense arrays, reset on exit.
void plan_tile (int nbRows, const int *rowNodes, const int *elemToNode, const int *row, const int *col,
                const int *n2eIndex, const int *n2eValue, const uint8_t *isInterface, const int *checkBounds, int nbNodes,
                const RingPlanLimits &lim,
                std::vector<int> &nodeLocal, std::vector<int> &colStamp, TileWork &w, TileResult &out)
{
    w.clear ();
    out = TileResult ();
    // ---- nodes: owned rows first, then every node of an element that touches an owned row ----
    for (int r = 0; r < nbRows; r++) { nodeLocal[rowNodes[r]] = r; w.nodes.push_back (rowNodes[r]); }
    for (int r = 0; r < nbRows; r++) {
        const int n = rowNodes[r];
        for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
            const int *en = elemToNode + (size_t)n2eValue[p] * kDimElem;
            for (int a = 0; a < kDimElem; a++) {
                for (int b = a + 1; b < kDimElem; b++) if (en[a] == en[b]) out.bad = 2;
                const int m = en[a] - 1;
                if (nodeLocal[m] < 0) { nodeLocal[m] = (int)w.nodes.size (); w.nodes.push_back (m); }
            }
        }
    }
    auto release = [&] () { for (int n : w.nodes) nodeLocal[n] = -1; };
    if ((int)w.nodes.size () > kRingMaxNodes) out.bad = 3;
    if (out.bad) { release (); return; }

    // ---- row table ------------------------------------------------------------------------
    int localStart = 0;
    for (int r = 0; r < nbRows; r++) {
        const int n = rowNodes[r], begin = row[n], end = row[n + 1];
        RingRow rr;
        memset (&rr, 0, sizeof rr);
        const bool intf = isInterface && isInterface[n];
        out.hasInterface |= intf;
        rr.node = n | (intf ? (int)0x80000000u : 0);
        if (checkBounds) {
            for (int c = 0; c < 3; c++) if (checkBounds[(size_t)c * nbNodes + n] != 0) rr.node |= 1 << (28 + c);
        }
        rr.valueStart = begin;
        rr.localStart = (uint16_t)localStart;
        rr.len = (uint16_t)(end - begin);
        rr.diagOff = 0xFFFF;
        for (int l = begin; l < end; l++) if (col[l] == n + 1) { rr.diagOff = (uint16_t)(l - begin); break; }
        w.rows.push_back (rr);
        localStart += (end - begin) + ring_row_padding (end - begin);
    }
    out.nbRows = nbRows;
    out.nbEntries = localStart;
    w.jobs.reserve ((size_t)localStart);

    // ---- jobs: one per mesh edge {i, j}, i owned; both blocks when j is owned too -------------
    auto first_position = [&] (int n, int target) {           // first l in row n with col[l] == target + 1, or -1
        for (int l = row[n]; l < row[n + 1]; l++) if (col[l] == target + 1) return l - row[n];
        return -1;
    };
    for (int r = 0; r < nbRows; r++) {
        const int n = rowNodes[r], begin = row[n], end = row[n + 1];
        for (int l = begin; l < end; l++) colStamp[col[l] - 1] = n + 1;
        for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {  // every node pair of every element must have its entry
            const int *en = elemToNode + (size_t)n2eValue[p] * kDimElem;
            for (int a = 0; a < kDimElem; a++) if (colStamp[en[a] - 1] != n + 1) out.bad = 1;
        }
        for (int l = begin; l < end; l++) {
            const int c = col[l] - 1, pos = l - begin;
            if (pos == w.rows[r].diagOff) continue;           // the diagonal entry comes from the row sum
            RingJob job;
            job.i = r;
            job.slotIJ = w.rows[r].localStart + pos;
            const bool firstSeen = first_position (n, c) == pos;
            if (!firstSeen || c == n) {
                // a column listed twice: the reference's search (src/assembly.cc:419-421) only ever
                // finds the first one, the others stay zero — a job without elements
                job.j = r;
                w.jobs.push_back (job);
                continue;
            }
            const int cl = nodeLocal[c];                       // >= 0: c shares an element with n or is owned
            if (cl >= 0 && cl < nbRows) {
                const int back = first_position (c, n);        // where row c keeps (c, n)
                if (cl < r && back >= 0) continue;             // written by row c's job as its transposed block
                if (back >= 0) job.slotJI = w.rows[cl].localStart + back;
            }
            w.linkP.clear (); w.linkQ.clear ();
            for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
                const int *en = elemToNode + (size_t)n2eValue[p] * kDimElem;
                int others[4], nOthers = 0;
                bool hasC = false;
                for (int a = 0; a < kDimElem; a++) {
                    if (en[a] - 1 == c) hasC = true;
                    else if (en[a] - 1 != n) others[nOthers++] = nodeLocal[en[a] - 1];
                }
                if (hasC && nOthers == 2) { w.linkP.push_back (others[0]); w.linkQ.push_back (others[1]); }
            }
            if (cl < 0) {                                      // entry without any element: stays zero
                job.j = r;
                w.jobs.push_back (job);
                continue;
            }
            job.j = cl;
            int breaks = 0;
            append_chains (w, job, breaks);
            out.breaks += breaks;
            out.ringSteps += (int64_t)w.linkP.size ();
            if (job.slotJI != 0xFFFF) out.symmetric++;
            w.jobs.push_back (job);
        }
    }
    out.jobs = (int64_t)w.jobs.size ();
    if (out.bad) { release (); return; }

    // ---- lanes: jobs sorted by ring length, 32 to a warp ---------------------------------------
    const int nbJobs = (int)w.jobs.size ();
    w.order.resize ((size_t)nbJobs);
    for (int k = 0; k < nbJobs; k++) w.order[k] = k;
    // longest rings first (a warp runs as many steps as its longest job); inside a length class the
    // jobs keep their row order: the rings of one row share their nodes, which the bank-aware
    // numbering below exploits, and consecutive lanes then write consecutive slab slots.
    // (Tried on the model: jobs with a transposed block first, and half-warps picked for distinct
    // slab banks — fewer slab conflicts, but more gather conflicts than they save.)
    std::stable_sort (w.order.begin (), w.order.end (), [&] (int a, int b) { return w.jobs[a].codeLen > w.jobs[b].codeLen; });
    while (w.order.size () % 32) w.order.push_back (-1);
    if (lim.bankAware) {
        // Lane order inside a half-warp is free as far as the gather goes (its conflicts depend on which
        // jobs share the half-warp, not on their lanes), so use it for the slab: split the 16 jobs into
        // the two quarter-warps so that the slots a quarter stores to differ mod 8 as far as possible.
        for (size_t h0 = 0; h0 < w.order.size (); h0 += 16) {
            int group[2][8], count[2] = {0, 0};
            unsigned loadIJ[2][8] = {{0}}, loadJI[2][8] = {{0}};
            for (int l = 0; l < 16; l++) {
                const int k = w.order[h0 + l];
                int best = count[0] <= count[1] ? 0 : 1;
                if (k >= 0) {
                    const RingJob &jb = w.jobs[k];
                    long bestCost = 1l << 60;
                    for (int q = 0; q < 2; q++) {
                        if (count[q] >= 8) continue;
                        long cost = 16l * loadIJ[q][jb.slotIJ & 7] + (jb.slotJI != 0xFFFF ? 16l * loadJI[q][jb.slotJI & 7] : 0) + count[q];
                        if (cost < bestCost) { bestCost = cost; best = q; }
                    }
                    loadIJ[best][jb.slotIJ & 7]++;
                    if (jb.slotJI != 0xFFFF) loadJI[best][jb.slotJI & 7]++;
                }
                else if (count[best] >= 8) best ^= 1;
                group[best][count[best]++] = k;
            }
            for (int q = 0; q < 2; q++) for (int l = 0; l < 8; l++) w.order[h0 + (size_t)q * 8 + l] = group[q][l];
        }
    }
    const int nbBatches = (int)w.order.size () / 32;
    out.nbBatches = nbBatches;

    // ---- tile-local node ids: id mod 16 = shared-memory bank (8-byte words) --------------------
    const int nbRef = (int)w.nodes.size ();
    w.newId.assign ((size_t)nbRef, -1);
    const int nbHalf = nbBatches * 2;
    if (lim.bankAware && nbRef > 0) {
        // nodes that one half-warp loads (ring nodes and the two end nodes of its jobs) should
        // sit in different banks: greedy, most constrained node first
        std::vector<std::vector<int>> halfOf ((size_t)nbRef);
        std::vector<int> appearances ((size_t)nbRef, 0);
        for (int h = 0; h < nbHalf; h++) {
            for (int l = 0; l < 16; l++) {
                const int k = w.order[(size_t)h * 16 + l];
                if (k < 0) continue;
                const RingJob &jb = w.jobs[k];
                auto touch = [&] (int n) {
                    appearances[n]++;
                    if (halfOf[n].empty () || halfOf[n].back () != h) halfOf[n].push_back (h);
                };
                touch (jb.i); touch (jb.j);
                for (int q = 0; q < jb.codeLen; q++) if (w.codes[(size_t)jb.codeStart + q] != kRingBreak) touch (w.codes[(size_t)jb.codeStart + q]);
            }
        }
        std::vector<int> byWeight ((size_t)nbRef);
        for (int n = 0; n < nbRef; n++) byWeight[n] = n;
        std::stable_sort (byWeight.begin (), byWeight.end (), [&] (int a, int b) { return appearances[a] > appearances[b]; });
        std::vector<uint16_t> used ((size_t)std::max (nbHalf, 1) * 16, 0);
        int population[16] = {0};
        for (int n : byWeight) {
            int bestClass = -1;
            long bestCost = 1l << 60;
            for (int c = 0; c < 16; c++) {
                if (c + 16 * population[c] >= kRingMaxNodes) continue;
                long cost = 0;
                for (int h : halfOf[n]) cost += used[(size_t)h * 16 + c];
                cost = cost * 1024 + population[c];
                if (cost < bestCost) { bestCost = cost; bestClass = c; }
            }
            w.newId[n] = bestClass + 16 * population[bestClass]++;
            for (int h : halfOf[n]) used[(size_t)h * 16 + bestClass]++;
        }
    }
    else {
        for (int n = 0; n < nbRef; n++) w.newId[n] = n;
    }
    // ---- ring rotation / direction per job, lane after lane of a half-warp ---------------------
    // step 0 loads node i, step 1 node j, step 2 + q the q-th code byte
    static thread_local HalfWarpBanks banksTls;
    HalfWarpBanks &banks = banksTls;          // bound once per tile: every access of the thread-local goes through its wrapper
    std::vector<uint8_t> best;
    std::vector<int> costOf;
    std::vector<int> idOf;
    int maxSteps = 0;
    for (int b = 0; b < nbBatches; b++) {
        int nbSteps = 0;
        for (int l = 0; l < 32; l++) if (w.order[(size_t)b * 32 + l] >= 0) nbSteps = std::max (nbSteps, w.jobs[w.order[(size_t)b * 32 + l]].codeLen);
        RingBatch rb;
        rb.codeBase = 0; rb.nbSteps = (uint16_t)nbSteps; rb.flags = 0;
        for (int l = 0; l < 32; l++) {
            const int k = w.order[(size_t)b * 32 + l];
            if (k < 0) continue;
            for (int q = 0; q < w.jobs[k].codeLen; q++) if (w.codes[(size_t)w.jobs[k].codeStart + q] == kRingBreak) rb.flags |= kRingBatchGeneral;
        }
        w.batches.push_back (rb);
        out.paddedSteps += (int64_t)32 * nbSteps;
        maxSteps = std::max (maxSteps, nbSteps);
    }
    auto rotate_pass = [&] () {
        out.gatherWf = out.gatherIdeal = out.slabWf = out.slabIdeal = 0;
        for (int h = 0; h < nbHalf; h++) {
            const int nbSteps = w.batches[h / 2].nbSteps;
            banks.reset (nbSteps + 2);
            // slab stores are 128-bit (entry stride 80 bytes): a quarter-warp of 8 lanes per wavefront,
            // conflict-free when its slots differ mod 8
            unsigned loadIJ[2][8] = {{0}}, loadJI[2][8] = {{0}};
            auto choose = [&] (RingJob &jb) {      // best start / direction of the job's ring against the banks taken so far
                uint8_t *codes = w.codes.data () + jb.codeStart;
                const int len = jb.codeLen;
                if (!(lim.bankAware && jb.single && len >= 2)) return;
                // candidates: the chain reversed; a closed ring of v elements (v + 1 bytes, first = last)
                // may also start at any of its v nodes
                const bool jobClosed = jb.closedSingle;
                const int v = len - 1, nbRot = jobClosed ? v : 1;
                // cost of having ring node m in step q, once for all candidates; a closed ring's row is stored twice in a
                // row (stride 2 v) so that a rotation is an offset, not a division per term
                const int nbPos = jobClosed ? v : len, stride = jobClosed ? 2 * v : len;
                idOf.resize ((size_t)nbPos);
                for (int m = 0; m < nbPos; m++) idOf[m] = w.newId[codes[m]];
                {   // the search below starts with the ring as it stands and stops at a free candidate: look at that one first
                    long asIs = 0;
                    for (int q = 0; q < len && asIs == 0; q++) asIs += banks.cost (2 + q, idOf[jobClosed && q == v ? 0 : q]);
                    if (asIs == 0) return;
                }
                costOf.resize ((size_t)len * stride);
                for (int q = 0; q < len; q++) {
                    int *rowCost = costOf.data () + (size_t)q * stride;
                    for (int m = 0; m < nbPos; m++) rowCost[m] = banks.cost (2 + q, idOf[m]);
                    if (jobClosed) for (int m = 0; m < v; m++) rowCost[v + m] = rowCost[m];
                }
                long bestCost = 1l << 60;
                int bestRot = 0, bestDir = 0;
                for (int rot = 0; rot < nbRot && bestCost > 0; rot++) {
                    for (int dir = 0; dir < 2 && bestCost > 0; dir++) {
                        long cost = 0;
                        if (jobClosed) {
                            // src = (rot + q) mod v, resp. (rot - q) mod v = index rot - q + v of the doubled row
                            if (dir == 0) for (int q = 0; q < len; q++) cost += costOf[(size_t)q * stride + rot + q];
                            else          for (int q = 0; q < len; q++) cost += costOf[(size_t)q * stride + rot - q + v];
                        }
                        else {
                            if (dir == 0) for (int q = 0; q < len; q++) cost += costOf[(size_t)q * stride + q];
                            else          for (int q = 0; q < len; q++) cost += costOf[(size_t)q * stride + len - 1 - q];
                        }
                        if (cost < bestCost) { bestCost = cost; bestRot = rot; bestDir = dir; }
                    }
                }
                best.resize ((size_t)len);
                for (int q = 0; q < len; q++) {
                    int src;
                    if (jb.closedSingle) src = bestDir ? ((bestRot - q) % v + v) % v : (bestRot + q) % v;
                    else src = bestDir ? len - 1 - q : q;
                    best[q] = codes[src];
                }
                memcpy (codes, best.data (), (size_t)len);
            };
            auto place = [&] (const RingJob &jb, bool add) {
                const uint8_t *codes = w.codes.data () + jb.codeStart;
                for (int q = 0; q < jb.codeLen; q++) {
                    if (codes[q] == kRingBreak) continue;
                    if (add) banks.add (2 + q, w.newId[codes[q]]); else banks.remove (2 + q, w.newId[codes[q]]);
                }
            };
            for (int l = 0; l < 16; l++) {
                const int k = w.order[(size_t)h * 16 + l];
                if (k < 0) continue;
                RingJob &jb = w.jobs[k];
                choose (jb);
                banks.add (0, w.newId[jb.i]);
                banks.add (1, w.newId[jb.j]);
                place (jb, true);
                loadIJ[l >> 3][jb.slotIJ & 7]++;
                if (jb.slotJI != 0xFFFF) loadJI[l >> 3][jb.slotJI & 7]++;
            }
            // coordinate descent: every lane in turn is taken out and re-placed against all the others
            for (int sweep = 0; sweep < (lim.bankAware ? lim.rotationSweeps : 0); sweep++) {
                for (int l = 0; l < 16; l++) {
                    const int k = w.order[(size_t)h * 16 + l];
                    if (k < 0) continue;
                    RingJob &jb = w.jobs[k];
                    if (!(jb.single && jb.codeLen >= 2)) continue;
                    place (jb, false);
                    choose (jb);
                    place (jb, true);
                }
            }
            for (int s = 0; s < banks.steps; s++) {
                const int wf = banks.wavefronts (s);
                out.gatherWf += 3 * wf;
                out.gatherIdeal += 3 * (wf > 0);
            }
            for (int quarter = 0; quarter < 2; quarter++) {
                unsigned mIJ = 0, mJI = 0;
                for (int c = 0; c < 8; c++) { mIJ = std::max (mIJ, loadIJ[quarter][c]); mJI = std::max (mJI, loadJI[quarter][c]); }
                out.slabWf += mIJ + mJI;
                out.slabIdeal += (mIJ ? 1 : 0) + (mJI ? 1 : 0);
            }
        }
    };
    // With the rotations fixed, the loads that really meet in one half-warp step are known: move
    // every node to the bank where it meets the fewest others, then rotate again.
    auto renumber_pass = [&] () {
        const int S = std::min (maxSteps + 2, (int)HalfWarpBanks::kMaxSteps);
        static thread_local std::vector<std::vector<int>> appTls;       // inner vectors keep their capacity from tile to tile
        std::vector<std::vector<int>> &app = appTls;
        if ((int)app.size () < nbRef) app.resize ((size_t)nbRef);
        for (int n = 0; n < nbRef; n++) app[n].clear ();
        for (int h = 0; h < nbHalf; h++) {
            for (int l = 0; l < 16; l++) {
                const int k = w.order[(size_t)h * 16 + l];
                if (k < 0) continue;
                const RingJob &jb = w.jobs[k];
                app[jb.i].push_back (h * S);
                app[jb.j].push_back (h * S + 1);
                for (int q = 0; q < jb.codeLen && q + 2 < S; q++) {
                    const int c = w.codes[(size_t)jb.codeStart + q];
                    if (c != kRingBreak) app[c].push_back (h * S + 2 + q);
                }
            }
        }
        std::vector<uint8_t> usage ((size_t)std::max (nbHalf, 1) * S * 16, 0);
        std::vector<int> cls ((size_t)nbRef), byWeight ((size_t)nbRef);
        int population[16] = {0};
        for (int n = 0; n < nbRef; n++) {
            std::sort (app[n].begin (), app[n].end ());
            app[n].erase (std::unique (app[n].begin (), app[n].end ()), app[n].end ());
            cls[n] = w.newId[n] & 15;
            population[cls[n]]++;
            for (int key : app[n]) usage[(size_t)key * 16 + cls[n]]++;
            byWeight[n] = n;
        }
        std::stable_sort (byWeight.begin (), byWeight.end (), [&] (int a, int b) { return app[a].size () > app[b].size (); });
        for (int n : byWeight) {
            const int c0 = cls[n];
            for (int key : app[n]) usage[(size_t)key * 16 + c0]--;
            population[c0]--;
            int bestClass = c0;
            long bestCost = 1l << 60;
            for (int c = 0; c < 16; c++) {
                if (population[c] >= (kRingMaxNodes - 1 - c) / 16 + 1) continue;
                long cost = 0;
                for (int key : app[n]) cost += usage[(size_t)key * 16 + c];
                cost = cost * 4096 + (c == c0 ? 0 : 1 + population[c]);
                if (cost < bestCost) { bestCost = cost; bestClass = c; }
            }
            cls[n] = bestClass;
            population[bestClass]++;
            for (int key : app[n]) usage[(size_t)key * 16 + bestClass]++;
        }
        int next[16] = {0};
        for (int n = 0; n < nbRef; n++) w.newId[n] = cls[n] + 16 * next[cls[n]]++;
    };
    rotate_pass ();
    for (int pass = 0; pass < (lim.bankAware ? lim.refinePasses : 0); pass++) {
        renumber_pass ();
        rotate_pass ();
    }
    int maxId = -1;
    for (int n = 0; n < nbRef; n++) maxId = std::max (maxId, w.newId[n]);
    out.nbNodes = maxId + 1;

    // ---- serialise ------------------------------------------------------------------------------
    uint32_t at = (uint32_t)sizeof (RingTileHeader) + (uint32_t)(w.rows.size () * sizeof (RingRow));
    at = align16 (at);
    const uint32_t offNodes = at;
    at = align16 (at + 4u * (uint32_t)out.nbNodes);
    const uint32_t headBytes = at;
    at = align16 (at + (uint32_t)(nbBatches * sizeof (RingBatch)));
    const uint32_t offJobs = at;
    at += 256u * (uint32_t)nbBatches;
    const uint32_t offCodes = at;
    uint32_t words = 0;
    for (RingBatch &rb : w.batches) { rb.codeBase = words * 32; words += (rb.nbSteps + 7u) / 8u; }
    at += 256u * words;
    out.headBytes = headBytes;
    out.blob.assign ((size_t)at, 0);
    uint8_t *base = out.blob.data ();
    RingTileHeader h;
    memset (&h, 0, sizeof h);
    h.nbRows = (uint16_t)nbRows; h.nbNodes = (uint16_t)out.nbNodes; h.nbBatches = (uint16_t)nbBatches;
    h.nbEntries = (uint16_t)out.nbEntries; h.hasInterface = out.hasInterface ? 1 : 0;
    h.offNodes = offNodes; h.headBytes = headBytes; h.offJobs = offJobs; h.offCodes = offCodes; h.blobBytes = at;
    memcpy (base, &h, sizeof h);
    if (!w.rows.empty ()) memcpy (base + sizeof h, w.rows.data (), w.rows.size () * sizeof (RingRow));
    int *nodesOut = reinterpret_cast<int*> (base + offNodes);
    for (int q = 0; q < out.nbNodes; q++) nodesOut[q] = w.nodes.empty () ? 0 : w.nodes[0];   // holes: any valid node
    for (int n = 0; n < nbRef; n++) nodesOut[w.newId[n]] = w.nodes[n];
    if (nbBatches > 0) memcpy (base + headBytes, w.batches.data (), (size_t)nbBatches * sizeof (RingBatch));
    uint64_t *jobsOut = reinterpret_cast<uint64_t*> (base + offJobs);
    uint64_t *codesOut = reinterpret_cast<uint64_t*> (base + offCodes);
    for (int b = 0; b < nbBatches; b++) {
        const RingBatch &rb = w.batches[b];
        const int nbWords = (rb.nbSteps + 7) / 8;
        for (int l = 0; l < 32; l++) {
            const int k = w.order[(size_t)b * 32 + l];
            if (k < 0) {
                jobsOut[(size_t)b * 32 + l] = ring_job (0, 0, 0xFFFF, 0xFFFF, 0);          // node 0 exists in every tile
                for (int wd = 0; wd < nbWords; wd++) codesOut[(size_t)rb.codeBase + (size_t)wd * 32 + l] = 0;
                continue;
            }
            const RingJob &jb = w.jobs[k];
            jobsOut[(size_t)b * 32 + l] = ring_job (w.newId[jb.i], w.newId[jb.j], jb.slotIJ, jb.slotJI, jb.codeLen);
            for (int wd = 0; wd < nbWords; wd++) {
                uint64_t word = 0;
                for (int q = 0; q < 8; q++) {
                    const int pos = wd * 8 + q;
                    int byte = w.newId[jb.i];                  // padding: any valid node, the step is masked
                    if (pos < jb.codeLen) {
                        const int c = w.codes[(size_t)jb.codeStart + pos];
                        byte = c == kRingBreak ? kRingBreak : w.newId[c];
                    }
                    word |= (uint64_t)byte << (8 * q);
                }
                codesOut[(size_t)rb.codeBase + (size_t)wd * 32 + l] = word;
            }
        }
    }
    release ();
}

}  // namespace

namespace {

// Recursive coordinate bisection: the nodes of [lo, hi) are split across the longest axis of their
// bounding box into two parts that hold a whole number of leaves of (nearly) equal size <= leafSize.
// Leaves come out box-shaped, which is what keeps mesh edges inside one tile (an interior edge is
// one job that writes both of its blocks); a run of the Morton curve of the same length is ragged.
void rcb_split (std::vector<int> &idx, int lo, int hi, int leafSize, const double *coord, std::vector<int> &leafStart, int depth = 0)
{
    const int count = hi - lo;
    if (count <= leafSize && leafSize > 0) { leafStart.push_back (lo); leafSize = 0; }   // below a leaf: keep bisecting, for the order only
    if (leafSize == 0 && count <= 2) return;
    double bmin[3], bmax[3];
    for (int a = 0; a < 3; a++) bmin[a] = bmax[a] = coord[(size_t)idx[lo] * 3 + a];
    for (int q = lo + 1; q < hi; q++) {
        for (int a = 0; a < 3; a++) {
            const double v = coord[(size_t)idx[q] * 3 + a];
            bmin[a] = std::min (bmin[a], v); bmax[a] = std::max (bmax[a], v);
        }
    }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (bmax[a] - bmin[a] > bmax[axis] - bmin[axis]) axis = a;
    int mid = lo + count / 2;
    if (leafSize > 0) {
        const int leaves = (count + leafSize - 1) / leafSize, leavesLeft = leaves / 2;
        mid = lo + (int)((int64_t)count * leavesLeft / leaves);
    }
    std::nth_element (idx.begin () + lo, idx.begin () + mid, idx.begin () + hi, [&] (int x, int y) {
        const double cx = coord[(size_t)x * 3 + axis], cy = coord[(size_t)y * 3 + axis];
        return cx < cy || (cx == cy && x < y);
    });
    if (depth < 6 && count > 20000) {
        // the two halves are disjoint ranges of idx: bisect them as tasks; the leaves of the right half follow those of
        // the left half, as in the sequential recursion
        std::vector<int> right;
        #pragma omp task shared(idx, leafStart) firstprivate(lo, mid, leafSize, depth)
        rcb_split (idx, lo, mid, leafSize, coord, leafStart, depth + 1);
        #pragma omp task shared(idx, right) firstprivate(mid, hi, leafSize, depth)
        rcb_split (idx, mid, hi, leafSize, coord, right, depth + 1);
        #pragma omp taskwait
        leafStart.insert (leafStart.end (), right.begin (), right.end ());
        return;
    }
    rcb_split (idx, lo, mid, leafSize, coord, leafStart, depth + 1);
    rcb_split (idx, mid, hi, leafSize, coord, leafStart, depth + 1);
}

// Tiles = the leaves of the bisection, cut further (greedily, in leaf order) wherever a leaf exceeds
// the entry / node caps.  Same outputs as cut_node_tiles.
int cut_node_tiles_rcb (int nbNodes, const int *elemToNode, const int *row, const int *col, const double *coord, const int *n2eIndex,
                        const int *n2eValue, const RingPlanLimits &lim, std::vector<int> &nodeOrder,
                        std::vector<int> &tileStart, std::string &error)
{
    nodeOrder.resize ((size_t)nbNodes);
    for (int n = 0; n < nbNodes; n++) nodeOrder[n] = n;
    std::vector<int> leafStart;
    if (nbNodes > 0) {
        #pragma omp parallel num_threads(plan_team_size ())
        #pragma omp single
        rcb_split (nodeOrder, 0, nbNodes, lim.maxRows, coord, leafStart);
    }
    leafStart.push_back (nbNodes);
    tileStart.clear ();
    const int nbLeaves = (int)leafStart.size () - 1;
    std::vector<std::vector<int>> tilesOfLeaf ((size_t)std::max (nbLeaves, 0));
    std::string firstError;
    bool anyFailed = false;
    // the leaves are cut independently of each other (every thread with its own stamps); their tiles are
    // concatenated in leaf order afterwards
    #pragma omp parallel num_threads(plan_team_size ())
    {
    std::vector<int> nodeStamp ((size_t)nbNodes, -1), ownedStamp ((size_t)nbNodes, -1), fresh;
    int stamp = 0;
    bool failed = false;
    std::string error;
    std::vector<int> *tileStartOut = nullptr;
    // does [lo, hi) of nodeOrder fit one tile?  (rows, referenced nodes, slab slots, jobs)
    auto fits = [&] (int lo, int hi) {
        stamp++;
        int rows = 0, refs = 0, entries = 0, jobs = 0;
        for (int at = lo; at < hi; at++) {
            const int n = nodeOrder[at], rowLen = row[n + 1] - row[n];
            int addRefs = (nodeStamp[n] != stamp) ? 1 : 0;
            fresh.clear ();
            for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
                const int *en = elemToNode + (size_t)n2eValue[p] * kDimElem;
                for (int k = 0; k < kDimElem; k++) {
                    const int m = en[k] - 1;
                    if (m != n && nodeStamp[m] != stamp && std::find (fresh.begin (), fresh.end (), m) == fresh.end ()) fresh.push_back (m);
                }
            }
            addRefs += (int)fresh.size ();
            const int slots = rowLen + ring_row_padding (rowLen);      // the slab keeps the padding too
            // an edge to a row the tile already owns is that row's job (it gains the transposed block)
            int addJobs = 0;
            for (int l = row[n]; l < row[n + 1]; l++) addJobs += (col[l] - 1 != n && ownedStamp[col[l] - 1] != stamp);
            rows++; refs += addRefs; entries += slots; jobs += addJobs;
            if (rows > lim.maxRows || refs > lim.maxNodes || entries > lim.maxEntries || (rows > 1 && jobs > lim.maxJobs)) return false;
            nodeStamp[n] = stamp;
            ownedStamp[n] = stamp;
            for (int m : fresh) nodeStamp[m] = stamp;
        }
        return true;
    };
    // a leaf that exceeds a cap is halved (its nodes are in bisection order below the leaf level too,
    // so both halves stay compact) instead of losing a ragged tail
    auto emit = [&] (auto &&self, int lo, int hi) -> void {
        if (failed || lo >= hi) return;
        if (fits (lo, hi)) { tileStartOut->push_back (lo); return; }
        if (hi - lo == 1) {
            const int n = nodeOrder[lo];
            error = "node " + std::to_string (n + 1) + " alone exceeds the tile caps (" + std::to_string (row[n + 1] - row[n]) + " entries)";
            failed = true;
            return;
        }
        const int mid = lo + (hi - lo + 1) / 2;
        self (self, lo, mid);
        self (self, mid, hi);
    };
    #pragma omp for schedule(dynamic, 64)
    for (int leaf = 0; leaf < nbLeaves; leaf++) {
        if (failed) continue;
        tileStartOut = &tilesOfLeaf[(size_t)leaf];
        emit (emit, leafStart[(size_t)leaf], leafStart[(size_t)leaf + 1]);
    }
    if (failed) {
        #pragma omp critical
        { if (!anyFailed) { anyFailed = true; firstError = error; } }
    }
    }
    if (anyFailed) { error = firstError; return -1; }
    for (int leaf = 0; leaf < nbLeaves; leaf++) tileStart.insert (tileStart.end (), tilesOfLeaf[(size_t)leaf].begin (), tilesOfLeaf[(size_t)leaf].end ());
    tileStart.push_back (nbNodes);
    if (nbNodes == 0) tileStart.assign (1, 0);
    return 0;
}

}  // namespace

int build_ring_plan (int nbNodes, int nbElem, const int *elemToNode, const int *row, const int *col,
                     const double *coord, const uint8_t *isInterface, const int *checkBounds,
                     const RingPlanLimits &lim, RingPlan &plan, std::string &error)
{
    plan = RingPlan ();
    if (lim.maxRows < 1 || lim.maxRows > 255 || lim.maxEntries < 1 || lim.maxEntries > 60000 ||
        lim.maxNodes < 4 || lim.maxNodes > kRingMaxNodes) {
        error = "ring plan limits out of range";
        return -1;
    }
    if (nbNodes > kRingMaxNodeId) { error = "more than 2^27 nodes in one subdomain (the row tags keep five flag bits)"; return -1; }
    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());
    std::vector<int> nodeOrder, tileStart;
    if (lim.bisection) {
        if (cut_node_tiles_rcb (nbNodes, elemToNode, row, col, coord, n2eIndex.data (), n2eValue.data (), lim,
                                nodeOrder, tileStart, error) != 0) return -1;
    }
    else {
        TileCutLimits cut = { lim.maxRows, 1 << 30, lim.maxNodes, lim.maxEntries };
        if (cut_node_tiles (nbNodes, nbElem, elemToNode, row, coord, n2eIndex.data (), n2eValue.data (), cut,
                            nodeOrder, tileStart, error) != 0) return -1;
    }
    const int nbTiles = (int)tileStart.size () - 1;
    std::vector<TileResult> results ((size_t)nbTiles);
    const int team = plan_team_size ();
    #pragma omp parallel num_threads(team)
    {
        std::vector<int> nodeLocal ((size_t)nbNodes, -1), colStamp ((size_t)nbNodes, 0);
        TileWork work;
        #pragma omp for schedule(dynamic, 16)
        for (int t = 0; t < nbTiles; t++) {
            plan_tile (tileStart[t + 1] - tileStart[t], nodeOrder.data () + tileStart[t], elemToNode, row, col,
                       n2eIndex.data (), n2eValue.data (), isInterface, checkBounds, nbNodes, lim, nodeLocal, colStamp, work, results[t]);
        }
    }
    // interface tiles first (they feed the halo exchange)
    std::vector<int> execOrder;
    for (int t = 0; t < nbTiles; t++) if (results[t].hasInterface) execOrder.push_back (t);
    plan.nbInterfaceTiles = (int)execOrder.size ();
    for (int t = 0; t < nbTiles; t++) if (!results[t].hasInterface) execOrder.push_back (t);
    plan.nbTiles = nbTiles;
    plan.tileOffset.assign ((size_t)nbTiles + 1, 0);
    for (int k = 0; k < nbTiles; k++) {
        const TileResult &r = results[execOrder[k]];
        if (r.bad == 1) { error = "CSR lacks a node pair of an element (tile " + std::to_string (execOrder[k]) + ")"; return -1; }
        if (r.bad == 2) { error = "an element names a node twice (tile " + std::to_string (execOrder[k]) + ")"; return -1; }
        if (r.bad == 3) { error = "tile references more than 254 nodes"; return -1; }
        plan.tileOffset[k + 1] = plan.tileOffset[k] + r.blob.size ();
        plan.maxBlobBytes = std::max (plan.maxBlobBytes, (uint32_t)r.blob.size ());
        plan.maxHeadBytes = std::max (plan.maxHeadBytes, r.headBytes);
        plan.maxTailBytes = std::max (plan.maxTailBytes, (uint32_t)r.blob.size () - r.headBytes);
        plan.maxRows = std::max (plan.maxRows, r.nbRows);
        plan.maxNodes = std::max (plan.maxNodes, r.nbNodes);
        plan.maxEntries = std::max (plan.maxEntries, r.nbEntries);
        plan.maxBatches = std::max (plan.maxBatches, r.nbBatches);
        plan.nbJobs += r.jobs; plan.nbSymmetricJobs += r.symmetric; plan.nbRingSteps += r.ringSteps;
        plan.nbPaddedSteps += r.paddedSteps; plan.nbBreaks += r.breaks;
        plan.gatherWavefronts += r.gatherWf; plan.gatherIdeal += r.gatherIdeal;
        plan.slabWriteWavefronts += r.slabWf; plan.slabWriteIdeal += r.slabIdeal;
    }
    plan.blob.assign ((size_t)plan.tileOffset[nbTiles], 0);
    #pragma omp parallel for schedule(dynamic, 64) num_threads(plan_team_size ())
    for (int k = 0; k < nbTiles; k++) {
        TileResult &r = results[execOrder[k]];
        memcpy (plan.blob.data () + plan.tileOffset[k], r.blob.data (), r.blob.size ());
        std::vector<uint8_t> ().swap (r.blob);
    }
    return 0;
}

int verify_ring_plan (const RingPlan &plan, int nbNodes, int nbElem, const int *elemToNode,
                      const int *row, const int *col, const int *checkBounds, std::string &error)
{
    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());
    std::vector<uint8_t> rowSeen ((size_t)nbNodes, 0);
    std::vector<uint16_t> pairSeen ((size_t)nbElem, 0);        // bit 4a+b per element
    std::vector<uint8_t> entrySeen (nbNodes > 0 ? (size_t)row[nbNodes] : 0, 0);
    int64_t rowsTotal = 0;
    bool interiorSeen = false;
    std::vector<int> elemsOfEdge;
    for (int t = 0; t < plan.nbTiles; t++) {
        const uint8_t *base = plan.blob.data () + plan.tileOffset[t];
        const RingTileHeader &h = *plan.header (t);
        if (h.blobBytes != plan.tileOffset[t + 1] - plan.tileOffset[t] || (plan.tileOffset[t] & 15) || (h.headBytes & 15) ||
            (h.offJobs & 15) || (h.offCodes & 15) || (h.blobBytes & 15)) { error = "record size / alignment"; return -1; }
        if (h.hasInterface && interiorSeen) { error = "interface tiles must come first"; return -1; }
        if (!h.hasInterface) interiorSeen = true;
        if ((t < plan.nbInterfaceTiles) != (h.hasInterface != 0)) { error = "nbInterfaceTiles mismatch"; return -1; }
        if (h.nbRows > plan.maxRows || h.nbNodes > plan.maxNodes || h.nbNodes > kRingMaxNodes || h.nbEntries > plan.maxEntries ||
            h.nbBatches > plan.maxBatches) { error = "tile exceeds the plan maxima"; return -1; }
        const RingRow *rows = reinterpret_cast<const RingRow*> (base + sizeof (RingTileHeader));
        const int *tileNodes = reinterpret_cast<const int*> (base + h.offNodes);
        const RingBatch *batches = reinterpret_cast<const RingBatch*> (base + h.headBytes);
        const uint64_t *jobs = reinterpret_cast<const uint64_t*> (base + h.offJobs);
        const uint64_t *codes = reinterpret_cast<const uint64_t*> (base + h.offCodes);
        for (int q = 0; q < h.nbNodes; q++) if (tileNodes[q] < 0 || tileNodes[q] >= nbNodes) { error = "node list out of range"; return -1; }
        // slab slot -> global CSR entry
        std::vector<int> slotEntry ((size_t)h.nbEntries, -1);
        int expectStart = 0;
        for (int r = 0; r < h.nbRows; r++) {
            const RingRow &rr = rows[r];
            const int n = rr.node & kRingNodeMask;
            if (n < 0 || n >= nbNodes || rowSeen[n]) { error = "row owned twice or out of range"; return -1; }
            if (rr.node & (1 << 27)) { error = "bit 27 of a row tag must be clear"; return -1; }
            if (checkBounds) {
                for (int c = 0; c < 3; c++) {
                    if (((rr.node >> (28 + c)) & 1) != (checkBounds[(size_t)c * nbNodes + n] != 0)) { error = "Dirichlet mask bits differ from checkBounds"; return -1; }
                }
            }
            rowSeen[n] = 1;
            rowsTotal++;
            if (rr.valueStart != row[n] || rr.len != row[n + 1] - row[n] || rr.localStart != expectStart) { error = "row table differs from nodeToNodeRow"; return -1; }
            for (int l = 0; l < rr.len; l++) slotEntry[(size_t)expectStart + l] = rr.valueStart + l;
            expectStart += rr.len + ring_row_padding (rr.len);
            if (expectStart > h.nbEntries) { error = "rows exceed the slab"; return -1; }
            int diag = 0xFFFF;
            for (int l = 0; l < rr.len; l++) if (col[rr.valueStart + l] == n + 1) { diag = l; break; }
            if (diag != rr.diagOff) { error = "diagOff is not the first diagonal entry"; return -1; }
            if (diag != 0xFFFF) {
                if (entrySeen[(size_t)rr.valueStart + diag]) { error = "diagonal entry written twice"; return -1; }
                entrySeen[(size_t)rr.valueStart + diag] = 1;
            }
        }
        if (expectStart != h.nbEntries) { error = "entry count mismatch"; return -1; }
        for (int b = 0; b < h.nbBatches; b++) {
            const RingBatch &rb = batches[b];
            const int nbWords = (rb.nbSteps + 7) / 8;
            bool batchHasBreak = false;
            for (int lane = 0; lane < 32; lane++) {
                const uint64_t job = jobs[(size_t)b * 32 + lane];
                const int li = (int)(job & 0xFF), lj = (int)((job >> 8) & 0xFF);
                const int sIJ = (int)((job >> 16) & 0xFFFF), sJI = (int)((job >> 32) & 0xFFFF), len = (int)(job >> 48);
                std::vector<int> bytes;
                for (int wd = 0; wd < nbWords; wd++) {
                    const uint64_t word = codes[(size_t)rb.codeBase + (size_t)wd * 32 + lane];
                    for (int q = 0; q < 8 && wd * 8 + q < rb.nbSteps; q++) bytes.push_back ((int)((word >> (8 * q)) & 0xFF));
                }
                if (len > rb.nbSteps) { error = "job longer than its batch"; return -1; }
                // bytes beyond the job's length are loaded and masked: they must name a node of the tile
                for (int q = len; q < rb.nbSteps; q++) if (bytes[q] >= h.nbNodes) { error = "padding byte is not a node of the tile"; return -1; }
                if (sIJ == 0xFFFF) {                                   // idle lane
                    if (sJI != 0xFFFF || len != 0) { error = "idle lane with a transposed slot or codes"; return -1; }
                    continue;
                }
                bytes.resize ((size_t)len);
                if (li >= h.nbNodes || lj >= h.nbNodes || sIJ >= h.nbEntries || (sJI != 0xFFFF && sJI >= h.nbEntries)) { error = "job out of range"; return -1; }
                const int gi = tileNodes[li], gj = tileNodes[lj];
                const int eIJ = slotEntry[sIJ];
                if (eIJ < 0) { error = "job writes a padding slot"; return -1; }
                if (entrySeen[eIJ]) { error = "CSR entry written twice"; return -1; }
                entrySeen[eIJ] = 1;
                const bool empty = len == 0;
                // the entry must sit in row i; with elements to add, its column must be j
                {
                    int r = 0;
                    while (r < h.nbRows && !(sIJ >= rows[r].localStart && sIJ < rows[r].localStart + rows[r].len)) r++;
                    if (r == h.nbRows || (rows[r].node & kRingNodeMask) != gi) { error = "job slot is not in row i"; return -1; }
                    if (!empty && col[eIJ] != gj + 1) { error = "job slot is not column j"; return -1; }
                }
                if (sJI != 0xFFFF) {
                    const int eJI = slotEntry[sJI];
                    if (eJI < 0) { error = "job writes a padding slot (transposed block)"; return -1; }
                    if (entrySeen[eJI]) { error = "CSR entry written twice (transposed block)"; return -1; }
                    entrySeen[eJI] = 1;
                    int r = 0;
                    while (r < h.nbRows && !(sJI >= rows[r].localStart && sJI < rows[r].localStart + rows[r].len)) r++;
                    if (r == h.nbRows || (rows[r].node & kRingNodeMask) != gj || col[eJI] != gi + 1) { error = "transposed slot is not entry (j, i)"; return -1; }
                }
                // walk the chains: consecutive nodes p, q name the element {i, j, p, q}
                elemsOfEdge.clear ();
                int prev = -1;
                for (int c : bytes) {
                    if (c == kRingBreak) { prev = -1; batchHasBreak = true; continue; }
                    if (c >= h.nbNodes) { error = "ring node out of range"; return -1; }
                    const int g = tileNodes[c];
                    if (prev >= 0) {
                        int found = -1, a = -1, bb = -1;
                        for (int p = n2eIndex[gi]; p < n2eIndex[gi + 1] && found < 0; p++) {
                            const int e = n2eValue[p];
                            const int *en = elemToNode + (size_t)e * kDimElem;
                            int hit = 0, ai = -1, aj = -1;
                            for (int k = 0; k < kDimElem; k++) {
                                if (en[k] - 1 == gi) { hit |= 1; ai = k; }
                                else if (en[k] - 1 == gj) { hit |= 2; aj = k; }
                                else if (en[k] - 1 == prev) hit |= 4;
                                else if (en[k] - 1 == g) hit |= 8;
                            }
                            if (hit == 15 && std::find (elemsOfEdge.begin (), elemsOfEdge.end (), e) == elemsOfEdge.end ()) { found = e; a = ai; bb = aj; }
                        }
                        if (found < 0) { error = "ring step names no (new) element of the edge"; return -1; }
                        elemsOfEdge.push_back (found);
                        const uint16_t bitIJ = (uint16_t)(1u << (4 * a + bb)), bitJI = (uint16_t)(1u << (4 * bb + a));
                        if (pairSeen[found] & bitIJ) { error = "element pair covered twice"; return -1; }
                        pairSeen[found] |= bitIJ;
                        if (sJI != 0xFFFF) {
                            if (pairSeen[found] & bitJI) { error = "element pair covered twice (transposed)"; return -1; }
                            pairSeen[found] |= bitJI;
                        }
                    }
                    prev = g;
                }
            }
            // the kernel's regular loop takes every byte for a node: a break needs the general loop
            if (batchHasBreak != ((rb.flags & kRingBatchGeneral) != 0)) { error = "batch flag does not match its codes"; return -1; }
        }
    }
    if (rowsTotal != nbNodes) { error = "not every node is owned by a tile"; return -1; }
    for (size_t l = 0; l < entrySeen.size (); l++) if (!entrySeen[l]) { error = "CSR entry " + std::to_string (l) + " is never written"; return -1; }
    for (int e = 0; e < nbElem; e++) {
        if (pairSeen[e] != 0x7BDE) { error = "element " + std::to_string (e) + " misses an off-diagonal pair"; return -1; }   // all 4a+b, a != b
    }
    return 0;
}

}  // namespace mfb
