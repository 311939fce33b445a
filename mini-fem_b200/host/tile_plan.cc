#include "tile_plan.h"

#include <omp.h>

#include <cstdlib>
#include <thread>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "mesh_topology.h"

namespace mfb {

int plan_team_size ()
{
    if (const char *v = getenv ("MFB_PLAN_THREADS")) return std::max (1, std::min (atoi (v), 256));
    int hw = (int)std::thread::hardware_concurrency ();
    if (hw <= 0) hw = omp_get_num_procs ();
    int local = 1;
    if (const char *v = getenv ("LOCAL_WORLD_SIZE")) local = std::max (1, atoi (v));
    return std::max (1, std::min (hw / local, 32));
}


// smallest s = 4 (mod 16) such that s - 4 (the Laplacian's plane pitch) holds maxElems + 3 ids
// and the 16 zero slots
int tile_plan_stride (int maxElems) { return ((maxElems + 3 + 16 + 15) / 16) * 16 + 4; }

namespace {

// Interleave the low 21 bits of x, y, z.
inline uint64_t spread21 (uint64_t v)
{
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8))  & 0x100F00F00F00F00Full;
    v = (v | (v << 4))  & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2))  & 0x1249249249249249ull;
    return v;
}

inline uint32_t align16 (uint32_t x) { return (x + 15u) & ~15u; }

struct TileScratch {
    std::vector<int> elems;                   // global element ids, first-touch order
    std::vector<int> nodes;                   // global node ids, owned first
    std::vector<uint16_t> elemNodes;          // 4 local ids per element
    std::vector<TileRow> rows;                // + sentinel
    std::vector<uint8_t> entryRow;
    std::vector<uint16_t> laneEntry;          // per off-diagonal-pass lane: tile-local entry or 0xFFFF
    int nbIds = 0;                            // element ids incl. holes of the coset numbering
    std::vector<TileBatch> batches;
    std::vector<uint16_t> pairCodes, diagCodes;
    int64_t contributions = 0, paddedSteps = 0;
    bool bad = false, hasInterface = false;
    uint32_t blobBytes = 0, headBytes = 0;
};

// Shared-memory bank pair (8-byte words, 16 per 128-byte line) of coefficient slot
// (local node a, element e); the planes start on a 128-byte boundary.
inline int bank_of (int a, int e, int stride) { return (a * stride + e) & 15; }

// Schedules the off-diagonal contributions of one ROW CHUNK (<= 16 consecutive entries of one
// CSR row = one half-warp).  The (up to) three lanes whose column nodes belong to one element
// are served in the same step: they read the same row-side slot (a broadcast) and three
// different column-side planes.  Greedy per step: the element covering the most lanes that
// cannot wait, then the most free lanes, at most four elements unless a lane without slack
// needs a fifth.  The bank classes of the elements are chosen AFTERWARDS (coset numbering)
// so that the elements of one step get different classes.
// out[step*16 + lane] = code or 0xFFFF.  Returns the number of steps.
int schedule_chunk (const std::vector<uint16_t> *lists, int nbLanes, bool bankAware, std::vector<uint16_t> &out)
{
    int T = 0, remaining[16], total = 0;
    for (int l = 0; l < nbLanes; l++) {
        remaining[l] = (int)lists[l].size ();
        T = std::max (T, remaining[l]);
        total += remaining[l];
    }
    out.clear ();
    if (!bankAware) {                               // plain element order, no gaps
        out.assign ((size_t)T * 16, 0xFFFF);
        for (int l = 0; l < nbLanes; l++) for (size_t t = 0; t < lists[l].size (); t++) out[t * 16 + l] = lists[l][t];
        return T;
    }
    // contributions grouped by element, groups in order of first appearance: group g owns the
    // (lane, code) pairs gPairs[gStart[g] .. gStart[g] + gCount[g]) that are still to be placed
    static thread_local std::vector<int> gElem, gStart, gCount;
    static thread_local std::vector<std::pair<int, uint16_t>> gPairs;
    static thread_local std::vector<int16_t> groupOf (4096, -1);      // element (12-bit field) -> group, -1 between calls
    gElem.clear (); gCount.clear ();
    for (int l = 0; l < nbLanes; l++) {
        for (uint16_t code : lists[l]) {
            const int e = code >> 4;
            if (groupOf[e] < 0) { groupOf[e] = (int16_t)gElem.size (); gElem.push_back (e); gCount.push_back (0); }
            gCount[groupOf[e]]++;
        }
    }
    const int nbGroups = (int)gElem.size ();
    gStart.assign ((size_t)nbGroups + 1, 0);
    for (int g = 0; g < nbGroups; g++) { gStart[g + 1] = gStart[g] + gCount[g]; gCount[g] = 0; }
    gPairs.resize ((size_t)total);
    for (int l = 0; l < nbLanes; l++) {
        for (uint16_t code : lists[l]) {
            const int g = groupOf[code >> 4];
            gPairs[(size_t)gStart[g] + gCount[g]++] = {l, code};
        }
    }
    for (int e : gElem) groupOf[e] = -1;
    out.reserve ((size_t)(T + 2) * 16);
    int step = 0;
    while (total > 0) {
        out.resize ((size_t)(step + 1) * 16, 0xFFFF);
        uint16_t *cur = &out[(size_t)step * 16];
        const int stepsLeft = std::max (T - step, 1);
        bool busy[16] = {false}, must[16];
        int nMustOpen = 0, nChosen = 0;
        for (int l = 0; l < nbLanes; l++) { must[l] = remaining[l] >= stepsLeft; nMustOpen += must[l]; }
        for (;;) {
            // element covering the most lanes that cannot wait, then the most free lanes,
            // in a bank class this step does not use yet
            int best = -1, bestScore = 0;
            for (int g = 0; g < nbGroups; g++) {
                const std::pair<int, uint16_t> *pairs = &gPairs[(size_t)gStart[g]];
                int nMust = 0, nFree = 0;
                for (int k = 0; k < gCount[g]; k++) if (!busy[pairs[k].first]) { nFree++; nMust += must[pairs[k].first]; }
                if (nFree == 0) continue;
                if (nChosen >= 4 && nMust == 0) continue;            // four bank classes per step
                if (nMustOpen == 0 && nFree < 2 && nChosen > 0) continue;   // do not open an element for one idle lane
                const int score = nMust * 64 + nFree * 4 + gCount[g];
                if (score > bestScore) { bestScore = score; best = g; }
            }
            if (best < 0) break;
            std::pair<int, uint16_t> *pairs = &gPairs[(size_t)gStart[best]];
            nChosen++;
            int kept = 0;
            for (int k = 0; k < gCount[best]; k++) {
                const int l = pairs[k].first;
                if (busy[l]) { pairs[kept++] = pairs[k]; continue; }
                cur[l] = pairs[k].second;
                busy[l] = true; remaining[l]--; total--;
                if (must[l]) nMustOpen--;
            }
            gCount[best] = kept;
        }
        step++;
        if (step > 4096) break;                     // cannot happen: every step places >= 1 code
    }
    return step;
}

// Same idea for the diagonal pass: 4 lanes share one row's list (lane `sub` takes codes
// sub, sub+4, ...), 4 rows per half-warp.
void order_diag_half_warp (std::vector<uint16_t> *rowLists, int nbRows, int stride)
{
    size_t steps = 0;
    for (int r = 0; r < nbRows; r++) steps = std::max (steps, (rowLists[r].size () + 3) / 4);
    std::vector<std::vector<uint16_t>> out ((size_t)nbRows);
    std::vector<std::vector<uint16_t>> rest (rowLists, rowLists + nbRows);
    for (size_t t = 0; t < steps; t++) {
        int cnt[16] = {0};
        // one code per (row, sub-lane); the row with the fewest codes left chooses first
        int orderRows[4] = {0, 1, 2, 3};
        for (int i = 1; i < nbRows; i++) {                       // insertion sort of <= 4 rows
            for (int j = i; j > 0 && rest[orderRows[j]].size () < rest[orderRows[j - 1]].size (); j--) std::swap (orderRows[j], orderRows[j - 1]);
        }
        uint16_t picked[4][4];
        int nPicked[4] = {0, 0, 0, 0};
        for (int sub = 0; sub < 4; sub++) {
            for (int oi = 0; oi < nbRows; oi++) {
                const int r = orderRows[oi];
                if (rest[r].empty ()) continue;
                // free bank first; among free banks the one this row still has most codes in
                // (keeps rare banks for the late steps, when choice runs out)
                int have[16] = {0};
                for (uint16_t c : rest[r]) have[bank_of (c & 3, c >> 2, stride)]++;
                int best = 0, bestCost = 1 << 30;
                for (size_t k = 0; k < rest[r].size (); k++) {
                    const int b = bank_of (rest[r][k] & 3, rest[r][k] >> 2, stride);
                    const int cost = cnt[b] * 1024 - have[b];
                    if (cost < bestCost) { bestCost = cost; best = (int)k; }
                }
                const uint16_t code = rest[r][best];
                rest[r].erase (rest[r].begin () + best);
                picked[r][nPicked[r]++] = code;
                cnt[bank_of (code & 3, code >> 2, stride)]++;
            }
        }
        for (int r = 0; r < nbRows; r++) for (int k = 0; k < nPicked[r]; k++) out[r].push_back (picked[r][k]);
    }
    for (int r = 0; r < nbRows; r++) rowLists[r].swap (out[r]);
}

}  // namespace

int cut_node_tiles (int nbNodes, int nbElem, const int *elemToNode, const int *row, const double *coord,
                    const int *n2eIndexPtr, const int *n2eValuePtr, const TileCutLimits &lim,
                    std::vector<int> &nodeOrder, std::vector<int> &tileStart, std::string &error)
{
    const int *n2eIndex = n2eIndexPtr, *n2eValue = n2eValuePtr;
    // ---- 1. spatial order of the nodes -------------------------------------------
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    if (nbNodes > 0) {
        for (int a = 0; a < 3; a++) lo[a] = hi[a] = coord[a];
        for (int n = 1; n < nbNodes; n++) {
            for (int a = 0; a < 3; a++) {
                lo[a] = std::min (lo[a], coord[(size_t)n * 3 + a]);
                hi[a] = std::max (hi[a], coord[(size_t)n * 3 + a]);
            }
        }
    }
    // one common scale keeps cells cubic; 2^21 cells along the longest axis
    double extent = std::max ({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
    const double scale = 2097151.0 / extent;
    std::vector<std::pair<uint64_t, int>> order ((size_t)nbNodes);
    #pragma omp parallel for schedule(static) num_threads(plan_team_size ())
    for (int n = 0; n < nbNodes; n++) {
        uint64_t q[3];
        for (int a = 0; a < 3; a++) q[a] = (uint64_t)((coord[(size_t)n * 3 + a] - lo[a]) * scale);
        order[n] = { spread21 (q[0]) | (spread21 (q[1]) << 1) | (spread21 (q[2]) << 2), n };
    }
    std::sort (order.begin (), order.end ());

    // ---- 2. greedy cut under the shared-memory caps -------------------------------
    tileStart.clear ();
    {
        std::vector<int> elemStamp ((size_t)nbElem, -1), nodeStamp ((size_t)nbNodes, -1);
        int rows = 0, elems = 0, refs = 0, entries = 0, tile = 0;
        std::vector<int> fresh;
        tileStart.push_back (0);
        for (int at = 0; at < nbNodes; at++) {
            const int n = order[at].second;
            const int rowLen = row[n + 1] - row[n];
            for (int attempt = 0; attempt < 2; attempt++) {
                int addElems = 0, addRefs = (nodeStamp[n] != tile) ? 1 : 0;
                fresh.clear ();
                for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
                    const int e = n2eValue[p];
                    if (elemStamp[e] == tile) continue;
                    addElems++;
                    for (int k = 0; k < kDimElem; k++) {
                        const int m = elemToNode[(size_t)e * kDimElem + k] - 1;
                        if (m != n && nodeStamp[m] != tile &&
                            std::find (fresh.begin (), fresh.end (), m) == fresh.end ()) fresh.push_back (m);
                    }
                }
                addRefs += (int)fresh.size ();
                const bool fits = rows + 1 <= lim.maxRows && elems + addElems <= lim.maxElems &&
                                  refs + addRefs <= lim.maxNodesRef && entries + rowLen <= lim.maxEntries;
                if (fits) {
                    nodeStamp[n] = tile;
                    for (int m : fresh) nodeStamp[m] = tile;
                    for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) elemStamp[n2eValue[p]] = tile;
                    rows++; elems += addElems; refs += addRefs; entries += rowLen;
                    break;
                }
                if (rows == 0 || attempt == 1) {
                    error = "node " + std::to_string (n + 1) + " alone exceeds the tile caps (" +
                            std::to_string (addElems) + " elements, " + std::to_string (rowLen) + " entries)";
                    return -1;
                }
                tile++;                       // close the tile, retry the node in a fresh one
                tileStart.push_back (at);
                rows = elems = refs = entries = 0;
            }
        }
        tileStart.push_back (nbNodes);
        if (nbNodes == 0) tileStart.assign (1, 0);
    }
    nodeOrder.resize ((size_t)nbNodes);
    for (int at = 0; at < nbNodes; at++) nodeOrder[at] = order[at].second;
    return 0;
}

int build_tile_plan (int nbNodes, int nbElem, const int *elemToNode, const int *row,
                     const int *col, const double *coord, const uint8_t *isInterface,
                     const TilePlanLimits &lim, TilePlan &plan, std::string &error)
{
    plan = TilePlan ();
    if (lim.maxRows < 1 || lim.maxRows > 255 || lim.maxElems < 1 || lim.maxElems > 4064 ||
        lim.maxNodesRef < 4 || lim.maxNodesRef > 65535 || lim.maxEntries < 1 || lim.maxEntries > 65535) {
        error = "tile plan limits out of range";
        return -1;
    }
    // plane stride = 4 (mod 16): the 4 local-node planes of element e occupy the 4 banks of the
    // coset e mod 4 (8-byte words, 16 bank pairs); room for the ids (<= maxElems + 3 with the
    // coset numbering's holes) and for the 16 all-zero slots the padding codes point at
    const int stride = tile_plan_stride (lim.maxElems);
    plan.elemStride = stride;
    plan.laplacian = lim.laplacian;

    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());

    // ---- 1 + 2. spatial order of the nodes, greedy cut under the shared-memory caps ----
    std::vector<int> nodeOrder, tileStart;
    {
        TileCutLimits cut = { lim.maxRows, lim.maxElems, lim.maxNodesRef, lim.maxEntries };
        if (cut_node_tiles (nbNodes, nbElem, elemToNode, row, coord, n2eIndex.data (), n2eValue.data (), cut,
                            nodeOrder, tileStart, error) != 0) return -1;
    }
    const int nbTiles = (int)tileStart.size () - 1;

    // ---- 3. per-tile tables (independent) ------------------------------------------
    std::vector<TileScratch> scratch ((size_t)nbTiles);
    // each thread keeps two dense global->local maps; cap the team so that the maps of a
    // 48 M-element mesh stay within a few GB on a many-core host
    const int team = plan_team_size ();
    #pragma omp parallel num_threads(team)
    {
        std::vector<int> nodeLocal ((size_t)nbNodes, -1), elemLocal ((size_t)nbElem, -1);
        std::vector<std::vector<uint16_t>> lists, diagLists;
        std::vector<int> meetStart, meetFill, meetList, setStart, setCount, setList;   // reused from tile to tile
        #pragma omp for schedule(dynamic, 16)
        for (int t = 0; t < nbTiles; t++) {
            TileScratch &s = scratch[t];
            const int first = tileStart[t], nbRows = tileStart[t + 1] - first;
            for (int r = 0; r < nbRows; r++) {                       // owned rows first
                const int n = nodeOrder[first + r];
                nodeLocal[n] = r;
                s.nodes.push_back (n);
            }
            for (int r = 0; r < nbRows; r++) {
                const int n = s.nodes[r];
                for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
                    const int e = n2eValue[p];
                    if (elemLocal[e] >= 0) continue;
                    elemLocal[e] = (int)s.elems.size ();
                    s.elems.push_back (e);
                    for (int k = 0; k < kDimElem; k++) {
                        const int m = elemToNode[(size_t)e * kDimElem + k] - 1;
                        if (nodeLocal[m] < 0) { nodeLocal[m] = (int)s.nodes.size (); s.nodes.push_back (m); }
                        s.elemNodes.push_back ((uint16_t)nodeLocal[m]);
                    }
                }
            }
            int nbEntries = 0;
            for (int r = 0; r < nbRows; r++) nbEntries += row[s.nodes[r] + 1] - row[s.nodes[r]];
            if ((int)lists.size () < nbEntries + 32) lists.resize (nbEntries + 32);
            for (int q = 0; q < nbEntries + 32; q++) lists[q].clear ();
            if ((int)diagLists.size () < nbRows + 8) diagLists.resize (nbRows + 8);
            for (int r = 0; r < nbRows + 8; r++) diagLists[r].clear ();
            s.entryRow.resize (nbEntries);

            const int nbTileElems = (int)s.elems.size ();
            int localStart = 0;
            for (int r = 0; r < nbRows; r++) {
                const int n = s.nodes[r], begin = row[n], end = row[n + 1];
                TileRow tr;
                const bool intf = isInterface && isInterface[n];
                s.hasInterface |= intf;
                tr.node = n | (intf ? (int)0x80000000u : 0);
                tr.valueStart = begin;
                tr.diagCodeBase = 0;
                tr.localStart = (uint16_t)localStart;
                tr.diagLocal = 0xFFFF;
                for (int l = begin; l < end; l++) {
                    s.entryRow[localStart + (l - begin)] = (uint8_t)r;
                    if (col[l] == n + 1 && tr.diagLocal == 0xFFFF) tr.diagLocal = (uint16_t)(localStart + (l - begin));
                }
                for (int p = n2eIndex[n]; p < n2eIndex[n + 1]; p++) {
                    const int e = n2eValue[p], el = elemLocal[e];
                    const int *nodes = elemToNode + (size_t)e * kDimElem;
                    int a = 0;
                    while (nodes[a] != n + 1) a++;
                    for (int b = 0; b < kDimElem; b++) {
                        if (b == a) {
                            diagLists[r].push_back ((uint16_t)((el << 2) | a));
                            continue;
                        }
                        int l = begin;
                        while (l < end && col[l] != nodes[b]) l++;
                        if (l == end) { s.bad = true; continue; }
                        lists[localStart + (l - begin)].push_back ((uint16_t)((el << 4) | (a << 2) | b));
                    }
                }
                s.rows.push_back (tr);
                localStart += end - begin;
            }

            // off-diagonal codes: every row starts on a half-warp boundary (chunks of <= 16
            // consecutive entries), chunks are scheduled independently, two chunks of similar
            // length share a warp; codes are stored transposed per warp [step][32 lanes]
            struct Chunk { int firstEntry, nbLanes, steps; std::vector<uint16_t> codes; };
            std::vector<Chunk> chunks;
            for (int r = 0; r < nbRows; r++) {
                const int len = row[s.nodes[r] + 1] - row[s.nodes[r]], first = s.rows[r].localStart;
                for (int c0 = 0; c0 < len; c0 += 16) {
                    Chunk ch;
                    ch.firstEntry = first + c0; ch.nbLanes = std::min (16, len - c0);
                    ch.steps = schedule_chunk (&lists[ch.firstEntry], ch.nbLanes, lim.bankAware, ch.codes);
                    for (int l = 0; l < ch.nbLanes; l++) s.contributions += (int64_t)lists[ch.firstEntry + l].size ();
                    chunks.push_back (std::move (ch));
                }
            }
            // -- coset numbering: element id mod 4 = class.  Elements that meet in one step of one
            //    chunk should differ in class; greedy colouring of that conflict graph, classes
            //    kept equally large so that the ids stay dense.
            std::vector<int> newId ((size_t)nbTileElems);
            int nbIds = nbTileElems;
            if (lim.bankAware && nbTileElems > 0) {
                // meets: CSR over the elements, filled in two sweeps over the scheduled steps
                meetStart.assign ((size_t)nbTileElems + 1, 0);
                auto for_each_step = [&] (auto &&visit) {
                    for (const Chunk &ch : chunks) {
                        for (int t = 0; t < ch.steps; t++) {
                            int present[16], nPresent = 0;
                            for (int l = 0; l < ch.nbLanes; l++) {
                                const uint16_t code = ch.codes[(size_t)t * 16 + l];
                                if (code == 0xFFFF) continue;
                                const int e = code >> 4;
                                bool known = false;
                                for (int u = 0; u < nPresent; u++) known |= present[u] == e;
                                if (!known) present[nPresent++] = e;
                            }
                            visit (present, nPresent);
                        }
                    }
                };
                for_each_step ([&] (const int *present, int nPresent) {
                    for (int u = 0; u < nPresent; u++) meetStart[(size_t)present[u] + 1] += nPresent - 1;
                });
                for (int el = 0; el < nbTileElems; el++) meetStart[(size_t)el + 1] += meetStart[el];
                meetFill.assign (meetStart.begin (), meetStart.end () - 1);
                meetList.resize ((size_t)meetStart[nbTileElems]);
                for_each_step ([&] (const int *present, int nPresent) {
                    for (int u = 0; u < nPresent; u++) for (int v = 0; v < nPresent; v++) if (u != v) meetList[(size_t)meetFill[present[u]]++] = present[v];
                });
                auto degree = [&] (int el) { return meetStart[(size_t)el + 1] - meetStart[el]; };
                std::vector<int> cls ((size_t)nbTileElems, -1), byDegree ((size_t)nbTileElems);
                for (int el = 0; el < nbTileElems; el++) byDegree[el] = el;
                std::stable_sort (byDegree.begin (), byDegree.end (), [&] (int x, int y) { return degree (x) > degree (y); });
                int classSize[4] = {0, 0, 0, 0};
                const int classCap = (nbTileElems + 3) / 4;
                for (int el : byDegree) {
                    int clash[4] = {0, 0, 0, 0};
                    for (int q = meetStart[el]; q < meetStart[(size_t)el + 1]; q++) if (cls[meetList[q]] >= 0) clash[cls[meetList[q]]]++;
                    int bestClass = -1, bestCost = 1 << 30;
                    for (int c = 0; c < 4; c++) {
                        if (classSize[c] >= classCap) continue;
                        const int cost = clash[c] * 4096 + classSize[c];
                        if (cost < bestCost) { bestCost = cost; bestClass = c; }
                    }
                    cls[el] = bestClass;
                    classSize[bestClass]++;
                }
                int next[4] = {0, 0, 0, 0};
                for (int el = 0; el < nbTileElems; el++) newId[el] = 4 * next[cls[el]]++ + cls[el];
                nbIds = 4 * std::max ({classSize[0], classSize[1], classSize[2], classSize[3]});
                std::vector<uint16_t> renum ((size_t)nbIds * 4, 0xFFFF);        // holes keep 0xFFFF
                for (int el = 0; el < nbTileElems; el++) {
                    for (int k = 0; k < 4; k++) renum[(size_t)newId[el] * 4 + k] = s.elemNodes[(size_t)el * 4 + k];
                }
                s.elemNodes.swap (renum);
                for (Chunk &ch : chunks) {
                    for (uint16_t &code : ch.codes) if (code != 0xFFFF) code = (uint16_t)((newId[code >> 4] << 4) | (code & 15));
                }
                for (int r = 0; r < nbRows; r++) {
                    for (uint16_t &code : diagLists[r]) code = (uint16_t)((newId[code >> 2] << 2) | (code & 3));
                }
            }
            s.nbIds = nbIds;

            // -- halo node numbering against bank conflicts of the coefficient phase: thread e reads
            //    the coordinates of its 4 nodes from planes indexed by local node id, 16 consecutive
            //    element ids per half-warp; owned nodes keep id = row, every other referenced node
            //    gets the residue (id mod 16) least used by the nodes it shares such a read with
            if (lim.bankAware && nbIds > 0) {
                const int nbRef = (int)s.nodes.size ();
                const int nbSets = ((nbIds + 15) / 16) * 4;
                // setList[setStart[n] .. + setCount[n]): the distinct read sets node n takes part in
                setStart.assign ((size_t)nbRef + 1, 0);
                for (int id = 0; id < nbIds; id++) {
                    if (s.elemNodes[(size_t)id * 4] == 0xFFFF) continue;
                    for (int k = 0; k < 4; k++) setStart[(size_t)s.elemNodes[(size_t)id * 4 + k] + 1]++;
                }
                for (int n = 0; n < nbRef; n++) setStart[(size_t)n + 1] += setStart[n];
                setCount.assign ((size_t)nbRef, 0);
                setList.resize ((size_t)setStart[nbRef]);
                for (int id = 0; id < nbIds; id++) {
                    if (s.elemNodes[(size_t)id * 4] == 0xFFFF) continue;
                    for (int k = 0; k < 4; k++) {
                        const int n = s.elemNodes[(size_t)id * 4 + k], set = (id / 16) * 4 + k;
                        int *v = &setList[(size_t)setStart[n]];
                        if (std::find (v, v + setCount[n], set) == v + setCount[n]) v[setCount[n]++] = set;
                    }
                }
                auto sets_of = [&] (int n, auto &&visit) {
                    const int *v = &setList[(size_t)setStart[n]];
                    for (int q = 0; q < setCount[n]; q++) visit (v[q]);
                };
                std::vector<uint8_t> used ((size_t)nbSets * 16, 0);
                std::vector<int> newNode ((size_t)nbRef, -1), nextFree (16);
                for (int n = 0; n < nbRows; n++) {
                    newNode[n] = n;
                    sets_of (n, [&] (int set) { used[(size_t)set * 16 + (n & 15)]++; });
                }
                for (int r = 0; r < 16; r++) { nextFree[r] = nbRows + ((r - nbRows) & 15); }
                int maxId = nbRows - 1;
                for (int n = nbRows; n < nbRef; n++) {
                    int bestRes = 0, bestCost = 1 << 30;
                    for (int r = 0; r < 16; r++) {
                        if (nextFree[r] >= lim.maxNodesRef) continue;
                        int cost = 0;
                        sets_of (n, [&] (int set) { cost += used[(size_t)set * 16 + r]; });
                        cost = cost * 65536 + nextFree[r];
                        if (cost < bestCost) { bestCost = cost; bestRes = r; }
                    }
                    newNode[n] = nextFree[bestRes];
                    nextFree[bestRes] += 16;
                    maxId = std::max (maxId, newNode[n]);
                    sets_of (n, [&] (int set) { used[(size_t)set * 16 + bestRes]++; });
                }
                std::vector<int> renumNodes ((size_t)maxId + 1, s.nodes.empty () ? 0 : s.nodes[0]);   // holes: any valid node
                for (int n = 0; n < nbRef; n++) renumNodes[newNode[n]] = s.nodes[n];
                s.nodes.swap (renumNodes);
                for (uint16_t &ln : s.elemNodes) if (ln != 0xFFFF) ln = (uint16_t)newNode[ln];
            }

            // diagonal codes, 4 rows (= one half-warp of the diagonal pass) at a time
            for (int r0 = 0; r0 < nbRows; r0 += 4) {
                if (lim.bankAware) order_diag_half_warp (&diagLists[r0], std::min (4, nbRows - r0), stride);
            }
            for (int r = 0; r < nbRows; r++) {
                s.rows[r].diagCodeBase = (int)s.diagCodes.size ();
                s.diagCodes.insert (s.diagCodes.end (), diagLists[r].begin (), diagLists[r].end ());
            }
            TileRow sentinel = {0, 0, (int)s.diagCodes.size (), 0, 0xFFFF};
            s.rows.push_back (sentinel);
            s.contributions += (int64_t)s.diagCodes.size ();

            std::vector<int> byLength (chunks.size ());
            for (size_t c = 0; c < chunks.size (); c++) byLength[c] = (int)c;
            std::stable_sort (byLength.begin (), byLength.end (), [&] (int x, int y) { return chunks[x].steps > chunks[y].steps; });
            // Idle lanes and gap steps get a padding code that names one of the 16 all-zero slots
            // [nbIds, nbIds + 16) (local nodes 1 / 0): it adds nothing.  Where the half-warp has a
            // live lane in that step, the slot is picked so that the padding lands in banks the
            // live lanes leave free: on the column side the bank of a live element's row-side
            // slot, on the row side a column-side bank of that element.
            const uint16_t zeroPad = (uint16_t)((nbIds << 4) | (1 << 2) | 0);   // local nodes 1 / 0 like every padding code
            for (size_t c = 0; c < byLength.size (); c += 2) {
                const Chunk *half[2] = { &chunks[byLength[c]], c + 1 < byLength.size () ? &chunks[byLength[c + 1]] : nullptr };
                const int steps = half[0]->steps;            // sorted: the first is the longer one
                TileBatch tb = { (int)s.pairCodes.size (), steps };
                s.batches.push_back (tb);
                s.pairCodes.resize (s.pairCodes.size () + (size_t)steps * 32, zeroPad);
                for (int h = 0; h < 2; h++) {
                    for (int l = 0; l < 16; l++) {
                        const bool liveLane = half[h] && l < half[h]->nbLanes;
                        s.laneEntry.push_back (liveLane ? (uint16_t)(half[h]->firstEntry + l) : (uint16_t)0xFFFF);
                    }
                    if (!half[h]) continue;
                    for (int t = 0; t < half[h]->steps; t++) {
                        uint16_t pad = zeroPad;
                        for (int l = 0; l < half[h]->nbLanes && pad == zeroPad; l++) {
                            const uint16_t code = half[h]->codes[(size_t)t * 16 + l];
                            if (code != 0xFFFF) {
                                // ela: the bank of the live lane's row-side slot; lap (one dot product
                                // per contribution, banks {c, c+4, c+8} of the class): c + 12
                                const int freeBank = lim.laplacian ? (((code >> 4) & 3) + 12)
                                                                   : ((4 * ((code >> 2) & 3) + (code >> 4)) & 15);
                                const int ez = nbIds + ((freeBank - nbIds) & 15);                  // ez = freeBank (mod 16)
                                pad = (uint16_t)((ez << 4) | (1 << 2) | 0);
                            }
                        }
                        for (int l = 0; l < 16; l++) {
                            const uint16_t code = l < half[h]->nbLanes ? half[h]->codes[(size_t)t * 16 + l] : (uint16_t)0xFFFF;
                            s.pairCodes[tb.codeBase + (size_t)t * 32 + h * 16 + l] = code != 0xFFFF ? code : pad;
                        }
                    }
                }
                s.paddedSteps += steps;
            }
            if (lim.laplacian) {                 // codes -> shared-memory slots of the dot products
                const int PS = stride - 4;
                for (uint16_t &code : s.pairCodes) code = (uint16_t)lap_pair_slot ((code >> 2) & 3, code & 3, code >> 4, PS);
                for (uint16_t &code : s.diagCodes) code = (uint16_t)lap_diag_slot (code & 3, code >> 2, PS);
            }
            uint32_t bytes = (uint32_t)sizeof (TileBlobHeader) + (uint32_t)(s.rows.size () * sizeof (TileRow));
            bytes = align16 (bytes) + align16 ((uint32_t)(s.nodes.size () * 4));
            bytes += align16 ((uint32_t)(s.elemNodes.size () * 2));
            s.headBytes = bytes;
            bytes += align16 ((uint32_t)s.entryRow.size ());
            bytes += align16 ((uint32_t)(s.laneEntry.size () * 2));
            bytes += align16 ((uint32_t)(s.batches.size () * sizeof (TileBatch)));
            bytes += align16 ((uint32_t)(s.diagCodes.size () * 2)) + align16 ((uint32_t)(s.pairCodes.size () * 2));
            s.blobBytes = bytes;
            for (int n : s.nodes) nodeLocal[n] = -1;
            for (int e : s.elems) elemLocal[e] = -1;
        }
    }

    // ---- 4. serialise: interface tiles first ----------------------------------------
    std::vector<int> execOrder;
    for (int t = 0; t < nbTiles; t++) if (scratch[t].hasInterface) execOrder.push_back (t);
    plan.nbInterfaceTiles = (int)execOrder.size ();
    for (int t = 0; t < nbTiles; t++) if (!scratch[t].hasInterface) execOrder.push_back (t);
    plan.nbTiles = nbTiles;
    plan.tileOffset.assign ((size_t)nbTiles + 1, 0);
    for (int k = 0; k < nbTiles; k++) {
        const TileScratch &s = scratch[execOrder[k]];
        if (s.bad) { error = "CSR lacks a node pair of an element (tile " + std::to_string (execOrder[k]) + ")"; return -1; }
        plan.tileOffset[k + 1] = plan.tileOffset[k] + s.blobBytes;
        plan.maxBlobBytes = std::max (plan.maxBlobBytes, s.blobBytes);
        plan.maxHeadBytes = std::max (plan.maxHeadBytes, s.headBytes);
        plan.maxTailBytes = std::max (plan.maxTailBytes, s.blobBytes - s.headBytes);
        plan.maxRows = std::max (plan.maxRows, (int)s.rows.size () - 1);
        plan.maxElems = std::max (plan.maxElems, s.nbIds);
        plan.maxNodesRef = std::max (plan.maxNodesRef, (int)s.nodes.size ());
        plan.maxEntries = std::max (plan.maxEntries, (int)s.entryRow.size ());
        plan.nbTileElems += (int64_t)s.elems.size ();
        plan.nbContributions += s.contributions;
        plan.nbPaddedSteps += s.paddedSteps;
    }
    plan.blob.assign ((size_t)plan.tileOffset[nbTiles], 0);
    #pragma omp parallel for schedule(dynamic, 64) num_threads(plan_team_size ())
    for (int k = 0; k < nbTiles; k++) {
        const TileScratch &s = scratch[execOrder[k]];
        uint8_t *base = plan.blob.data () + plan.tileOffset[k];
        TileBlobHeader h;
        memset (&h, 0, sizeof h);
        h.nbRows = (uint16_t)(s.rows.size () - 1); h.nbNodesRef = (uint16_t)s.nodes.size ();
        h.nbElems = (uint16_t)s.nbIds; h.nbEntries = (uint16_t)s.entryRow.size ();
        h.nbBatches = (uint16_t)s.batches.size (); h.hasInterface = s.hasInterface ? 1 : 0;
        uint32_t at = (uint32_t)sizeof (TileBlobHeader);
        memcpy (base + at, s.rows.data (), s.rows.size () * sizeof (TileRow));
        at = align16 (at + (uint32_t)(s.rows.size () * sizeof (TileRow)));
        h.offNodes = at; memcpy (base + at, s.nodes.data (), s.nodes.size () * 4);
        at += align16 ((uint32_t)(s.nodes.size () * 4));
        h.offElems = at; memcpy (base + at, s.elemNodes.data (), s.elemNodes.size () * 2);
        at += align16 ((uint32_t)(s.elemNodes.size () * 2));
        h.offEntryRow = at; memcpy (base + at, s.entryRow.data (), s.entryRow.size ());
        at += align16 ((uint32_t)s.entryRow.size ());
        h.offLaneEntry = at; memcpy (base + at, s.laneEntry.data (), s.laneEntry.size () * 2);
        at += align16 ((uint32_t)(s.laneEntry.size () * 2));
        h.offBatches = at; memcpy (base + at, s.batches.data (), s.batches.size () * sizeof (TileBatch));
        at += align16 ((uint32_t)(s.batches.size () * sizeof (TileBatch)));
        h.offDiag = at; memcpy (base + at, s.diagCodes.data (), s.diagCodes.size () * 2);
        at += align16 ((uint32_t)(s.diagCodes.size () * 2));
        h.offPair = at; memcpy (base + at, s.pairCodes.data (), s.pairCodes.size () * 2);
        at += align16 ((uint32_t)(s.pairCodes.size () * 2));
        h.blobBytes = at;
        memcpy (base, &h, sizeof h);
    }
    return 0;
}

int verify_tile_plan (const TilePlan &plan, int nbNodes, int nbElem, const int *elemToNode,
                      const int *row, const int *col, std::string &error)
{
    std::vector<uint8_t> rowSeen ((size_t)nbNodes, 0);
    std::vector<uint16_t> tripleSeen ((size_t)nbElem, 0);      // bit 4j+k per element
    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());
    auto mark = [&] (int e, int j, int k) -> bool {
        const uint16_t bit = (uint16_t)(1u << (4 * j + k));
        if (tripleSeen[e] & bit) return false;
        tripleSeen[e] |= bit;
        return true;
    };
    int64_t rowsTotal = 0;
    bool interiorSeen = false;
    for (int t = 0; t < plan.nbTiles; t++) {
        const uint8_t *base = plan.blob.data () + plan.tileOffset[t];
        const TileBlobHeader &h = *plan.header (t);
        if (h.blobBytes != plan.tileOffset[t + 1] - plan.tileOffset[t] || (plan.tileOffset[t] & 15)) {
            error = "blob size / alignment"; return -1;
        }
        if (h.hasInterface && interiorSeen) { error = "interface tiles must come first"; return -1; }
        if (!h.hasInterface) interiorSeen = true;
        if ((t < plan.nbInterfaceTiles) != (h.hasInterface != 0)) { error = "nbInterfaceTiles mismatch"; return -1; }
        const TileRow *rows = reinterpret_cast<const TileRow*> (base + sizeof (TileBlobHeader));
        const int *tileNodes = reinterpret_cast<const int*> (base + h.offNodes);
        const uint16_t *tileElems = reinterpret_cast<const uint16_t*> (base + h.offElems);
        const uint8_t *entryRow = base + h.offEntryRow;
        const uint16_t *laneEntry = reinterpret_cast<const uint16_t*> (base + h.offLaneEntry);
        const TileBatch *batches = reinterpret_cast<const TileBatch*> (base + h.offBatches);
        const uint16_t *diagCodes = reinterpret_cast<const uint16_t*> (base + h.offDiag);
        const uint16_t *pairCodes = reinterpret_cast<const uint16_t*> (base + h.offPair);
        if (h.nbElems > plan.maxElems || h.nbElems >= plan.elemStride || h.nbRows > plan.maxRows ||
            h.nbNodesRef > plan.maxNodesRef) {
            error = "tile exceeds the plan maxima"; return -1;
        }
        // tile-local element -> global element, by matching its node quadruple
        std::vector<int> globalElem (h.nbElems, -1);
        for (int el = 0; el < h.nbElems; el++) {
            const uint16_t *ln = tileElems + (size_t)el * 4;
            if (ln[0] == 0xFFFF) continue;                       // hole of the coset numbering
            int g[4];
            for (int k = 0; k < 4; k++) {
                if (ln[k] >= h.nbNodesRef) { error = "local node index out of range"; return -1; }
                g[k] = tileNodes[ln[k]] + 1;
            }
            for (int p = n2eIndex[g[0] - 1]; p < n2eIndex[g[0]]; p++) {
                const int *cand = elemToNode + (size_t)n2eValue[p] * 4;
                if (cand[0] == g[0] && cand[1] == g[1] && cand[2] == g[2] && cand[3] == g[3] &&
                    std::find (globalElem.begin (), globalElem.end (), n2eValue[p]) == globalElem.end ()) {
                    globalElem[el] = n2eValue[p];
                    break;
                }
            }
            if (globalElem[el] < 0) { error = "tile element matches no mesh element"; return -1; }
        }
        int expectStart = 0;
        for (int r = 0; r < h.nbRows; r++) {
            const TileRow &tr = rows[r];
            const int n = tr.node & 0x7fffffff;
            if (n < 0 || n >= nbNodes || rowSeen[n]) { error = "row owned twice or out of range"; return -1; }
            rowSeen[n] = 1;
            rowsTotal++;
            if (tileNodes[r] != n) { error = "owned rows must lead the tile's node list"; return -1; }
            if (tr.valueStart != row[n] || tr.localStart != expectStart) { error = "row offsets differ from nodeToNodeRow"; return -1; }
            const int len = row[n + 1] - row[n];
            expectStart += len;
            for (int l = 0; l < len; l++) {
                if (entryRow[tr.localStart + l] != r) { error = "entryRow mismatch"; return -1; }
            }
            if (tr.diagLocal != 0xFFFF && col[tr.valueStart + (tr.diagLocal - tr.localStart)] != n + 1) {
                error = "diagLocal is not the diagonal entry"; return -1;
            }
            for (int k = tr.diagCodeBase; k < rows[r + 1].diagCodeBase; k++) {
                int code = diagCodes[k];
                if (plan.laplacian) {                                // slot -> (element << 2 | a)
                    const int PS = plan.elemStride - 4, a = code / PS - 6;
                    if (a < 0 || a > 3) { error = "diagonal slot outside the diagonal planes"; return -1; }
                    code = ((code - (6 + a) * PS - 4 * a) << 2) | a;
                }
                const int el = code >> 2, a = code & 3;
                if (el >= h.nbElems || globalElem[el] < 0) { error = "diagonal code names a foreign element"; return -1; }
                if (tileNodes[tileElems[(size_t)el * 4 + a]] != n) { error = "diagonal code: wrong local node"; return -1; }
                if (tr.diagLocal == 0xFFFF) { error = "diagonal contribution on a row without diagonal entry"; return -1; }
                if (!mark (globalElem[el], a, a)) { error = "diagonal contribution listed twice"; return -1; }
            }
        }
        if (expectStart != h.nbEntries) { error = "entry count mismatch"; return -1; }
        std::vector<uint8_t> entrySeen (h.nbEntries, 0);
        for (int b = 0; b < h.nbBatches; b++) {
            const TileBatch &tb = batches[b];
            for (int lane = 0; lane < 32; lane++) {
                const int q = laneEntry[b * 32 + lane];
                if (q != 0xFFFF) {
                    if (q >= h.nbEntries || entrySeen[q]) { error = "entry mapped to two lanes"; return -1; }
                    entrySeen[q] = 1;
                    // a half-warp holds consecutive entries of ONE row
                    if ((lane & 15) && laneEntry[b * 32 + lane - 1] != 0xFFFF &&
                        (laneEntry[b * 32 + lane - 1] != q - 1 || entryRow[q - 1] != entryRow[q])) {
                        error = "half-warp lanes are not consecutive entries of one row"; return -1;
                    }
                }
                for (int s = 0; s < tb.steps; s++) {
                    int code = pairCodes[(size_t)tb.codeBase + (size_t)s * 32 + lane];
                    if (plan.laplacian) {                            // slot -> (element << 4 | a << 2 | b)
                        static const int kPair[6][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}, {0, 3}, {1, 2}};
                        const int PS = plan.elemStride - 4, pl = code / PS;
                        if (pl > 5) { error = "pair slot outside the pair planes"; return -1; }
                        const int el = code - pl * PS - 4 * (pl / 2);
                        int a = kPair[pl][0], b = kPair[pl][1];
                        if (el >= h.nbElems) { a = 1; b = 0; }       // padding slots live in plane 0 = pair (1,0)
                        else if (q != 0xFFFF && globalElem[el] >= 0 &&
                                 tileNodes[tileElems[(size_t)el * 4 + a]] != (rows[entryRow[q]].node & 0x7fffffff)) std::swap (a, b);
                        code = (el << 4) | (a << 2) | b;
                    }
                    const int el = code >> 4, a = (code >> 2) & 3, bb = code & 3;
                    if (el >= h.nbElems) {                           // padding: one of the 16 zero slots
                        // local nodes must be (1, 0): the kernels zero exactly those slots (and the
                        // Laplacian's pair plane of (1, 0))
                        if (el >= h.nbElems + 16 || el >= plan.elemStride - 4 || a != 1 || bb != 0) { error = "bad padding code"; return -1; }
                        continue;
                    }
                    if (el >= h.nbElems || q == 0xFFFF || globalElem[el] < 0) { error = "pair code out of range"; return -1; }
                    const TileRow &tr = rows[entryRow[q]];
                    const int n = tr.node & 0x7fffffff;
                    const uint16_t *ln = tileElems + (size_t)el * 4;
                    const int l = tr.valueStart + (q - tr.localStart);
                    if (tileNodes[ln[a]] != n || col[l] != tileNodes[ln[bb]] + 1) { error = "pair code lands on the wrong CSR entry"; return -1; }
                    if (!mark (globalElem[el], a, bb)) { error = "pair contribution listed twice"; return -1; }
                }
            }
        }
        for (int q = 0; q < h.nbEntries; q++) if (!entrySeen[q]) { error = "entry without a lane"; return -1; }
    }
    if (rowsTotal != nbNodes) { error = "not every node is owned by a tile"; return -1; }
    for (int e = 0; e < nbElem; e++) {
        if (tripleSeen[e] != 0xFFFF) { error = "element " + std::to_string (e) + " misses a contribution"; return -1; }
    }
    return 0;
}

}  // namespace mfb
