#include "tile_plan.h"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "mesh_topology.h"

namespace mfb {

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

struct TileScratch {
    std::vector<int> elems;                   // global element ids, first-touch order
    std::vector<int> nodes;                   // global node ids, owned first
    std::vector<uint16_t> elemNodes;          // 4 local ids per element
    std::vector<TileRow> rows;
    std::vector<uint8_t> entryRow;
    std::vector<TileBatch> batches;
    std::vector<uint16_t> pairCodes, diagCodes;
    int64_t contributions = 0;
    bool bad = false;
};

}  // namespace

int64_t TilePlan::bytes () const
{
    return (int64_t)tiles.size () * sizeof (TileHeader) + tileNodes.size () * sizeof (int) +
           tileElems.size () * sizeof (uint16_t) + rows.size () * sizeof (TileRow) +
           entryRow.size () + batches.size () * sizeof (TileBatch) +
           pairCodes.size () * sizeof (uint16_t) + diagCodes.size () * sizeof (uint16_t);
}

int build_tile_plan (int nbNodes, int nbElem, const int *elemToNode, const int *row,
                     const int *col, const double *coord, const uint8_t *isInterface,
                     const TilePlanLimits &lim, TilePlan &plan, std::string &error)
{
    plan = TilePlan ();
    if (lim.maxRows < 1 || lim.maxRows > 255 || lim.maxElems < 1 || lim.maxElems > 4094 ||
        lim.maxNodesRef < 4 || lim.maxNodesRef > 65535 || lim.maxEntries < 1 || lim.maxEntries > 65535) {
        error = "tile plan limits out of range";
        return -1;
    }
    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());

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
    #pragma omp parallel for schedule(static)
    for (int n = 0; n < nbNodes; n++) {
        uint64_t q[3];
        for (int a = 0; a < 3; a++) q[a] = (uint64_t)((coord[(size_t)n * 3 + a] - lo[a]) * scale);
        order[n] = { spread21 (q[0]) | (spread21 (q[1]) << 1) | (spread21 (q[2]) << 2), n };
    }
    std::sort (order.begin (), order.end ());

    // ---- 2. greedy cut under the shared-memory caps -------------------------------
    std::vector<int> tileStart;          // offsets into `order`
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
    const int nbTiles = (int)tileStart.size () - 1;

    // ---- 3. per-tile tables (independent) ------------------------------------------
    std::vector<TileScratch> scratch ((size_t)nbTiles);
    // each thread keeps two dense global->local maps; cap the team so that the maps of a
    // 48 M-element mesh stay within a few GB on a many-core host
    const int team = std::max (1, std::min (omp_get_max_threads (), 16));
    #pragma omp parallel num_threads(team)
    {
        std::vector<int> nodeLocal ((size_t)nbNodes, -1), elemLocal ((size_t)nbElem, -1);
        std::vector<std::vector<uint16_t>> lists;
        #pragma omp for schedule(dynamic, 16)
        for (int t = 0; t < nbTiles; t++) {
            TileScratch &s = scratch[t];
            const int first = tileStart[t], nbRows = tileStart[t + 1] - first;
            for (int r = 0; r < nbRows; r++) {                       // owned rows first
                const int n = order[first + r].second;
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
            if ((int)lists.size () < nbEntries) lists.resize (nbEntries);
            for (int q = 0; q < nbEntries; q++) lists[q].clear ();
            s.entryRow.resize (nbEntries);

            int localStart = 0;
            for (int r = 0; r < nbRows; r++) {
                const int n = s.nodes[r], begin = row[n], end = row[n + 1];
                TileRow tr;
                tr.node = n | ((isInterface && isInterface[n]) ? (int)0x80000000u : 0);
                tr.valueStart = begin;
                tr.diagCodeBase = (int)s.diagCodes.size ();
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
                            s.diagCodes.push_back ((uint16_t)((el << 2) | a));
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
            s.contributions = (int64_t)s.diagCodes.size ();
            const int nbBatches = (nbEntries + 31) / 32;
            for (int b = 0; b < nbBatches; b++) {
                int steps = 0;
                for (int lane = 0; lane < 32 && b * 32 + lane < nbEntries; lane++) {
                    steps = std::max (steps, (int)lists[b * 32 + lane].size ());
                }
                TileBatch tb = { (int)s.pairCodes.size (), steps };
                s.batches.push_back (tb);
                s.pairCodes.resize (s.pairCodes.size () + (size_t)steps * 32, (uint16_t)(s.elems.size () << 4));
                for (int lane = 0; lane < 32 && b * 32 + lane < nbEntries; lane++) {
                    const std::vector<uint16_t> &l = lists[b * 32 + lane];
                    for (size_t k = 0; k < l.size (); k++) s.pairCodes[tb.codeBase + k * 32 + lane] = l[k];
                    s.contributions += (int64_t)l.size ();
                }
            }
            for (int n : s.nodes) nodeLocal[n] = -1;
            for (int e : s.elems) elemLocal[e] = -1;
        }
    }

    // ---- 4. concatenate ---------------------------------------------------------
    plan.nbTiles = nbTiles;
    plan.tiles.resize ((size_t)nbTiles);
    size_t nNodes = 0, nElems = 0, nRows = 0, nEntries = 0, nBatches = 0, nPair = 0, nDiag = 0;
    for (int t = 0; t < nbTiles; t++) {
        const TileScratch &s = scratch[t];
        if (s.bad) { error = "CSR lacks a node pair of an element (tile " + std::to_string (t) + ")"; return -1; }
        TileHeader &h = plan.tiles[t];
        h.nodeBase = (int)nNodes; h.elemBase = (int)nElems; h.rowBase = (int)nRows;
        h.entryBase = (int)nEntries; h.batchBase = (int)nBatches;
        h.nbRows = (uint16_t)s.rows.size (); h.nbNodesRef = (uint16_t)s.nodes.size ();
        h.nbElems = (uint16_t)s.elems.size (); h.nbEntries = (uint16_t)s.entryRow.size ();
        h.pad = 0;
        nNodes += s.nodes.size (); nElems += s.elems.size (); nRows += s.rows.size ();
        nEntries += s.entryRow.size (); nBatches += s.batches.size ();
        nPair += s.pairCodes.size (); nDiag += s.diagCodes.size ();
        plan.maxRows = std::max (plan.maxRows, (int)h.nbRows);
        plan.maxElems = std::max (plan.maxElems, (int)h.nbElems);
        plan.maxNodesRef = std::max (plan.maxNodesRef, (int)h.nbNodesRef);
        plan.maxEntries = std::max (plan.maxEntries, (int)h.nbEntries);
        plan.nbContributions += s.contributions;
    }
    if (nPair > (size_t)INT32_MAX || nDiag > (size_t)INT32_MAX || nElems > (size_t)INT32_MAX) {
        error = "tile plan exceeds 32-bit offsets";
        return -1;
    }
    plan.nbTileElems = (int64_t)nElems;
    plan.tileNodes.resize (nNodes); plan.tileElems.resize (nElems * 4); plan.rows.resize (nRows + 1);
    plan.entryRow.resize (nEntries); plan.batches.resize (nBatches);
    plan.pairCodes.resize (nPair); plan.diagCodes.resize (nDiag);
    std::vector<size_t> pairBase ((size_t)nbTiles + 1, 0), diagBase ((size_t)nbTiles + 1, 0);
    for (int t = 0; t < nbTiles; t++) {
        pairBase[t + 1] = pairBase[t] + scratch[t].pairCodes.size ();
        diagBase[t + 1] = diagBase[t] + scratch[t].diagCodes.size ();
    }
    #pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < nbTiles; t++) {
        const TileScratch &s = scratch[t];
        const TileHeader &h = plan.tiles[t];
        std::copy (s.nodes.begin (), s.nodes.end (), plan.tileNodes.begin () + h.nodeBase);
        std::copy (s.elemNodes.begin (), s.elemNodes.end (), plan.tileElems.begin () + (size_t)h.elemBase * 4);
        std::copy (s.entryRow.begin (), s.entryRow.end (), plan.entryRow.begin () + h.entryBase);
        std::copy (s.pairCodes.begin (), s.pairCodes.end (), plan.pairCodes.begin () + pairBase[t]);
        std::copy (s.diagCodes.begin (), s.diagCodes.end (), plan.diagCodes.begin () + diagBase[t]);
        for (size_t r = 0; r < s.rows.size (); r++) {
            TileRow tr = s.rows[r];
            tr.diagCodeBase += (int)diagBase[t];
            plan.rows[h.rowBase + r] = tr;
        }
        for (size_t b = 0; b < s.batches.size (); b++) {
            TileBatch tb = s.batches[b];
            tb.codeBase += (int)pairBase[t];
            plan.batches[h.batchBase + b] = tb;
        }
    }
    TileRow sentinel = {0, 0, (int)nDiag, 0, 0xFFFF};
    plan.rows[nRows] = sentinel;

    // tiles that own interface nodes first: their diagonal blocks feed the halo exchange
    if (isInterface) {
        auto owns = [&] (const TileHeader &h) {
            for (int r = 0; r < h.nbRows; r++) if (plan.rows[h.rowBase + r].node < 0) return true;
            return false;
        };
        auto mid = std::stable_partition (plan.tiles.begin (), plan.tiles.end (), owns);
        plan.nbInterfaceTiles = (int)(mid - plan.tiles.begin ());
    }
    return 0;
}

}  // namespace mfb

namespace mfb {

int verify_tile_plan (const TilePlan &plan, int nbNodes, int nbElem, const int *elemToNode,
                      const int *row, const int *col, std::string &error)
{
    std::vector<uint8_t> rowSeen ((size_t)nbNodes, 0);
    std::vector<uint16_t> tripleSeen ((size_t)nbElem, 0);      // bit 4j+k per element
    int64_t rowsTotal = 0;
    for (int t = 0; t < plan.nbTiles; t++) {
        const TileHeader &h = plan.tiles[t];
        // global element of each tile-local element, recovered from its 4 node ids
        for (int r = 0; r < h.nbRows; r++) {
            const TileRow &tr = plan.rows[h.rowBase + r];
            const int n = tr.node & 0x7fffffff;
            if (n < 0 || n >= nbNodes || rowSeen[n]) { error = "row owned twice or out of range"; return -1; }
            rowSeen[n] = 1;
            rowsTotal++;
            if (plan.tileNodes[h.nodeBase + r] != n) { error = "owned rows must lead the tile's node list"; return -1; }
            if (tr.valueStart != row[n]) { error = "row offset differs from nodeToNodeRow"; return -1; }
            const int len = row[n + 1] - row[n];
            for (int l = 0; l < len; l++) {
                if (plan.entryRow[h.entryBase + tr.localStart + l] != r) { error = "entryRow mismatch"; return -1; }
            }
            // diagonal codes: one per incident element, local index a must be this node
            const int dEnd = plan.rows[h.rowBase + r + 1].diagCodeBase;
            for (int k = tr.diagCodeBase; k < dEnd; k++) {
                const int code = plan.diagCodes[k], el = code >> 2, a = code & 3;
                if (el >= h.nbElems) { error = "diagonal code names a foreign element"; return -1; }
                const uint16_t *ln = &plan.tileElems[((size_t)h.elemBase + el) * 4];
                if (plan.tileNodes[h.nodeBase + ln[a]] != n) { error = "diagonal code: wrong local node"; return -1; }
                if (tr.diagLocal == 0xFFFF) { error = "diagonal contribution on a row without diagonal entry"; return -1; }
            }
        }
        // off-diagonal codes, batch by batch
        const int nbBatches = (h.nbEntries + 31) / 32;
        for (int b = 0; b < nbBatches; b++) {
            const TileBatch &tb = plan.batches[h.batchBase + b];
            for (int lane = 0; lane < 32; lane++) {
                const int q = b * 32 + lane;
                for (int s = 0; s < tb.steps; s++) {
                    const int code = plan.pairCodes[(size_t)tb.codeBase + (size_t)s * 32 + lane];
                    const int el = code >> 4, a = (code >> 2) & 3, bb = code & 3;
                    if (el == h.nbElems) { if (a || bb) { error = "bad padding code"; return -1; } continue; }
                    if (el > h.nbElems || q >= h.nbEntries) { error = "pair code out of range"; return -1; }
                    const int r = plan.entryRow[h.entryBase + q];
                    const TileRow &tr = plan.rows[h.rowBase + r];
                    const int n = tr.node & 0x7fffffff;
                    const uint16_t *ln = &plan.tileElems[((size_t)h.elemBase + el) * 4];
                    const int na = plan.tileNodes[h.nodeBase + ln[a]], nb = plan.tileNodes[h.nodeBase + ln[bb]];
                    const int l = tr.valueStart + (q - tr.localStart);
                    if (na != n || col[l] != nb + 1) { error = "pair code lands on the wrong CSR entry"; return -1; }
                }
            }
        }
    }
    if (rowsTotal != nbNodes) { error = "not every node is owned by a tile"; return -1; }

    // every (element, j, k) exactly once: count per tile through the element identity
    // (tile-local element -> global element by matching its node quadruple among the
    // elements incident to its first node)
    std::vector<int> n2eIndex ((size_t)nbNodes + 1), n2eValue ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, n2eIndex.data (), n2eValue.data ());
    auto mark = [&] (int e, int j, int k) -> bool {
        const uint16_t bit = (uint16_t)(1u << (4 * j + k));
        if (tripleSeen[e] & bit) return false;
        tripleSeen[e] |= bit;
        return true;
    };
    for (int t = 0; t < plan.nbTiles; t++) {
        const TileHeader &h = plan.tiles[t];
        std::vector<int> globalElem (h.nbElems, -1);
        for (int el = 0; el < h.nbElems; el++) {
            const uint16_t *ln = &plan.tileElems[((size_t)h.elemBase + el) * 4];
            int g[4];
            for (int k = 0; k < 4; k++) g[k] = plan.tileNodes[h.nodeBase + ln[k]] + 1;
            for (int p = n2eIndex[g[0] - 1]; p < n2eIndex[g[0]]; p++) {
                const int *cand = elemToNode + (size_t)n2eValue[p] * 4;
                if (cand[0] == g[0] && cand[1] == g[1] && cand[2] == g[2] && cand[3] == g[3]) {
                    // duplicates of one quadruple are interchangeable; take the first not yet used by this tile
                    if (std::find (globalElem.begin (), globalElem.end (), n2eValue[p]) == globalElem.end ()) {
                        globalElem[el] = n2eValue[p];
                        break;
                    }
                }
            }
            if (globalElem[el] < 0) { error = "tile element matches no mesh element"; return -1; }
        }
        for (int r = 0; r < h.nbRows; r++) {
            const TileRow &tr = plan.rows[h.rowBase + r];
            const int dEnd = plan.rows[h.rowBase + r + 1].diagCodeBase;
            for (int k = tr.diagCodeBase; k < dEnd; k++) {
                const int code = plan.diagCodes[k];
                if (!mark (globalElem[code >> 2], code & 3, code & 3)) { error = "diagonal contribution listed twice"; return -1; }
            }
        }
        const int nbBatches = (h.nbEntries + 31) / 32;
        for (int b = 0; b < nbBatches; b++) {
            const TileBatch &tb = plan.batches[h.batchBase + b];
            for (int s = 0; s < tb.steps; s++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int code = plan.pairCodes[(size_t)tb.codeBase + (size_t)s * 32 + lane];
                    if ((code >> 4) == h.nbElems) continue;
                    if (!mark (globalElem[code >> 4], (code >> 2) & 3, code & 3)) { error = "pair contribution listed twice"; return -1; }
                }
            }
        }
    }
    for (int e = 0; e < nbElem; e++) {
        if (tripleSeen[e] != 0xFFFF) { error = "element " + std::to_string (e) + " misses a contribution"; return -1; }
    }
    return 0;
}

}  // namespace mfb
