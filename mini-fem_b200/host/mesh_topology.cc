#include "mesh_topology.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mfb {

void node_to_elem (const int *elemToNode, int nbElem, int nbNodes, int *index, int *value)
{
    std::vector<int> cursor ((size_t)nbNodes + 1, 0);
    const int64_t nbInc = (int64_t)nbElem * kDimElem;
    for (int64_t k = 0; k < nbInc; k++) cursor[elemToNode[k]]++;
    index[0] = 0;
    for (int n = 0; n < nbNodes; n++) index[n + 1] = index[n] + cursor[n + 1];
    for (int n = 0; n < nbNodes; n++) cursor[n] = index[n];
    for (int e = 0; e < nbElem; e++) {
        const int *nodes = elemToNode + (size_t)e * kDimElem;
        for (int k = 0; k < kDimElem; k++) value[cursor[nodes[k] - 1]++] = e;
    }
}

// One sweep over the nodes; `emit(i, col)` sees the columns of row i in first-seen order.
template <class Emit>
static int64_t sweep_rows (const int *elemToNode, int nbElem, int nbNodes, int *row, Emit emit)
{
    std::vector<int> index ((size_t)nbNodes + 1), value ((size_t)nbElem * kDimElem);
    node_to_elem (elemToNode, nbElem, nbNodes, index.data (), value.data ());
    std::vector<int> stamp ((size_t)nbNodes, -1);
    int64_t total = 0;
    for (int i = 0; i < nbNodes; i++) {
        if (row) row[i] = (int)total;
        for (int p = index[i]; p < index[i + 1]; p++) {
            const int *nodes = elemToNode + (size_t)value[p] * kDimElem;
            for (int k = 0; k < kDimElem; k++) {
                int cand = nodes[k];
                if (stamp[cand - 1] != i) {
                    stamp[cand - 1] = i;
                    emit (total, cand);
                    total++;
                }
            }
        }
    }
    if (row) row[nbNodes] = (int)total;
    return total;
}

int64_t count_csr_entries (const int *elemToNode, int nbElem, int nbNodes)
{
    return sweep_rows (elemToNode, nbElem, nbNodes, nullptr, [] (int64_t, int) {});
}

int64_t build_csr (const int *elemToNode, int nbElem, int nbNodes, int *row, int *col)
{
    return sweep_rows (elemToNode, nbElem, nbNodes, row,
                       [col] (int64_t at, int id) { col[at] = id; });
}

int build_elem_to_edge (const int *row, const int *col, const int *elemToNode,
                        int *elemToEdge, int nbElem)
{
    int missing = 0;
    #pragma omp parallel for schedule(static) reduction(+ : missing)
    for (int e = 0; e < nbElem; e++) {
        const int *nodes = elemToNode + (size_t)e * kDimElem;
        int *out = elemToEdge + (size_t)e * kValuesPerElem;
        for (int j = 0; j < kDimElem; j++) {
            const int begin = row[nodes[j] - 1], end = row[nodes[j]];
            for (int k = 0; k < kDimElem; k++) {
                int found = -1;
                for (int l = begin; l < end; l++) {
                    if (col[l] == nodes[k]) { found = l; break; }
                }
                if (found < 0) missing++;
                out[4 * j + k] = found;
            }
        }
    }
    return missing ? -1 : 0;
}

int color_elements (const int *elemToNode, int nbElem, int nbNodes, int *colorPart,
                    int *colorToElem, int *colorPerm)
{
    // usedAt[n] = colours already taken by the coloured elements around node n.  The
    // union over an element's 4 nodes is exactly the union over its neighbour elements.
    std::vector<unsigned __int128> usedAt ((size_t)nbNodes, 0);
    int highest = 0;
    for (int e = 0; e < nbElem; e++) {
        const int *nodes = elemToNode + (size_t)e * kDimElem;
        unsigned __int128 taken = usedAt[nodes[0] - 1] | usedAt[nodes[1] - 1]
                                | usedAt[nodes[2] - 1] | usedAt[nodes[3] - 1];
        uint64_t lo = ~(uint64_t)taken, hi = ~(uint64_t)(taken >> 64);
        int color;
        if (lo)      color = __builtin_ctzll (lo);
        else if (hi) color = 64 + __builtin_ctzll (hi);
        else return -1;
        unsigned __int128 bit = (unsigned __int128)1 << color;
        for (int k = 0; k < kDimElem; k++) usedAt[nodes[k] - 1] |= bit;
        colorPart[e] = color;
        if (color > highest) highest = color;
    }
    const int nbColors = highest + 1;

    std::vector<int> slot (kMaxColor + 1, 0);
    for (int e = 0; e < nbElem; e++) slot[colorPart[e] + 1]++;
    for (int c = 0; c < kMaxColor; c++) slot[c + 1] += slot[c];
    for (int c = 0; c <= nbColors; c++) colorToElem[c] = slot[c];
    for (int e = 0; e < nbElem; e++) colorPerm[e] = slot[colorPart[e]]++;
    return nbColors;
}

void permute_rows (int *tab, const int *perm, int nbItem, int dim)
{
    std::vector<int> old (tab, tab + (size_t)nbItem * dim);
    for (int i = 0; i < nbItem; i++) {
        memcpy (tab + (size_t)perm[i] * dim, old.data () + (size_t)i * dim, sizeof (int) * dim);
    }
}

int boundary_mask (const int *boundNodesCode, int nbNodes, int *checkBounds)
{
    memset (checkBounds, 0, sizeof (int) * (size_t)nbNodes * kDimNode);
    int *mx = checkBounds, *my = checkBounds + nbNodes, *mz = checkBounds + 2 * (size_t)nbNodes;
    int nbBound = 0;
    for (int n = 0; n < nbNodes; n++) {
        switch (boundNodesCode[n]) {
            case 0:  break;
            case 52: mx[n] = 1; nbBound++; break;
            case 53: my[n] = 1; nbBound++; break;
            case 54: mz[n] = 1; nbBound++; break;
            default: mx[n] = my[n] = mz[n] = 1; nbBound++; break;
        }
    }
    return nbBound;
}

double double_norm (const double *tab, int64_t size)
{
    double acc = 0;
    for (int64_t i = 0; i < size; i++) acc += tab[i] * tab[i];
    return std::sqrt (acc);
}

namespace {

// recursive bisection of element centroids into leaves of <= leafSize elements (in `idx` order)
void bisect_elements (std::vector<int> &idx, int lo, int hi, int leafSize, const std::vector<float> &centroid, std::vector<int> &leafStart)
{
    const int count = hi - lo;
    if (count <= leafSize) { leafStart.push_back (lo); return; }
    float bmin[3], bmax[3];
    for (int a = 0; a < 3; a++) bmin[a] = bmax[a] = centroid[(size_t)idx[lo] * 3 + a];
    for (int q = lo + 1; q < hi; q++) {
        for (int a = 0; a < 3; a++) {
            const float v = centroid[(size_t)idx[q] * 3 + a];
            bmin[a] = std::min (bmin[a], v); bmax[a] = std::max (bmax[a], v);
        }
    }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (bmax[a] - bmin[a] > bmax[axis] - bmin[axis]) axis = a;
    const int leaves = (count + leafSize - 1) / leafSize;
    const int mid = lo + (int)((int64_t)count * (leaves / 2) / leaves);
    std::nth_element (idx.begin () + lo, idx.begin () + mid, idx.begin () + hi, [&] (int x, int y) {
        const float cx = centroid[(size_t)x * 3 + axis], cy = centroid[(size_t)y * 3 + axis];
        return cx < cy || (cx == cy && x < y);
    });
    bisect_elements (idx, lo, mid, leafSize, centroid, leafStart);
    bisect_elements (idx, mid, hi, leafSize, centroid, leafStart);
}

}  // namespace

int build_block_coloring (const int *elemToNode, int nbElem, int nbNodes, const double *coord, int blockElems,
                          BlockColoring &out)
{
    out = BlockColoring ();
    out.launchStart.assign (1, 0);
    out.localIndex.assign (1, 0);
    out.localStart.assign (1, 0);
    if (nbElem <= 0) return 0;
    blockElems = std::max (blockElems, 1);
    // ---- blocks: leaves of the bisection of the centroids ------------------------------------------------------
    std::vector<float> centroid ((size_t)nbElem * 3);
    for (int e = 0; e < nbElem; e++) {
        for (int a = 0; a < 3; a++) {
            double sum = 0.0;
            for (int k = 0; k < kDimElem; k++) sum += coord[(size_t)(elemToNode[(size_t)e * kDimElem + k] - 1) * 3 + a];
            centroid[(size_t)e * 3 + a] = (float)(0.25 * sum);
        }
    }
    std::vector<int> idx ((size_t)nbElem), leafStart;
    for (int e = 0; e < nbElem; e++) idx[e] = e;
    bisect_elements (idx, 0, nbElem, blockElems, centroid, leafStart);
    leafStart.push_back (nbElem);
    const int nbBlocks = (int)leafStart.size () - 1;
    // ---- block colours: first colour none of the block's nodes has seen (64-bit masks per node) ------------------
    std::vector<uint64_t> nodeBlockMask ((size_t)nbNodes, 0);
    std::vector<int> blockColor ((size_t)nbBlocks, 0);
    int nbBlockColors = 0;
    for (int b = 0; b < nbBlocks; b++) {
        uint64_t forbidden = 0;
        for (int q = leafStart[b]; q < leafStart[b + 1]; q++) {
            const int *en = elemToNode + (size_t)idx[q] * kDimElem;
            for (int k = 0; k < kDimElem; k++) forbidden |= nodeBlockMask[en[k] - 1];
        }
        if (~forbidden == 0) return -2;
        const int color = __builtin_ctzll (~forbidden);
        blockColor[b] = color;
        nbBlockColors = std::max (nbBlockColors, color + 1);
        for (int q = leafStart[b]; q < leafStart[b + 1]; q++) {
            const int *en = elemToNode + (size_t)idx[q] * kDimElem;
            for (int k = 0; k < kDimElem; k++) nodeBlockMask[en[k] - 1] |= 1ull << color;
        }
    }
    // ---- blocks in colour order (stable), local colours inside every block ------------------------------------------
    std::vector<int> blocksByColor ((size_t)nbBlocks), countOfColor ((size_t)nbBlockColors + 1, 0);
    for (int b = 0; b < nbBlocks; b++) countOfColor[blockColor[b] + 1]++;
    for (int c = 0; c < nbBlockColors; c++) countOfColor[c + 1] += countOfColor[c];
    out.launchStart.assign (countOfColor.begin (), countOfColor.end ());
    {
        std::vector<int> cursor (countOfColor.begin (), countOfColor.end () - 1);
        for (int b = 0; b < nbBlocks; b++) blocksByColor[cursor[blockColor[b]]++] = b;
    }
    out.elemOrder.reserve ((size_t)nbElem);
    out.localIndex.clear (); out.localStart.clear ();
    std::vector<uint64_t> maskLo ((size_t)nbNodes, 0), maskHi ((size_t)nbNodes, 0);
    std::vector<int> elems, localColor;
    for (int nb = 0; nb < nbBlocks; nb++) {
        const int b = blocksByColor[nb];
        elems.assign (idx.begin () + leafStart[b], idx.begin () + leafStart[b + 1]);
        std::sort (elems.begin (), elems.end ());                  // greedy in element order, like coloring.cc
        localColor.assign (elems.size (), 0);
        int nbLocal = 0;
        for (size_t q = 0; q < elems.size (); q++) {
            const int *en = elemToNode + (size_t)elems[q] * kDimElem;
            uint64_t lo = 0, hi = 0;
            for (int k = 0; k < kDimElem; k++) { lo |= maskLo[en[k] - 1]; hi |= maskHi[en[k] - 1]; }
            int color;
            if (~lo != 0) color = __builtin_ctzll (~lo);
            else if (~hi != 0) color = 64 + __builtin_ctzll (~hi);
            else return -1;
            localColor[q] = color;
            nbLocal = std::max (nbLocal, color + 1);
            for (int k = 0; k < kDimElem; k++) {
                if (color < 64) maskLo[en[k] - 1] |= 1ull << color; else maskHi[en[k] - 1] |= 1ull << (color - 64);
            }
        }
        for (size_t q = 0; q < elems.size (); q++) {               // the masks are per block: clear what this block set
            const int *en = elemToNode + (size_t)elems[q] * kDimElem;
            for (int k = 0; k < kDimElem; k++) { maskLo[en[k] - 1] = 0; maskHi[en[k] - 1] = 0; }
        }
        out.localIndex.push_back ((int)out.localStart.size ());
        for (int c = 0; c < nbLocal; c++) {
            out.localStart.push_back ((int)out.elemOrder.size ());
            for (size_t q = 0; q < elems.size (); q++) if (localColor[q] == c) out.elemOrder.push_back (elems[q]);
        }
        out.localStart.push_back ((int)out.elemOrder.size ());     // end of the block's last colour
        out.maxLocalColors = std::max (out.maxLocalColors, nbLocal);
    }
    out.localIndex.push_back ((int)out.localStart.size ());
    out.nbBlocks = nbBlocks; out.nbBlockColors = nbBlockColors;
    return 0;
}

}  // namespace mfb
