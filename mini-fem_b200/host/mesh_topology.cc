#include "mesh_topology.h"

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

}  // namespace mfb
