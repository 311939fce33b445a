/* TEST INFRASTRUCTURE — CPU oracle for the Mini-FEM assembly hot path.
 *
 * A plain-C restatement of what the reference computes on the path
 *     assembly -> prec_init -> halo sum -> prec_inversion      (src/FEM.cc:177-257)
 * and of the once-per-run structures that path consumes (CSR, elemToEdge, colours,
 * Dirichlet mask).  Each function cites the reference lines it follows
 * (paths relative to /root/reference).
 *
 * Who may use this file: tests/, __graft_entry__.smoke() and bench.py's CPU baseline
 * legs — as the CHECKER only.  The product (mini-fem_b200/, include/) never links,
 * imports or executes it.
 *
 * Pinning: the reference ships no golden vectors for this path (its "checkings" files
 * live in an absent data/ tree), so this restatement is pinned against the reference's
 * OWN sources compiled here (oracle/_ref, see oracle/Makefile) by tests/test_oracle_*.py
 * and against the fixtures those sources produced (tests/golden/).  Three third-party
 * pieces stay PARITY UNPINNED because they are not in the reference tree: DC-lib's
 * helper semantics (marked [inferred]), MKL's DGETRF/DGETRI, and DefMesh's files.
 *
 * Arithmetic follows the reference's expression order and is compiled with
 * -ffp-contract=off, like the reference's -msse/-mavx builds (no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lapack3.h"

#define DIM_ELEM 4          /* src/headers/globals.h:22 */
#define DIM_NODE 3          /* src/headers/globals.h:23 */
#define VALUES_PER_ELEM 16  /* src/headers/globals.h:24 */
#define MAX_COLOR 128       /* src/coloring.cc:22 */

/* ---------------------------------------------------------------- DC-lib helpers */

/* DC_create_nodeToElem as used at src/main.cc:247 and src/coloring.cc:90.
 * [inferred] counting sort of (node, element) incidences: elements of a node appear in
 * increasing element id; elemToNode holds 1-based node ids; index/value are 0-based. */
void orc_node_to_elem (const int *elemToNode, int nbElem, int nbNodes, int *index,
                       int *value)
{
    int *cursor = (int*)calloc ((size_t)nbNodes + 1, sizeof (int));
    for (long k = 0; k < (long)nbElem * DIM_ELEM; k++) cursor[elemToNode[k]]++;
    index[0] = 0;
    for (int n = 0; n < nbNodes; n++) index[n + 1] = index[n] + cursor[n + 1];
    for (int n = 0; n < nbNodes; n++) cursor[n] = index[n];
    for (int e = 0; e < nbElem; e++) {
        for (int k = 0; k < DIM_ELEM; k++) {
            int n = elemToNode[e * DIM_ELEM + k] - 1;
            value[cursor[n]++] = e;
        }
    }
    free (cursor);
}

/* DC_create_permutation (src/coloring.cc:107) + DC_permute_int_2d_array
 * (src/main.cc:229).  [inferred] stable counting sort; perm[i] is where item i goes. */
void orc_create_permutation (int *perm, const int *part, int size, int nbPart)
{
    int *slot = (int*)calloc ((size_t)nbPart + 1, sizeof (int));
    for (int i = 0; i < size; i++) slot[part[i] + 1]++;
    for (int p = 0; p < nbPart; p++) slot[p + 1] += slot[p];
    for (int i = 0; i < size; i++) perm[i] = slot[part[i]]++;
    free (slot);
}

void orc_permute_int_2d (int *tab, const int *perm, int nbItem, int dimItem)
{
    size_t bytes = sizeof (int) * (size_t)nbItem * dimItem;
    int *old = (int*)malloc (bytes);
    memcpy (old, tab, bytes);
    for (int i = 0; i < nbItem; i++) {
        memcpy (tab + (size_t)perm[i] * dimItem, old + (size_t)i * dimItem,
                sizeof (int) * dimItem);
    }
    free (old);
}

/* ------------------------------------------------------------ layout: CSR, edges */

/* create_nodeToNode, src/matrix.cc:55-91.  Row i lists, in first-seen order, the
 * 1-based ids of every node of every element incident to node i (elements in
 * nodeToElem order, nodes in local order).  row is 0-based offsets.  Pass col == NULL
 * to count only.  Returns the number of entries (row[nbNodes]). */
int orc_create_nodeToNode (const int *elemToNode, int nbElem, int nbNodes, int *row,
                           int *col)
{
    int *idx = (int*)malloc (sizeof (int) * ((size_t)nbNodes + 1));
    int *val = (int*)malloc (sizeof (int) * (size_t)nbElem * DIM_ELEM);
    orc_node_to_elem (elemToNode, nbElem, nbNodes, idx, val);

    int total = 0, cap = 64;
    int *seen = (int*)malloc (sizeof (int) * cap);
    for (int i = 0; i < nbNodes; i++) {
        int nbSeen = 0, need = (idx[i + 1] - idx[i]) * DIM_ELEM;
        if (need > cap) { cap = need; seen = (int*)realloc (seen, sizeof (int) * cap); }
        if (row) row[i] = total;
        for (int j = idx[i]; j < idx[i + 1]; j++) {
            int e = val[j];
            for (int k = 0; k < DIM_ELEM; k++) {
                int cand = elemToNode[e * DIM_ELEM + k];
                int fresh = 1;
                for (int l = 0; l < nbSeen; l++) if (seen[l] == cand) fresh = 0;
                if (fresh) {
                    if (col) col[total] = cand;
                    seen[nbSeen++] = cand;
                    total++;
                }
            }
        }
    }
    if (row) row[nbNodes] = total;
    free (seen); free (val); free (idx);
    return total;
}

/* create_elemToEdge, src/matrix.cc:25-52: elemToEdge[e*16 + 4j + k] = CSR index of
 * (node_j, node_k). */
void orc_create_elemToEdge (const int *row, const int *col, const int *elemToNode,
                            int *elemToEdge, int nbElem)
{
    for (int e = 0; e < nbElem; e++) {
        int ctr = 0;
        for (int j = 0; j < DIM_ELEM; j++) {
            int n1 = elemToNode[e * DIM_ELEM + j] - 1;
            for (int k = 0; k < DIM_ELEM; k++) {
                int n2 = elemToNode[e * DIM_ELEM + k] - 1;
                for (int l = row[n1]; l < row[n1 + 1]; l++) {
                    if (col[l] == n2 + 1) {
                        elemToEdge[(size_t)e * VALUES_PER_ELEM + ctr] = l;
                        ctr++;
                        break;
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------- colouring */

/* coloring_creation, src/coloring.cc:84-109, with its helpers
 * create_longest_color_part (:46-81) and fill_color_index (:27-42).
 * Greedy first fit in element order over 128-bit colour masks; neighbours = elements
 * sharing a node (DC_create_elemToElem, [inferred]; only the set matters because
 * uncoloured elements carry mask 0).  colorToElem needs MAX_COLOR+1 ints.
 * Returns the number of colours, or -1 where the reference prints "Not enough colors"
 * and exits (:66-69). */
int orc_coloring (const int *elemToNode, int nbElem, int nbNodes, int *colorPart,
                  int *colorToElem, int *colorPerm)
{
    int *idx = (int*)malloc (sizeof (int) * ((size_t)nbNodes + 1));
    int *val = (int*)malloc (sizeof (int) * (size_t)nbElem * DIM_ELEM);
    orc_node_to_elem (elemToNode, nbElem, nbNodes, idx, val);

    unsigned __int128 *elemToColor =
        (unsigned __int128*)calloc ((size_t)nbElem > 0 ? nbElem : 1, sizeof (unsigned __int128));
    int nbColors = 0;
    for (int i = 0; i < nbElem; i++) {
        unsigned __int128 mask = 1, neighborColor = 0;
        int color = 0;
        for (int k = 0; k < DIM_ELEM; k++) {
            int n = elemToNode[i * DIM_ELEM + k] - 1;
            for (int p = idx[n]; p < idx[n + 1]; p++) neighborColor |= elemToColor[val[p]];
        }
        while (neighborColor & mask) { neighborColor >>= 1; color++; }
        if (color >= MAX_COLOR) {
            free (elemToColor); free (val); free (idx);
            return -1;
        }
        elemToColor[i] = mask << color;
        colorPart[i] = color;
        if (color > nbColors) nbColors = color;
    }
    nbColors++;
    free (elemToColor); free (val); free (idx);

    /* fill_color_index, offset 0 */
    int *count = (int*)calloc ((size_t)nbColors, sizeof (int));
    for (int i = 0; i < nbElem; i++) count[colorPart[i]]++;
    colorToElem[0] = 0;
    for (int c = 1; c <= nbColors; c++) colorToElem[c] = colorToElem[c - 1] + count[c - 1];
    free (count);

    orc_create_permutation (colorPerm, colorPart, nbElem, MAX_COLOR);
    return nbColors;
}

/* ------------------------------------------------------------- Dirichlet mask */

/* dqmrd4_ (src/Fortran/qdmrd4.f:1-23) then e_essbcm_ (src/Fortran/e_cgmelissa.F:1-51)
 * as called at src/main.cc:343-345 with NDIM = 3: nodes with a non-zero code; code 52
 * masks x, 53 y, 54 z (DefMesh_jlog.h:17-20), any other non-zero code all three.
 * checkBounds is component-major: checkBounds[comp*nbNodes + node]. */
void orc_boundary_mask (const int *boundNodesCode, int nbNodes, int *checkBounds)
{
    for (long k = 0; k < (long)nbNodes * DIM_NODE; k++) checkBounds[k] = 0;
    for (int n = 0; n < nbNodes; n++) {
        int code = boundNodesCode[n];
        if (code == 0) continue;
        if (code == 52)      checkBounds[0 * (size_t)nbNodes + n] = 1;
        else if (code == 53) checkBounds[1 * (size_t)nbNodes + n] = 1;
        else if (code == 54) checkBounds[2 * (size_t)nbNodes + n] = 1;
        else {
            checkBounds[0 * (size_t)nbNodes + n] = 1;
            checkBounds[1 * (size_t)nbNodes + n] = 1;
            checkBounds[2 * (size_t)nbNodes + n] = 1;
        }
    }
}

/* ------------------------------------------------------------------ assembly */

/* elem_coef_seq, src/assembly.cc:85-121.  Gradient coefficients of one P1 tetrahedron:
 * edge vectors from node 3 to nodes 0, 2, 1 (a, b, c), cross products, row 3 = minus
 * the sum, everything times (1./vol) with vol = a . row0. */
void orc_elem_coef (const double *coord, const int *elemToNode, int elem, double *c /*[4][3]*/)
{
    double p[DIM_ELEM][DIM_NODE];
    for (int i = 0; i < DIM_ELEM; i++) {
        int n = elemToNode[elem * DIM_ELEM + i] - 1;
        for (int j = 0; j < DIM_NODE; j++) p[i][j] = coord[(size_t)n * DIM_NODE + j];
    }
    double xa = p[0][0] - p[3][0], xb = p[2][0] - p[3][0], xc = p[1][0] - p[3][0];
    double ya = p[0][1] - p[3][1], yb = p[2][1] - p[3][1], yc = p[1][1] - p[3][1];
    double za = p[0][2] - p[3][2], zb = p[2][2] - p[3][2], zc = p[1][2] - p[3][2];
    c[0] = yb * zc - yc * zb;  c[1] = zb * xc - zc * xb;  c[2]  = xb * yc - xc * yb;
    c[3] = ya * zb - yb * za;  c[4] = za * xb - zb * xa;  c[5]  = xa * yb - xb * ya;
    c[6] = yc * za - ya * zc;  c[7] = zc * xa - za * xc;  c[8]  = xc * ya - xa * yc;
    c[9]  = - (c[0] + c[3] + c[6]);
    c[10] = - (c[1] + c[4] + c[7]);
    c[11] = - (c[2] + c[5] + c[8]);
    double vol = xa * c[0] + ya * c[1] + za * c[2];
    double inv = 1. / vol;
    for (int k = 0; k < DIM_ELEM * DIM_NODE; k++) c[k] *= inv;
}

static int find_edge (const int *row, const int *col, int n1, int n2OneBased)
{
    for (int l = row[n1]; l < row[n1 + 1]; l++) if (col[l] == n2OneBased) return l;
    return -1;
}

/* assembly_lap_seq, src/assembly.cc:485-588 (OPTIMIZED :533-544, search :547-562). */
static void lap_interval (const double *coord, double *values, const int *row,
                          const int *col, const int *elemToNode, const int *elemToEdge,
                          int first, int last)
{
    double c[12];
    for (int e = first; e <= last; e++) {
        orc_elem_coef (coord, elemToNode, e, c);
        for (int j = 0; j < DIM_ELEM; j++) {
            for (int k = 0; k < DIM_ELEM; k++) {
                int l = elemToEdge ? elemToEdge[(size_t)e * VALUES_PER_ELEM + 4 * j + k]
                                   : find_edge (row, col, elemToNode[e * DIM_ELEM + j] - 1,
                                                elemToNode[e * DIM_ELEM + k]);
                if (l < 0) continue;
                values[l] += (c[3*j+0] * c[3*k+0] + c[3*j+1] * c[3*k+1] + c[3*j+2] * c[3*k+2]);
            }
        }
    }
}

/* assembly_ela_seq, src/assembly.cc:332-479 (OPTIMIZED :380-412, search :415-451).
 * 3x3 block per node pair, row-major; diagonal entries weight their own component by
 * 2.25, off-diagonal entries are 1.25 * c_a * d_b. */
static void ela_interval (const double *coord, double *values, const int *row,
                          const int *col, const int *elemToNode, const int *elemToEdge,
                          int first, int last)
{
    double c[12];
    for (int e = first; e <= last; e++) {
        orc_elem_coef (coord, elemToNode, e, c);
        for (int j = 0; j < DIM_ELEM; j++) {
            for (int k = 0; k < DIM_ELEM; k++) {
                int l = elemToEdge ? elemToEdge[(size_t)e * VALUES_PER_ELEM + 4 * j + k]
                                   : find_edge (row, col, elemToNode[e * DIM_ELEM + j] - 1,
                                                elemToNode[e * DIM_ELEM + k]);
                if (l < 0) continue;
                double *v = values + (size_t)l * 9;
                const double *a = c + 3 * j, *b = c + 3 * k;
                v[0] += a[0] * b[0] * 2.25 + a[1] * b[1] + a[2] * b[2];
                v[1] += a[0] * b[1] * 1.25;
                v[2] += a[0] * b[2] * 1.25;
                v[3] += a[1] * b[0] * 1.25;
                v[4] += a[0] * b[0] + a[1] * b[1] * 2.25 + a[2] * b[2];
                v[5] += a[1] * b[2] * 1.25;
                v[6] += a[2] * b[0] * 1.25;
                v[7] += a[2] * b[1] * 1.25;
                v[8] += a[0] * b[0] + a[1] * b[1] + a[2] * b[2] * 2.25;
            }
        }
    }
}

/* assembly, src/assembly.cc:615-720: zero the values (:649-651 / :663-666), then one
 * interval [0, nbElem-1] (REF, :653-659) or one interval per colour
 * (coloring_assembly, :593-611) when colorToElem != NULL.  elemToEdge != NULL selects
 * the OPTIMIZED index path. */
void orc_assembly (const double *coord, double *values, const int *row, const int *col,
                   const int *elemToNode, const int *elemToEdge, int nbElem, int nbEdges,
                   int operatorID, const int *colorToElem, int nbTotalColors)
{
    int operatorDim = operatorID == 0 ? 1 : DIM_NODE * DIM_NODE;
    for (long i = 0; i < (long)nbEdges * operatorDim; i++) values[i] = 0;
    if (colorToElem == NULL) {
        if (operatorID == 0) lap_interval (coord, values, row, col, elemToNode, elemToEdge, 0, nbElem - 1);
        else                 ela_interval (coord, values, row, col, elemToNode, elemToEdge, 0, nbElem - 1);
        return;
    }
    for (int color = 0; color < nbTotalColors; color++) {
        int first = colorToElem[color], last = colorToElem[color + 1] - 1;
        if (operatorID == 0) lap_interval (coord, values, row, col, elemToNode, elemToEdge, first, last);
        else                 ela_interval (coord, values, row, col, elemToNode, elemToEdge, first, last);
    }
}

/* ------------------------------------------------------------- preconditioner */

/* prec_init, src/preconditioner.cc:52-87: zero prec, then copy each node's diagonal
 * CSR block (first column equal to the node itself). */
void orc_prec_init (double *prec, const double *values, const int *row, const int *col,
                    int nbNodes, int operatorDim)
{
    for (long i = 0; i < (long)nbNodes * operatorDim; i++) prec[i] = 0;
    for (int i = 0; i < nbNodes; i++) {
        for (int j = row[i]; j < row[i + 1]; j++) {
            if (col[j] - 1 == i) {
                for (int k = 0; k < operatorDim; k++) {
                    prec[(size_t)i * operatorDim + k] = values[(size_t)j * operatorDim + k];
                }
                break;
            }
        }
    }
}

/* ela_invert_prec, src/Fortran/elasclpr.f:2-56, for one 1-based node `cur`.
 * prec(ki,kj,i) is column-major over the C block: blk[kj*3 + ki]. */
static void ela_invert_one (int nbNodes, const int *row, const int *col, double *prec,
                            const int *checkBounds, int cur)
{
    double *blk = prec + (size_t)(cur - 1) * 9;
    for (int ki = 0; ki < 3; ki++) {
        if (checkBounds[(size_t)ki * nbNodes + (cur - 1)] != 0) {           /* :19-27 */
            for (int kj = 0; kj < 3; kj++) { blk[kj * 3 + ki] = 0.; blk[ki * 3 + kj] = 0.; }
            blk[ki * 3 + ki] = 1.;
        }
    }
    int hasDiag = 0;                                                        /* :29-32 */
    for (int m = row[cur - 1] + 1; m <= row[cur]; m++) {
        if (col[m - 1] == cur) { hasDiag = 1; break; }
    }
    if (!hasDiag) return;
    double a[9];
    int ipiv[3];
    memcpy (a, blk, sizeof a);                                              /* :34-38 */
    l3_getrf (3, a, ipiv);                                                  /* :39 */
    l3_getri (3, a, ipiv);                                                  /* :44 */
    memcpy (blk, a, sizeof a);                                              /* :49-53 */
}

/* prec_inversion, src/preconditioner.cc:25-49: scalar reciprocal (lap, :40; the mask
 * is not used) or masked 3x3 inverse per node (ela, :44-46). */
void orc_prec_inversion (double *prec, const int *row, const int *col,
                         const int *checkBounds, int nbNodes, int operatorID)
{
    for (int i = 0; i < nbNodes; i++) {
        if (operatorID == 0) prec[i] = 1.0 / prec[i];
        else ela_invert_one (nbNodes, row, col, prec, checkBounds, i + 1);
    }
}

/* ------------------------------------------------------------------- halo sum */

/* MPI_halo_exchange, src/halo.cc:39-122, for ALL subdomains at once: every rank packs
 * its pre-exchange interface values (:77-80), "sends" segment i to neighborsList[i]-1
 * (:85-93), and adds what it received from that neighbour, position by position, into
 * its own interface nodes (:113-116).  The j-th node of rank r's list for neighbour s
 * is the j-th node of s's list for r.  No-op when nranks < 2 (:44).
 * Returns 0, or -1 if two facing segments disagree in length. */
int orc_halo_exchange (int nranks, double **prec, int **intfIndex, int **intfNodes,
                       int **neighborsList, const int *nbIntf, int operatorDim)
{
    if (nranks < 2) return 0;
    double **sendbuf = (double**)calloc ((size_t)nranks, sizeof (double*));
    for (int r = 0; r < nranks; r++) {
        int total = intfIndex[r][nbIntf[r]];
        sendbuf[r] = (double*)malloc (sizeof (double) * (size_t)(total > 0 ? total : 1) * operatorDim);
        for (int j = 0; j < total; j++) {
            int node = intfNodes[r][j] - 1;
            for (int k = 0; k < operatorDim; k++) {
                sendbuf[r][(size_t)j * operatorDim + k] = prec[r][(size_t)node * operatorDim + k];
            }
        }
    }
    int status = 0;
    for (int r = 0; r < nranks; r++) {
        for (int i = 0; i < nbIntf[r]; i++) {
            int s = neighborsList[r][i] - 1, begin = intfIndex[r][i], end = intfIndex[r][i + 1];
            int back = -1;                       /* s's segment that faces r */
            for (int q = 0; q < nbIntf[s]; q++) if (neighborsList[s][q] - 1 == r) back = q;
            if (back < 0 || intfIndex[s][back + 1] - intfIndex[s][back] != end - begin) {
                status = -1;
                continue;
            }
            const double *incoming = sendbuf[s] + (size_t)intfIndex[s][back] * operatorDim;
            for (int j = begin; j < end; j++) {
                int node = intfNodes[r][j] - 1;
                for (int k = 0; k < operatorDim; k++) {
                    prec[r][(size_t)node * operatorDim + k] += incoming[(size_t)(j - begin) * operatorDim + k];
                }
            }
        }
    }
    for (int r = 0; r < nranks; r++) free (sendbuf[r]);
    free (sendbuf);
    return status;
}

/* compute_double_norm, src/FEM.cc:48-56: serial sum of pow(x,2), then sqrt. */
double orc_norm (const double *tab, long size)
{
    double norm = 0;
    for (long i = 0; i < size; i++) norm += pow (tab[i], 2);
    return sqrt (norm);
}
