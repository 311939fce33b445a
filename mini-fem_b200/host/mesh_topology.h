// Host-side mesh structures of the assembly path: the layouts the GPU kernels consume
// and that must stay bit-identical to what the reference builds once per run.
//
//   node_to_elem        DC-lib DC_create_nodeToElem   (call sites main.cc:247, coloring.cc:90)
//   build_csr           create_nodeToNode             (src/matrix.cc:55-91)
//   build_elem_to_edge  create_elemToEdge             (src/matrix.cc:25-52)
//   color_elements      coloring_creation             (src/coloring.cc:84-109)
//   permute_rows        DC_permute_int_2d_array       (call site main.cc:229)
//   boundary_mask       dqmrd4_ + e_essbcm_           (src/Fortran/qdmrd4.f, e_cgmelissa.F)
//
// Same outputs, different algorithms: stamps instead of O(deg^2) list scans, per-node
// colour masks instead of a materialised element-to-element graph.
#ifndef MFB_MESH_TOPOLOGY_H
#define MFB_MESH_TOPOLOGY_H

#include <cstdint>
#include <vector>

namespace mfb {

constexpr int kDimElem = 4;        // nodes per tetrahedron
constexpr int kDimNode = 3;        // coordinates per node
constexpr int kValuesPerElem = 16; // node pairs per element
constexpr int kMaxColor = 128;     // colours representable in the reference's 128-bit mask

// node -> incident elements, CSR form.  index has nbNodes+1 offsets, value 4*nbElem
// element ids in increasing order per node.  elemToNode holds 1-based node ids.
void node_to_elem (const int *elemToNode, int nbElem, int nbNodes, int *index, int *value);

// Number of CSR entries create_nodeToNode would produce.
int64_t count_csr_entries (const int *elemToNode, int nbElem, int nbNodes);

// Node-to-node CSR: row = 0-based offsets (nbNodes+1), col = 1-based node ids in
// first-seen order (NOT sorted).  col must hold count_csr_entries() ints.
int64_t build_csr (const int *elemToNode, int nbElem, int nbNodes, int *row, int *col);

// elemToEdge[e*16 + 4j + k] = CSR index of (node_j(e), node_k(e)).
// Returns 0, or -1 if a pair is missing from the CSR.
int build_elem_to_edge (const int *row, const int *col, const int *elemToNode,
                        int *elemToEdge, int nbElem);

// Greedy first-fit colouring in element order; colorPart[e] = colour,
// colorToElem[c] = first element of colour c after sorting (kMaxColor+1 ints),
// colorPerm[e] = position of element e in colour-sorted order (stable).
// Returns the number of colours or -1 if more than kMaxColor would be needed.
int color_elements (const int *elemToNode, int nbElem, int nbNodes, int *colorPart,
                    int *colorToElem, int *colorPerm);

// Locality-blocked colouring (the idea behind the reference's D&C variant,
// src/assembly.cc:123-244: colour leaves of a spatial decomposition instead of single elements).  The elements
// are cut into blocks of <= blockElems by recursive bisection of their centroids; BLOCKS are coloured so that two
// blocks of one colour share no node (one launch per block colour, one CTA per block), and inside a block the
// elements are coloured greedily in element order (the CTA walks its local colours with a barrier in between).
// A CSR row is then touched by few CTAs, each of which adds to it many times while it sits in L1 / L2 — where a
// colour of the element-wise colouring touches every row of the matrix once.
struct BlockColoring {
    std::vector<int> elemOrder;    // new position -> old element id: sorted by (block colour, block, local colour, id)
    std::vector<int> launchStart;  // nbBlockColors + 1: blocks [launchStart[c], launchStart[c+1]) have block colour c
    std::vector<int> localIndex;   // nbBlocks + 1: block b's local-colour offsets are localStart[localIndex[b] .. localIndex[b+1]]
    std::vector<int> localStart;   // element positions (new numbering); local colour q of block b =
                                   // [localStart[localIndex[b] + q], localStart[localIndex[b] + q + 1])
    int nbBlocks = 0, nbBlockColors = 0, maxLocalColors = 0;
};
// Returns 0, -1 if a block needs more than kMaxColor local colours, -2 if the blocks need more than 64 colours.
int build_block_coloring (const int *elemToNode, int nbElem, int nbNodes, const double *coord, int blockElems,
                          BlockColoring &out);

// tab[perm[i]] <- tab[i] for rows of `dim` ints.
void permute_rows (int *tab, const int *perm, int nbItem, int dim);

// checkBounds[comp*nbNodes + node] (component-major) from per-node boundary codes:
// 52 -> x, 53 -> y, 54 -> z, any other non-zero code -> all three.
// Returns the number of nodes with a non-zero code.
int boundary_mask (const int *boundNodesCode, int nbNodes, int *checkBounds);

// sqrt of the serial sum of squares, as compute_double_norm (src/FEM.cc:48-56).
double double_norm (const double *tab, int64_t size);

}  // namespace mfb

#endif
