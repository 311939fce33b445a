// Tile plan of the write-once ("TILED") assembly path.
//
// The reference scatters: every element adds 16 node-pair blocks into the CSR values
// (src/assembly.cc:382-412).  On a GPU that means a zero-fill plus read-modify-write
// of a value array ~9x larger than L2.  The plan turns the scatter into a gather that
// writes every CSR entry exactly once:
//
//   * nodes (= CSR rows) are grouped into spatially compact TILES (Morton order of the
//     coordinates, cut greedily under shared-memory caps); one CTA works on one tile;
//   * a tile lists every element touching its rows (elements on tile borders appear in
//     several tiles and their 12 gradient coefficients are recomputed there);
//   * each CSR entry (i,j) of an owned row carries the list of (element, a, b) triples
//     that contribute to it, a/b = local index of i/j in the element;
//   * the diagonal entry (i,i) has one contribution per incident element and is handled
//     by 4 lanes per row;
//   * in the off-diagonal pass every row starts on a half-warp boundary (chunks of <= 16
//     entries), two chunks share a warp, and the contribution codes are stored transposed
//     [step][32 lanes].  Tile-local element ids follow a coset numbering (id mod 4 = class,
//     classes balanced around every row) and each chunk is scheduled so that one step touches
//     elements of different classes: with a plane stride of 4 (mod 16) their coefficient
//     slots then sit in disjoint shared-memory banks.
//
// Everything one tile needs is ONE contiguous, 16-byte aligned record ("blob") so that a
// single bulk copy (TMA, cp.async.bulk) stages it into shared memory.
#ifndef MFB_TILE_PLAN_H
#define MFB_TILE_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

namespace mfb {

// Threads of the plan builders.  Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank, which would
// make a plan of an EIB-size subdomain take half a minute; the builders are private to this library, so they
// size their own team: MFB_PLAN_THREADS if set, else the host's hardware threads divided by the ranks of this
// node (LOCAL_WORLD_SIZE), between 1 and 32.
int plan_team_size ();


// First 48 bytes of a tile blob.  Offsets are in bytes from the start of the blob and
// multiples of 16.  The row table follows the header immediately.
struct TileBlobHeader {
    uint16_t nbRows, nbNodesRef, nbElems, nbEntries, nbBatches, hasInterface;   // nbElems counts ids (holes included)
    uint32_t offNodes;      // int[nbNodesRef]: 0-based global ids, owned rows first
    uint32_t offElems;      // ushort4[nbElems]: tile-local node indices of each element
    uint32_t offEntryRow;   // uint8[nbEntries]: owned-row index of each tile-local entry
    uint32_t offLaneEntry;  // uint16[32*nbBatches]: tile-local entry of each off-diagonal-pass lane, 0xFFFF = idle
    uint32_t offBatches;    // TileBatch[nbBatches]
    uint32_t offDiag;       // uint16[]: (element<<2 | a) per diagonal contribution
    uint32_t offPair;       // uint16[]: (element<<4 | a<<2 | b), transposed [step][lane]
    uint32_t blobBytes;
    uint32_t pad[1];
};

// 16 bytes per owned row, plus one sentinel whose diagCodeBase closes the last row.
struct TileRow {
    int node;             // 0-based global node id; bit 31 set = interface node
    int valueStart;       // nodeToNodeRow[node]
    int diagCodeBase;     // first diagonal code of this row (index into the blob's diag section)
    uint16_t localStart;  // tile-local index of the row's first entry
    uint16_t diagLocal;   // tile-local index of the diagonal entry (0xFFFF if none)
};

// 8 bytes per warp batch of 32 consecutive tile-local entries.
struct TileBatch {
    int codeBase;         // first code of the batch (index into the blob's pair section)
    int steps;            // longest contribution list in the batch
};

struct TilePlan {
    int nbTiles = 0, maxRows = 0, maxElems = 0, maxNodesRef = 0, maxEntries = 0;
    int nbInterfaceTiles = 0;            // tiles owning >= 1 interface node come first
    int elemStride = 0;                  // shared-memory stride between the 4 local-node planes
    bool laplacian = false;              // codes are shared-memory slots of dot products (see lap_pair_slot)
    int64_t nbTileElems = 0, nbContributions = 0, nbPaddedSteps = 0;
    uint32_t maxBlobBytes = 0;
    uint32_t maxHeadBytes = 0, maxTailBytes = 0;   // head = header..elements (bytes [0, offEntryRow)), tail = the rest
    std::vector<uint64_t> tileOffset;    // nbTiles + 1 byte offsets into `blob`
    std::vector<uint8_t> blob;
    int64_t bytes () const { return (int64_t)blob.size () + (int64_t)tileOffset.size () * 8; }
    const TileBlobHeader *header (int t) const { return reinterpret_cast<const TileBlobHeader*> (blob.data () + tileOffset[t]); }
};

struct TilePlanLimits {
    int maxRows = 36;        // <= 255 (entryRow is a byte); 36 / 384 keeps a tile under 75 KB of shared
                             // memory, i.e. three co-resident CTAs per SM (measured fastest on B200)
    int maxElems = 384;      // <= 4064 (12-bit element field, slack for numbering holes + padding slot)
    int maxNodesRef = 384;   // shared-memory coordinate staging
    int maxEntries = 2048;   // <= 65535
    bool laplacian = false;  // scalar operator: codes become dot-product slots, padding goes to bank c+12
    bool bankAware = true;   // coset numbering + conflict-aware step schedule (false: element order)
};

// Padding codes name an element id in [nbElems, nbElems + 16): the kernel keeps all-zero
// coefficient vectors there, so padded steps need neither a branch nor a select.
int tile_plan_stride (int maxElems);   // shared-memory stride between the 4 local-node planes

// Laplacian tiles keep, per element, the 10 distinct dot products of its gradient rows instead
// of the rows themselves (one shared-memory load per contribution instead of six), and their
// codes ARE the shared-memory slots.  Plane of the pair (a,b), a != b: 2*colour + (a && b),
// colour = (a ^ b) - 1 — the proper 3-edge-colouring of K4 ({01,23}, {02,13}, {03,12}), so the
// three pairs of one local node get three different colours; plane of (a,a): 6 + a.  Plane p
// starts at p*PS + 4*residue (PS = stride - 4, a multiple of 16; residue = colour, or a for the
// diagonal planes): with the coset numbering an element of class c keeps its off-diagonal
// dots in banks {c, c+4, c+8}.
#ifdef __CUDACC__
#define MFB_HD __host__ __device__ __forceinline__
#else
#define MFB_HD inline
#endif
MFB_HD int lap_pair_slot (int a, int b, int e, int PS)
{
    const int colour = (a ^ b) - 1;
    return (2 * colour + ((a && b) ? 1 : 0)) * PS + 4 * colour + e;
}
MFB_HD int lap_diag_slot (int a, int e, int PS) { return (6 + a) * PS + 4 * a + e; }
// Steps 1-2 of every write-once plan: nodes in Morton order of their coordinates, cut greedily
// into tiles under the caps.  nodeOrder[tileStart[t] .. tileStart[t+1]) = the rows tile t owns.
// n2eIndex / n2eValue = node_to_elem (mesh_topology.h).
struct TileCutLimits { int maxRows, maxElems, maxNodesRef, maxEntries; };
int cut_node_tiles (int nbNodes, int nbElem, const int *elemToNode, const int *row, const double *coord,
                    const int *n2eIndex, const int *n2eValue, const TileCutLimits &limits,
                    std::vector<int> &nodeOrder, std::vector<int> &tileStart, std::string &error);

//
// isInterface may be null.  Returns 0, or -1 with `error` set (e.g. one node alone
// exceeds a cap, or the CSR lacks a pair).
int build_tile_plan (int nbNodes, int nbElem, const int *elemToNode, const int *row,
                     const int *col, const double *coord, const uint8_t *isInterface,
                     const TilePlanLimits &limits, TilePlan &plan, std::string &error);

// Replays the plan on the host and compares it with the reference's double loop over
// (element, j, k) (src/assembly.cc:382-412): every triple must appear exactly once, on the
// entry whose row is node_j and whose column is node_k; rows, tiles and batches must tile
// the CSR exactly.  Returns 0 or -1 with `error` set.
int verify_tile_plan (const TilePlan &plan, int nbNodes, int nbElem, const int *elemToNode,
                      const int *row, const int *col, std::string &error);

}  // namespace mfb

#endif
