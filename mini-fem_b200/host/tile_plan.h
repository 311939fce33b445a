// Tile plan of the write-once ("TILED") assembly path.
//
// The reference scatters: every element adds 16 node-pair blocks into the CSR values
// (src/assembly.cc:382-412).  On a GPU that means a zero-fill plus read-modify-write
// of a value array ~9x larger than L2.  The plan turns the scatter into a gather that
// writes every CSR entry exactly once:
//
//   * nodes (= CSR rows) are grouped into spatially compact TILES (Morton order of the
//     coordinates, cut greedily under shared-memory caps); one CTA owns one tile;
//   * a tile lists every element touching its rows (elements on tile borders appear in
//     several tiles and their 12 gradient coefficients are recomputed there);
//   * each CSR entry (i,j) of an owned row carries the list of (element, a, b) triples
//     that contribute to it, a/b = local index of i/j in the element, elements in
//     increasing id (the reference REF build's summation order);
//   * the diagonal entry (i,i) has one contribution per incident element and is handled
//     by 4 lanes per row; the off-diagonal lists are stored transposed per 32 entries
//     (one warp) so that a warp reads one coalesced 64-byte line per step.
//
// All arrays are flat so that they can be copied to the device verbatim.
#ifndef MFB_TILE_PLAN_H
#define MFB_TILE_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

namespace mfb {

// 32 bytes per tile.
struct TileHeader {
    int nodeBase;     // first referenced node in TilePlan::tileNodes
    int elemBase;     // first element in TilePlan::tileElems
    int rowBase;      // first row in TilePlan::rows
    int entryBase;    // first entry in TilePlan::entryRow
    int batchBase;    // first 32-entry batch in TilePlan::batches
    uint16_t nbRows, nbNodesRef, nbElems, nbEntries;
    int pad;
};

// 16 bytes per owned row.
struct TileRow {
    int node;             // 0-based global node id; bit 31 set = interface node
    int valueStart;       // nodeToNodeRow[node]
    int diagCodeBase;     // first (element<<2 | a) code of this row in TilePlan::diagCodes
    uint16_t localStart;  // tile-local index of the row's first entry
    uint16_t diagLocal;   // tile-local index of the diagonal entry (0xFFFF if none)
};

// 8 bytes per warp batch of 32 consecutive tile-local entries.
struct TileBatch {
    int codeBase;         // first code in TilePlan::pairCodes (transposed: [step][lane])
    int steps;            // longest contribution list in the batch
};

struct TilePlan {
    int nbTiles = 0, maxRows = 0, maxElems = 0, maxNodesRef = 0, maxEntries = 0;
    int nbInterfaceTiles = 0;            // tiles owning >= 1 interface node come first
    int64_t nbTileElems = 0, nbContributions = 0;
    std::vector<TileHeader> tiles;
    std::vector<int> tileNodes;          // 0-based global ids; owned rows first, then halo nodes
    std::vector<uint16_t> tileElems;     // 4 tile-local node indices per element
    std::vector<TileRow> rows;           // + 1 sentinel: rows[r+1].diagCodeBase ends row r's codes
    std::vector<uint8_t> entryRow;       // per tile-local entry: owned-row index
    std::vector<TileBatch> batches;
    std::vector<uint16_t> pairCodes;     // (element<<4 | a<<2 | b); padding = (nbElems<<4), a
                                         // slot the kernel keeps at zero, so no branch is needed
    std::vector<uint16_t> diagCodes;     // (element<<2 | a)
    int64_t bytes () const;
};

struct TilePlanLimits {
    int maxRows = 64;        // <= 255 (entryRow is a byte)
    int maxElems = 704;      // <= 4094 (12-bit element field, one slot kept for padding)
    int maxNodesRef = 512;   // shared-memory coordinate staging
    int maxEntries = 2048;   // <= 65535
};

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
