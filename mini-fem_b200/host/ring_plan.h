// Plan of the RING path: write-once assembly that walks the ring of elements around every
// mesh edge straight from the node coordinates.
//
// The TILED path (tile_plan.h) computes the 12 gradient coefficients of every tile element into
// shared-memory planes and lets one lane per CSR entry gather 48 bytes per contribution; ncu
// shows it bound by shared-memory bandwidth.  The RING path keeps the tiles
// (same Morton cut, cut_node_tiles) but removes the coefficient planes:
//
//   * a JOB is one mesh edge {i, j} with i owned by the tile.  Its lane walks the nodes of the
//     edge's link (the polygon r_0, r_1, ... around the edge: element k = (i, j, r_k, r_k+1)),
//     loading ONE node (24 bytes) per element, and rebuilds the two gradients it needs from the
//     coordinates: with d = x_j - x_i, u = x_p - x_i, w = x_q - x_i,
//         n_j = u x w,  n_i = n_j + (w - u) x d,  det = n_j . d  (= 6 V, signed),
//         grad(lambda_j) = n_j / det,  grad(lambda_i) = -n_i / det      (src/assembly.cc:85-121
//     computes exactly these gradients as cross products over vol), so the element adds
//         A_ij += -(n_i n_j^T) / det^2                                   (src/assembly.cc:386-409)
//     and the CSR entry is K_ij = 1.25 A_ij + tr(A_ij) I as on the TILED path;
//   * when j is owned by the same tile the job also writes K_ji = K_ij^T: every interior edge
//     is computed once;
//   * jobs are independent of rows, so they are sorted by ring length and packed 32 to a warp;
//   * finished blocks go to a tile-wide shared slab laid out like the tile's CSR rows; the
//     write-out streams each row to global memory as one contiguous run and sums it on the way:
//     the element matrices have zero row sums (the four gradients of an element add up to
//     zero), hence K_ii = -sum_{j != i} K_ij — the diagonal entry and the preconditioner block
//     come from the run that is being written anyway.
//
// Per tile the plan is one contiguous record: HEAD = header, row table, node list (needed one
// tile ahead for the coordinate prefetch and during the write-out), TAIL = batches, jobs, ring
// codes (needed by the job phase only).
#ifndef MFB_RING_PLAN_H
#define MFB_RING_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

namespace mfb {

constexpr int kRingBreak = 0xFE;   // code byte: the chain of elements is interrupted, the next node starts a new one
constexpr int kRingBatchGeneral = 1; // RingBatch::flags: some job of the batch holds a break (the kernel's general loop)
constexpr int kRingMaxNodes = 254; // tile-local node ids 0 .. 253

struct RingTileHeader {            // 32 bytes
    uint16_t nbRows, nbNodes, nbBatches, nbEntries, hasInterface, pad0;   // nbEntries = slab slots, padding included
    uint32_t offNodes;             // int[nbNodes]: 0-based global ids by tile-local id (holes name a valid node)
    uint32_t headBytes;            // bytes [0, headBytes) = header + rows + nodes; the tail starts here with RingBatch[nbBatches]
    uint32_t offJobs;              // uint64[32 * nbBatches]
    uint32_t offCodes;             // uint64[]: per batch [word][32 lanes], 8 code bytes per word, low byte first;
                                   // a job's bytes beyond its length name a valid node (they are loaded and masked)
    uint32_t blobBytes;
};

// RingRow::node: 0-based global node id in the low 27 bits; bit 27 is free for the kernel (it tags rows that
// hold a diagonal entry); bits 28..30 = Dirichlet mask of the components x, y, z (checkBounds[c * nbNodes + node]
// != 0: src/Fortran/e_cgmelissa.F, applied by ela_invert_prec, src/Fortran/elasclpr.f:19-27); bit 31 = interface
// node (its block leaves the fused kernel raw, for the halo sum).
constexpr int kRingNodeMask = 0x07FFFFFF;
constexpr int kRingMaxNodeId = kRingNodeMask;

struct RingRow {                   // 16 bytes per owned row, right after the header
    int node;                      // see above
    int valueStart;                // nodeToNodeRow[node]
    uint16_t localStart;           // slab slot of the row's first entry; consecutive rows start 1 (mod 8) slots apart
                                   // (ring_row_padding), so that three rows can stream out side by side
    uint16_t len;                  // entries of the row
    uint16_t diagOff;              // position of the diagonal entry inside the row, 0xFFFF = none
    uint16_t pad0;
};

// Idle slab slots after a row of `len` entries: the next row starts at a slot = 1 (mod 8) past this row's
// start.  The write-out lets three groups of ten lanes copy three consecutive rows at once, entry q of each
// in the same instruction; with 80-byte slab entries the three 72-byte pieces then fall into disjoint banks.
inline int ring_row_padding (int len) { return ((1 - len) % 8 + 8) % 8; }

struct RingBatch {                 // 8 bytes per warp batch of 32 jobs
    uint32_t codeBase;             // first code word of the batch (index into the codes section)
    uint16_t nbSteps;              // code bytes to walk (longest job of the batch); ceil (nbSteps / 8) code words per lane
    uint16_t flags;                // kRingBatchGeneral
};

// job word: i | j << 8 | slotIJ << 16 | slotJI << 32 | len << 48; slots are tile-local entry indices
// (slab position = slot * slab stride), 0xFFFF = no such block; len = code bytes of the job (byte 0 names the
// first node of a chain, every further node byte adds one element); an idle lane has both slots 0xFFFF and len 0.
inline uint64_t ring_job (int i, int j, int slotIJ, int slotJI, int len)
{
    return (uint64_t)(i & 0xFF) | ((uint64_t)(j & 0xFF) << 8) | ((uint64_t)(slotIJ & 0xFFFF) << 16) |
           ((uint64_t)(slotJI & 0xFFFF) << 32) | ((uint64_t)(len & 0xFFFF) << 48);
}

struct RingPlanLimits {
    int maxRows = 36;              // rows per tile (<= 255)
    int maxEntries = 640;          // slab slots per tile: CSR entries plus the padding between rows (with the Morton
                                   // cut the cap applies to the entries alone); 640 x 80 bytes keeps three CTAs per SM
    int maxNodes = kRingMaxNodes;  // tile-local nodes (<= 254: one code byte per node)
    int maxJobs = 1 << 30;         // jobs (mesh edges with an owned end) per tile: a multiple of 32 x the CTA's warps lets the
                                   // job phase finish in whole rounds of warp batches
    bool bankAware = true;         // node numbering + ring rotation chosen against bank conflicts
    int rotationSweeps = 1;        // coordinate-descent sweeps over the lanes of a half-warp after the greedy rotation choice
    bool bisection = true;         // tiles = leaves of a recursive coordinate bisection (false: runs of the Morton curve, as TILED)
    int refinePasses = 1;          // renumber-and-rotate rounds after the first numbering.  Modelled gather conflict
                                   // factor / plan build time on a 64^3 Kuhn mesh: (0 rounds, 0 sweeps) 1.34 / 1.2 s,
                                   // (0, 1) 1.27 / 1.9 s, (1, 0) 1.28 / 1.8 s, (1, 1) 1.22 / 3.0 s, (2, 1) 1.22 / 4.3 s
};

struct RingPlan {
    int nbTiles = 0, nbInterfaceTiles = 0;
    int maxRows = 0, maxNodes = 0, maxEntries = 0, maxBatches = 0;
    uint32_t maxBlobBytes = 0, maxHeadBytes = 0, maxTailBytes = 0;
    int64_t nbJobs = 0, nbSymmetricJobs = 0, nbRingSteps = 0, nbPaddedSteps = 0, nbBreaks = 0;
    // shared-memory model, in wavefronts of 128 bytes.  gather: the LDS.64 of the coordinate planes, one
    // wavefront per half-warp and plane without conflicts.  slab: per quarter-warp (8 lanes) and per 128-bit
    // store of a finished block (an elasticity block = 4 such stores + one 64-bit one)
    int64_t gatherWavefronts = 0, gatherIdeal = 0, slabWriteWavefronts = 0, slabWriteIdeal = 0;
    std::vector<uint64_t> tileOffset;   // nbTiles + 1 byte offsets into `blob`
    std::vector<uint8_t> blob;
    const RingTileHeader *header (int t) const { return reinterpret_cast<const RingTileHeader*> (blob.data () + tileOffset[t]); }
};

// isInterface and checkBounds (int[3][nbNodes], component-major: main.cc:343-345) may be null.
// Returns 0, or -1 with `error` set.
int build_ring_plan (int nbNodes, int nbElem, const int *elemToNode, const int *row, const int *col,
                     const double *coord, const uint8_t *isInterface, const int *checkBounds,
                     const RingPlanLimits &limits, RingPlan &plan, std::string &error);

// Structural replay: rows tile the CSR; every off-diagonal CSR entry is written by exactly one job
// (as its (i,j) or its transposed (j,i) block); the consecutive node pairs of a job's chains are
// distinct elements containing {i, j, p, q}; every element's 12 ordered off-diagonal node pairs
// (src/assembly.cc:382-412) are covered exactly once.  Returns 0 or -1 with `error` set.
int verify_ring_plan (const RingPlan &plan, int nbNodes, int nbElem, const int *elemToNode,
                      const int *row, const int *col, const int *checkBounds, std::string &error);

}  // namespace mfb

#endif
