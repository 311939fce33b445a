// Launchers of every kernel on the path (definitions in kernels_*.cu).
#ifndef MFB_KERNELS_CUH
#define MFB_KERNELS_CUH

#include <cuda_runtime.h>
#include <cstdint>

#include "../host/tile_plan.h"
#include "../host/ring_plan.h"

namespace mfb {

// ATOMIC / COLOR element kernels over an element interval [firstElem, firstElem+count).
cudaError_t launch_scatter (int operatorID, bool atomic, const double *coord, const int *elemToNode,
                            const int *elemToEdge, double *values, int firstElem, int count,
                            cudaStream_t stream);
// BLOCKCOLOR: the blocks [firstBlock, firstBlock + nbBlocks) of one block colour, one CTA each (host/mesh_topology.h).
cudaError_t launch_scatter_blocks (int operatorID, const double *coord, const int *elemToNode, const int *elemToEdge,
                                   double *values, const int *localIndex, const int *localStart, int firstBlock,
                                   int nbBlocks, cudaStream_t stream);
cudaError_t launch_elem_to_edge (const int *row, const int *col, const int *elemToNode,
                                 int *elemToEdge, int nbElem, int *missing, cudaStream_t stream);
cudaError_t launch_diag_index (const int *row, const int *col, int *diagIndex, int nbNodes,
                               cudaStream_t stream);

cudaError_t launch_prec_init (int operatorDim, double *prec, const double *values,
                              const int *diagIndex, int nbNodes, cudaStream_t stream);
cudaError_t launch_prec_inversion (int operatorID, double *prec, const int *diagIndex,
                                   const int *checkBounds, int nbNodes, cudaStream_t stream);
cudaError_t launch_prec_inversion_list (int operatorID, double *prec, const int *diagIndex,
                                        const int *checkBounds, int nbNodes, const int *nodes,
                                        int count, cudaStream_t stream);
cudaError_t launch_halo_pack (double *sendBuf, const double *prec, const int *intfNodes, int dim,
                              int nbIntfNodes, cudaStream_t stream);
cudaError_t launch_halo_add (double *prec, const double *recvBuf, const int *uniqNodes,
                             const int *slotIndex, const int *slots, int dim, int nbUniq,
                             cudaStream_t stream);

// Peer-to-peer halo exchange fused with the interface sum and inversion (kernels_halo_p2p.cu); runs next to the RING
// assembly kernel of the same iteration.  All pointers are device pointers of THIS process; peerRecv / peerFlag
// point into the neighbours' windows (cudaIpc mappings or, inside one process, plain device pointers).
struct HaloP2PState {                 // device memory, zeroed once
    unsigned intfDone;                // + 1 per (write-out warp, interface tile) of the assembly kernel
    unsigned packDone, finished;      // CTAs of the exchange kernel past their pack / past their sum
    unsigned epoch;                   // exchanges completed so far
};
struct HaloP2PArgs {
    HaloP2PState *state;
    unsigned intfTarget;              // value of state->intfDone when every interface tile is written
    unsigned *status;                 // mapped host word: bit 0 = the assembly kernel's signal timed out, bit 1 = a neighbour's flag
    double *prec;
    int nbIntf, nbUniq, nbNodes;
    const int *intfIndex, *intfNodes; // halo.cc: interface i holds intfNodes[intfIndex[i] .. intfIndex[i+1]) (1-based ids)
    double *const *peerRecv;          // [2 * nbIntf]: where interface i's blocks go in the neighbour's window, per epoch parity
    unsigned *const *peerFlag;        // [nbIntf]: this subdomain's flag in the neighbour's window
    const unsigned *localFlags;       // [nbIntf]: the neighbours' flags in this window
    const double *localRecv[2];       // this window's receive buffers, laid out like bufferRecv (segment i at intfIndex[i] * dim)
    const int *uniqNodes, *slotIndex, *slots, *diagIndex, *checkBounds;
};
cudaError_t launch_halo_p2p (const HaloP2PArgs &args, int operatorDim, int ctas, cudaStream_t stream);

// compute_double_norm (FEM.cc:48-56) of a device array; `partials` holds double_norm_scratch_doubles().
int double_norm_scratch_doubles ();
cudaError_t launch_double_norm (const double *x, int64_t n, double *partials, double *out, cudaStream_t stream);

// GPU layout builders (kernels_topology.cu); every pointer is a device pointer unless noted.
// *badIds = node ids outside [1, nbNodes] (nothing else is written then).
cudaError_t device_node_to_elem (const int *dElemToNode, int nbElem, int nbNodes, int *dIndex, int *dValue,
                                 int *badIds, cudaStream_t stream);
// dRow holds nbNodes+1 ints; *dColOut is cudaMalloc'ed here (the count is not known before).
cudaError_t device_build_csr (const int *dElemToNode, const int *dIndex, const int *dValue, int nbElem,
                              int nbNodes, int *dRow, int **dColOut, int64_t *nbEdges, cudaStream_t stream);
// colorToElem: HOST array of 129 ints.  *nbColors = -1: more than 128 colours; -2: an element
// names a node twice.
cudaError_t device_color_elements (const int *dElemToNode, const int *dIndex, const int *dValue, int nbElem,
                                   int nbNodes, int *dColorPart, int *dColorPerm, int *colorToElem,
                                   int *nbColors, cudaStream_t stream);

// Device copy of a TilePlan (host/tile_plan.h): one blob per tile.
struct DeviceTilePlan {
    const uint8_t *blob = nullptr;
    const uint64_t *tileOffset = nullptr;     // nbTiles + 1
    int nbTiles = 0, nbInterfaceTiles = 0;
    int maxRows = 0, elemStride = 0, maxNodesRef = 0;
    unsigned maxBlobBytes = 0, maxHeadBytes = 0, maxTailBytes = 0;
};

// fusePrec: 0 = values only; 1 = also write prec: the raw diagonal block for interface
// nodes (they still need the halo sum), the masked + inverted block for all others.
size_t tiled_smem_bytes (int operatorID, const DeviceTilePlan &plan, int threads);
// Pipelined variant (one persistent 512-thread CTA per SM, warps specialised by role).
size_t tiled_pipeline_smem_bytes (int operatorID, const DeviceTilePlan &plan);
// Prefetching variant (default when it fits): next tile's head record and coordinates land while the
// current tile is in its off-diagonal pass.  Selected with threads = 0 / 256 and prefetch = true.
size_t tiled_prefetch_smem_bytes (int operatorID, const DeviceTilePlan &plan, int threads);
int tiled_pipeline_threads ();
cudaError_t tiled_configure (int operatorID, size_t smemBytes);
// `ctas` CTAs walk the tiles [firstTile, firstTile + nbTiles) with stride `ctas`.
cudaError_t launch_tiled (int operatorID, const DeviceTilePlan &plan, int firstTile, int nbTiles, int ctas,
                          int threads, size_t smemBytes, const double *coord, double *values,
                          double *prec, const int *checkBounds, int nbNodes, int fusePrec,
                          cudaStream_t stream, bool prefetch = false);

// Device copy of a RingPlan (host/ring_plan.h) and the RING kernel (kernels_ring.cu): same
// persistent-grid contract as launch_tiled.  tileOffset entries carry (head bytes / 16) in their
// top 16 bits.
struct DeviceRingPlan {
    const uint8_t *blob = nullptr;
    const uint64_t *tileOffset = nullptr;     // nbTiles + 1
    int nbTiles = 0, nbInterfaceTiles = 0;
    int maxRows = 0, maxNodes = 0, maxEntries = 0;
    unsigned maxHeadBytes = 0, maxTailBytes = 0;
};
size_t ring_smem_bytes (int operatorID, const DeviceRingPlan &plan);
cudaError_t ring_configure (int operatorID);
bool ring_threads_supported (int threads);          // CTA sizes the RING kernel is instantiated for
// resident CTAs per SM of the RING kernel for this CTA size and dynamic shared memory (occupancy API)
cudaError_t ring_ctas_per_sm (int operatorID, int threads, size_t smemBytes, int *ctas);
cudaError_t launch_ring (int operatorID, const DeviceRingPlan &plan, int firstTile, int nbTiles, int ctas,
                         int threads, size_t smemBytes, const double *coord, double *values, double *prec,
                         int fusePrec, cudaStream_t stream,       // the Dirichlet mask travels in the plan (RingRow::node)
                         unsigned *intfDone = nullptr);           // peer-to-peer halo: counter of finished (write-out warp, interface tile) pairs
// write-out warps per CTA: an interface tile adds that many to *intfDone
int ring_write_out_warps (int operatorID, int threads);

}  // namespace mfb

#endif
