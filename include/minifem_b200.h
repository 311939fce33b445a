/* minifem_b200 — C ABI of the B200-native Mini-FEM assembly path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Each
 * entry point names the reference interface it replaces (paths relative to the
 * Mini-FEM tree).  INTEGRATION.md shows the binding a Mini-FEM maintainer would add.
 *
 * Conventions kept from the reference: elemToNode / nodeToNodeColumn / intfNodes /
 * neighborsList hold 1-based ids; nodeToNodeRow is 0-based; values are row-major
 * operatorDim-sized blocks per CSR entry (operatorDim = 1 lap, 9 ela); checkBounds is
 * component-major [comp*nbNodes + node].
 *
 * Every function returns MFB_OK (0) or a negative MFB_ERR_* code; mfb_last_error()
 * gives the text.  The reference's stage functions return void and exit(EXIT_FAILURE)
 * on fatal errors — the host driver (mini-fem_b200/host/fem_driver.cc) applies that
 * convention on top of these codes.  There is no CPU fallback: without a CUDA device
 * the mfb_ctx_* functions fail with MFB_ERR_CUDA.
 */
#ifndef MINIFEM_B200_H
#define MINIFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_OK            0
#define MFB_ERR_ARG      -1   /* bad argument or inconsistent sizes            */
#define MFB_ERR_CUDA     -2   /* CUDA runtime / no device / kernel failure     */
#define MFB_ERR_COMM     -3   /* NCCL missing or failed                        */
#define MFB_ERR_IO       -4   /* file could not be read / written              */
#define MFB_ERR_COLORS   -5   /* more than 128 colours (coloring.cc:66-69)     */
#define MFB_ERR_STATE    -6   /* call order (e.g. comm not initialised)        */

/* Scatter-add strategies for the assembly stage. */
#define MFB_PATH_TILED    0   /* write-once gather over node tiles staged in shared memory */
#define MFB_PATH_ATOMIC   1   /* memset + native FP64 atomicAdd per contribution            */
#define MFB_PATH_COLOR    2   /* memset + one conflict-free launch per reference colour     */
#define MFB_PATH_RING     3   /* write-once like TILED, but every mesh edge walks the ring of elements
                                 around it from the node coordinates: no coefficient planes in shared
                                 memory, interior edges computed once for both (i,j) and (j,i), the
                                 diagonal block = minus the row sum.  The default of the driver and of
                                 bench.py, and the fastest path: a warp-specialised kernel,
                                 768 threads per CTA, one CTA per SM (mfb_options.threads = 0 or 768; 384 = two CTAs per SM) */

#define MFB_PATH_BLOCKCOLOR 4  /* memset + one launch per BLOCK colour: the elements are cut into spatial blocks, blocks of one
                                 colour share no node, one CTA walks the local colours of its block (locality-blocked
                                 colouring, the idea of the reference's D&C variant, src/assembly.cc:123-244).
                                 Keeps its own element order: mfb_problem.elemToEdge / colorToElem are not used,
                                 mfb_options.tileElems = elements per block (0 = 1024) */

const char *mfb_last_error (void);
const char *mfb_version (void);

/* ------------------------------------------------------------------------------
 * Host-side layout builders (run once per mesh; results are bit-identical to the
 * reference's).  No GPU needed.
 * ------------------------------------------------------------------------------ */

/* DC-lib DC_create_nodeToElem as called at main.cc:247 / coloring.cc:90.
 * index[nbNodes+1], value[4*nbElem]. */
int mfb_node_to_elem (const int *elemToNode, int nbElem, int nbNodes, int *index, int *value);

/* Number of CSR entries create_nodeToNode (matrix.cc:55-91) produces; the reference
 * trusts the nbEdges of the input file header (IO.cc:77). */
int64_t mfb_count_edges (const int *elemToNode, int nbElem, int nbNodes);

/* create_nodeToNode, matrix.cc:55-91.  nodeToNodeColumn must hold mfb_count_edges()
 * ints.  *nbEdgesOut receives nodeToNodeRow[nbNodes]. */
int mfb_create_nodeToNode (const int *elemToNode, int nbElem, int nbNodes,
                           int *nodeToNodeRow, int *nodeToNodeColumn, int *nbEdgesOut);

/* create_elemToEdge, matrix.cc:25-52. */
int mfb_create_elemToEdge (const int *nodeToNodeRow, const int *nodeToNodeColumn,
                           const int *elemToNode, int *elemToEdge, int nbElem);

/* coloring_creation, coloring.cc:84-109 (colorPart + colorToElem[129] + colorPerm), and
 * DC_permute_int_2d_array as applied at main.cc:229. */
int mfb_coloring_creation (const int *elemToNode, int nbElem, int nbNodes, int *colorPart,
                           int *colorToElem, int *colorPerm, int *nbTotalColors);
int mfb_permute_int_2d (int *tab, const int *perm, int nbItem, int dimItem);

/* dqmrd4_ + e_essbcm_ as called at main.cc:343-345 (Fortran: qdmrd4.f, e_cgmelissa.F). */
int mfb_boundary_mask (const int *boundNodesCode, int nbNodes, int *checkBounds,
                       int *nbBoundNodes);

/* compute_double_norm, FEM.cc:48-56. */
double mfb_double_norm (const double *tab, int64_t size);

/* ------------------------------------------------------------------------------
 * Input files (IO.cc) and the synthetic stand-in for the absent data/ tree.
 * ------------------------------------------------------------------------------ */

typedef struct mfb_mesh mfb_mesh;   /* one subdomain's input arrays, owned by the library */

typedef struct {
    int nbElem, nbNodes, nbEdges, nbIntf, nbIntfNodes, nbBoundNodes;
    double *coord;
    int *elemToNode, *neighborsList, *intfIndex, *intfNodes, *boundNodesCode;
    int64_t *globalNode;            /* generator only, else NULL */
} mfb_mesh_view;

/* Structured Kuhn mesh of nx*ny*nz cubes, block `rank` of a px*py*pz partition. */
int mfb_mesh_generate (int nx, int ny, int nz, int px, int py, int pz, int rank,
                       uint64_t seed, mfb_mesh **out);
/* read_input_data / store_input_data_, IO.cc:61-96 / :99-127 (same byte layout). */
int mfb_mesh_read (const char *file, mfb_mesh **out);
int mfb_mesh_write (const mfb_mesh *mesh, const char *file);
int mfb_mesh_get (const mfb_mesh *mesh, mfb_mesh_view *view);
void mfb_mesh_free (mfb_mesh *mesh);
void mfb_choose_blocks (int nx, int ny, int nz, int maxRanks, int *px, int *py, int *pz);
/* read_ref_assembly / store_ref_assembly_, IO.cc:26-39 / :42-58. */
int mfb_checking_write (const char *file, double matrixNorm, double precNorm);
int mfb_checking_read (const char *file, double *matrixNorm, double *precNorm);

/* ------------------------------------------------------------------------------
 * GPU context: the four stages FEM_loop calls every iteration (FEM.cc:177-257).
 * ------------------------------------------------------------------------------ */

typedef struct mfb_ctx mfb_ctx;

/* Everything assembly()/prec_init()/MPI_halo_exchange()/prec_inversion() borrow from
 * main (main.cc:243-246,341-342; IO.cc:82-87).  Arrays are copied to the device at
 * creation; the caller keeps ownership of the host copies. */
typedef struct {
    int operatorID;                 /* 0 = lap (operatorDim 1), 1 = ela (operatorDim 9): main.cc:131-138 */
    int nbElem, nbNodes, nbEdges;
    const double *coord;            /* nbNodes*3 */
    const int *elemToNode;          /* nbElem*4 (colour-sorted already when path = COLOR) */
    const int *nodeToNodeRow;       /* nbNodes+1 */
    const int *nodeToNodeColumn;    /* nbEdges */
    const int *elemToEdge;          /* nbElem*16 or NULL (OPTIMIZED, main.cc:325); built on demand */
    const int *checkBounds;         /* nbNodes*3 or NULL (treated as all zero) */
    const int *colorToElem;         /* nbTotalColors+1 or NULL (globals.h:43) */
    int nbTotalColors;
    int nbBlocks, rank;             /* MPI world size / rank of the reference */
    int nbIntf, nbIntfNodes;
    const int *intfIndex;           /* nbIntf+1 */
    const int *intfNodes;           /* nbIntfNodes */
    const int *neighborsList;       /* >= nbIntf */
} mfb_problem;

typedef struct {
    int path;                       /* MFB_PATH_* */
    int device;                     /* CUDA device ordinal */
    int tileRows;                   /* TILED: max rows per tile (0 = default) */
    int tileElems;                  /* TILED: max elements per tile; RING: max slab slots per tile = CSR entries + row padding (0 = default) */
    int threads;                    /* TILED: threads per CTA (0 = default 256); RING: 0 = default (elasticity 768, Laplacian 1024), 640 / 768 / 896 / 1024 (one CTA per SM) or 384 (two) */
    int useGraph;                   /* capture mfb_ctx_iteration in a CUDA graph */
    int ctas;                       /* TILED: CTAs walking the tiles (0 = default, -1 = one per tile) */
    int bankAware;                  /* TILED: order inside each contribution list: 0 / 1 = chosen against
                                       shared-memory bank conflicts (default), -1 = increasing element id
                                       (the REF build's summation order) */
} mfb_options;

int mfb_ctx_create (const mfb_problem *problem, const mfb_options *options, mfb_ctx **out);
void mfb_ctx_destroy (mfb_ctx *ctx);

/* assembly(), assembly.h:61-63 / assembly.cc:615-720.  Asynchronous on the context's
 * stream; mfb_ctx_sync() or any *_host call waits. */
int mfb_ctx_assembly (mfb_ctx *ctx);
/* assembly_{lap,ela}_seq(userArgs, firstElem, lastElem), assembly.h:44,51: scatter-add of
 * an INCLUSIVE element interval without zeroing first (ATOMIC / COLOR element kernels).  On the COLOR
 * path an interval that spans colours is cut at the colorToElem boundaries (one launch per piece). */
int mfb_ctx_assembly_interval (mfb_ctx *ctx, int firstElem, int lastElem);
int mfb_ctx_zero_values (mfb_ctx *ctx);
/* prec_init(), preconditioner.h:31-32 / preconditioner.cc:52-87. */
int mfb_ctx_prec_init (mfb_ctx *ctx);
/* MPI_halo_exchange(), halo.h:41-43 / halo.cc:39-122.  Needs mfb_ctx_comm_init when
 * nbBlocks > 1; no-op when nbBlocks < 2 (halo.cc:44). */
int mfb_ctx_halo_exchange (mfb_ctx *ctx);
/* prec_inversion(), preconditioner.h:27-28 / preconditioner.cc:25-49 with
 * ela_invert_prec (elasclpr.f:2-56). */
int mfb_ctx_prec_inversion (mfb_ctx *ctx);
/* One whole iteration of FEM_loop (FEM.cc:183-233): the four stages, fused where the
 * path allows (TILED: assembly + prec_init + inversion of non-interface nodes in one
 * kernel; halo + interface inversion overlapped on a second stream). */
int mfb_ctx_iteration (mfb_ctx *ctx);
int mfb_ctx_sync (mfb_ctx *ctx);

/* Results back to caller-owned host buffers (check_results needs them, main.cc:376).
 * Either pointer may be NULL. */
int mfb_ctx_download (mfb_ctx *ctx, double *nodeToNodeValue, double *prec);
/* Refresh the device coordinates (moving mesh / e2e measurement). */
int mfb_ctx_upload_coord (mfb_ctx *ctx, const double *coord);
/* Host-buffer step: upload coord, run one iteration, download values and prec, and
 * wait.  Pointers should be pinned (mfb_host_alloc) for full PCIe speed. */
int mfb_ctx_iteration_host (mfb_ctx *ctx, const double *coord, double *nodeToNodeValue,
                            double *prec);

/* check_results' two norms (FEM.cc:68-76: compute_double_norm of nodeToNodeValue and of prec)
 * computed on the device; only 16 bytes travel.  The summation order differs from the
 * reference's serial loop (relative difference ~1e-15). */
int mfb_ctx_norms (mfb_ctx *ctx, double *matrixNorm, double *precNorm);
/* Device-resident step for callers that consume the matrix on the GPU: upload coord, run one
 * iteration, return the two norms.  norms[0] = matrix, norms[1] = prec. */
int mfb_ctx_iteration_norms_host (mfb_ctx *ctx, const double *coord, double norms[2]);

/* Device pointers of the results (for callers that keep working on the GPU). */
int mfb_ctx_device_ptrs (mfb_ctx *ctx, void **values, void **prec);
/* The stream the stages are launched on (cudaStream_t as void*). */
int mfb_ctx_stream (mfb_ctx *ctx, void **stream);

/* Device time of the last completed call of each stage, in ms:
 * [0] assembly [1] prec_init [2] halo [3] prec_inversion [4] fused iteration. */
int mfb_ctx_stage_ms (mfb_ctx *ctx, float ms[5]);
/* Kernel launches issued by this context so far. */
int64_t mfb_ctx_launch_count (mfb_ctx *ctx);
/* Bytes of the device-resident plan and mesh (reporting). */
int mfb_ctx_device_bytes (mfb_ctx *ctx, int64_t *meshBytes, int64_t *planBytes);
/* Tile statistics of the TILED plan: [0] tiles [1] tile elements (with duplicates)
 * [2] contributions [3] max rows [4] max elems [5] shared memory bytes per CTA
 * [6] padded lane-steps of the off-diagonal pass [7] largest tile record in bytes.
 * RING plan: [1] jobs (mesh edges) [2] ring steps (element visits) [4] max tile-local nodes
 * [6] padded lane-steps of the job phase; the others as above. */
int mfb_ctx_plan_stats (mfb_ctx *ctx, int64_t stats[8]);

/* Multi-GPU: one context per process / GPU, NCCL over NVLink for the interface sum. */
#define MFB_COMM_ID_BYTES 128
int mfb_comm_unique_id (unsigned char id[MFB_COMM_ID_BYTES]);              /* rank 0, then broadcast */
int mfb_ctx_comm_init (mfb_ctx *ctx, const unsigned char id[MFB_COMM_ID_BYTES]);

/* Peer-to-peer exchange for the fused RING iteration (csrc/kernels_halo_p2p.cu), replacing the NCCL send / recv
 * group by stores into the neighbours' receive windows over NVLink plus an epoch flag — the write + notify scheme
 * of the reference's GASPI variant, src/halo.cc:127-226 (segments created in FEM.cc / main.cc under -DGASPI).
 * Every subdomain publishes a card (its window's cudaIpc handle, neighbour list and interface offsets); the caller
 * gathers the cards of all nbBlocks subdomains in rank order (torch.distributed all_gather, files, MPI ...) and
 * hands them to connect.  Subdomains living in one process are connected through plain device pointers.  All
 * subdomains must switch together: call mfb_ctx_p2p_enable (ctx, 0) everywhere if connect failed anywhere (NCCL
 * then carries the exchange, mfb_ctx_comm_init).  A timed-out wait surfaces as MFB_ERR_COMM from mfb_ctx_sync. */
#define MFB_P2P_CARD_BYTES 1024
int mfb_ctx_p2p_card (mfb_ctx *ctx, unsigned char card[MFB_P2P_CARD_BYTES]);
int mfb_ctx_p2p_connect (mfb_ctx *ctx, const unsigned char *cards /* nbBlocks x MFB_P2P_CARD_BYTES */);
int mfb_ctx_p2p_enable (mfb_ctx *ctx, int on);
int mfb_ctx_p2p_active (mfb_ctx *ctx);

/* The two halves of the halo exchange with the transport left to the caller (tests, or a
 * host MPI): pack = halo.cc:77-80 into sendBuf[nbIntfNodes*operatorDim]; add = halo.cc:113-116
 * from recvBuf laid out like bufferRecv (segment i at intfIndex[i]*operatorDim). */
int mfb_ctx_halo_pack_host (mfb_ctx *ctx, double *sendBuf);
int mfb_ctx_halo_add_host (mfb_ctx *ctx, const double *recvBuf);
/* TILED only: the fused kernel of mfb_ctx_iteration without the exchange — values, plus
 * prec holding the inverted block of every non-interface node and the raw diagonal block of
 * every interface node (assembly + prec_init + the interior part of prec_inversion). */
int mfb_ctx_assembly_fused (mfb_ctx *ctx);
/* Inversion restricted to the interface nodes (what the fused iteration leaves undone
 * until the halo sum has arrived). */
int mfb_ctx_prec_inversion_interface (mfb_ctx *ctx);

/* `steps` back-to-back fused iterations bracketed by CUDA events on the context's stream;
 * *ms = total device time.  Used by bench.py for the kernel-time roofline. */
int mfb_ctx_run_timed (mfb_ctx *ctx, int steps, float *ms);

/* Builds the TILED plan on the host and verifies its invariants without a GPU: every
 * (element, j, k) contribution of the reference's double loop (assembly.cc:382-412) appears
 * exactly once, on the CSR entry elemToEdge names.  stats: [0] tiles [1] tile elements
 * [2] contributions [3] max rows [4] max elems [5] plan bytes. */
int mfb_tile_plan_selfcheck (const mfb_problem *problem, int tileRows, int tileElems,
                             int64_t stats[6]);

/* Same for the RING plan (host/ring_plan.h): rows tile the CSR, every off-diagonal entry is written
 * by exactly one job, the consecutive nodes of a job's chains name distinct elements around its edge,
 * every element's 12 off-diagonal node pairs are covered exactly once.  tileRows / tileEntries 0 =
 * defaults.  stats: [0] tiles [1] jobs [2] jobs that also write the transposed block [3] ring steps
 * [4] padded lane-steps [5] chain breaks [6] modelled gather wavefronts [7] conflict-free count
 * [8] modelled slab-store wavefronts per block component [9] conflict-free count [10] plan bytes
 * [11] interface tiles. */
int mfb_ring_plan_selfcheck (const mfb_problem *problem, int tileRows, int tileEntries,
                             int64_t stats[12]);

/* The layout of MFB_PATH_BLOCKCOLOR, built on the host (no GPU needed; the context builds the same internally):
 * elemOrder[nbElem] = element ids sorted by (block colour, block, local colour, id); blocks
 * [launchStart[c], launchStart[c+1]) have block colour c and share no node; local colour q of block b is the
 * positions [localStart[localIndex[b] + q], localStart[localIndex[b] + q + 1]) of elemOrder, its elements share no
 * node.  Capacities: launchStart 65, localIndex nbElem + 2, localStart 2 * nbElem + 2.  counts = {blocks, block
 * colours, max local colours of a block, entries of localStart}.  blockElems 0 = 1024. */
int mfb_block_coloring (const int *elemToNode, int nbElem, int nbNodes, const double *coord, int blockElems,
                        int *elemOrder, int *launchStart, int *localIndex, int *localStart, int counts[4]);

/* Pinned host memory for the *_host calls. */
int mfb_host_alloc (void **ptr, int64_t bytes);
void mfb_host_free (void *ptr);

int mfb_device_count (void);

/* ------------------------------------------------------------------------------
 * The layout builders on the GPU.  HOST pointers in and
 * out, results bit-identical to the host builders above and therefore to the reference:
 * create_nodeToNode matrix.cc:55-91 (columns in first-seen order), create_elemToEdge
 * matrix.cc:25-52, coloring_creation coloring.cc:46-109 (greedy first-fit in element order,
 * evaluated front by front) + the stable permutation of coloring.cc:107.
 * ------------------------------------------------------------------------------ */

/* nodeToNodeColumn may be NULL (count only).  *nbEdgesOut always receives the entry count;
 * MFB_ERR_ARG if columnCapacity is smaller. */
int mfb_device_create_nodeToNode (const int *elemToNode, int nbElem, int nbNodes,
                                  int *nodeToNodeRow, int *nodeToNodeColumn,
                                  int64_t columnCapacity, int *nbEdgesOut, int device);
int mfb_device_create_elemToEdge (const int *nodeToNodeRow, const int *nodeToNodeColumn,
                                  const int *elemToNode, int *elemToEdge, int nbElem,
                                  int nbNodes, int device);
/* colorPart[nbElem], colorToElem[129], colorPerm[nbElem] as mfb_coloring_creation. */
int mfb_device_coloring_creation (const int *elemToNode, int nbElem, int nbNodes,
                                  int *colorPart, int *colorToElem, int *colorPerm,
                                  int *nbTotalColors, int device);

#ifdef __cplusplus
}
#endif
#endif
