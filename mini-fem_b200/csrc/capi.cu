// C ABI of the B200 assembly path (include/minifem_b200.h): host-side layout builders,
// mesh files, and the GPU context that runs the four stages of FEM_loop
// (src/FEM.cc:177-257).  No CPU fallback: every mfb_ctx_* call needs a CUDA device.
#include "../../include/minifem_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/mesh_data.h"
#include "../host/mesh_topology.h"
#include "../host/tile_plan.h"
#include "kernels.cuh"

using namespace mfb;

// ------------------------------------------------------------------------ errors

static thread_local std::string g_lastError;

static int fail (int code, const std::string &msg)
{
    g_lastError = msg;
    return code;
}

#define MFB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            return fail (MFB_ERR_CUDA, std::string (#call) + ": " + cudaGetErrorString (e_)); \
        }                                                                                \
    } while (0)

extern "C" const char *mfb_last_error (void) { return g_lastError.c_str (); }
extern "C" const char *mfb_version (void) { return "minifem_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ host builders

extern "C" int mfb_node_to_elem (const int *elemToNode, int nbElem, int nbNodes, int *index, int *value)
{
    if (!elemToNode || !index || !value || nbElem < 0 || nbNodes < 0) return fail (MFB_ERR_ARG, "mfb_node_to_elem: bad argument");
    node_to_elem (elemToNode, nbElem, nbNodes, index, value);
    return MFB_OK;
}

extern "C" int64_t mfb_count_edges (const int *elemToNode, int nbElem, int nbNodes)
{
    if ((!elemToNode && nbElem > 0) || nbElem < 0 || nbNodes < 0) return fail (MFB_ERR_ARG, "mfb_count_edges: bad argument");
    return count_csr_entries (elemToNode, nbElem, nbNodes);
}

extern "C" int mfb_create_nodeToNode (const int *elemToNode, int nbElem, int nbNodes,
                                      int *nodeToNodeRow, int *nodeToNodeColumn, int *nbEdgesOut)
{
    if (!nodeToNodeRow || (!nodeToNodeColumn && nbElem > 0) || nbElem < 0 || nbNodes < 0) {
        return fail (MFB_ERR_ARG, "mfb_create_nodeToNode: bad argument");
    }
    int64_t n = build_csr (elemToNode, nbElem, nbNodes, nodeToNodeRow, nodeToNodeColumn);
    if (n > INT32_MAX) return fail (MFB_ERR_ARG, "mfb_create_nodeToNode: more than 2^31 entries");
    if (nbEdgesOut) *nbEdgesOut = (int)n;
    return MFB_OK;
}

extern "C" int mfb_create_elemToEdge (const int *nodeToNodeRow, const int *nodeToNodeColumn,
                                      const int *elemToNode, int *elemToEdge, int nbElem)
{
    if (!nodeToNodeRow || !nodeToNodeColumn || !elemToNode || !elemToEdge || nbElem < 0) {
        return fail (MFB_ERR_ARG, "mfb_create_elemToEdge: bad argument");
    }
    if (build_elem_to_edge (nodeToNodeRow, nodeToNodeColumn, elemToNode, elemToEdge, nbElem) != 0) {
        return fail (MFB_ERR_ARG, "mfb_create_elemToEdge: a node pair is missing from the CSR");
    }
    return MFB_OK;
}

extern "C" int mfb_coloring_creation (const int *elemToNode, int nbElem, int nbNodes, int *colorPart,
                                      int *colorToElem, int *colorPerm, int *nbTotalColors)
{
    if (!elemToNode || !colorPart || !colorToElem || !colorPerm || !nbTotalColors || nbElem < 0) {
        return fail (MFB_ERR_ARG, "mfb_coloring_creation: bad argument");
    }
    int n = color_elements (elemToNode, nbElem, nbNodes, colorPart, colorToElem, colorPerm);
    if (n < 0) return fail (MFB_ERR_COLORS, "Error: Not enough colors.");
    *nbTotalColors = n;
    return MFB_OK;
}

extern "C" int mfb_permute_int_2d (int *tab, const int *perm, int nbItem, int dimItem)
{
    if (!tab || !perm || nbItem < 0 || dimItem < 1) return fail (MFB_ERR_ARG, "mfb_permute_int_2d: bad argument");
    permute_rows (tab, perm, nbItem, dimItem);
    return MFB_OK;
}

extern "C" int mfb_boundary_mask (const int *boundNodesCode, int nbNodes, int *checkBounds, int *nbBoundNodes)
{
    if (!boundNodesCode || !checkBounds || nbNodes < 0) return fail (MFB_ERR_ARG, "mfb_boundary_mask: bad argument");
    int n = boundary_mask (boundNodesCode, nbNodes, checkBounds);
    if (nbBoundNodes) *nbBoundNodes = n;
    return MFB_OK;
}

extern "C" double mfb_double_norm (const double *tab, int64_t size) { return double_norm (tab, size); }

// ------------------------------------------------------------------------ meshes

struct mfb_mesh { SubMesh m; };

extern "C" int mfb_mesh_generate (int nx, int ny, int nz, int px, int py, int pz, int rank,
                                  uint64_t seed, mfb_mesh **out)
{
    if (!out) return fail (MFB_ERR_ARG, "mfb_mesh_generate: out is NULL");
    mfb_mesh *mesh = new mfb_mesh ();
    if (generate_block (nx, ny, nz, px, py, pz, rank, seed, mesh->m) != 0) {
        delete mesh;
        return fail (MFB_ERR_ARG, "mfb_mesh_generate: invalid grid / partition (or counts exceed 32-bit indices)");
    }
    *out = mesh;
    return MFB_OK;
}

extern "C" int mfb_mesh_read (const char *file, mfb_mesh **out)
{
    if (!file || !out) return fail (MFB_ERR_ARG, "mfb_mesh_read: bad argument");
    mfb_mesh *mesh = new mfb_mesh ();
    if (read_input (file, mesh->m) != 0) {
        delete mesh;
        return fail (MFB_ERR_IO, std::string ("Error: cannot read input data: ") + file);
    }
    *out = mesh;
    return MFB_OK;
}

extern "C" int mfb_mesh_write (const mfb_mesh *mesh, const char *file)
{
    if (!mesh || !file) return fail (MFB_ERR_ARG, "mfb_mesh_write: bad argument");
    if (write_input (file, mesh->m) != 0) return fail (MFB_ERR_IO, std::string ("Error: cannot store input data: ") + file);
    return MFB_OK;
}

extern "C" int mfb_mesh_get (const mfb_mesh *mesh, mfb_mesh_view *v)
{
    if (!mesh || !v) return fail (MFB_ERR_ARG, "mfb_mesh_get: bad argument");
    SubMesh &m = const_cast<SubMesh&> (mesh->m);
    v->nbElem = m.nbElem; v->nbNodes = m.nbNodes; v->nbEdges = m.nbEdges; v->nbIntf = m.nbIntf;
    v->nbIntfNodes = m.nbIntfNodes; v->nbBoundNodes = m.nbBoundNodes;
    v->coord = m.coord.data (); v->elemToNode = m.elemToNode.data ();
    v->neighborsList = m.neighborsList.data (); v->intfIndex = m.intfIndex.data ();
    v->intfNodes = m.intfNodes.data (); v->boundNodesCode = m.boundNodesCode.data ();
    v->globalNode = m.globalNode.empty () ? nullptr : m.globalNode.data ();
    return MFB_OK;
}

extern "C" void mfb_mesh_free (mfb_mesh *mesh) { delete mesh; }

extern "C" void mfb_choose_blocks (int nx, int ny, int nz, int maxRanks, int *px, int *py, int *pz)
{
    int a = 1, b = 1, c = 1;
    choose_blocks (nx, ny, nz, std::max (maxRanks, 1), a, b, c);
    if (px) *px = a;
    if (py) *py = b;
    if (pz) *pz = c;
}

extern "C" int mfb_checking_write (const char *file, double matrixNorm, double precNorm)
{
    if (!file || write_checking (file, matrixNorm, precNorm) != 0) return fail (MFB_ERR_IO, "Error: cannot store reference checking.");
    return MFB_OK;
}

extern "C" int mfb_checking_read (const char *file, double *matrixNorm, double *precNorm)
{
    double a = 0, b = 0;
    if (!file || read_checking (file, a, b) != 0) {
        return fail (MFB_ERR_IO, std::string ("Error: cannot read reference checking: ") + (file ? file : "(null)") + ".");
    }
    if (matrixNorm) *matrixNorm = a;
    if (precNorm) *precNorm = b;
    return MFB_OK;
}

// -------------------------------------------------------------------------- NCCL
// Loaded lazily with dlopen so that the library has no link-time NCCL dependency and,
// inside a PyTorch process, shares the libnccl.so.2 torch already mapped.

namespace {

struct NcclId { char internal[MFB_COMM_ID_BYTES]; };
typedef void *NcclComm;

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId) (NcclId*) = nullptr;
    int (*CommInitRank) (NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy) (NcclComm) = nullptr;
    int (*Send) (const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv) (void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart) () = nullptr;
    int (*GroupEnd) () = nullptr;
    const char *(*GetErrorString) (int) = nullptr;
};

const int kNcclFloat64 = 8;   // ncclDouble

NcclApi *nccl_api (std::string &why)
{
    static NcclApi api;
    static bool tried = false;
    static std::string error;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.lib = dlopen (n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { error = std::string ("NCCL not found: ") + dlerror (); }
        else {
            bool ok = true;
            auto sym = [&] (const char *name) { void *p = dlsym (api.lib, name); if (!p) ok = false; return p; };
            api.GetUniqueId = (int (*) (NcclId*))sym ("ncclGetUniqueId");
            api.CommInitRank = (int (*) (NcclComm*, int, NcclId, int))sym ("ncclCommInitRank");
            api.CommDestroy = (int (*) (NcclComm))sym ("ncclCommDestroy");
            api.Send = (int (*) (const void*, size_t, int, int, NcclComm, cudaStream_t))sym ("ncclSend");
            api.Recv = (int (*) (void*, size_t, int, int, NcclComm, cudaStream_t))sym ("ncclRecv");
            api.GroupStart = (int (*) ())sym ("ncclGroupStart");
            api.GroupEnd = (int (*) ())sym ("ncclGroupEnd");
            api.GetErrorString = (const char *(*) (int))sym ("ncclGetErrorString");
            if (!ok) { error = "NCCL library lacks a required symbol"; api.lib = nullptr; }
        }
    }
    why = error;
    return api.lib ? &api : nullptr;
}

template <class T>
cudaError_t upload (T **dst, const T *src, size_t count, int64_t &bytes)
{
    *dst = nullptr;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc ((void**)dst, count * sizeof (T));
    if (e != cudaSuccess) return e;
    bytes += (int64_t)(count * sizeof (T));
    return src ? cudaMemcpy (*dst, src, count * sizeof (T), cudaMemcpyHostToDevice) : cudaSuccess;
}

}  // namespace

// ----------------------------------------------------------------------- context

namespace {
constexpr size_t kP2PFlagBytes = 256;                 // 64 flags of 4 bytes
constexpr int kP2PMaxIntf = 64;
inline size_t p2p_align (size_t bytes) { return (bytes + 255) & ~(size_t)255; }
}

struct mfb_ctx {
    int path = MFB_PATH_TILED, device = 0, threads = 256, useGraph = 0;
    int operatorID = 0, operatorDim = 1;
    int nbElem = 0, nbNodes = 0, nbEdges = 0, nbBlocks = 1, rank = 0;
    int nbIntf = 0, nbIntfNodes = 0, nbUniqIntf = 0, nbTotalColors = 0;
    std::vector<int> colorToElem, intfIndex, neighbors;
    std::vector<int> blockLaunchStart;      // MFB_PATH_BLOCKCOLOR: blocks per block colour (host/mesh_topology.h)
    int *dLocalIndex = nullptr, *dLocalStart = nullptr;
    int blockStats[3] = {0, 0, 0};          // blocks, block colours, max local colours

    cudaStream_t stream = nullptr, commStream = nullptr;
    cudaEvent_t evStart[5] = {}, evStop[5] = {}, evIntfDone = nullptr, evCommDone = nullptr;
    bool stageRan[5] = {};

    double *dCoord = nullptr, *dValues = nullptr, *dPrec = nullptr, *dSend = nullptr, *dRecv = nullptr;
    int *dElemToNode = nullptr, *dRow = nullptr, *dCol = nullptr, *dElemToEdge = nullptr,
        *dCheckBounds = nullptr, *dDiagIndex = nullptr, *dIntfNodes = nullptr, *dUniqNodes = nullptr,
        *dSlotIndex = nullptr, *dSlots = nullptr;
    std::vector<void*> planAllocs;
    DeviceTilePlan plan;
    TilePlan hostPlanStats;       // counters only (vectors released after upload)
    size_t tiledSmem = 0;
    int tiledCtas = 1;
    bool tiledPrefetch = false;
    bool ring = false;              // MFB_PATH_RING: same iteration structure as TILED, other plan + kernel
    DeviceRingPlan ringPlan;
    RingPlan ringStats;             // counters only
    double *dNorm = nullptr;        // [partials][2 results], allocated by the first mfb_ctx_norms
    int haloCoresident = 0;         // multi-GPU overlap scheme, see do_iteration
    int haloReserveCtas = 8;       // RING: CTAs the persistent interior grid leaves free for the kernels of the exchange (4 SMs)
    int eagerIterations = 0;       // fused iterations run before the multi-GPU graph is captured
    bool multiGraph = false;       // MFB_MULTI_GPU_GRAPH=1: capture the two-stream iteration with its NCCL group (opt-in)
    int interiorTilesPerCta = 4;   // measured at N=2: 1 -> 0.71 ms, 4 -> 0.645, 8 -> 0.647 (kernel alone 0.636 in that build)
    int64_t meshBytes = 0, planBytes = 0, launches = 0, graphLaunches = 0;

    NcclComm comm = nullptr;
    cudaGraphExec_t graphExec = nullptr;

    // peer-to-peer halo (kernels_halo_p2p.cu): this context's window = [64 flags][recv parity 0][recv parity 1]
    unsigned char *p2pWindow = nullptr;
    HaloP2PState *p2pState = nullptr;
    unsigned *p2pStatusHost = nullptr, *p2pStatusDev = nullptr;      // one mapped word
    double **dPeerRecv = nullptr;
    unsigned **dPeerFlag = nullptr;
    int *dIntfIndex = nullptr;
    std::vector<void*> p2pOpened;  // cudaIpcOpenMemHandle mappings to close
    bool p2pReady = false;
    int p2pCtas = 2, p2pReserveCtas = 2;
    int deviceCtas = 148;          // RING CTAs the device holds at once
};

namespace {

template <class T>
cudaError_t put_plan (mfb_ctx *c, const T **dst, const std::vector<T> &vec)
{
    T *d = nullptr;
    cudaError_t e = upload (&d, vec.data (), vec.size (), c->planBytes);
    if (e == cudaSuccess) { *dst = d; c->planAllocs.push_back ((void*)d); }
    return e;
}

int build_device_plan (mfb_ctx *c, const mfb_problem *p, const mfb_options *o)
{
    std::vector<uint8_t> isIntf;
    if (p->nbBlocks > 1 && p->nbIntfNodes > 0) {
        isIntf.assign ((size_t)p->nbNodes, 0);
        for (int j = 0; j < p->nbIntfNodes; j++) isIntf[p->intfNodes[j] - 1] = 1;
    }
    TilePlanLimits lim;
    if (o && o->tileRows > 0) lim.maxRows = o->tileRows;
    if (o && o->tileElems > 0) lim.maxElems = o->tileElems;
    lim.maxNodesRef = std::min (65535, std::max (lim.maxElems, 64));   // 24 B of staging per referenced node
    lim.maxEntries = 65535;
    lim.bankAware = !(o && o->bankAware < 0);
    lim.laplacian = p->operatorID == 0;
    TilePlan hp;
    std::string err;
    if (build_tile_plan (p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn,
                         p->coord, isIntf.empty () ? nullptr : isIntf.data (), lim, hp, err) != 0) {
        return fail (MFB_ERR_ARG, "tile plan: " + err);
    }
    if (hp.blob.empty ()) hp.blob.resize (16);
    MFB_CUDA (put_plan (c, &c->plan.blob, hp.blob));
    {   // device table: byte offset in the low 48 bits, (head bytes / 16) above
        std::vector<uint64_t> packed (hp.tileOffset);
        for (int t = 0; t < hp.nbTiles; t++) packed[t] |= (uint64_t)(hp.header (t)->offEntryRow >> 4) << 48;
        MFB_CUDA (put_plan (c, &c->plan.tileOffset, packed));
    }
    c->plan.nbTiles = hp.nbTiles; c->plan.nbInterfaceTiles = hp.nbInterfaceTiles;
    c->plan.maxRows = std::max (hp.maxRows, 1); c->plan.elemStride = hp.elemStride;
    c->plan.maxNodesRef = std::max (hp.maxNodesRef, 4); c->plan.maxBlobBytes = std::max (hp.maxBlobBytes, 16u);
    c->plan.maxHeadBytes = std::max (hp.maxHeadBytes, 16u); c->plan.maxTailBytes = std::max (hp.maxTailBytes, 16u);
    c->hostPlanStats.nbTiles = hp.nbTiles; c->hostPlanStats.nbTileElems = hp.nbTileElems;
    c->hostPlanStats.nbContributions = hp.nbContributions; c->hostPlanStats.maxRows = hp.maxRows;
    c->hostPlanStats.maxElems = hp.maxElems; c->hostPlanStats.nbPaddedSteps = hp.nbPaddedSteps;
    c->hostPlanStats.maxBlobBytes = hp.maxBlobBytes;
    const bool pipelined = c->threads == tiled_pipeline_threads ();
    // kernel variant: 0 / default = prefetching kernel, MFB_TILED_VARIANT=plain selects the simpler one
    const char *variant = getenv ("MFB_TILED_VARIANT");
    c->tiledPrefetch = !pipelined && c->threads == 256 && !(variant && std::string (variant) == "plain");
    if (const char *v = getenv ("MFB_HALO_OVERLAP")) c->haloCoresident = std::string (v) == "coresident";
    if (const char *v = getenv ("MFB_INTERIOR_TILES_PER_CTA")) c->interiorTilesPerCta = std::max (atoi (v), 1);
    c->tiledSmem = pipelined ? tiled_pipeline_smem_bytes (c->operatorID, c->plan)
                 : c->tiledPrefetch ? tiled_prefetch_smem_bytes (c->operatorID, c->plan, c->threads)
                                    : tiled_smem_bytes (c->operatorID, c->plan, c->threads);
    if (c->tiledSmem > 227 * 1024) return fail (MFB_ERR_ARG, "tile plan needs more than 227 KB of shared memory per CTA; lower tileElems");
    MFB_CUDA (tiled_configure (c->operatorID, c->tiledSmem));
    // persistent grid: as many CTAs as fit on the device at once, each walking tiles with that stride
    cudaDeviceProp prop;
    MFB_CUDA (cudaGetDeviceProperties (&prop, c->device));
    const int perSM = pipelined ? 1 : std::max (1, std::min ((int)(prop.sharedMemPerMultiprocessor / (c->tiledSmem + 1024)), 2048 / c->threads));
    c->tiledCtas = (o && o->ctas > 0) ? o->ctas : (o && o->ctas == -1) ? (1 << 30) : prop.multiProcessorCount * perSM;
    return MFB_OK;
}

int build_device_ring_plan (mfb_ctx *c, const mfb_problem *p, const mfb_options *o)
{
    std::vector<uint8_t> isIntf;
    if (p->nbBlocks > 1 && p->nbIntfNodes > 0) {
        isIntf.assign ((size_t)p->nbNodes, 0);
        for (int j = 0; j < p->nbIntfNodes; j++) isIntf[p->intfNodes[j] - 1] = 1;
    }
    // two slabs per CTA: 30 rows / 510 slots keep two 384-thread CTAs per SM, 64 rows / 1100 slots fit the
    // single 768-thread CTA
    RingPlanLimits lim;
    lim.maxRows = c->threads >= 640 ? 64 : 30;
    lim.maxEntries = c->threads >= 640 ? 1100 : 510;
    if (o && o->tileRows > 0) lim.maxRows = o->tileRows;
    if (o && o->tileElems > 0) lim.maxEntries = o->tileElems;          // RING: tileElems caps the slab slots of a tile
    lim.bankAware = !(o && o->bankAware < 0);
    // experiment knobs: MFB_RING_CUT=morton (tiles = runs of the Morton curve, like TILED),
    // MFB_RING_REFINE=n (renumber-and-rotate rounds of the bank-aware numbering), MFB_RING_SWEEPS=n
    if (const char *v = getenv ("MFB_RING_CUT")) lim.bisection = std::string (v) != "morton";
    if (const char *v = getenv ("MFB_RING_REFINE")) lim.refinePasses = std::max (atoi (v), 0);
    if (const char *v = getenv ("MFB_RING_SWEEPS")) lim.rotationSweeps = std::max (atoi (v), 0);
    if (const char *v = getenv ("MFB_RING_MAXJOBS")) lim.maxJobs = std::max (atoi (v), 32);
    RingPlan hp;
    std::string err;
    if (build_ring_plan (p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn, p->coord,
                         isIntf.empty () ? nullptr : isIntf.data (), p->checkBounds, lim, hp, err) != 0) {
        return fail (MFB_ERR_ARG, "ring plan: " + err);
    }
    if (hp.blob.empty ()) hp.blob.resize (16);
    MFB_CUDA (put_plan (c, &c->ringPlan.blob, hp.blob));
    {   // device table: byte offset in the low 48 bits, (head bytes / 16) above
        std::vector<uint64_t> packed (hp.tileOffset);
        for (int t = 0; t < hp.nbTiles; t++) packed[t] |= (uint64_t)(hp.header (t)->headBytes >> 4) << 48;
        MFB_CUDA (put_plan (c, &c->ringPlan.tileOffset, packed));
    }
    c->ringPlan.nbTiles = hp.nbTiles; c->ringPlan.nbInterfaceTiles = hp.nbInterfaceTiles;
    c->ringPlan.maxRows = std::max (hp.maxRows, 1); c->ringPlan.maxNodes = std::max (hp.maxNodes, 4);
    c->ringPlan.maxEntries = std::max (hp.maxEntries, 1);
    c->ringPlan.maxHeadBytes = std::max (hp.maxHeadBytes, 16u); c->ringPlan.maxTailBytes = std::max (hp.maxTailBytes, 16u);
    c->plan.nbTiles = hp.nbTiles; c->plan.nbInterfaceTiles = hp.nbInterfaceTiles;     // do_iteration splits on these
    std::vector<uint64_t> ().swap (hp.tileOffset);
    std::vector<uint8_t> ().swap (hp.blob);
    c->ringStats = hp;
    if (const char *v = getenv ("MFB_HALO_OVERLAP")) c->haloCoresident = std::string (v) == "coresident";
    if (const char *v = getenv ("MFB_INTERIOR_TILES_PER_CTA")) c->interiorTilesPerCta = std::max (atoi (v), 1);
    c->haloReserveCtas = c->threads >= 640 ? 4 : 8;            // four SMs either way
    if (const char *v = getenv ("MFB_HALO_RESERVE_CTAS")) c->haloReserveCtas = std::max (atoi (v), 0);
    if (const char *v = getenv ("MFB_MULTI_GPU_GRAPH")) c->multiGraph = atoi (v) != 0;
    c->tiledSmem = ring_smem_bytes (c->operatorID, c->ringPlan);
    if (c->tiledSmem > 227 * 1024) return fail (MFB_ERR_ARG, "ring plan needs more than 227 KB of shared memory per CTA; lower tileRows / tileElems");
    MFB_CUDA (ring_configure (c->operatorID));
    cudaDeviceProp prop;
    MFB_CUDA (cudaGetDeviceProperties (&prop, c->device));
    int perSM = 1;         // registers limit the 384-thread kernel to two CTAs per SM whatever the shared memory allows
    MFB_CUDA (ring_ctas_per_sm (c->operatorID, c->threads, c->tiledSmem, &perSM));
    perSM = std::max (perSM, 1);
    c->tiledCtas = (o && o->ctas > 0) ? o->ctas : (o && o->ctas == -1) ? (1 << 30) : prop.multiProcessorCount * perSM;
    c->deviceCtas = prop.multiProcessorCount * perSM;
    return MFB_OK;
}

// The write-once kernel of the context (TILED or RING) over the tiles [firstTile, firstTile + nbTiles).
cudaError_t launch_write_once (mfb_ctx *c, int firstTile, int nbTiles, int ctas, int fusePrec, cudaStream_t stream)
{
    if (c->ring) {
        return launch_ring (c->operatorID, c->ringPlan, firstTile, nbTiles, ctas, c->threads, c->tiledSmem, c->dCoord,
                            c->dValues, c->dPrec, fusePrec, stream);
    }
    return launch_tiled (c->operatorID, c->plan, firstTile, nbTiles, ctas, c->threads, c->tiledSmem, c->dCoord,
                         c->dValues, c->dPrec, c->dCheckBounds, c->nbNodes, fusePrec, stream, c->tiledPrefetch);
}

int record (mfb_ctx *c, int stage, bool start)
{
    MFB_CUDA (cudaEventRecord (start ? c->evStart[stage] : c->evStop[stage], c->stream));
    if (!start) c->stageRan[stage] = true;
    return MFB_OK;
}

int do_zero (mfb_ctx *c)
{
    MFB_CUDA (cudaMemsetAsync (c->dValues, 0, sizeof (double) * (size_t)c->nbEdges * c->operatorDim, c->stream));
    return MFB_OK;
}

int do_scatter_interval (mfb_ctx *c, int first, int count)
{
    MFB_CUDA (launch_scatter (c->operatorID, c->path == MFB_PATH_ATOMIC, c->dCoord, c->dElemToNode,
                              c->dElemToEdge, c->dValues, first, count, c->stream));
    if (count > 0) c->launches++;
    return MFB_OK;
}

// assembly(): src/assembly.cc:615-720
int do_assembly (mfb_ctx *c, int fusePrec)
{
    if (c->path == MFB_PATH_TILED) {
        MFB_CUDA (launch_write_once (c, 0, c->plan.nbTiles, c->tiledCtas, fusePrec, c->stream));
        if (c->plan.nbTiles > 0) c->launches++;
        return MFB_OK;
    }
    int rc = do_zero (c);                                  // :649-651 / :663-666
    if (rc) return rc;
    if (c->path == MFB_PATH_ATOMIC) return do_scatter_interval (c, 0, c->nbElem);   // :653-659
    if (c->path == MFB_PATH_BLOCKCOLOR) {
        for (size_t bcol = 0; bcol + 1 < c->blockLaunchStart.size (); bcol++) {
            const int first = c->blockLaunchStart[bcol], count = c->blockLaunchStart[bcol + 1] - first;
            MFB_CUDA (launch_scatter_blocks (c->operatorID, c->dCoord, c->dElemToNode, c->dElemToEdge, c->dValues,
                                             c->dLocalIndex, c->dLocalStart, first, count, c->stream));
            if (count > 0) c->launches++;
        }
        return MFB_OK;
    }
    for (int color = 0; color < c->nbTotalColors; color++) {                        // :593-611
        rc = do_scatter_interval (c, c->colorToElem[color], c->colorToElem[color + 1] - c->colorToElem[color]);
        if (rc) return rc;
    }
    return MFB_OK;
}

int do_prec_init (mfb_ctx *c)
{
    MFB_CUDA (launch_prec_init (c->operatorDim, c->dPrec, c->dValues, c->dDiagIndex, c->nbNodes, c->stream));
    if (c->nbNodes > 0) c->launches++;
    return MFB_OK;
}

// MPI_halo_exchange(): src/halo.cc:39-122, on `s`
int do_halo (mfb_ctx *c, cudaStream_t s)
{
    if (c->nbBlocks < 2 || c->nbIntf == 0) return MFB_OK;         // halo.cc:44
    if (!c->comm) return fail (MFB_ERR_STATE, "mfb_ctx_halo_exchange: call mfb_ctx_comm_init first (nbBlocks > 1)");
    std::string why;
    NcclApi *api = nccl_api (why);
    if (!api) return fail (MFB_ERR_COMM, why);
    const int dim = c->operatorDim;
    MFB_CUDA (launch_halo_pack (c->dSend, c->dPrec, c->dIntfNodes, dim, c->nbIntfNodes, s));
    c->launches++;
    int rc = api->GroupStart ();
    for (int i = 0; i < c->nbIntf && rc == 0; i++) {
        const size_t begin = (size_t)c->intfIndex[i] * dim, count = (size_t)(c->intfIndex[i + 1] - c->intfIndex[i]) * dim;
        const int peer = c->neighbors[i] - 1;
        rc = api->Recv (c->dRecv + begin, count, kNcclFloat64, peer, c->comm, s);
        if (rc == 0) rc = api->Send (c->dSend + begin, count, kNcclFloat64, peer, c->comm, s);
    }
    int rc2 = api->GroupEnd ();
    if (rc == 0) rc = rc2;
    if (rc != 0) return fail (MFB_ERR_COMM, std::string ("NCCL halo exchange: ") + api->GetErrorString (rc));
    MFB_CUDA (launch_halo_add (c->dPrec, c->dRecv, c->dUniqNodes, c->dSlotIndex, c->dSlots, dim, c->nbUniqIntf, s));
    c->launches++;
    return MFB_OK;
}

int do_prec_inversion (mfb_ctx *c)
{
    MFB_CUDA (launch_prec_inversion (c->operatorID, c->dPrec, c->dDiagIndex, c->dCheckBounds, c->nbNodes, c->stream));
    if (c->nbNodes > 0) c->launches++;
    return MFB_OK;
}

// One FEM_loop iteration (src/FEM.cc:183-233) with no per-stage timing.
int do_iteration (mfb_ctx *c)
{
    int rc;
    if (c->path != MFB_PATH_TILED) {
        if ((rc = do_assembly (c, 0))) return rc;
        if ((rc = do_prec_init (c))) return rc;
        if ((rc = do_halo (c, c->stream))) return rc;
        return do_prec_inversion (c);
    }
    const bool exchange = c->nbBlocks > 1 && c->nbIntf > 0;
    if (!exchange) return do_assembly (c, 1);
    if (!c->comm && !(c->ring && c->p2pReady)) return fail (MFB_ERR_STATE, "mfb_ctx_iteration: call mfb_ctx_comm_init or mfb_ctx_p2p_connect first (nbBlocks > 1)");
    const int nIntfTiles = c->plan.nbInterfaceTiles, nInterior = c->plan.nbTiles - nIntfTiles;
    if (c->ring && c->p2pReady) {
        // Peer-to-peer exchange: TWO launches per iteration.  The exchange kernel (a few CTAs on the high-priority
        // stream) waits for the assembly kernel's interface tiles — the first tiles of every CTA —, stores their
        // blocks into the neighbours' windows over NVLink, waits for theirs and sums + inverts; the assembly kernel
        // takes ALL tiles on the rest of the device (it never waits for the exchange, so the pair cannot deadlock:
        // its grid leaves p2pReserveCtas free and the exchange kernel asks for no more SMs than that).
        HaloP2PArgs a;
        a.state = c->p2pState;
        a.intfTarget = (unsigned)nIntfTiles * (unsigned)ring_write_out_warps (c->operatorID, c->threads);
        a.status = c->p2pStatusDev;
        a.prec = c->dPrec;
        a.nbIntf = c->nbIntf; a.nbUniq = c->nbUniqIntf; a.nbNodes = c->nbNodes;
        a.intfIndex = c->dIntfIndex; a.intfNodes = c->dIntfNodes;
        a.peerRecv = c->dPeerRecv; a.peerFlag = c->dPeerFlag;
        a.localFlags = reinterpret_cast<const unsigned*> (c->p2pWindow);
        const size_t bufBytes = sizeof (double) * (size_t)c->nbIntfNodes * c->operatorDim;
        a.localRecv[0] = reinterpret_cast<const double*> (c->p2pWindow + kP2PFlagBytes);
        a.localRecv[1] = reinterpret_cast<const double*> (c->p2pWindow + kP2PFlagBytes + p2p_align (bufBytes));
        a.uniqNodes = c->dUniqNodes; a.slotIndex = c->dSlotIndex; a.slots = c->dSlots;
        a.diagIndex = c->dDiagIndex; a.checkBounds = c->dCheckBounds;
        MFB_CUDA (cudaEventRecord (c->evIntfDone, c->stream));                 // everything queued so far
        MFB_CUDA (cudaStreamWaitEvent (c->commStream, c->evIntfDone, 0));
        MFB_CUDA (launch_halo_p2p (a, c->operatorDim, c->p2pCtas, c->commStream));
        c->launches++;
        const int ctas = std::max (std::min (c->tiledCtas, c->deviceCtas - c->p2pReserveCtas), 1);
        MFB_CUDA (launch_ring (c->operatorID, c->ringPlan, 0, c->plan.nbTiles, ctas, c->threads, c->tiledSmem, c->dCoord,
                               c->dValues, c->dPrec, 1, c->stream, &c->p2pState->intfDone));
        if (c->plan.nbTiles > 0) c->launches++;
        MFB_CUDA (cudaEventRecord (c->evCommDone, c->commStream));
        MFB_CUDA (cudaStreamWaitEvent (c->stream, c->evCommDone, 0));
        return MFB_OK;
    }
    if (c->haloCoresident) {
        // MFB_HALO_OVERLAP=coresident: two co-resident kernels.  The high-priority stream takes the
        // tiles that own interface nodes on a small persistent grid, the main stream the interior
        // tiles on the rest of the device.  Fastest with one neighbour (N=2: 0.59 ms per EIB block);
        // with 7 neighbours NCCL's kernel needs more room than the interface CTAs leave and waits
        // for the interior kernel to end (N=8: 0.96 ms) — not the default.
        const int intfCtas = std::max (c->tiledCtas / 16, 1), interiorCtas = std::max (c->tiledCtas - intfCtas, 1);
        MFB_CUDA (cudaEventRecord (c->evIntfDone, c->stream));                 // everything queued so far
        MFB_CUDA (cudaStreamWaitEvent (c->commStream, c->evIntfDone, 0));
        MFB_CUDA (launch_write_once (c, 0, nIntfTiles, intfCtas, 1, c->commStream));
        if (nIntfTiles > 0) c->launches++;
        MFB_CUDA (launch_write_once (c, nIntfTiles, nInterior, interiorCtas, 1, c->stream));
        if (nInterior > 0) c->launches++;
    }
    else {
        // Default: the tiles that own interface nodes come first, on the whole device; their raw
        // diagonal blocks then travel on the high-priority stream (pack, NCCL, add, invert) while
        // the interior tiles assemble.  The interior launch is NOT one persistent grid: its CTAs
        // take a few tiles each, so SM resources free up continuously and the halo kernels are
        // scheduled as soon as their inputs are ready, however many CTAs NCCL asks for.
        MFB_CUDA (launch_write_once (c, 0, nIntfTiles, c->tiledCtas, 1, c->stream));
        if (nIntfTiles > 0) c->launches++;
        MFB_CUDA (cudaEventRecord (c->evIntfDone, c->stream));
        // RING: the warp-specialised kernel pipelines plan records three tiles ahead and wants a persistent grid;
        // it leaves a few SMs free instead (MFB_HALO_RESERVE_CTAS, default 8 CTAs = 4 SMs) so that the pack, NCCL
        // and add kernels of the exchange start at once.
        const int interiorCtas = c->ring ? std::max (c->tiledCtas - c->haloReserveCtas, 1)
                                         : std::max ((nInterior + c->interiorTilesPerCta - 1) / c->interiorTilesPerCta, 1);
        MFB_CUDA (launch_write_once (c, nIntfTiles, nInterior, interiorCtas, 1, c->stream));
        if (nInterior > 0) c->launches++;
        MFB_CUDA (cudaStreamWaitEvent (c->commStream, c->evIntfDone, 0));
    }
    if ((rc = do_halo (c, c->commStream))) return rc;
    MFB_CUDA (launch_prec_inversion_list (c->operatorID, c->dPrec, c->dDiagIndex, c->dCheckBounds, c->nbNodes,
                                          c->dUniqNodes, c->nbUniqIntf, c->commStream));
    if (c->nbUniqIntf > 0) c->launches++;
    MFB_CUDA (cudaEventRecord (c->evCommDone, c->commStream));
    MFB_CUDA (cudaStreamWaitEvent (c->stream, c->evCommDone, 0));
    return MFB_OK;
}

}  // namespace

// A bounded wait of the exchange kernel ran out: the results of that iteration are garbage.
static int p2p_status (mfb_ctx *c)
{
    if (!c->p2pStatusHost || *c->p2pStatusHost == 0) return MFB_OK;
    const unsigned code = *c->p2pStatusHost;
    *c->p2pStatusHost = 0;
    return fail (MFB_ERR_COMM, std::string ("peer-to-peer halo exchange timed out waiting for ") +
                 ((code & 1u) ? "the interface tiles of the assembly kernel" : "a neighbour's flag (do all ranks run the same iterations?)"));
}

extern "C" int mfb_device_count (void)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
    return n;
}

extern "C" void mfb_ctx_destroy (mfb_ctx *c)
{
    if (!c) return;
    cudaSetDevice (c->device);
    if (c->stream) cudaStreamSynchronize (c->stream);
    if (c->commStream) cudaStreamSynchronize (c->commStream);
    if (c->comm) { std::string why; NcclApi *api = nccl_api (why); if (api) api->CommDestroy (c->comm); }
    if (c->graphExec) cudaGraphExecDestroy (c->graphExec);
    for (void *p : c->p2pOpened) cudaIpcCloseMemHandle (p);
    if (c->p2pStatusHost) cudaFreeHost (c->p2pStatusHost);
    void *p2pPtrs[] = {c->p2pWindow, c->p2pState, c->dPeerRecv, c->dPeerFlag, c->dIntfIndex};
    for (void *p : p2pPtrs) if (p) cudaFree (p);
    void *ptrs[] = {c->dCoord, c->dValues, c->dPrec, c->dSend, c->dRecv, c->dElemToNode, c->dRow, c->dCol,
                    c->dElemToEdge, c->dCheckBounds, c->dDiagIndex, c->dIntfNodes, c->dUniqNodes,
                    c->dSlotIndex, c->dSlots, c->dNorm, c->dLocalIndex, c->dLocalStart};
    for (void *p : ptrs) if (p) cudaFree (p);
    for (void *p : c->planAllocs) if (p) cudaFree (p);
    for (int s = 0; s < 5; s++) {
        if (c->evStart[s]) cudaEventDestroy (c->evStart[s]);
        if (c->evStop[s]) cudaEventDestroy (c->evStop[s]);
    }
    if (c->evIntfDone) cudaEventDestroy (c->evIntfDone);
    if (c->evCommDone) cudaEventDestroy (c->evCommDone);
    if (c->stream) cudaStreamDestroy (c->stream);
    if (c->commStream) cudaStreamDestroy (c->commStream);
    delete c;
}

// Node ids of the caller's arrays are used as indices by the plan builders: range-check them before.
static const char *check_problem_ids (const mfb_problem *p)
{
    if (p->nbElem > 0 && !p->elemToNode) return "missing elemToNode";
    for (size_t k = 0; k < (size_t)std::max (p->nbElem, 0) * 4; k++) {
        if (p->elemToNode[k] < 1 || p->elemToNode[k] > p->nbNodes) return "elemToNode id out of range";
    }
    if (p->nbBlocks > 1 && p->nbIntf > 0 && p->nbIntfNodes > 0) {
        if (!p->intfIndex || !p->intfNodes) return "missing interface arrays";
        if (p->intfIndex[0] != 0 || p->intfIndex[p->nbIntf] != p->nbIntfNodes) return "intfIndex / nbIntfNodes mismatch";
        for (int j = 0; j < p->nbIntfNodes; j++) {
            if (p->intfNodes[j] < 1 || p->intfNodes[j] > p->nbNodes) return "interface node id out of range";
        }
    }
    return nullptr;
}

static int ctx_create_impl (const mfb_problem *p, const mfb_options *o, mfb_ctx *c)
{
    c->path = o ? o->path : MFB_PATH_TILED;
    c->device = o ? o->device : 0;
    c->threads = (o && o->threads > 0) ? o->threads : 256;
    c->useGraph = o ? o->useGraph : 0;
    if (c->path < MFB_PATH_TILED || c->path > MFB_PATH_BLOCKCOLOR) return fail (MFB_ERR_ARG, "mfb_ctx_create: unknown path");
    if (c->path == MFB_PATH_RING) {          // a write-once path like TILED: same stages, same fused iteration
        c->ring = true;
        c->path = MFB_PATH_TILED;
        // default CTA: 768 threads (13 job + 11 write-out warps) for elasticity; the Laplacian, whose write-out is an eighth
        // of the bytes, runs 24 job + 8 write-out warps in a 1024-thread CTA with per-warpgroup register counts
        // (EIB, ms per iteration: elasticity 768: 0.406, 896: 0.431, 1024: 0.437, 640: 0.426; Laplacian 768: 0.240, 1024: 0.225)
        if (!(o && o->threads > 0)) c->threads = p->operatorID == 0 ? 1024 : 768;
        if (!ring_threads_supported (c->threads)) return fail (MFB_ERR_ARG, "mfb_ctx_create: the RING kernel runs 384 (two CTAs per SM), 640, 768, 896 or 1024 (one) threads per CTA");
    }
    if (!c->ring && c->threads != tiled_pipeline_threads () && (c->threads % 32 || c->threads < 32 || c->threads > 256)) {
        return fail (MFB_ERR_ARG, "mfb_ctx_create: threads must be a multiple of 32 in [32, 256], or the pipelined kernel's CTA size");
    }
    if (p->operatorID != 0 && p->operatorID != 1) return fail (MFB_ERR_ARG, "mfb_ctx_create: operatorID must be 0 (lap) or 1 (ela)");
    if (p->nbElem < 0 || p->nbNodes < 0 || p->nbEdges < 0) return fail (MFB_ERR_ARG, "mfb_ctx_create: negative size");
    if ((p->nbNodes > 0 && (!p->coord || !p->nodeToNodeRow)) || (p->nbElem > 0 && !p->elemToNode) ||
        (p->nbEdges > 0 && !p->nodeToNodeColumn)) return fail (MFB_ERR_ARG, "mfb_ctx_create: missing array");
    if (p->nbNodes > 0 && p->nodeToNodeRow[p->nbNodes] != p->nbEdges) {
        return fail (MFB_ERR_ARG, "mfb_ctx_create: nbEdges does not match nodeToNodeRow[nbNodes]");
    }
    if (c->path == MFB_PATH_COLOR && (!p->colorToElem || p->nbTotalColors < 1)) {
        return fail (MFB_ERR_ARG, "mfb_ctx_create: the COLOR path needs colorToElem / nbTotalColors (coloring.cc)");
    }
    if (p->nbBlocks > 1 && p->nbIntf > 0 && (!p->intfIndex || !p->intfNodes || !p->neighborsList)) {
        return fail (MFB_ERR_ARG, "mfb_ctx_create: missing interface arrays");
    }
    // ids are used as indices by the plan builders below: range-check them first
    if (const char *bad = check_problem_ids (p)) return fail (MFB_ERR_ARG, std::string ("mfb_ctx_create: ") + bad);
    c->operatorID = p->operatorID;
    c->operatorDim = p->operatorID == 0 ? 1 : 9;
    c->nbElem = p->nbElem; c->nbNodes = p->nbNodes; c->nbEdges = p->nbEdges;
    c->nbBlocks = std::max (p->nbBlocks, 1); c->rank = p->rank;
    c->nbIntf = c->nbBlocks > 1 ? p->nbIntf : 0;
    c->nbIntfNodes = c->nbBlocks > 1 ? p->nbIntfNodes : 0;

    int count = 0;
    if (cudaGetDeviceCount (&count) != cudaSuccess || count == 0) {
        cudaGetLastError ();
        return fail (MFB_ERR_CUDA, "mfb_ctx_create: no CUDA device (this path has no CPU fallback)");
    }
    MFB_CUDA (cudaSetDevice (c->device));
    MFB_CUDA (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
    // the halo chain (pack, NCCL, add, interface inversion) must slip in between the CTAs of the
    // interior assembly, not queue behind them
    int leastPriority = 0, greatestPriority = 0;
    MFB_CUDA (cudaDeviceGetStreamPriorityRange (&leastPriority, &greatestPriority));
    MFB_CUDA (cudaStreamCreateWithPriority (&c->commStream, cudaStreamNonBlocking, greatestPriority));
    for (int s = 0; s < 5; s++) { MFB_CUDA (cudaEventCreate (&c->evStart[s])); MFB_CUDA (cudaEventCreate (&c->evStop[s])); }
    MFB_CUDA (cudaEventCreateWithFlags (&c->evIntfDone, cudaEventDisableTiming));
    MFB_CUDA (cudaEventCreateWithFlags (&c->evCommDone, cudaEventDisableTiming));

    MFB_CUDA (upload (&c->dCoord, p->coord, (size_t)p->nbNodes * 3, c->meshBytes));
    MFB_CUDA (upload (&c->dRow, p->nodeToNodeRow, (size_t)p->nbNodes + 1, c->meshBytes));
    MFB_CUDA (upload (&c->dCol, p->nodeToNodeColumn, (size_t)p->nbEdges, c->meshBytes));
    MFB_CUDA (upload (&c->dCheckBounds, p->checkBounds, p->checkBounds ? (size_t)p->nbNodes * 3 : 0, c->meshBytes));
    MFB_CUDA (upload<double> (&c->dValues, nullptr, std::max<size_t> ((size_t)p->nbEdges * c->operatorDim, 1), c->meshBytes));
    MFB_CUDA (upload<double> (&c->dPrec, nullptr, std::max<size_t> ((size_t)p->nbNodes * c->operatorDim, 1), c->meshBytes));
    MFB_CUDA (upload<int> (&c->dDiagIndex, nullptr, std::max<size_t> ((size_t)p->nbNodes, 1), c->meshBytes));
    MFB_CUDA (launch_diag_index (c->dRow, c->dCol, c->dDiagIndex, c->nbNodes, c->stream));

    if (c->path != MFB_PATH_TILED) {
        const int *elemToEdgeHost = p->elemToEdge;
        if (c->path == MFB_PATH_BLOCKCOLOR) {
            // the library's own element order: by (block colour, block, local colour); elemToEdge is rebuilt for it below
            for (int e = 0; e < p->nbElem; e++) {              // the kernel updates the four entries of a row of the element matrix together
                const int *en = p->elemToNode + (size_t)e * 4;
                if (en[0] == en[1] || en[0] == en[2] || en[0] == en[3] || en[1] == en[2] || en[1] == en[3] || en[2] == en[3]) {
                    return fail (MFB_ERR_ARG, "mfb_ctx_create: an element names a node twice (element " + std::to_string (e) + ")");
                }
            }
            BlockColoring bc;
            int blockElems = 1024;
            if (o && o->tileElems > 0) blockElems = o->tileElems;
            const int rcb = build_block_coloring (p->elemToNode, p->nbElem, p->nbNodes, p->coord, blockElems, bc);
            if (rcb == -1) return fail (MFB_ERR_COLORS, "mfb_ctx_create: a block needs more than 128 local colours");
            if (rcb == -2) return fail (MFB_ERR_COLORS, "mfb_ctx_create: the blocks need more than 64 colours");
            std::vector<int> permuted ((size_t)p->nbElem * 4);
            for (int e = 0; e < p->nbElem; e++) memcpy (&permuted[(size_t)e * 4], p->elemToNode + (size_t)bc.elemOrder[e] * 4, 4 * sizeof (int));
            MFB_CUDA (upload (&c->dElemToNode, permuted.data (), permuted.size (), c->meshBytes));
            MFB_CUDA (upload (&c->dLocalIndex, bc.localIndex.data (), bc.localIndex.size (), c->meshBytes));
            MFB_CUDA (upload (&c->dLocalStart, bc.localStart.data (), bc.localStart.size (), c->meshBytes));
            c->blockLaunchStart = bc.launchStart;
            c->blockStats[0] = bc.nbBlocks; c->blockStats[1] = bc.nbBlockColors; c->blockStats[2] = bc.maxLocalColors;
            elemToEdgeHost = nullptr;
        }
        else MFB_CUDA (upload (&c->dElemToNode, p->elemToNode, (size_t)p->nbElem * 4, c->meshBytes));
        MFB_CUDA (upload (&c->dElemToEdge, elemToEdgeHost, (size_t)p->nbElem * 16, c->meshBytes));
        if (!elemToEdgeHost && p->nbElem > 0) {            // create_elemToEdge on the device
            int *dMissing = nullptr, missing = 0;
            MFB_CUDA (cudaMalloc ((void**)&dMissing, sizeof (int)));
            MFB_CUDA (cudaMemsetAsync (dMissing, 0, sizeof (int), c->stream));
            MFB_CUDA (launch_elem_to_edge (c->dRow, c->dCol, c->dElemToNode, c->dElemToEdge, c->nbElem, dMissing, c->stream));
            MFB_CUDA (cudaMemcpyAsync (&missing, dMissing, sizeof (int), cudaMemcpyDeviceToHost, c->stream));
            MFB_CUDA (cudaStreamSynchronize (c->stream));
            cudaFree (dMissing);
            if (missing) return fail (MFB_ERR_ARG, "mfb_ctx_create: the CSR lacks a node pair of an element");
        }
        if (c->path == MFB_PATH_COLOR) {
            c->nbTotalColors = p->nbTotalColors;
            c->colorToElem.assign (p->colorToElem, p->colorToElem + p->nbTotalColors + 1);
            if (c->colorToElem[0] != 0 || c->colorToElem[p->nbTotalColors] != p->nbElem) {
                return fail (MFB_ERR_ARG, "mfb_ctx_create: colorToElem does not cover [0, nbElem)");
            }
        }
    }
    else {
        int rc = c->ring ? build_device_ring_plan (c, p, o) : build_device_plan (c, p, o);
        if (rc) return rc;
    }

    if (c->nbIntf > 0) {
        c->intfIndex.assign (p->intfIndex, p->intfIndex + p->nbIntf + 1);
        c->neighbors.assign (p->neighborsList, p->neighborsList + p->nbIntf);
        if (c->intfIndex[p->nbIntf] != p->nbIntfNodes) return fail (MFB_ERR_ARG, "mfb_ctx_create: intfIndex / nbIntfNodes mismatch");
        for (int j = 0; j < p->nbIntfNodes; j++) {
            if (p->intfNodes[j] < 1 || p->intfNodes[j] > p->nbNodes) return fail (MFB_ERR_ARG, "mfb_ctx_create: interface node id out of range");
        }
        // unique interface nodes and, for each, its positions j in increasing order
        std::vector<std::pair<int, int>> byNode ((size_t)p->nbIntfNodes);
        for (int j = 0; j < p->nbIntfNodes; j++) byNode[j] = {p->intfNodes[j] - 1, j};
        std::sort (byNode.begin (), byNode.end ());
        std::vector<int> uniq, slotIndex (1, 0), slots;
        for (size_t k = 0; k < byNode.size (); k++) {
            if (k == 0 || byNode[k].first != byNode[k - 1].first) {
                if (k) slotIndex.push_back ((int)slots.size ());
                uniq.push_back (byNode[k].first);
            }
            slots.push_back (byNode[k].second);
        }
        slotIndex.push_back ((int)slots.size ());
        c->nbUniqIntf = (int)uniq.size ();
        const size_t bufDoubles = (size_t)p->nbIntfNodes * c->operatorDim;
        MFB_CUDA (upload (&c->dIntfNodes, p->intfNodes, (size_t)p->nbIntfNodes, c->meshBytes));
        MFB_CUDA (upload (&c->dUniqNodes, uniq.data (), uniq.size (), c->meshBytes));
        MFB_CUDA (upload (&c->dSlotIndex, slotIndex.data (), slotIndex.size (), c->meshBytes));
        MFB_CUDA (upload (&c->dSlots, slots.data (), slots.size (), c->meshBytes));
        MFB_CUDA (upload<double> (&c->dSend, nullptr, bufDoubles, c->meshBytes));
        MFB_CUDA (upload<double> (&c->dRecv, nullptr, bufDoubles, c->meshBytes));
    }
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    return MFB_OK;
}

extern "C" int mfb_ctx_create (const mfb_problem *problem, const mfb_options *options, mfb_ctx **out)
{
    if (!problem || !out) return fail (MFB_ERR_ARG, "mfb_ctx_create: NULL argument");
    *out = nullptr;
    mfb_ctx *c = new mfb_ctx ();
    int rc = ctx_create_impl (problem, options, c);
    if (rc != MFB_OK) {
        std::string keep = g_lastError;
        mfb_ctx_destroy (c);
        g_lastError = keep;
        return rc;
    }
    *out = c;
    return MFB_OK;
}

#define CTX_ENTER(c)                                                        \
    if (!(c)) return fail (MFB_ERR_ARG, "NULL context");                    \
    MFB_CUDA (cudaSetDevice ((c)->device))

extern "C" int mfb_ctx_zero_values (mfb_ctx *c)
{
    CTX_ENTER (c);
    return do_zero (c);
}

extern "C" int mfb_ctx_assembly_interval (mfb_ctx *c, int firstElem, int lastElem)
{
    CTX_ENTER (c);
    if (c->path == MFB_PATH_TILED) return fail (MFB_ERR_STATE, "mfb_ctx_assembly_interval: element intervals exist on the ATOMIC / COLOR paths only");
    if (c->path == MFB_PATH_BLOCKCOLOR) return fail (MFB_ERR_STATE, "mfb_ctx_assembly_interval: the BLOCKCOLOR path keeps its own element order; use ATOMIC or COLOR");
    if (firstElem < 0 || lastElem >= c->nbElem) return fail (MFB_ERR_ARG, "mfb_ctx_assembly_interval: interval out of range");
    if (c->path == MFB_PATH_COLOR) {
        // The plain += of the COLOR kernel is only conflict-free inside one colour (coloring.cc): an interval that
        // spans colours — (0, nbElem - 1), a D&C leaf — is cut at the colorToElem boundaries, one launch per piece,
        // in colour order like coloring_assembly (src/assembly.cc:593-611).
        for (int color = 0; color < c->nbTotalColors; color++) {
            const int lo = std::max (firstElem, c->colorToElem[color]), hi = std::min (lastElem, c->colorToElem[color + 1] - 1);
            if (lo > hi) continue;
            const int rc = do_scatter_interval (c, lo, hi - lo + 1);
            if (rc) return rc;
        }
        return MFB_OK;
    }
    return do_scatter_interval (c, firstElem, lastElem - firstElem + 1);
}

extern "C" int mfb_ctx_assembly (mfb_ctx *c)
{
    CTX_ENTER (c);
    int rc = record (c, 0, true);
    if (!rc) rc = do_assembly (c, 0);
    if (!rc) rc = record (c, 0, false);
    return rc;
}

extern "C" int mfb_ctx_prec_init (mfb_ctx *c)
{
    CTX_ENTER (c);
    int rc = record (c, 1, true);
    if (!rc) rc = do_prec_init (c);
    if (!rc) rc = record (c, 1, false);
    return rc;
}

extern "C" int mfb_ctx_halo_exchange (mfb_ctx *c)
{
    CTX_ENTER (c);
    int rc = record (c, 2, true);
    if (!rc) rc = do_halo (c, c->stream);
    if (!rc) rc = record (c, 2, false);
    return rc;
}

extern "C" int mfb_ctx_prec_inversion (mfb_ctx *c)
{
    CTX_ENTER (c);
    int rc = record (c, 3, true);
    if (!rc) rc = do_prec_inversion (c);
    if (!rc) rc = record (c, 3, false);
    return rc;
}

extern "C" int mfb_ctx_iteration (mfb_ctx *c)
{
    CTX_ENTER (c);
    int rc = record (c, 4, true);
    if (rc) return rc;
    // The whole iteration as one CUDA graph.  With several subdomains the graph holds both
    // streams — interface tiles, interior tiles, pack, the NCCL group, add, interface inversion — joined by the two
    // events; NCCL is given two eager iterations first (it allocates on first use), and a capture it refuses
    // falls back to eager launches for good.
    const bool multi = c->nbBlocks > 1 && c->nbIntf > 0;
    const bool graphable = c->useGraph && (!multi || (c->multiGraph && c->eagerIterations >= 2));
    if (graphable) {
        if (!c->graphExec) {
            cudaGraph_t graph = nullptr;
            const int64_t before = c->launches;
            MFB_CUDA (cudaStreamBeginCapture (c->stream, multi ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
            rc = do_iteration (c);
            cudaError_t e = cudaStreamEndCapture (c->stream, &graph);
            c->graphLaunches = c->launches - before;      // kernels one replay runs
            c->launches = before;
            if (!rc && e == cudaSuccess) e = cudaGraphInstantiate (&c->graphExec, graph, 0);
            if (graph) cudaGraphDestroy (graph);
            if (rc || e != cudaSuccess) {
                if (!multi) { if (rc) return rc; MFB_CUDA (e); }
                cudaGetLastError ();
                c->graphExec = nullptr;
                c->useGraph = 0;                          // eager from now on
                g_lastError = "mfb_ctx_iteration: graph capture of the multi-GPU iteration refused, running eagerly";
                rc = do_iteration (c);
                if (rc) return rc;
                return record (c, 4, false);
            }
        }
        MFB_CUDA (cudaGraphLaunch (c->graphExec, c->stream));
        c->launches += c->graphLaunches;
    }
    else {
        rc = do_iteration (c);
        if (rc) return rc;
        c->eagerIterations++;
    }
    return record (c, 4, false);
}

extern "C" int mfb_ctx_sync (mfb_ctx *c)
{
    CTX_ENTER (c);
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    MFB_CUDA (cudaStreamSynchronize (c->commStream));
    return p2p_status (c);
}

extern "C" int mfb_ctx_download (mfb_ctx *c, double *nodeToNodeValue, double *prec)
{
    CTX_ENTER (c);
    if (nodeToNodeValue && c->nbEdges > 0) {
        MFB_CUDA (cudaMemcpyAsync (nodeToNodeValue, c->dValues, sizeof (double) * (size_t)c->nbEdges * c->operatorDim,
                                   cudaMemcpyDeviceToHost, c->stream));
    }
    if (prec && c->nbNodes > 0) {
        MFB_CUDA (cudaMemcpyAsync (prec, c->dPrec, sizeof (double) * (size_t)c->nbNodes * c->operatorDim,
                                   cudaMemcpyDeviceToHost, c->stream));
    }
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    return MFB_OK;
}

extern "C" int mfb_ctx_upload_coord (mfb_ctx *c, const double *coord)
{
    CTX_ENTER (c);
    if (!coord) return fail (MFB_ERR_ARG, "mfb_ctx_upload_coord: NULL");
    if (c->nbNodes > 0) {
        MFB_CUDA (cudaMemcpyAsync (c->dCoord, coord, sizeof (double) * (size_t)c->nbNodes * 3, cudaMemcpyHostToDevice, c->stream));
    }
    return MFB_OK;
}

extern "C" int mfb_ctx_iteration_host (mfb_ctx *c, const double *coord, double *nodeToNodeValue, double *prec)
{
    int rc = mfb_ctx_upload_coord (c, coord);
    if (!rc) rc = mfb_ctx_iteration (c);
    if (!rc) rc = mfb_ctx_download (c, nodeToNodeValue, prec);
    return rc;
}

// check_results' two norms (FEM.cc:68-76) without moving the arrays to the host.
extern "C" int mfb_ctx_norms (mfb_ctx *c, double *matrixNorm, double *precNorm)
{
    CTX_ENTER (c);
    if (!matrixNorm || !precNorm) return fail (MFB_ERR_ARG, "mfb_ctx_norms: NULL");
    const int scratch = double_norm_scratch_doubles ();
    if (!c->dNorm) MFB_CUDA (cudaMalloc (&c->dNorm, sizeof (double) * (size_t)(scratch + 2)));
    MFB_CUDA (launch_double_norm (c->dValues, (int64_t)c->nbEdges * c->operatorDim, c->dNorm, c->dNorm + scratch, c->stream));
    MFB_CUDA (launch_double_norm (c->dPrec, (int64_t)c->nbNodes * c->operatorDim, c->dNorm, c->dNorm + scratch + 1, c->stream));
    c->launches += 4;
    double host[2];
    MFB_CUDA (cudaMemcpyAsync (host, c->dNorm + scratch, sizeof (host), cudaMemcpyDeviceToHost, c->stream));
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    *matrixNorm = host[0];
    *precNorm = host[1];
    return MFB_OK;
}

extern "C" int mfb_ctx_iteration_norms_host (mfb_ctx *c, const double *coord, double norms[2])
{
    if (!norms) return fail (MFB_ERR_ARG, "mfb_ctx_iteration_norms_host: NULL");
    int rc = mfb_ctx_upload_coord (c, coord);
    if (!rc) rc = mfb_ctx_iteration (c);
    if (!rc) rc = mfb_ctx_norms (c, &norms[0], &norms[1]);
    return rc;
}

extern "C" int mfb_ctx_device_ptrs (mfb_ctx *c, void **values, void **prec)
{
    if (!c) return fail (MFB_ERR_ARG, "NULL context");
    if (values) *values = c->dValues;
    if (prec) *prec = c->dPrec;
    return MFB_OK;
}

extern "C" int mfb_ctx_stream (mfb_ctx *c, void **stream)
{
    if (!c || !stream) return fail (MFB_ERR_ARG, "NULL argument");
    *stream = (void*)c->stream;
    return MFB_OK;
}

extern "C" int mfb_ctx_stage_ms (mfb_ctx *c, float ms[5])
{
    CTX_ENTER (c);
    for (int s = 0; s < 5; s++) {
        ms[s] = 0.f;
        if (!c->stageRan[s]) continue;
        MFB_CUDA (cudaEventSynchronize (c->evStop[s]));
        MFB_CUDA (cudaEventElapsedTime (&ms[s], c->evStart[s], c->evStop[s]));
    }
    return MFB_OK;
}

extern "C" int64_t mfb_ctx_launch_count (mfb_ctx *c) { return c ? c->launches : 0; }

extern "C" int mfb_ctx_device_bytes (mfb_ctx *c, int64_t *meshBytes, int64_t *planBytes)
{
    if (!c) return fail (MFB_ERR_ARG, "NULL context");
    if (meshBytes) *meshBytes = c->meshBytes;
    if (planBytes) *planBytes = c->planBytes;
    return MFB_OK;
}

extern "C" int mfb_ctx_plan_stats (mfb_ctx *c, int64_t stats[8])
{
    if (!c || !stats) return fail (MFB_ERR_ARG, "NULL argument");
    if (c->path == MFB_PATH_BLOCKCOLOR) {       // [0] blocks [1] block colours (= launches per assembly) [2] max local colours of a block
        for (int k = 0; k < 8; k++) stats[k] = k < 3 ? c->blockStats[k] : 0;
        return MFB_OK;
    }
    if (c->ring) {              // RING: [1] jobs (mesh edges) [2] ring steps [4] max nodes [6] padded lane-steps
        stats[0] = c->ringStats.nbTiles; stats[1] = c->ringStats.nbJobs; stats[2] = c->ringStats.nbRingSteps;
        stats[3] = c->ringStats.maxRows; stats[4] = c->ringStats.maxNodes; stats[5] = (int64_t)c->tiledSmem;
        stats[6] = c->ringStats.nbPaddedSteps; stats[7] = c->ringStats.maxBlobBytes;
        return MFB_OK;
    }
    stats[0] = c->hostPlanStats.nbTiles; stats[1] = c->hostPlanStats.nbTileElems;
    stats[2] = c->hostPlanStats.nbContributions; stats[3] = c->hostPlanStats.maxRows;
    stats[4] = c->hostPlanStats.maxElems; stats[5] = (int64_t)c->tiledSmem;
    stats[6] = c->hostPlanStats.nbPaddedSteps * 32; stats[7] = c->hostPlanStats.maxBlobBytes;
    return MFB_OK;
}

extern "C" int mfb_comm_unique_id (unsigned char id[MFB_COMM_ID_BYTES])
{
    std::string why;
    NcclApi *api = nccl_api (why);
    if (!api) return fail (MFB_ERR_COMM, why);
    NcclId nid;
    int rc = api->GetUniqueId (&nid);
    if (rc != 0) return fail (MFB_ERR_COMM, std::string ("ncclGetUniqueId: ") + api->GetErrorString (rc));
    memcpy (id, nid.internal, MFB_COMM_ID_BYTES);
    return MFB_OK;
}

extern "C" int mfb_ctx_comm_init (mfb_ctx *c, const unsigned char id[MFB_COMM_ID_BYTES])
{
    CTX_ENTER (c);
    if (c->nbBlocks < 2) return MFB_OK;
    std::string why;
    NcclApi *api = nccl_api (why);
    if (!api) return fail (MFB_ERR_COMM, why);
    NcclId nid;
    memcpy (nid.internal, id, MFB_COMM_ID_BYTES);
    int rc = api->CommInitRank (&c->comm, c->nbBlocks, nid, c->rank);
    if (rc != 0) { c->comm = nullptr; return fail (MFB_ERR_COMM, std::string ("ncclCommInitRank: ") + api->GetErrorString (rc)); }
    return MFB_OK;
}

// ---- peer-to-peer windows (kernels_halo_p2p.cu) -----------------------------------------------------------------
namespace {
struct P2PCard {                       // what a subdomain publishes; MFB_P2P_CARD_BYTES on the wire
    uint32_t magic;
    int32_t rank, device, nbIntf, dim, hasHandle;
    int64_t pid;
    uint64_t rawPtr, windowBytes;
    cudaIpcMemHandle_t handle;
    int32_t neighbors[kP2PMaxIntf];    // 1-based ranks, as neighborsList
    int32_t intfIndex[kP2PMaxIntf + 1];
};
static_assert (sizeof (P2PCard) <= MFB_P2P_CARD_BYTES, "card does not fit its wire size");
constexpr uint32_t kP2PMagic = 0x4D465032u;   // "MFP2"
}

extern "C" int mfb_ctx_p2p_card (mfb_ctx *c, unsigned char card[MFB_P2P_CARD_BYTES])
{
    CTX_ENTER (c);
    if (!card) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_card: NULL card");
    if (!c->ring) return fail (MFB_ERR_STATE, "mfb_ctx_p2p_card: the peer-to-peer exchange belongs to the fused RING iteration");
    if (c->nbIntf > kP2PMaxIntf) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_card: more than 64 interfaces; use the NCCL exchange");
    const size_t bufBytes = sizeof (double) * (size_t)c->nbIntfNodes * c->operatorDim;
    const size_t windowBytes = kP2PFlagBytes + 2 * p2p_align (bufBytes);
    if (!c->p2pWindow) {
        MFB_CUDA (cudaMalloc ((void**)&c->p2pWindow, windowBytes));
        MFB_CUDA (cudaMemset (c->p2pWindow, 0, windowBytes));
        MFB_CUDA (cudaMalloc ((void**)&c->p2pState, sizeof (HaloP2PState)));
        MFB_CUDA (cudaMemset (c->p2pState, 0, sizeof (HaloP2PState)));
        MFB_CUDA (cudaHostAlloc ((void**)&c->p2pStatusHost, sizeof (unsigned), cudaHostAllocMapped));
        *c->p2pStatusHost = 0;
        MFB_CUDA (cudaHostGetDevicePointer ((void**)&c->p2pStatusDev, c->p2pStatusHost, 0));
        int64_t bytes = 0;
        MFB_CUDA (upload (&c->dIntfIndex, c->intfIndex.data (), c->intfIndex.size (), bytes));
        c->meshBytes += (int64_t)windowBytes + bytes;
    }
    P2PCard k;
    memset (&k, 0, sizeof k);
    k.magic = kP2PMagic; k.rank = c->rank; k.device = c->device; k.nbIntf = c->nbIntf; k.dim = c->operatorDim;
    k.pid = (int64_t)getpid (); k.rawPtr = (uint64_t)(uintptr_t)c->p2pWindow; k.windowBytes = windowBytes;
    if (cudaIpcGetMemHandle (&k.handle, c->p2pWindow) == cudaSuccess) k.hasHandle = 1;
    else cudaGetLastError ();                              // same-process peers still work through rawPtr
    for (int i = 0; i < c->nbIntf; i++) k.neighbors[i] = c->neighbors[i];
    for (int i = 0; i <= c->nbIntf; i++) k.intfIndex[i] = c->intfIndex.empty () ? 0 : c->intfIndex[i];
    memset (card, 0, MFB_P2P_CARD_BYTES);
    memcpy (card, &k, sizeof k);
    return MFB_OK;
}

extern "C" int mfb_ctx_p2p_connect (mfb_ctx *c, const unsigned char *cards)
{
    CTX_ENTER (c);
    if (!cards) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: NULL cards");
    if (!c->p2pWindow) return fail (MFB_ERR_STATE, "mfb_ctx_p2p_connect: call mfb_ctx_p2p_card first");
    c->p2pReady = false;
    std::vector<double*> peerRecv ((size_t)2 * c->nbIntf, nullptr);
    std::vector<unsigned*> peerFlag ((size_t)c->nbIntf, nullptr);
    std::vector<unsigned char*> baseOfRank ((size_t)c->nbBlocks, nullptr);
    for (int i = 0; i < c->nbIntf; i++) {
        const int q = c->neighbors[i] - 1;
        if (q < 0 || q >= c->nbBlocks) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: neighbour rank out of range");
        P2PCard k;
        memcpy (&k, cards + (size_t)q * MFB_P2P_CARD_BYTES, sizeof k);
        if (k.magic != kP2PMagic || k.rank != q) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: rank " + std::to_string (q) + " published no card");
        if (k.dim != c->operatorDim) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: operator differs between neighbours");
        int iq = -1;
        for (int t = 0; t < k.nbIntf; t++) if (k.neighbors[t] - 1 == c->rank) { iq = t; break; }
        if (iq < 0) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: rank " + std::to_string (q) + " does not list this subdomain as a neighbour");
        const int mine = c->intfIndex[i + 1] - c->intfIndex[i], theirs = k.intfIndex[iq + 1] - k.intfIndex[iq];
        if (mine != theirs) return fail (MFB_ERR_ARG, "mfb_ctx_p2p_connect: interface sizes differ between the two sides");
        if (!baseOfRank[q]) {
            if (k.pid == (int64_t)getpid ()) {
                // both subdomains in one process (tests, or a driver that owns several GPUs): plain device pointers
                if (k.device != c->device) {
                    int can = 0;
                    MFB_CUDA (cudaDeviceCanAccessPeer (&can, c->device, k.device));
                    if (!can) return fail (MFB_ERR_COMM, "mfb_ctx_p2p_connect: no peer access between devices " + std::to_string (c->device) + " and " + std::to_string (k.device));
                    cudaError_t e = cudaDeviceEnablePeerAccess (k.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MFB_CUDA (e);
                    cudaGetLastError ();
                }
                baseOfRank[q] = reinterpret_cast<unsigned char*> ((uintptr_t)k.rawPtr);
            }
            else {
                if (!k.hasHandle) return fail (MFB_ERR_COMM, "mfb_ctx_p2p_connect: rank " + std::to_string (q) + " could not export its window (cudaIpcGetMemHandle)");
                void *ptr = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle (&ptr, k.handle, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) { cudaGetLastError (); return fail (MFB_ERR_COMM, std::string ("mfb_ctx_p2p_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString (e)); }
                c->p2pOpened.push_back (ptr);
                baseOfRank[q] = static_cast<unsigned char*> (ptr);
            }
        }
        const size_t peerBuf = p2p_align (sizeof (double) * (size_t)k.intfIndex[k.nbIntf] * k.dim);
        for (int par = 0; par < 2; par++) {
            peerRecv[(size_t)2 * i + par] = reinterpret_cast<double*> (baseOfRank[q] + kP2PFlagBytes + (size_t)par * peerBuf) +
                                            (size_t)k.intfIndex[iq] * k.dim;
        }
        peerFlag[i] = reinterpret_cast<unsigned*> (baseOfRank[q]) + iq;
    }
    if (c->dPeerRecv) { cudaFree (c->dPeerRecv); c->dPeerRecv = nullptr; }
    if (c->dPeerFlag) { cudaFree (c->dPeerFlag); c->dPeerFlag = nullptr; }
    int64_t bytes = 0;
    MFB_CUDA (upload (&c->dPeerRecv, peerRecv.data (), peerRecv.size (), bytes));
    MFB_CUDA (upload (&c->dPeerFlag, peerFlag.data (), peerFlag.size (), bytes));
    // The assembly grid leaves MFB_P2P_RESERVE_SMS SMs (default 2) to the exchange kernel, which runs one CTA on each
    // (MFB_P2P_CTAS lowers that): never more SMs than are left free, or the pair could wait for each other.  Two instead
    // of the four SMs the NCCL chain gets: the EIB block then takes 111 instead of 113 rounds of tiles per CTA.
    const int ctasPerSM = c->threads >= 640 ? 1 : 2;
    int reserveSMs = 2;
    if (const char *v = getenv ("MFB_P2P_RESERVE_SMS")) reserveSMs = std::min (std::max (atoi (v), 1), 16);
    c->p2pReserveCtas = reserveSMs * ctasPerSM;
    c->p2pCtas = reserveSMs;
    if (const char *v = getenv ("MFB_P2P_CTAS")) c->p2pCtas = std::min (std::max (atoi (v), 1), reserveSMs);
    if (c->deviceCtas - c->p2pReserveCtas < 1) return fail (MFB_ERR_STATE, "mfb_ctx_p2p_connect: the device has too few SMs to run the exchange kernel beside the assembly kernel");
    c->p2pReady = true;
    return MFB_OK;
}

extern "C" int mfb_ctx_p2p_enable (mfb_ctx *c, int on)
{
    CTX_ENTER (c);
    if (on && !c->dPeerRecv && c->nbIntf > 0) return fail (MFB_ERR_STATE, "mfb_ctx_p2p_enable: not connected");
    if (on && !c->p2pWindow) return fail (MFB_ERR_STATE, "mfb_ctx_p2p_enable: not connected");
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    MFB_CUDA (cudaStreamSynchronize (c->commStream));
    c->p2pReady = on != 0;
    return MFB_OK;
}

extern "C" int mfb_ctx_p2p_active (mfb_ctx *c) { return c && c->p2pReady ? 1 : 0; }

extern "C" int mfb_ctx_halo_pack_host (mfb_ctx *c, double *sendBuf)
{
    CTX_ENTER (c);
    if (c->nbIntfNodes == 0) return MFB_OK;
    if (!sendBuf) return fail (MFB_ERR_ARG, "mfb_ctx_halo_pack_host: NULL buffer");
    MFB_CUDA (launch_halo_pack (c->dSend, c->dPrec, c->dIntfNodes, c->operatorDim, c->nbIntfNodes, c->stream));
    c->launches++;
    MFB_CUDA (cudaMemcpyAsync (sendBuf, c->dSend, sizeof (double) * (size_t)c->nbIntfNodes * c->operatorDim,
                               cudaMemcpyDeviceToHost, c->stream));
    MFB_CUDA (cudaStreamSynchronize (c->stream));
    return MFB_OK;
}

extern "C" int mfb_ctx_halo_add_host (mfb_ctx *c, const double *recvBuf)
{
    CTX_ENTER (c);
    if (c->nbIntfNodes == 0) return MFB_OK;
    if (!recvBuf) return fail (MFB_ERR_ARG, "mfb_ctx_halo_add_host: NULL buffer");
    MFB_CUDA (cudaMemcpyAsync (c->dRecv, recvBuf, sizeof (double) * (size_t)c->nbIntfNodes * c->operatorDim,
                               cudaMemcpyHostToDevice, c->stream));
    MFB_CUDA (launch_halo_add (c->dPrec, c->dRecv, c->dUniqNodes, c->dSlotIndex, c->dSlots, c->operatorDim,
                               c->nbUniqIntf, c->stream));
    c->launches++;
    return MFB_OK;
}

extern "C" int mfb_ctx_assembly_fused (mfb_ctx *c)
{
    CTX_ENTER (c);
    if (c->path != MFB_PATH_TILED) return fail (MFB_ERR_STATE, "mfb_ctx_assembly_fused: only the TILED path fuses the preconditioner into assembly");
    int rc = record (c, 0, true);
    if (!rc) rc = do_assembly (c, 1);
    if (!rc) rc = record (c, 0, false);
    return rc;
}

extern "C" int mfb_ctx_prec_inversion_interface (mfb_ctx *c)
{
    CTX_ENTER (c);
    MFB_CUDA (launch_prec_inversion_list (c->operatorID, c->dPrec, c->dDiagIndex, c->dCheckBounds, c->nbNodes,
                                          c->dUniqNodes, c->nbUniqIntf, c->stream));
    if (c->nbUniqIntf > 0) c->launches++;
    return MFB_OK;
}

extern "C" int mfb_ctx_run_timed (mfb_ctx *c, int steps, float *ms)
{
    CTX_ENTER (c);
    if (steps < 1 || !ms) return fail (MFB_ERR_ARG, "mfb_ctx_run_timed: bad argument");
    MFB_CUDA (cudaStreamSynchronize (c->commStream));
    MFB_CUDA (cudaEventRecord (c->evStart[4], c->stream));
    for (int k = 0; k < steps; k++) {
        int rc;
        if (c->useGraph && c->nbBlocks < 2 && c->graphExec) {
            MFB_CUDA (cudaGraphLaunch (c->graphExec, c->stream));
            c->launches += c->graphLaunches;
        }
        else if ((rc = do_iteration (c))) return rc;
    }
    MFB_CUDA (cudaEventRecord (c->evStop[4], c->stream));
    MFB_CUDA (cudaEventSynchronize (c->evStop[4]));
    MFB_CUDA (cudaEventElapsedTime (ms, c->evStart[4], c->evStop[4]));
    c->stageRan[4] = true;
    return p2p_status (c);
}

// Host builder of the locality-blocked colouring, exposed for the CPU tests (no GPU): elemOrder[nbElem], and the three
// index arrays with their sizes in counts[4] = {blocks, block colours, max local colours, localStart entries};
// launchStart holds up to 65 ints, localIndex nbElem + 2, localStart 2 * nbElem + 2 (upper bounds).
extern "C" int mfb_block_coloring (const int *elemToNode, int nbElem, int nbNodes, const double *coord, int blockElems,
                                   int *elemOrder, int *launchStart, int *localIndex, int *localStart, int counts[4])
{
    if ((nbElem > 0 && (!elemToNode || !coord)) || !elemOrder || !launchStart || !localIndex || !localStart || !counts) return fail (MFB_ERR_ARG, "mfb_block_coloring: NULL argument");
    for (int64_t q = 0; q < (int64_t)nbElem * 4; q++) if (elemToNode[q] < 1 || elemToNode[q] > nbNodes) return fail (MFB_ERR_ARG, "mfb_block_coloring: node id out of range");
    BlockColoring bc;
    const int rc = build_block_coloring (elemToNode, nbElem, nbNodes, coord, blockElems > 0 ? blockElems : 1024, bc);
    if (rc) return fail (MFB_ERR_COLORS, rc == -1 ? "mfb_block_coloring: a block needs more than 128 local colours" : "mfb_block_coloring: the blocks need more than 64 colours");
    std::copy (bc.elemOrder.begin (), bc.elemOrder.end (), elemOrder);
    std::copy (bc.launchStart.begin (), bc.launchStart.end (), launchStart);
    std::copy (bc.localIndex.begin (), bc.localIndex.end (), localIndex);
    std::copy (bc.localStart.begin (), bc.localStart.end (), localStart);
    counts[0] = bc.nbBlocks; counts[1] = bc.nbBlockColors; counts[2] = bc.maxLocalColors; counts[3] = (int)bc.localStart.size ();
    return MFB_OK;
}

extern "C" int mfb_tile_plan_selfcheck (const mfb_problem *p, int tileRows, int tileElems, int64_t stats[6])
{
    if (!p || !stats) return fail (MFB_ERR_ARG, "mfb_tile_plan_selfcheck: NULL argument");
    if (const char *bad = check_problem_ids (p)) return fail (MFB_ERR_ARG, std::string ("mfb_tile_plan_selfcheck: ") + bad);
    TilePlanLimits lim;
    if (tileRows > 0) lim.maxRows = tileRows;
    if (tileElems > 0) lim.maxElems = tileElems;
    lim.maxNodesRef = std::min (65535, std::max (lim.maxElems, 64));
    lim.maxEntries = 65535;
    lim.laplacian = p->operatorID == 0;
    std::vector<uint8_t> isIntf;
    if (p->nbBlocks > 1 && p->nbIntfNodes > 0) {
        isIntf.assign ((size_t)p->nbNodes, 0);
        for (int j = 0; j < p->nbIntfNodes; j++) isIntf[p->intfNodes[j] - 1] = 1;
    }
    TilePlan plan;
    std::string err;
    if (build_tile_plan (p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn, p->coord,
                         isIntf.empty () ? nullptr : isIntf.data (), lim, plan, err) != 0) {
        return fail (MFB_ERR_ARG, "tile plan: " + err);
    }
    if (verify_tile_plan (plan, p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn, err) != 0) {
        return fail (MFB_ERR_STATE, "tile plan self-check: " + err);
    }
    stats[0] = plan.nbTiles; stats[1] = plan.nbTileElems; stats[2] = plan.nbContributions;
    stats[3] = plan.maxRows; stats[4] = plan.maxElems; stats[5] = plan.bytes ();
    return MFB_OK;
}

extern "C" int mfb_ring_plan_selfcheck (const mfb_problem *p, int tileRows, int tileEntries, int64_t stats[12])
{
    if (!p || !stats) return fail (MFB_ERR_ARG, "mfb_ring_plan_selfcheck: NULL argument");
    if (const char *bad = check_problem_ids (p)) return fail (MFB_ERR_ARG, std::string ("mfb_ring_plan_selfcheck: ") + bad);
    RingPlanLimits lim;
    if (tileRows > 0) lim.maxRows = tileRows;
    if (tileEntries > 0) lim.maxEntries = tileEntries;
    std::vector<uint8_t> isIntf;
    if (p->nbBlocks > 1 && p->nbIntfNodes > 0) {
        isIntf.assign ((size_t)p->nbNodes, 0);
        for (int j = 0; j < p->nbIntfNodes; j++) isIntf[p->intfNodes[j] - 1] = 1;
    }
    RingPlan plan;
    std::string err;
    if (build_ring_plan (p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn, p->coord,
                         isIntf.empty () ? nullptr : isIntf.data (), p->checkBounds, lim, plan, err) != 0) {
        return fail (MFB_ERR_ARG, "ring plan: " + err);
    }
    if (verify_ring_plan (plan, p->nbNodes, p->nbElem, p->elemToNode, p->nodeToNodeRow, p->nodeToNodeColumn, p->checkBounds, err) != 0) {
        return fail (MFB_ERR_STATE, "ring plan self-check: " + err);
    }
    stats[0] = plan.nbTiles; stats[1] = plan.nbJobs; stats[2] = plan.nbSymmetricJobs; stats[3] = plan.nbRingSteps;
    stats[4] = plan.nbPaddedSteps; stats[5] = plan.nbBreaks; stats[6] = plan.gatherWavefronts; stats[7] = plan.gatherIdeal;
    stats[8] = plan.slabWriteWavefronts; stats[9] = plan.slabWriteIdeal; stats[10] = (int64_t)plan.blob.size ();
    stats[11] = plan.nbInterfaceTiles;
    return MFB_OK;
}

extern "C" int mfb_host_alloc (void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail (MFB_ERR_ARG, "mfb_host_alloc: bad argument");
    MFB_CUDA (cudaMallocHost (ptr, (size_t)std::max<int64_t> (bytes, 1)));
    return MFB_OK;
}

extern "C" void mfb_host_free (void *ptr) { if (ptr) cudaFreeHost (ptr); }

// ------------------------------------------------------------- layout builders on the GPU

namespace {

struct DevArray {
    void *p = nullptr;
    ~DevArray () { if (p) cudaFree (p); }
    template <class T> T *as () { return static_cast<T*> (p); }
};

int upload_ints (DevArray &dst, const int *src, size_t count)
{
    MFB_CUDA (cudaMalloc (&dst.p, sizeof (int) * std::max<size_t> (count, 1)));
    if (count) MFB_CUDA (cudaMemcpy (dst.p, src, sizeof (int) * count, cudaMemcpyHostToDevice));
    return MFB_OK;
}

// elemToNode on the device plus its node -> element lists
struct DeviceIncidence {
    DevArray elemToNode, index, value;
    int build (const char *who, const int *hostElemToNode, int nbElem, int nbNodes, int device)
    {
        if (mfb_device_count () <= device || device < 0) return fail (MFB_ERR_CUDA, std::string (who) + ": no such CUDA device");
        MFB_CUDA (cudaSetDevice (device));
        int rc = upload_ints (elemToNode, hostElemToNode, (size_t)nbElem * 4);
        if (rc != MFB_OK) return rc;
        MFB_CUDA (cudaMalloc (&index.p, sizeof (int) * ((size_t)nbNodes + 1)));
        MFB_CUDA (cudaMalloc (&value.p, sizeof (int) * std::max<size_t> ((size_t)nbElem * 4, 1)));
        int badIds = 0;
        MFB_CUDA (device_node_to_elem (elemToNode.as<int> (), nbElem, nbNodes, index.as<int> (), value.as<int> (), &badIds, 0));
        if (badIds) return fail (MFB_ERR_ARG, std::string (who) + ": elemToNode holds ids outside [1, nbNodes]");
        return MFB_OK;
    }
};

}  // namespace

extern "C" int mfb_device_create_nodeToNode (const int *elemToNode, int nbElem, int nbNodes,
                                             int *nodeToNodeRow, int *nodeToNodeColumn,
                                             int64_t columnCapacity, int *nbEdgesOut, int device)
{
    if ((!elemToNode && nbElem > 0) || !nodeToNodeRow || !nbEdgesOut || nbElem < 0 || nbNodes < 0) {
        return fail (MFB_ERR_ARG, "mfb_device_create_nodeToNode: bad argument");
    }
    DeviceIncidence inc;
    int rc = inc.build ("mfb_device_create_nodeToNode", elemToNode, nbElem, nbNodes, device);
    if (rc != MFB_OK) return rc;
    DevArray row, col;
    MFB_CUDA (cudaMalloc (&row.p, sizeof (int) * ((size_t)nbNodes + 1)));
    int *dCol = nullptr;
    int64_t total = 0;
    MFB_CUDA (device_build_csr (inc.elemToNode.as<int> (), inc.index.as<int> (), inc.value.as<int> (), nbElem, nbNodes,
                                row.as<int> (), &dCol, &total, 0));
    col.p = dCol;
    *nbEdgesOut = (int)total;
    MFB_CUDA (cudaMemcpy (nodeToNodeRow, row.p, sizeof (int) * ((size_t)nbNodes + 1), cudaMemcpyDeviceToHost));
    if (!nodeToNodeColumn) return MFB_OK;
    if (columnCapacity < total) return fail (MFB_ERR_ARG, "mfb_device_create_nodeToNode: nodeToNodeColumn is too small");
    if (total) MFB_CUDA (cudaMemcpy (nodeToNodeColumn, col.p, sizeof (int) * (size_t)total, cudaMemcpyDeviceToHost));
    return MFB_OK;
}

extern "C" int mfb_device_create_elemToEdge (const int *nodeToNodeRow, const int *nodeToNodeColumn,
                                             const int *elemToNode, int *elemToEdge, int nbElem,
                                             int nbNodes, int device)
{
    if (!nodeToNodeRow || !nodeToNodeColumn || !elemToNode || !elemToEdge || nbElem < 0 || nbNodes < 0) {
        return fail (MFB_ERR_ARG, "mfb_device_create_elemToEdge: bad argument");
    }
    if (mfb_device_count () <= device || device < 0) return fail (MFB_ERR_CUDA, "mfb_device_create_elemToEdge: no such CUDA device");
    MFB_CUDA (cudaSetDevice (device));
    for (int64_t k = 0; k < (int64_t)nbElem * 4; k++) {
        if (elemToNode[k] < 1 || elemToNode[k] > nbNodes) return fail (MFB_ERR_ARG, "mfb_device_create_elemToEdge: elemToNode holds ids outside [1, nbNodes]");
    }
    DevArray row, col, conn, out, missing;
    int rc;
    if ((rc = upload_ints (row, nodeToNodeRow, (size_t)nbNodes + 1)) != MFB_OK) return rc;
    if ((rc = upload_ints (col, nodeToNodeColumn, (size_t)nodeToNodeRow[nbNodes])) != MFB_OK) return rc;
    if ((rc = upload_ints (conn, elemToNode, (size_t)nbElem * 4)) != MFB_OK) return rc;
    MFB_CUDA (cudaMalloc (&out.p, sizeof (int) * std::max<size_t> ((size_t)nbElem * 16, 1)));
    MFB_CUDA (cudaMalloc (&missing.p, sizeof (int)));
    MFB_CUDA (cudaMemset (missing.p, 0, sizeof (int)));
    if (nbElem > 0) MFB_CUDA (launch_elem_to_edge (row.as<int> (), col.as<int> (), conn.as<int> (), out.as<int> (), nbElem, missing.as<int> (), 0));
    int nbMissing = 0;
    MFB_CUDA (cudaMemcpy (&nbMissing, missing.p, sizeof (int), cudaMemcpyDeviceToHost));
    if (nbElem > 0) MFB_CUDA (cudaMemcpy (elemToEdge, out.p, sizeof (int) * (size_t)nbElem * 16, cudaMemcpyDeviceToHost));
    if (nbMissing) return fail (MFB_ERR_ARG, "mfb_device_create_elemToEdge: a node pair is missing from the CSR");
    return MFB_OK;
}

extern "C" int mfb_device_coloring_creation (const int *elemToNode, int nbElem, int nbNodes,
                                             int *colorPart, int *colorToElem, int *colorPerm,
                                             int *nbTotalColors, int device)
{
    if (!elemToNode || !colorPart || !colorToElem || !colorPerm || !nbTotalColors || nbElem < 0 || nbNodes < 0) {
        return fail (MFB_ERR_ARG, "mfb_device_coloring_creation: bad argument");
    }
    DeviceIncidence inc;
    int rc = inc.build ("mfb_device_coloring_creation", elemToNode, nbElem, nbNodes, device);
    if (rc != MFB_OK) return rc;
    DevArray part, perm;
    MFB_CUDA (cudaMalloc (&part.p, sizeof (int) * std::max<size_t> (nbElem, 1)));
    MFB_CUDA (cudaMalloc (&perm.p, sizeof (int) * std::max<size_t> (nbElem, 1)));
    int colors = 0;
    MFB_CUDA (device_color_elements (inc.elemToNode.as<int> (), inc.index.as<int> (), inc.value.as<int> (), nbElem, nbNodes,
                                     part.as<int> (), perm.as<int> (), colorToElem, &colors, 0));
    if (colors == -1) return fail (MFB_ERR_COLORS, "Error: Not enough colors.");
    if (colors < 0) return fail (MFB_ERR_ARG, "mfb_device_coloring_creation: an element names the same node twice");
    if (nbElem > 0) {
        MFB_CUDA (cudaMemcpy (colorPart, part.p, sizeof (int) * (size_t)nbElem, cudaMemcpyDeviceToHost));
        MFB_CUDA (cudaMemcpy (colorPerm, perm.p, sizeof (int) * (size_t)nbElem, cudaMemcpyDeviceToHost));
    }
    *nbTotalColors = colors;
    return MFB_OK;
}
