// RING path: write-once assembly (+ fused preconditioner) that walks the ring of elements around
// every mesh edge straight from the node coordinates (host/ring_plan.h, csrc/ring_math.h).
//
// A lane owns a mesh EDGE {i, j}: it keeps x_i and d = x_j - x_i in registers, loads ONE node (24 bytes)
// per element of the ring, rebuilds the two gradients from the coordinates (two cross products, one
// reciprocal) and accumulates A_ij; an edge inside the tile yields both K_ij and K_ji = K_ij^T.  There are no
// coefficient planes and no diagonal pass: the diagonal block of a row is minus the sum of the row's
// off-diagonal blocks (element matrices have zero row sums).
//
// Second version (round 2), shaped by the ncu capture of the first (profiles/r2_ring_ela_ncu_full.txt: a
// quarter of the warp samples parked at two block barriers per tile, 230 instructions per three rows in a
// write-out that copied the slab to global memory entry by entry, 9 % in a separate preconditioner pass):
//   * the slab is a byte image of the tile's CSR rows and there are TWO of them: tile t's rows leave for
//     global memory as TMA bulk stores (cp.async.bulk.global.shared::cta, one per run of rows that are
//     consecutive in the matrix) while the job phase of tile t+1 already fills the other slab — ONE block
//     barrier per tile (between the job phase and the write-out), no waiting for stores;
//   * the write-out only sums: three rows per warp side by side, nine lanes each; the diagonal block goes
//     into its slab slot, and the same nine lanes mask and invert it on the spot (cofactors through warp
//     shuffles, invert3_adj) and write the preconditioner block coalesced — no second pass, no staging;
//   * everything a tile needs arrives ahead of time: the plan HEAD three tiles ahead (TMA, four buffers),
//     the TAIL two tiles ahead (TMA, two buffers: with a write-out this short a tail requested at the
//     barrier would not be in when the next job phase starts), the node coordinates one tile ahead
//     (cp.async, two plane sets).
// Per tile:  [job phase: one lane per edge, 32 jobs per warp batch]  barrier  [write-out: row sums,
// diagonal + preconditioner blocks, bulk stores].  Every CSR entry is written exactly once; the summation
// order is fixed by the plan.
#include "kernels.cuh"
#include "device_math.cuh"
#include "ring_math.h"

namespace mfb {

namespace {

static_assert (sizeof (RingTileHeader) == 48 && sizeof (RingRow) == 16 && sizeof (RingBatch) == 8,
               "plan records are copied to the device verbatim");

constexpr int kRingHeadBuffers = 4;

// Byte offsets of the kernel's shared-memory sections, computed once on the host and passed as kernel
// arguments (constant bank): the compiler otherwise rebuilds them from the plan maxima inside the loops.
struct RingSmemLayout {
    unsigned headBytes;          // size of one head buffer; head b at b * headBytes
    unsigned tail, tailBytes, planes, slab, slabBytes, bars, total;
};

struct RingArgs {
    DeviceRingPlan plan;
    RingSmemLayout smem;
    const double *coord;
    double *values;
    double *prec;
    int fusePrec;
    int firstTile, lastTile;      // [firstTile, lastTile)
};

// tileOffset entries: byte offset of the record in the low 48 bits, (head bytes / 16) above
__device__ __forceinline__ uint64_t ring_record_offset (uint64_t packed) { return packed & 0xFFFFFFFFFFFFull; }
__device__ __forceinline__ unsigned ring_record_head_bytes (uint64_t packed) { return (unsigned)(packed >> 48) << 4; }

#ifdef MFB_RING_HOST_EMULATION
// tools/ring_kernel_host.cc compiles this file with g++ and runs the kernel below on host threads
// (tools/cuda_cta_emulation.h): the PTX helpers become their emulated counterparts.
inline void ring_mbar_init (uint64_t *bar, unsigned) { cta_emu::mbar_init (bar); }
inline void ring_mbar_expect_tx (uint64_t *bar, unsigned bytes) { cta_emu::mbar_expect_tx (bar, bytes); }
inline void ring_mbar_arrive (uint64_t *bar) { cta_emu::mbar_expect_tx (bar, 0); }
inline void ring_mbar_wait (uint64_t *bar, unsigned parity) { cta_emu::mbar_wait (bar, parity); }
inline void ring_bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar) { cta_emu::bulk_load (dst, src, bytes, bar); }
inline void ring_bulk_store (void *dst, const void *src, unsigned bytes) { cta_emu::bulk_store (dst, src, bytes); }
inline void ring_bulk_commit () {}
inline void ring_bulk_wait_read () {}
inline void ring_bulk_wait_all () {}
inline void ring_fence_async () {}
inline void ring_cp_async_f64 (double *dst, const double *src) { *dst = *src; }
inline void ring_cp_async_wait_all () {}
#else
__device__ __forceinline__ unsigned ring_smem_u32 (const void *p) { return (unsigned)__cvta_generic_to_shared (p); }

__device__ __forceinline__ void ring_mbar_init (uint64_t *bar, unsigned count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ring_smem_u32 (bar)), "r"(count) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void ring_mbar_expect_tx (uint64_t *bar, unsigned bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ring_smem_u32 (bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void ring_mbar_arrive (uint64_t *bar)
{
    asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ring_smem_u32 (bar)) : "memory");
}

// Bounded wait: a protocol bug must trap instead of hanging the device.
__device__ __forceinline__ void ring_mbar_wait (uint64_t *bar, unsigned parity)
{
    unsigned done = 0;
    for (long spin = 0; spin < (1l << 22); spin++) {
        asm volatile (
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(ring_smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap ();
}

// TMA bulk copy global -> shared, completion counted on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void ring_bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(ring_smem_u32 (dst)), "l"(src), "r"(bytes), "r"(ring_smem_u32 (bar)) : "memory");
}

// TMA bulk copy shared -> global (both addresses and the size multiples of 16 bytes), tracked by the
// issuing thread's bulk async-group.
__device__ __forceinline__ void ring_bulk_store (void *dst, const void *src, unsigned bytes)
{
    asm volatile ("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                  :: "l"(dst), "r"(ring_smem_u32 (src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_bulk_commit () { asm volatile ("cp.async.bulk.commit_group;" ::: "memory"); }
// this thread's bulk stores have READ their shared-memory source (the slab may be overwritten)
__device__ __forceinline__ void ring_bulk_wait_read () { asm volatile ("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ring_bulk_wait_all () { asm volatile ("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the TMA unit)
__device__ __forceinline__ void ring_fence_async () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void ring_cp_async_f64 (double *dst, const double *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(ring_smem_u32 (dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ring_cp_async_wait_all () { asm volatile ("cp.async.wait_all;" ::: "memory"); }

#endif

__host__ __device__ __forceinline__ unsigned ring_align128 (unsigned x) { return (x + 127u) & ~127u; }

// doubles per coordinate plane: tile-local node ids are one byte (kRingMaxNodes = 254)
constexpr int kRingPlane = 256;

// THREADS / MINB: 384 threads, two CTAs per SM (tiles of 12 warp batches and 12 row groups) or 256
// threads, three CTAs per SM (tiles of 8 / 8) — 24 warps per SM either way.
inline RingSmemLayout ring_smem_layout (int operatorID, const DeviceRingPlan &plan)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    RingSmemLayout L;
    L.headBytes = ring_align128 (plan.maxHeadBytes);
    L.tail = kRingHeadBuffers * L.headBytes;
    L.tailBytes = ring_align128 (plan.maxTailBytes);
    L.planes = L.tail + 2 * L.tailBytes;
    L.slab = L.planes + 2 * 3 * kRingPlane * (unsigned)sizeof (double);
    L.slabBytes = ring_align128 ((unsigned)plan.maxEntries * opDim * (unsigned)sizeof (double));
    L.bars = L.slab + 2 * L.slabBytes;
    L.total = L.bars + (kRingHeadBuffers + 2) * (unsigned)sizeof (uint64_t);
    return L;
}

template <int OPDIM, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ring_assembly_kernel (const RingArgs args)
{
#ifdef MFB_RING_HOST_EMULATION
    unsigned char *smemRaw = cta_emu::dynamic_smem ();
#else
    extern __shared__ __align__(128) unsigned char smemRaw[];
#endif
    const DeviceRingPlan &P = args.plan;
    const int tid = threadIdx.x, nThreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nWarps = nThreads >> 5;

    // shared memory: [head 0..3][tail 0, 1][X Y Z of even tiles][X Y Z of odd tiles][slab 0][slab 1][6 mbarriers]
    const RingSmemLayout &L = args.smem;
    const unsigned headBytes = L.headBytes;
    constexpr int planeStride = kRingPlane;                           // planes start on a 128-byte line: bank = id mod 16;
                                                                      // a constant, so that Y and Z are immediate offsets from X
    unsigned char *sHead0 = smemRaw, *sTail0 = smemRaw + L.tail;
    double *planes0 = reinterpret_cast<double*> (smemRaw + L.planes);
    double *slab0 = reinterpret_cast<double*> (smemRaw + L.slab);
    const unsigned slabDoubles = L.slabBytes / (unsigned)sizeof (double);
    uint64_t *bars = reinterpret_cast<uint64_t*> (smemRaw + L.bars);
    uint64_t *headFull = bars, *tailFull = bars + kRingHeadBuffers;   // headFull[4], tailFull[2]

    const int firstTile = args.firstTile + blockIdx.x, tileStep = gridDim.x;
    const int nbMine = firstTile < args.lastTile ? (args.lastTile - firstTile + tileStep - 1) / tileStep : 0;
    if (nbMine == 0) return;

    if (tid == 0) {
        for (int b = 0; b < kRingHeadBuffers; b++) ring_mbar_init (headFull + b, 1);
        ring_mbar_init (tailFull, 1); ring_mbar_init (tailFull + 1, 1);
    }
    __syncthreads ();

    auto head_of = [&] (int k) { return sHead0 + (unsigned)(k & (kRingHeadBuffers - 1)) * headBytes; };
    auto fetch_head = [&] (uint64_t packed, int k) {                  // thread 0 only
        const unsigned bytes = ring_record_head_bytes (packed);
        uint64_t *bar = headFull + (k & (kRingHeadBuffers - 1));
        ring_mbar_expect_tx (bar, bytes);
        ring_bulk_load (head_of (k), P.blob + ring_record_offset (packed), bytes, bar);
    };
    auto wait_head = [&] (int k) { ring_mbar_wait (headFull + (k & (kRingHeadBuffers - 1)), (unsigned)(k / kRingHeadBuffers) & 1u); };
    auto fetch_tail = [&] (int k) {                                   // thread 0 only, head k has landed
        const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head_of (k));
        const unsigned bytes = h.blobBytes - h.headBytes;
        uint64_t *bar = tailFull + (k & 1);
        if (bytes) {
            ring_mbar_expect_tx (bar, bytes);
            ring_bulk_load (sTail0 + (unsigned)(k & 1) * L.tailBytes, P.blob + h.selfOffset + h.headBytes, bytes, bar);
        }
        else ring_mbar_arrive (bar);                                    // a tile without jobs: the phase completes at once
    };
    auto gather_coords = [&] (const unsigned char *head, double *planes) {   // all threads, asynchronous
        const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head);
        const int *nodes = reinterpret_cast<const int*> (head + h.offNodes);
        for (int n = tid; n < h.nbNodes; n += nThreads) {
            const double *q = args.coord + (size_t)nodes[n] * 3;
            ring_cp_async_f64 (planes + n, q); ring_cp_async_f64 (planes + planeStride + n, q + 1);
            ring_cp_async_f64 (planes + 2 * planeStride + n, q + 2);
        }
    };

    // prologue: heads of this CTA's first three tiles, tail and coordinates of the first.  Thread 0 keeps the
    // packed offset of the next head to fetch in a register, loaded a whole tile before it is needed.
    uint64_t offAhead = 0;
    if (tid == 0) {
        for (int k = 0; k < 3 && k < nbMine; k++) fetch_head (P.tileOffset[firstTile + k * tileStep], k);
        if (nbMine > 3) offAhead = P.tileOffset[firstTile + 3 * tileStep];
    }
    wait_head (0);
    if (tid == 0) {
        fetch_tail (0);
        if (nbMine > 1) { wait_head (1); fetch_tail (1); }
    }
    gather_coords (sHead0, planes0);
    ring_cp_async_wait_all ();
    __syncthreads ();

    for (int k = 0; k < nbMine; k++) {
        const unsigned char *sHead = head_of (k);
        const RingTileHeader &hdr = *reinterpret_cast<const RingTileHeader*> (sHead);
        const int nbRows = hdr.nbRows, nbBatches = hdr.nbBatches;
        const RingRow *sRows = reinterpret_cast<const RingRow*> (sHead + sizeof (RingTileHeader));
        const unsigned char *sTail = sTail0 + (unsigned)(k & 1) * L.tailBytes;
        const RingBatch *batches = reinterpret_cast<const RingBatch*> (sTail);
        const uint64_t *jobs = reinterpret_cast<const uint64_t*> (sTail + (hdr.offJobs - hdr.headBytes));
        const uint64_t *codes = reinterpret_cast<const uint64_t*> (sTail + (hdr.offCodes - hdr.headBytes));
        const double *sX = planes0 + (k & 1) * (3 * planeStride), *sY = sX + planeStride, *sZ = sY + planeStride;
        double *slab = slab0 + (k & 1) * slabDoubles;

        // ---- 0. the next tile's coordinates start travelling into the other plane set -------------------
        // (its head came in two tiles ago; the readers of that plane set, the jobs of tile k - 1, are behind
        // the previous block barrier)
        if (k + 1 < nbMine) {
            wait_head (k + 1);
            gather_coords (head_of (k + 1), planes0 + ((k + 1) & 1) * (3 * planeStride));
        }

        // ---- 1. job phase: one lane per mesh edge ----------------------------------------------
        ring_mbar_wait (tailFull + (k & 1), (unsigned)(k >> 1) & 1u);
        for (int b = warp; b < nbBatches; b += nWarps) {
            const RingBatch rb = batches[b];
            const uint64_t job = jobs[b * 32 + lane];
            const int i = (int)(job & 0xFF), j = (int)((job >> 8) & 0xFF);
            const int sIJ = (int)((job >> 16) & 0xFFFF), sJI = (int)((job >> 32) & 0xFFFF);
            const double xi[3] = {sX[i], sY[i], sZ[i]};
            const double d[3] = {sX[j] - xi[0], sY[j] - xi[1], sZ[j] - xi[2]};
            const int len = (int)(job >> 48), nbSteps = rb.nbSteps;
            double acc[OPDIM], u[3] = {0.0, 0.0, 0.0};
            #pragma unroll
            for (int q = 0; q < OPDIM; q++) acc[q] = 0.0;
            const uint64_t *cw = codes + rb.codeBase + lane;
            uint64_t word = 0;
            if (rb.flags == 0) {
                // regular batch (every mesh without non-manifold edges): one chain per job.  Byte 0 names its
                // first node, every further byte adds one element; bytes beyond the job's length name a valid
                // node and their contribution is masked — no branch inside the step.
                if (nbSteps > 0) {
                    word = cw[0];
                    const int id = (int)(word & 0xFF);
                    u[0] = sX[id] - xi[0]; u[1] = sY[id] - xi[1]; u[2] = sZ[id] - xi[2];
                }
                // two steps per trip, u and w changing roles: no register copies between steps
                double w[3];
                auto load_node = [&] (int s, double v[3]) {
                    if ((s & 7) == 0) word = cw[(s >> 3) * 32]; else word >>= 8;
                    const int id = (int)(word & 0xFF);
                    v[0] = sX[id] - xi[0]; v[1] = sY[id] - xi[1]; v[2] = sZ[id] - xi[2];
                };
                int s = 1;
                for (; s + 1 < nbSteps; s += 2) {
                    load_node (s, w);
                    ring_accumulate<OPDIM> (d, u, w, acc, s < len);
                    load_node (s + 1, u);
                    ring_accumulate<OPDIM> (d, w, u, acc, s + 1 < len);
                }
                if (s < nbSteps) {
                    load_node (s, w);
                    ring_accumulate<OPDIM> (d, u, w, acc, s < len);
                }
            }
            else {
                // general batch: chains separated by breaks
                bool have = false;
                for (int s = 0; s < nbSteps; s++) {
                    if ((s & 7) == 0) word = cw[(s >> 3) * 32]; else word >>= 8;
                    const int id = (int)(word & 0xFF);
                    if (s >= len) continue;
                    if (id == kRingBreak) { have = false; continue; }
                    const double w[3] = {sX[id] - xi[0], sY[id] - xi[1], sZ[id] - xi[2]};
                    if (have) ring_accumulate<OPDIM> (d, u, w, acc);
                    u[0] = w[0]; u[1] = w[1]; u[2] = w[2];
                    have = true;
                }
            }
            if (sIJ != 0xFFFF) {
                if (OPDIM == 1) {
                    slab[sIJ] = acc[0];
                    if (sJI != 0xFFFF) slab[sJI] = acc[0];
                }
                else {
                    double blk[9];
                    ring_block (acc, blk);
                    double *dst = slab + sIJ * 9;
                    #pragma unroll
                    for (int q = 0; q < 9; q++) dst[q] = blk[q];
                    if (sJI != 0xFFFF) {                    // K_ji = K_ij^T
                        double *dstT = slab + sJI * 9;
                        #pragma unroll
                        for (int q = 0; q < 9; q++) dstT[ring_transposed (q)] = blk[q];
                    }
                }
            }
        }
        if (OPDIM == 9) ring_fence_async ();   // this thread's slab stores, for the bulk stores of the write-out
        ring_cp_async_wait_all ();             // this thread's share of the next tile's coordinates is in
        ring_bulk_wait_read ();                // this thread's bulk stores of the previous tile have read the other slab
        __syncthreads ();      // the slab is complete; tail and coordinates of this tile are dead; the other slab is free

        // ---- 2. the tail two tiles ahead and the head three tiles ahead start travelling (TMA) ---------
        if (tid == 0) {
            if (k + 2 < nbMine) { wait_head (k + 2); fetch_tail (k + 2); }     // into the buffer of this tile's tail
            if (k + 3 < nbMine) {
                fetch_head (offAhead, k + 3);            // buffer of tile k - 1, whose write-out ended before the barrier
                if (k + 4 < nbMine) offAhead = P.tileOffset[firstTile + (k + 4) * tileStep];
            }
        }

        // ---- 3. write-out: the diagonal entry of a row is minus the sum of the row's run ------------
        if (OPDIM == 1) {
            // Laplacian: one lane per row walks its entries (rows are short and the whole matrix is an eighth of
            // the elasticity one: 8-byte stores to 32 different rows per instruction are affordable).
            for (int r = warp * 32 + lane; r < nbRows; r += nWarps * 32) {
                const RingRow rr = sRows[r];
                const int len = rr.len, diagOff = rr.diagOff;         // 0xFFFF never equals a position
                double *out = args.values + (size_t)rr.valueStart;
                const double *src = slab + (size_t)rr.localStart;
                double a = 0.0;
                for (int q = 0; q < len; q++) {
                    if (q != diagOff) { const double v = src[q]; a += v; out[q] = v; }
                }
                const double diag = 0.0 - a;
                if (diagOff != 0xFFFF) out[diagOff] = diag;
                if (args.fusePrec) args.prec[rr.node & kRingNodeMask] = rr.node < 0 ? diag : 1.0 / diag;
            }
        }
        else {
            // Elasticity: three consecutive rows per warp side by side, nine lanes each (lanes 0..26); a lane
            // walks the entries of its row and sums its component, so the row sum needs no exchange between
            // lanes.  Lanes 27..29 issue the bulk stores of the three rows.
            const bool worker = lane < 27;
            const int grp = worker ? lane / 9 : lane - 27;            // row of the group this lane works for (3, 4: none)
            const int comp = worker ? lane - 9 * grp : 0;
            const int ca = comp / 3, cb = comp - 3 * ca;              // component (ca, cb) of the 3x3 block
            const int base = worker ? 9 * grp : lane;                 // first lane of the row's nine; idle lanes read themselves
            for (int r0 = warp * kRingStoreGroup; r0 < nbRows; r0 += nWarps * kRingStoreGroup) {
                const int r = r0 + grp;
                const bool rowOk = grp < kRingStoreGroup && r < nbRows;
                int4 rw = make_int4 (0, 0, 0, 0xFFFF);                // node, valueStart, localStart | len << 16, diagOff | segEntries << 16
                if (rowOk) rw = *reinterpret_cast<const int4*> (sRows + r);
                const int node = rw.x, len = (int)((unsigned)rw.z >> 16), diagOff = rw.w & 0xFFFF;
                double *row = slab + (unsigned)(rw.z & 0xFFFF) * 9u;
                double diag = 0.0;
                if (worker) {
                    double *sp = row + comp;
                    if (diagOff != 0xFFFF) sp[diagOff * 9] = 0.0;     // no job writes the diagonal slot: summed as a zero
                    double a = 0.0;
                    #pragma unroll 4
                    for (int q = 0; q < len; q++) a += sp[q * 9];
                    diag = 0.0 - a;
                    if (diagOff != 0xFFFF) sp[diagOff * 9] = diag;
                }
                if (args.fusePrec) {
                    // prec_init + prec_inversion (src/preconditioner.cc:25-87, src/Fortran/elasclpr.f:19-53) by the nine
                    // lanes that hold the block: Dirichlet rows / columns to identity, then cofactor / determinant
                    const bool masked = ((node >> (28 + ca)) | (node >> (28 + cb))) & 1;
                    const double m = masked ? (ca == cb ? 1.0 : 0.0) : diag;
                    const double x1 = __shfl_sync (0xffffffffu, m, base + adj_src (cb + 1, ca + 1));
                    const double x2 = __shfl_sync (0xffffffffu, m, base + adj_src (cb + 2, ca + 2));
                    const double x3 = __shfl_sync (0xffffffffu, m, base + adj_src (cb + 1, ca + 2));
                    const double x4 = __shfl_sync (0xffffffffu, m, base + adj_src (cb + 2, ca + 1));
                    const double m0 = __shfl_sync (0xffffffffu, m, base + ca);      // M(0, ca), used where cb == 0
                    const double t = x3 * x4;
                    const double cof = fma (x1, x2, -t);
                    const double p = m0 * cof;
                    const double p0 = __shfl_sync (0xffffffffu, p, base);
                    const double p1 = __shfl_sync (0xffffffffu, p, base + 3);
                    const double p2 = __shfl_sync (0xffffffffu, p, base + 6);
                    const double det = (p0 + p1) + p2;
                    const bool isInterface = node < 0;
                    const bool invert = worker && rowOk && !isInterface && diagOff != 0xFFFF;
                    const bool regular = fabs (det) > 0.0 && fabs (det) < 1.0e300;
                    double out = isInterface ? diag : m;
                    if (invert && regular) out = cof * ring_rcp (det);
                    if (__any_sync (0xffffffffu, invert && !regular)) {
                        // a singular or non-finite block: the infinities and NaNs of LAPACK's LU, not of the cofactors
                        double blk[9];
                        #pragma unroll
                        for (int q = 0; q < 9; q++) blk[q] = __shfl_sync (0xffffffffu, m, base + q);
                        if (invert && !regular) {
                            invert3_lu (blk);
                            #pragma unroll
                            for (int q = 0; q < 9; q++) if (q == comp) out = blk[q];
                        }
                    }
                    if (worker && rowOk) args.prec[(size_t)(node & kRingNodeMask) * 9 + comp] = out;
                }
                ring_fence_async ();           // the diagonal blocks, for the async proxy
                __syncwarp ();
                const unsigned seg = (unsigned)rw.w >> 16;
                if (!worker && seg) {
                    // one run of rows, consecutive in the matrix and in the slab: 8-byte pieces at the ends where
                    // the run does not start / end on a 16-byte boundary, one bulk store in between
                    double *g = args.values + (size_t)rw.y * 9;
                    const double *s = row;
                    unsigned n = seg * 9u;
                    if (rw.y & 1) { *g = *s; g++; s++; n--; }
                    if (n & 1) { g[n - 1] = s[n - 1]; n--; }
                    if (n) ring_bulk_store (g, s, n * (unsigned)sizeof (double));
                }
                ring_bulk_commit ();
            }
        }
        // no barrier: the next tile's jobs fill the other slab
    }
    ring_bulk_wait_all ();      // shared memory must outlive the bulk stores
}

#ifndef MFB_RING_HOST_EMULATION
template <class K>
cudaError_t ring_opt_in (K kernel)
{
    // The attribute belongs to the kernel, not to a context: always opt in to the device maximum.
    return cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
#endif

}  // namespace

size_t ring_smem_bytes (int operatorID, const DeviceRingPlan &plan)
{
    return ring_smem_layout (operatorID, plan).total;
}

#ifndef MFB_RING_HOST_EMULATION
cudaError_t ring_configure (int operatorID)
{
    cudaError_t e;
    if (operatorID == 0) {
        if ((e = ring_opt_in (ring_assembly_kernel<1, 256, 3>)) != cudaSuccess) return e;
        return ring_opt_in (ring_assembly_kernel<1, 384, 2>);
    }
    if ((e = ring_opt_in (ring_assembly_kernel<9, 256, 3>)) != cudaSuccess) return e;
    return ring_opt_in (ring_assembly_kernel<9, 384, 2>);
}

cudaError_t ring_ctas_per_sm (int operatorID, int threads, size_t smemBytes, int *ctas)
{
    if (threads == 384) {
        return operatorID == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor (ctas, ring_assembly_kernel<1, 384, 2>, 384, smemBytes)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor (ctas, ring_assembly_kernel<9, 384, 2>, 384, smemBytes);
    }
    return operatorID == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor (ctas, ring_assembly_kernel<1, 256, 3>, 256, smemBytes)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor (ctas, ring_assembly_kernel<9, 256, 3>, 256, smemBytes);
}

cudaError_t launch_ring (int operatorID, const DeviceRingPlan &plan, int firstTile, int nbTiles, int ctas,
                         int threads, size_t smemBytes, const double *coord, double *values, double *prec,
                         int fusePrec, cudaStream_t stream)
{
    if (nbTiles <= 0) return cudaSuccess;
    RingArgs args;
    args.plan = plan; args.smem = ring_smem_layout (operatorID, plan);
    args.coord = coord; args.values = values; args.prec = prec; args.fusePrec = fusePrec;
    args.firstTile = firstTile; args.lastTile = firstTile + nbTiles;
    const int grid = std::max (1, std::min (ctas, nbTiles));
    if (threads == 384) {
        if (operatorID == 0) ring_assembly_kernel<1, 384, 2><<<grid, 384, smemBytes, stream>>> (args);
        else                 ring_assembly_kernel<9, 384, 2><<<grid, 384, smemBytes, stream>>> (args);
    }
    else {
        if (operatorID == 0) ring_assembly_kernel<1, 256, 3><<<grid, 256, smemBytes, stream>>> (args);
        else                 ring_assembly_kernel<9, 256, 3><<<grid, 256, smemBytes, stream>>> (args);
    }
    return cudaGetLastError ();
}
#endif

}  // namespace mfb
