// RING path: write-once assembly (+ fused preconditioner) that walks the ring of elements around
// every mesh edge straight from the node coordinates (host/ring_plan.h, csrc/ring_math.h).
//
// The TILED kernel is bound by shared-memory bandwidth: 12 coefficients stored per tile element,
// 48 bytes gathered per contribution, every border element recomputed in ~2.2 tiles.  Here a lane
// owns a mesh EDGE {i, j}: it keeps x_i and d = x_j - x_i in registers, loads ONE node (24 bytes)
// per element of the ring, rebuilds the two gradients from the coordinates (two cross products,
// one reciprocal) and accumulates A_ij; an edge inside the tile yields both K_ij and K_ji = K_ij^T.
// There are no coefficient planes and no diagonal pass: the diagonal block of a row is minus the sum
// of the row's off-diagonal blocks (element matrices have zero row sums) and is formed while the row
// streams out of the slab.
//
// Per tile:
//   0. plan record by TMA (cp.async.bulk + mbarrier): the HEAD of tile t+1 (header, row table, node
//      list) is fetched at the start of tile t, the TAIL (batches, jobs, ring codes) and the
//      coordinates (cp.async) of tile t+1 while tile t is in its write-out;
//   1. job phase: one lane per edge, 32 jobs per warp batch, ring codes 8 to a 64-bit word;
//      finished 3x3 blocks go to the tile-wide slab at the slots of their CSR entries (entry stride
//      80 bytes, so a block leaves as four 128-bit stores and one 64-bit store);
//   2. write-out: three consecutive rows per warp, ten lanes each (9 components + one idle lane): a lane
//      walks its row entry by entry, copies its component to global memory (72 contiguous bytes per
//      group and instruction) and sums it, then stores its component of the diagonal entry; row starts
//      in the slab are padded so that the three pieces read together fall into disjoint banks;
//   3. fused mode: the diagonal blocks wait in shared memory until the tile is done; the last two warps
//      of the CTA (they get the fewest job batches) then mask / invert them into prec, one lane per
//      row, at the head of the next tile (prec_init + prec_inversion, src/preconditioner.cc:25-87,
//      src/Fortran/elasclpr.f:19-53) — two warps with full lanes instead of every warp with four.
// Every CSR entry is written exactly once by a plain store; the summation order is fixed by the plan.
#include "kernels.cuh"
#include "device_math.cuh"
#include "ring_math.h"

namespace mfb {

namespace {

static_assert (sizeof (RingTileHeader) == 32 && sizeof (RingRow) == 16 && sizeof (RingBatch) == 8,
               "plan records are copied to the device verbatim");

// Byte offsets of the kernel's shared-memory sections, computed once on the host and passed as kernel
// arguments (constant bank): the compiler otherwise rebuilds them from the plan maxima inside the loops.
struct RingSmemLayout {
    unsigned headBytes;          // size of one head buffer; head 0 at 0, head 1 at headBytes
    unsigned tail, planes, slab, diag, meta, bars, total;
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

__device__ __forceinline__ void ring_cp_async_f64 (double *dst, const double *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(ring_smem_u32 (dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ring_cp_async_wait_all () { asm volatile ("cp.async.wait_all;" ::: "memory"); }

#endif

__host__ __device__ __forceinline__ unsigned ring_align128 (unsigned x) { return (x + 127u) & ~127u; }

// doubles per coordinate plane: tile-local node ids are one byte (kRingMaxNodes = 254)
constexpr int kRingPlane = 256;

// doubles per slab entry: an elasticity block is padded to 80 bytes so that it is 16-byte aligned
__host__ __device__ constexpr int ring_slab_stride (int opDim) { return opDim == 9 ? 10 : 1; }

// THREADS / MINB: 256 threads, three CTAs per SM (default) or 384 threads, two CTAs per SM (larger tiles:
// mfb_options.threads = 384 with tileRows / tileElems raised, e.g. 54 / 960) — 24 warps per SM either way.
inline RingSmemLayout ring_smem_layout (int operatorID, const DeviceRingPlan &plan)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    RingSmemLayout L;
    L.headBytes = ring_align128 (plan.maxHeadBytes);
    L.tail = 2 * L.headBytes;
    L.planes = L.tail + ring_align128 (plan.maxTailBytes);
    L.slab = L.planes + 3 * kRingPlane * (unsigned)sizeof (double);
    L.diag = L.slab + ring_align128 ((unsigned)plan.maxEntries * ring_slab_stride (opDim) * (unsigned)sizeof (double));
    L.meta = L.diag + (((unsigned)plan.maxRows * opDim * (unsigned)sizeof (double) + 15u) & ~15u);
    L.bars = L.meta + (((unsigned)plan.maxRows * (unsigned)sizeof (int) + 15u) & ~15u);
    L.total = L.bars + 3 * (unsigned)sizeof (uint64_t);
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

    // shared memory: [head 0][head 1][tail][X Y Z][slab][diagonal blocks][row tags][3 mbarriers]
    const RingSmemLayout &L = args.smem;
    const unsigned headBytes = L.headBytes;
    constexpr int planeStride = kRingPlane;                           // planes start on a 128-byte line: bank = id mod 16;
                                                                      // a constant, so that Y and Z are immediate offsets from X
    constexpr int SLAB = ring_slab_stride (OPDIM);
    unsigned char *sHead0 = smemRaw, *sTail = smemRaw + L.tail;
    double *sX = reinterpret_cast<double*> (smemRaw + L.planes), *sY = sX + planeStride, *sZ = sY + planeStride;
    double *slab = reinterpret_cast<double*> (smemRaw + L.slab);
    double *sDiag = reinterpret_cast<double*> (smemRaw + L.diag);
    int *sMeta = reinterpret_cast<int*> (smemRaw + L.meta);           // node | interface << 31 | hasDiag << 30
    uint64_t *bars = reinterpret_cast<uint64_t*> (smemRaw + L.bars);
    uint64_t *headFull = bars, *tailFull = bars + 2;                  // headFull[2], tailFull

    if (tid == 0) { ring_mbar_init (headFull, 1); ring_mbar_init (headFull + 1, 1); ring_mbar_init (tailFull, 1); }
    __syncthreads ();

    const int firstTile = args.firstTile + blockIdx.x, tileStep = gridDim.x;
    auto fetch_head = [&] (uint64_t packed, int k) {                  // thread 0 only
        const unsigned bytes = ring_record_head_bytes (packed);
        ring_mbar_expect_tx (headFull + (k & 1), bytes);
        ring_bulk_load (sHead0 + (k & 1) * headBytes, P.blob + ring_record_offset (packed), bytes, headFull + (k & 1));
    };
    auto fetch_tail = [&] (uint64_t packed, const unsigned char *head) {   // thread 0 only, `head` has landed
        const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head);
        const unsigned bytes = h.blobBytes - h.headBytes;
        if (bytes) {
            ring_mbar_expect_tx (tailFull, bytes);
            ring_bulk_load (sTail, P.blob + ring_record_offset (packed) + h.headBytes, bytes, tailFull);
        }
        else ring_mbar_arrive (tailFull);                               // a tile without jobs: the phase completes at once
    };
    // Node coordinates of a tile: thread n loads node n (three 8-byte loads from global memory) and later
    // stores it into the planes with the warp's 32 nodes side by side — two wavefronts per store.  (An 8-byte
    // cp.async per coordinate costs one shared-memory wavefront EACH when the data come back: ~20 M of the
    // 77 M wavefronts of an EIB iteration in the round-1 kernel, profiles/r2_ring_ela_ncu_full.txt.)
    static_assert (THREADS >= kRingMaxNodes, "one node per thread");
    auto load_coords = [&] (const unsigned char *head, double c[3]) {
        const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head);
        const int *nodes = reinterpret_cast<const int*> (head + h.offNodes);
        if (tid < h.nbNodes) {
            const double *q = args.coord + (size_t)nodes[tid] * 3;
            c[0] = __ldg (q); c[1] = __ldg (q + 1); c[2] = __ldg (q + 2);
        }
    };
    auto store_coords = [&] (const unsigned char *head, const double c[3]) {
        const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head);
        if (tid < h.nbNodes) { sX[tid] = c[0]; sY[tid] = c[1]; sZ[tid] = c[2]; }
    };

    // thread 0 keeps the packed offsets of this tile and the next in registers and loads the one after
    // that a whole tile ahead, so that issuing the copies never waits on global memory
    uint64_t offCur = 0, offNext = 0;
    if (tid == 0 && firstTile < args.lastTile) {
        offCur = P.tileOffset[firstTile];
        if (firstTile + tileStep < args.lastTile) offNext = P.tileOffset[firstTile + tileStep];
    }
    // prologue: head, tail and coordinates of this CTA's first tile
    if (firstTile < args.lastTile) {
        if (tid == 0) fetch_head (offCur, 0);
        ring_mbar_wait (headFull, 0);
        if (tid == 0) fetch_tail (offCur, sHead0);
        double c[3] = {0.0, 0.0, 0.0};
        load_coords (sHead0, c);
        store_coords (sHead0, c);
    }
    __syncthreads ();

    // Fused preconditioner of one finished tile: its diagonal blocks and row tags are in sDiag / sMeta.
    // Run by the last two warps, one lane per row.
    auto prec_pass = [&] (int nbRowsDone) {
        for (int r = (warp - (nWarps - 2)) * 32 + lane; r < nbRowsDone; r += 64) {
            const int meta = sMeta[r];                                // RingRow::node | hasDiag << 27
            const int node = meta & kRingNodeMask;
            const bool isInterface = meta < 0, hasDiag = (meta & (1 << 27)) != 0;
            if (OPDIM == 1) {
                const double dgl = sDiag[r];
                args.prec[node] = isInterface ? dgl : 1.0 / dgl;
            }
            else {
                double b[9];
                #pragma unroll
                for (int q = 0; q < 9; q++) b[q] = sDiag[r * 9 + q];
                if (!isInterface) {
                    mask_block (b, (meta >> 28) & 1, (meta >> 29) & 1, (meta >> 30) & 1);
                    if (hasDiag) {
                        // cofactors over the determinant; a singular or non-finite block takes LAPACK's LU instead
                        // (its infinities and NaNs are the reference's)
                        double inv[9];
                        if (invert3_adj (b, inv, [] (double x) { return ring_rcp (x); })) {
                            #pragma unroll
                            for (int q = 0; q < 9; q++) b[q] = inv[q];
                        }
                        else invert3_lu (b);
                    }
                }
                double *dst = args.prec + (size_t)node * 9;
                #pragma unroll
                for (int q = 0; q < 9; q++) dst[q] = b[q];
            }
        }
    };
    const bool precWarp = args.fusePrec && warp >= nWarps - 2;
    int rowsDone = 0;             // rows of the previous tile whose preconditioner blocks are still to be written

    int k = 0;
    for (int tile = firstTile; tile < args.lastTile; tile += tileStep, k++) {
        const unsigned char *sHead = sHead0 + (k & 1) * headBytes;
        const RingTileHeader &hdr = *reinterpret_cast<const RingTileHeader*> (sHead);
        const bool hasNext = tile + tileStep < args.lastTile;
        // ---- 0. the next tile's head starts travelling ----------------------------------------
        uint64_t offAfter = 0;
        if (tid == 0) {
            if (tile + 2 * tileStep < args.lastTile) offAfter = P.tileOffset[tile + 2 * tileStep];   // used next iteration
            if (hasNext) fetch_head (offNext, k + 1);
        }
        const int nbRows = hdr.nbRows, nbBatches = hdr.nbBatches;
        const RingRow *sRows = reinterpret_cast<const RingRow*> (sHead + sizeof (RingTileHeader));
        const RingBatch *batches = reinterpret_cast<const RingBatch*> (sTail);
        const uint64_t *jobs = reinterpret_cast<const uint64_t*> (sTail + (hdr.offJobs - hdr.headBytes));
        const uint64_t *codes = reinterpret_cast<const uint64_t*> (sTail + (hdr.offCodes - hdr.headBytes));

        // ---- 0b. preconditioner blocks of the previous tile (its write-out ended at the block barrier) ----
        if (precWarp) prec_pass (rowsDone);
        rowsDone = nbRows;

        // ---- 1. job phase: one lane per mesh edge ----------------------------------------------
        ring_mbar_wait (tailFull, k & 1);
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
                auto load_node = [&] (int k, double v[3]) {
                    if ((k & 7) == 0) word = cw[(k >> 3) * 32]; else word >>= 8;
                    const int id = (int)(word & 0xFF);
                    v[0] = sX[id] - xi[0]; v[1] = sY[id] - xi[1]; v[2] = sZ[id] - xi[2];
                };
                int k = 1;
                for (; k + 1 < nbSteps; k += 2) {
                    load_node (k, w);
                    ring_accumulate<OPDIM> (d, u, w, acc, k < len);
                    load_node (k + 1, u);
                    ring_accumulate<OPDIM> (d, w, u, acc, k + 1 < len);
                }
                if (k < nbSteps) {
                    load_node (k, w);
                    ring_accumulate<OPDIM> (d, u, w, acc, k < len);
                }
            }
            else {
                // general batch: chains separated by breaks
                bool have = false;
                for (int k = 0; k < nbSteps; k++) {
                    if ((k & 7) == 0) word = cw[(k >> 3) * 32]; else word >>= 8;
                    const int id = (int)(word & 0xFF);
                    if (k >= len) continue;
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
                    double2 *dst = reinterpret_cast<double2*> (slab + sIJ * SLAB);
                    dst[0] = make_double2 (blk[0], blk[1]); dst[1] = make_double2 (blk[2], blk[3]);
                    dst[2] = make_double2 (blk[4], blk[5]); dst[3] = make_double2 (blk[6], blk[7]);
                    slab[sIJ * SLAB + 8] = blk[8];
                    if (sJI != 0xFFFF) {                    // K_ji = K_ij^T
                        double2 *dstT = reinterpret_cast<double2*> (slab + sJI * SLAB);
                        dstT[0] = make_double2 (blk[0], blk[3]); dstT[1] = make_double2 (blk[6], blk[1]);
                        dstT[2] = make_double2 (blk[4], blk[7]); dstT[3] = make_double2 (blk[2], blk[5]);
                        slab[sJI * SLAB + 8] = blk[8];
                    }
                }
            }
        }
        __syncthreads ();      // the slab is complete; tail and coordinates of this tile are dead

        // ---- 2. next tile's tail (TMA) and coordinates (plain loads) start travelling -------------
        double nextCoord[3] = {0.0, 0.0, 0.0};
        if (hasNext) {
            const unsigned char *nextHead = sHead0 + ((k + 1) & 1) * headBytes;
            ring_mbar_wait (headFull + ((k + 1) & 1), ((k + 1) >> 1) & 1);
            if (tid == 0) fetch_tail (offNext, nextHead);
            load_coords (nextHead, nextCoord);               // stored at the end of the write-out
        }

        // ---- 3. write-out: the diagonal entry of a row is minus the sum of the row's run ------------
        if (OPDIM == 1) {
            // Laplacian: one lane per row walks its entries (rows are short and the whole matrix is an eighth of
            // the elasticity one: 8-byte stores to 32 different rows per instruction are affordable; the four
            // entries of a sector come from the same lane in consecutive trips).  Row starts 1 (mod 8) slots
            // apart keep the slab reads of a half-warp in different banks.
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
                sDiag[r] = diag;
                sMeta[r] = rr.node | (diagOff != 0xFFFF ? (1 << 27) : 0);
            }
        }
        else {
            // Elasticity: three consecutive rows per warp side by side, ten lanes each (nine components and an
            // idle lane); a lane walks the entries of its row, so the row sum needs no exchange between lanes.
            // The slab starts of consecutive rows are 1 (mod 8) slots apart (ring_row_padding): the three
            // 72-byte pieces read in one instruction fall into disjoint banks.
            const int grp = lane / 10, comp = lane - 10 * grp;        // lanes 9, 19, 29, 30, 31 idle
            for (int r0 = warp * 3; r0 < nbRows; r0 += nWarps * 3) {
                const int r = r0 + grp;
                if (grp < 3 && comp < 9 && r < nbRows) {
                    const RingRow rr = sRows[r];
                    const int len = rr.len, diagOff = rr.diagOff;     // 0xFFFF never equals a position
                    const double *sp = slab + (size_t)rr.localStart * SLAB + comp;
                    double *out = args.values + (size_t)rr.valueStart * 9 + comp, *op = out;
                    double a = 0.0;
                    for (int q = 0; q < len; q++, sp += SLAB, op += 9) {
                        if (q != diagOff) { const double v = *sp; a += v; *op = v; }
                    }
                    const double diag = 0.0 - a;
                    if (diagOff != 0xFFFF) out[diagOff * 9] = diag;
                    sDiag[r * 9 + comp] = diag;
                    if (comp == 0) sMeta[r] = rr.node | (diagOff != 0xFFFF ? (1 << 27) : 0);
                }
            }
        }

        offCur = offNext; offNext = offAfter;
        if (hasNext) store_coords (sHead0 + ((k + 1) & 1) * headBytes, nextCoord);   // this tile's jobs, the planes' readers, are behind the barrier
        __syncthreads ();           // every reader of this tile's slab / head is done; sDiag / sMeta and the next coordinates are complete
    }
    if (precWarp) prec_pass (rowsDone);     // the CTA's last tile
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
    args.coord = coord; args.values = values; args.prec = prec;
    args.fusePrec = fusePrec;
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
