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
// Per tile (job warps and write-out warps work side by side on different tiles, see ring_assembly_kernel):
//   0. plan record by TMA (cp.async.bulk + mbarrier): the HEAD (header, row table, node list) three tiles ahead, the
//      TAIL (batches, jobs, ring codes) and the coordinates (8-byte cp.async copies that signal the tail's barrier)
//      two tiles ahead, as soon as the job warps have released the buffers;
//   1. job warps: one lane per edge, 32 jobs per warp batch, ring codes 8 to a 64-bit word; finished 3x3 blocks go to
//      one of two tile-wide slabs at the slots of their CSR entries (entry stride 80 bytes, so a block leaves as four
//      128-bit stores and one 64-bit store);
//   2. write-out warps: three consecutive rows per warp, ten lanes each (9 components + one idle lane): a lane walks
//      its row entry by entry, copies its component to global memory (72 contiguous bytes per group and instruction)
//      and sums it, then stores its component of the diagonal entry; row starts in the slab are padded so that the
//      three pieces read together fall into disjoint banks;
//   3. fused mode: the nine lanes that hold a row's diagonal block mask and invert it through warp shuffles (cofactors
//      over the determinant; LAPACK's LU for a singular block) and write prec (prec_init + prec_inversion,
//      src/preconditioner.cc:25-87, src/Fortran/elasclpr.f:19-53); interface rows leave raw, for the halo sum.
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
    unsigned pollNs;              // sleep between two polls of a barrier (0: poll back to back)
    unsigned *intfDone;           // peer-to-peer halo (kernels_halo_p2p.cu): every write-out warp adds one after it has
                                  // written its rows of a tile that owns interface nodes (null: no signal)
};

// tileOffset entries: byte offset of the record in the low 48 bits, (head bytes / 16) above
__device__ __forceinline__ uint64_t ring_record_offset (uint64_t packed) { return packed & 0xFFFFFFFFFFFFull; }
__device__ __forceinline__ unsigned ring_record_head_bytes (uint64_t packed) { return (unsigned)(packed >> 48) << 4; }

#ifdef MFB_RING_HOST_EMULATION
// tools/ring_kernel_host.cc compiles this file with g++ and runs the kernel below on host threads
// (tools/cuda_cta_emulation.h): the PTX helpers become their emulated counterparts.
inline void ring_mbar_init (uint64_t *bar, unsigned count) { cta_emu::mbar_init (bar, count); }
inline void ring_mbar_expect_tx (uint64_t *bar, unsigned bytes) { cta_emu::mbar_expect_tx (bar, bytes); }
inline void ring_mbar_arrive (uint64_t *bar) { cta_emu::mbar_arrive (bar); }
inline void ring_mbar_wait (uint64_t *bar, unsigned parity, unsigned = 0) { cta_emu::mbar_wait (bar, parity); }
inline void ring_bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar) { cta_emu::bulk_load (dst, src, bytes, bar); }
inline void ring_bar_sync (int id, int count) { cta_emu::bar_sync (id, count); }
inline void ring_cp_async_f64 (double *dst, const double *src) { *dst = *src; }
inline void ring_cp_async_wait_all () {}
inline void ring_cp_async_arrive (uint64_t *bar) { cta_emu::mbar_arrive (bar); }
template <int REGS> inline void ring_regs_inc () {}
template <int REGS> inline void ring_regs_dec () {}
inline void ring_signal_add (unsigned *counter) { __atomic_fetch_add (counter, 1u, __ATOMIC_RELEASE); }
inline void ring_fence_gpu () { __atomic_thread_fence (__ATOMIC_SEQ_CST); }
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

// Bounded wait: a protocol bug must trap instead of hanging the device.  try_wait suspends the warp until the next
// event on the barrier or the hint runs out; at speed the loops take 1.8 % of the issue slots (PC samples,
// profiles/r2_ring_wait_loops.txt — the instrumented instruction counts of ncu's source page overstate them 10x).
// pollNs (MFB_RING_POLL_NS, default 0) adds a plain sleep between two polls.
__device__ __forceinline__ void ring_mbar_wait (uint64_t *bar, unsigned parity, unsigned pollNs = 0)
{
    unsigned done = 0;
    #pragma unroll 1
    for (long spin = 0; spin < (1l << 18); spin++) {
        asm volatile (
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"     // suspended (no issue slots) until an event
            "selp.u32 %0, 1, 0, p;\n"                                          // or the hint (ns) runs out
            "}\n" : "=r"(done) : "r"(ring_smem_u32 (bar)), "r"(parity), "r"(20000u) : "memory");
        if (done) return;
        if (pollNs) __nanosleep (pollNs);
    }
    __trap ();
}

// TMA bulk copy global -> shared, completion counted on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void ring_bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(ring_smem_u32 (dst)), "l"(src), "r"(bytes), "r"(ring_smem_u32 (bar)) : "memory");
}

// named barrier among `count` threads (the write-out warps)
__device__ __forceinline__ void ring_bar_sync (int id, int count)
{
    asm volatile ("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void ring_cp_async_f64 (double *dst, const double *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(ring_smem_u32 (dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ring_cp_async_wait_all () { asm volatile ("cp.async.wait_all;" ::: "memory"); }
// one arrival on `bar` (counted in its initial count) once this thread's earlier cp.async copies have landed
__device__ __forceinline__ void ring_cp_async_arrive (uint64_t *bar)
{
    asm volatile ("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(ring_smem_u32 (bar)) : "memory");
}

// Register reallocation between the warpgroups of a CTA (setmaxnreg, sm_90a+; SASS: USETMAXREG): the 1024-thread
// kernel is launched with 64 registers per thread; its write-out warpgroups give registers back, its job
// warpgroups take them.  ptxas allocates each branch within the count named there.
template <int REGS> __device__ __forceinline__ void ring_regs_inc () { if (REGS > 0) asm volatile ("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REGS > 0 ? REGS : 24)); }
template <int REGS> __device__ __forceinline__ void ring_regs_dec () { if (REGS > 0) asm volatile ("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REGS > 0 ? REGS : 24)); }

__device__ __forceinline__ void ring_fence_gpu () { __threadfence (); }
// one more write-out warp is done with a tile that owns interface nodes: its stores to prec are ordered before the add
__device__ __forceinline__ void ring_signal_add (unsigned *counter)
{
    asm volatile ("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
}

#endif

__host__ __device__ __forceinline__ unsigned ring_align128 (unsigned x) { return (x + 127u) & ~127u; }

// doubles per coordinate plane: tile-local node ids are one byte (kRingMaxNodes = 254)
constexpr int kRingPlane = 256;

// doubles per slab entry: an elasticity block is padded to 80 bytes so that it is 16-byte aligned
__host__ __device__ constexpr int ring_slab_stride (int opDim) { return opDim == 9 ? 10 : 1; }

// Warp-specialised kernel (round 2).  The ncu captures of the round-1 kernel (profiles/r2_ring_ela_*.txt) show
// no pipe above 55 %: 19 % of the warp time is spent at the two block barriers of a tile, 11 % waiting for the
// next tile's coordinates, and because all warps of a CTA are in the same phase the FP64 pipe idles during
// every write-out and the load/store unit during every job phase; a second and a third CTA per SM recover
// only part of it (0.81 / 0.54 / 0.46 ms with one / two / three CTAs per SM).  Here the two phases run side
// by side, each on its own warps, decoupled by mbarriers and a double-buffered slab:
//   * JOB warps (two thirds of the CTA) only walk rings: tile k's jobs fill slab k & 1; a job warp arrives
//     on full[k & 1] after its last batch of the tile and goes on to tile k + 1 without waiting for anyone;
//   * WRITE-OUT warps wait for full[k & 1], stream the tile's rows to global memory (row sums -> diagonal
//     entry -> fused preconditioner block, as before) and arrive on ready[k & 1]; they also feed the job
//     warps: at the start of tile k's write-out they request the tail (TMA) and load the coordinates of tile
//     k + 2, whose buffers (those of tile k) the job warps have just released — a whole job phase ahead of
//     their use — and the plan head of tile k + 3.
// No block barrier after the prologue.  THREADS = 768 (default): 13 job warps + 11 write-out warps (elasticity), one CTA per SM,
// tiles of <= 64 rows; THREADS = 384: 7 + 5, two CTAs per SM, tiles of <= 30 rows.
// Coordinates of tile k + 2: 0 = 8-byte cp.async copies signalled on the tail barrier (no registers, no wait, but every
// copy costs a shared-memory wavefront on arrival: ~11 M of the 74 M of an EIB iteration); 1 = plain loads issued at the
// end of the previous tile's write-out, stored side by side (two wavefronts per 32 nodes) once the buffer is free.
#ifndef MFB_RING_STAGE_LDG
#define MFB_RING_STAGE_LDG 0
#endif

// Elasticity write-out: 0 (shipped) = three rows side by side, every lane copying its component while it sums it;
// 1 = the diagonal entry is first stored into the slab and the rows leave as runs of consecutive doubles.  A pure store
// kernel writes the matrix in 0.30 ms with the first shape and in 0.22 ms with the second (tools/microbench/
// store_probe.cu), but the second pass over the slab and the longer instruction stream of the write-out warps cost
// more than the stores save: 0.534 against 0.406 ms per EIB iteration (profiles/r2_experiments.md).
#ifndef MFB_RING_COALESCED_ROWS
#define MFB_RING_COALESCED_ROWS 0
#endif
// the copy loop of the first shape split at the diagonal entry (no comparison per entry)
#ifndef MFB_RING_SPLIT_DIAG
#define MFB_RING_SPLIT_DIAG 0
#endif

// Warps per role.  job warps out of 24, measured on the EIB mesh (ms per iteration, final pipeline): elasticity, 768 threads: 12: 0.421,
// 13: 0.414, 15: 0.439, 16: 0.449; 384 threads (x 2 CTAs): 6 of 12: 0.458, 7: 0.430, 8: 0.482.  The Laplacian has an
// eighth of the write-out work per row and wants more job warps (768 threads, 20 + 4: 0.252).
// 1024 threads: warpgroups (4 warps) change their register count after the prologue — job warps 72, write-out warps 56
// (elasticity, 16 + 16) resp. 40 (Laplacian, 24 + 8): 32 warps instead of the 24 that 80 registers per thread allow.
__host__ __device__ constexpr int ring_job_warps (int opDim, int threads)
{
#ifdef MFB_RING_JOB_WARPS_OF_24
    return threads / 32 * MFB_RING_JOB_WARPS_OF_24 / 24;
#else
#ifdef MFB_RING_768_REGS      // experiment: 768 threads as 12 job warps at 96 registers + 12 write-out warps at 64
    if (threads == 768 && opDim == 9) return 12;
#endif
    return threads == 1024 ? (opDim == 1 ? 24 : 16)
         : threads == 896  ? (opDim == 1 ? 20 : 16)
         : threads == 640  ? (opDim == 1 ? 16 : 11)
                           : threads / 32 * (opDim == 1 ? (threads == 768 ? 20 : 16) : (threads == 768 ? 13 : 14)) / 24;
#endif
}
#ifndef MFB_RING_JOB_REGS
#define MFB_RING_JOB_REGS 72
#endif
#ifndef MFB_RING_OUT_REGS
#define MFB_RING_OUT_REGS 56
#endif
// registers per thread after the prologue (0: as launched).  1024 threads are launched with 64: 16 x 72 + 16 x 56 (Laplacian
// 24 x 72 + 8 x 40); 896 threads with 72: 16 x 80 + 12 x 56 (Laplacian 20 x 80 + 8 x 40)
#ifdef MFB_RING_768_REGS
__host__ __device__ constexpr int ring_job_regs (int opDim, int threads) { return threads == 768 && opDim == 9 ? 96 : 0; }
__host__ __device__ constexpr int ring_out_regs (int opDim, int threads) { return threads == 768 && opDim == 9 ? 64 : 0; }
#else
__host__ __device__ constexpr int ring_job_regs (int opDim, int threads) { return threads == 1024 ? (opDim == 1 ? 72 : MFB_RING_JOB_REGS) : threads == 896 ? 80 : 0; }
__host__ __device__ constexpr int ring_out_regs (int opDim, int threads) { return threads == 1024 ? (opDim == 1 ? 40 : MFB_RING_OUT_REGS) : threads == 896 ? (opDim == 1 ? 40 : 56) : 0; }
#endif

constexpr unsigned kRingPollNs = 0;      // default sleep between two polls of a barrier
constexpr int kRingHeadBuffers = 5;      // heads of tiles k - 1 .. k + 3 are alive while the write-out warps work on tile k

inline RingSmemLayout ring_smem_layout (int operatorID, const DeviceRingPlan &plan)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    RingSmemLayout L;
    L.headBytes = ring_align128 (plan.maxHeadBytes);
    L.tail = kRingHeadBuffers * L.headBytes;
    L.tailBytes = ring_align128 (plan.maxTailBytes);
    L.planes = L.tail + 2 * L.tailBytes;
    L.slab = L.planes + 2 * 3 * kRingPlane * (unsigned)sizeof (double);
    L.slabBytes = ring_align128 ((unsigned)plan.maxEntries * ring_slab_stride (opDim) * (unsigned)sizeof (double));
    L.bars = L.slab + 2 * L.slabBytes;
    L.total = L.bars + (kRingHeadBuffers + 6) * (unsigned)sizeof (uint64_t);
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
    constexpr int NWARPS = THREADS / 32, NJOB = ring_job_warps (OPDIM, THREADS), NOUT = NWARPS - NJOB;
    static_assert (NJOB > 0 && NOUT > 0 && (ring_job_regs (OPDIM, THREADS) == 0 || (NJOB % 4 == 0 && NOUT % 4 == 0 && 32 * (NJOB * ring_job_regs (OPDIM, THREADS) + NOUT * ring_out_regs (OPDIM, THREADS)) <= 65536)),
                   "setmaxnreg works on whole warpgroups and within the register file");
    const DeviceRingPlan &P = args.plan;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // shared memory: [head 0..3][tail 0, 1][X Y Z even tiles][X Y Z odd tiles][slab 0][slab 1][10 mbarriers]
    const RingSmemLayout &L = args.smem;
    const unsigned headBytes = L.headBytes;
    constexpr int planeStride = kRingPlane;                           // planes start on a 128-byte line: bank = id mod 16;
                                                                      // a constant, so that Y and Z are immediate offsets from X
    constexpr int SLAB = ring_slab_stride (OPDIM);
    unsigned char *sHead0 = smemRaw, *sTail0 = smemRaw + L.tail;
    double *planes0 = reinterpret_cast<double*> (smemRaw + L.planes);
    double *slab0 = reinterpret_cast<double*> (smemRaw + L.slab);
    const unsigned slabDoubles = L.slabBytes / (unsigned)sizeof (double);
    uint64_t *bars = reinterpret_cast<uint64_t*> (smemRaw + L.bars);
    uint64_t *headFull = bars, *tailFull = bars + kRingHeadBuffers, *full = tailFull + 2, *ready = full + 2;

    const int firstTile = args.firstTile + blockIdx.x, tileStep = gridDim.x;
    const int nbMine = firstTile < args.lastTile ? (args.lastTile - firstTile + tileStep - 1) / tileStep : 0;
    if (nbMine == 0) return;

    if (tid == 0) {
        for (int b = 0; b < kRingHeadBuffers; b++) ring_mbar_init (headFull + b, 1);
        // tailFull: the loader's TMA (one arrival + its bytes) and every write-out thread's share of the coordinates
        for (int b = 0; b < 2; b++) { ring_mbar_init (tailFull + b, 1 + (MFB_RING_STAGE_LDG ? NOUT : NOUT * 32)); ring_mbar_init (full + b, NJOB); ring_mbar_init (ready + b, NOUT); }
    }
    __syncthreads ();          // the only block barrier

    auto head_of = [&] (int k) { return sHead0 + (unsigned)((k + kRingHeadBuffers) % kRingHeadBuffers) * headBytes; };
    auto wait_head = [&] (int k) { ring_mbar_wait (headFull + k % kRingHeadBuffers, (unsigned)(k / kRingHeadBuffers) & 1u); };

    if (warp < NJOB) {
        // =============================== job warps ===============================================
        ring_regs_inc<ring_job_regs (OPDIM, THREADS)> ();
        for (int k = 0; k < nbMine; k++) {
            wait_head (k);
            const unsigned char *sHead = head_of (k);
            const RingTileHeader &hdr = *reinterpret_cast<const RingTileHeader*> (sHead);
            const int nbBatches = hdr.nbBatches;
            const unsigned char *sTail = sTail0 + (unsigned)(k & 1) * L.tailBytes;
            const RingBatch *batches = reinterpret_cast<const RingBatch*> (sTail);
            const uint64_t *jobs = reinterpret_cast<const uint64_t*> (sTail + (hdr.offJobs - hdr.headBytes));
            const uint64_t *codes = reinterpret_cast<const uint64_t*> (sTail + (hdr.offCodes - hdr.headBytes));
            const double *sX = planes0 + (k & 1) * (3 * planeStride), *sY = sX + planeStride, *sZ = sY + planeStride;
            double *slab = slab0 + (k & 1) * slabDoubles;
            const unsigned phase = (unsigned)(k >> 1) & 1u;
            ring_mbar_wait (tailFull + (k & 1), phase, args.pollNs);     // tail and coordinates of tile k are in (requested two tiles ago)
            bool slabFree = false;                          // ready[k & 1] (slab drained by the write-out of tile k - 2) is only
                                                            // needed when the first block is stored, a whole batch from now
            // the batches of a tile go round the job warps, starting where the previous tile stopped
            for (int b = (warp + NJOB - (k * 5) % NJOB) % NJOB; b < nbBatches; b += NJOB) {
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
                if (!slabFree) { ring_mbar_wait (ready + (k & 1), phase, args.pollNs); slabFree = true; }
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
            if (!slabFree) ring_mbar_wait (ready + (k & 1), phase, args.pollNs);   // also without a batch: full[] then implies that every
                                                                      // write-out warp is done with tile k - 2
            __syncwarp ();
            if (lane == 0) ring_mbar_arrive (full + (k & 1));   // this warp's share of tile k is in the slab; it no longer
                                                                // reads the tile's tail, coordinates or head
        }
    }
    else {
        // =============================== write-out warps =========================================
        ring_regs_dec<ring_out_regs (OPDIM, THREADS)> ();
        const int ow = warp - NJOB, otid = tid - NJOB * 32;          // 0 .. NOUT * 32 - 1
        auto tile_of = [&] (int k) { return firstTile + k * tileStep; };
        auto fetch_head = [&] (uint64_t packed, int k) {              // one thread
            const unsigned bytes = ring_record_head_bytes (packed);
            uint64_t *bar = headFull + k % kRingHeadBuffers;
            ring_mbar_expect_tx (bar, bytes);
            ring_bulk_load (head_of (k), P.blob + ring_record_offset (packed), bytes, bar);
        };
        auto fetch_tail = [&] (uint64_t packed, int k) {              // one thread, head k has landed
            const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (head_of (k));
            const unsigned bytes = h.blobBytes - h.headBytes;
            uint64_t *bar = tailFull + (k & 1);
            if (bytes) {
                ring_mbar_expect_tx (bar, bytes);
                ring_bulk_load (sTail0 + (unsigned)(k & 1) * L.tailBytes, P.blob + ring_record_offset (packed) + h.headBytes, bytes, bar);
            }
            else ring_mbar_arrive (bar);                                // a tile without jobs: the phase completes at once
        };
        // the loader (first write-out thread) keeps the packed record offsets of tiles k + 2 and k + 3 in
        // registers and loads the next one a whole tile before it is needed
        uint64_t offA = 0, offB = 0;
        if (otid == 0) {
            for (int q = 0; q < 3 && q < nbMine; q++) fetch_head (P.tileOffset[tile_of (q)], q);
            offA = P.tileOffset[tile_of (0)];
            if (nbMine > 1) offB = P.tileOffset[tile_of (1)];
        }
#if MFB_RING_STAGE_LDG
        constexpr int NPT = (kRingMaxNodes + NOUT * 32 - 1) / (NOUT * 32);     // nodes per write-out thread
        double stage[NPT][3];
        auto load_coords = [&] (int kq) {                                   // the coordinates tile kq will need, into registers
            if (kq >= nbMine) return;
            wait_head (kq);
            const unsigned char *qh = head_of (kq);
            const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (qh);
            const int *nodes = reinterpret_cast<const int*> (qh + h.offNodes);
            #pragma unroll
            for (int q = 0; q < NPT; q++) {
                const int n = otid + q * (NOUT * 32);
                if (n < h.nbNodes) {
                    const double *g = args.coord + (size_t)nodes[n] * 3;
                    stage[q][0] = g[0]; stage[q][1] = g[1]; stage[q][2] = g[2];
                }
            }
        };
        load_coords (0);
#endif
        // k = -2, -1: nothing to write out yet, only the first two tiles to stage
        for (int k = -2; k < nbMine; k++) {
            // (no barrier among the write-out warps: full[k & 1] implies that every one of them has finished tile
            // k - 2 — the job warps could not have stored tile k otherwise — and the head fetched below replaces that of
            // tile k - 2)
            if (k >= 0) ring_mbar_wait (full + (k & 1), (unsigned)(k >> 1) & 1u, args.pollNs);   // the job warps are done with tile k
            // ---- staging for tile k + 2 (into the buffers of tile k) -------------------------------------
            const int kn = k + 2;
            if (kn < nbMine) {
                wait_head (kn);
                const unsigned char *nh = head_of (kn);
                const RingTileHeader &h = *reinterpret_cast<const RingTileHeader*> (nh);
                if (otid == 0) fetch_tail (offA, kn);
                double *pl = planes0 + (kn & 1) * (3 * planeStride);
#if MFB_RING_STAGE_LDG
                // coordinates: loaded into registers at the end of the previous write-out (load_coords), stored side by
                // side now that the job warps have released the planes
                #pragma unroll
                for (int q = 0; q < NPT; q++) {
                    const int n = otid + q * (NOUT * 32);
                    if (n < h.nbNodes) { pl[n] = stage[q][0]; pl[planeStride + n] = stage[q][1]; pl[2 * planeStride + n] = stage[q][2]; }
                }
                __syncwarp ();
                if (lane == 0) ring_mbar_arrive (tailFull + (kn & 1));
#else
                // coordinates: asynchronous 8-byte copies that signal tailFull[kn & 1] when they have landed (each costs a
                // shared-memory wavefront on arrival, but no register, no wait)
                const int *nodes = reinterpret_cast<const int*> (nh + h.offNodes);
                for (int n = otid; n < h.nbNodes; n += NOUT * 32) {
                    const double *g = args.coord + (size_t)nodes[n] * 3;
                    ring_cp_async_f64 (pl + n, g); ring_cp_async_f64 (pl + planeStride + n, g + 1);
                    ring_cp_async_f64 (pl + 2 * planeStride + n, g + 2);
                }
                ring_cp_async_arrive (tailFull + (kn & 1));
#endif
            }
            if (otid == 0) {
                // head of tile k + 3 into the buffer of tile k - 2 (five buffers)
                if (k + 3 >= 3 && k + 3 < nbMine) fetch_head (offB, k + 3);
                offA = offB;
                offB = k + 4 < nbMine ? P.tileOffset[tile_of (k + 4)] : 0;
            }
            // ---- write-out of tile k: the diagonal entry of a row is minus the sum of the row's run ---------
            if (k >= 0) {
                const unsigned char *sHead = head_of (k);
                const RingTileHeader &hdr = *reinterpret_cast<const RingTileHeader*> (sHead);
                const int nbRows = hdr.nbRows;
                const RingRow *sRows = reinterpret_cast<const RingRow*> (sHead + sizeof (RingTileHeader));
                double *slabW = slab0 + (k & 1) * slabDoubles;
                const double *slab = slabW;
                if (OPDIM == 1) {
                    // Laplacian: one lane per row walks its entries (rows are short and the whole matrix is an eighth
                    // of the elasticity one).  Row starts 1 (mod 8) slots apart keep the slab reads in different banks.
                    for (int r = ow * 32 + lane; r < nbRows; r += NOUT * 32) {
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
                    // Elasticity: three consecutive rows per warp side by side, ten lanes each (nine components and an
                    // idle lane); a lane walks the entries of its row, so the row sum needs no exchange between lanes.
                    // The slab starts of consecutive rows are 1 (mod 8) slots apart (ring_row_padding): the three
                    // 72-byte pieces read in one instruction fall into disjoint banks.
                    const int grp = lane / 10, comp = lane - 10 * grp;        // lanes 9, 19, 29, 30, 31 idle
                    const bool worker = grp < 3 && comp < 9;
                    const int ca = comp / 3, cb = comp - 3 * ca;              // component (ca, cb) of the 3x3 block
                    const int base = worker ? 10 * grp : lane;                // first lane of the row's nine; idle lanes read themselves
                    // (896 / 1024 threads: 21 row groups of a 63-row tile over 12 / 16 warps — the warps that take two change from tile to tile)
                    const int ow3 = (THREADS >= 896 ? (ow + NOUT - (k * 5) % NOUT) % NOUT : ow) * 3;
                    for (int r0 = ow3; r0 < nbRows; r0 += NOUT * 3) {
                        const int r = r0 + grp;
                        const bool rowOk = worker && r < nbRows;
                        int node = 0, diagOff = 0xFFFF;
                        double diag = 0.0;
#if MFB_RING_COALESCED_ROWS
                        if (rowOk) {
                            // row sum only; the diagonal entry joins the row in the slab, the row leaves below
                            const RingRow rr = sRows[r];
                            const int len = rr.len;
                            node = rr.node; diagOff = rr.diagOff;             // 0xFFFF never equals a position
                            double *sp = slabW + (size_t)rr.localStart * SLAB + comp;
                            double a = 0.0;
                            for (int q = 0; q < len; q++) {
                                if (q != diagOff) a += sp[q * SLAB];
                            }
                            diag = 0.0 - a;
                            if (diagOff != 0xFFFF) sp[diagOff * SLAB] = diag;
                        }
                        __syncwarp ();
                        // The three rows leave one after the other, every store instruction of the warp writing 32
                        // consecutive doubles (a row is one contiguous run of len x 72 bytes in nodeToNodeValue).  A pure
                        // store kernel writes the matrix in 0.30 ms with the old shape (three 72-byte pieces per instruction,
                        // 27 lanes) and in 0.19 ms with consecutive doubles (tools/microbench/store_probe.cu).
                        #pragma unroll
                        for (int g = 0; g < 3; g++) {
                            if (r0 + g < nbRows) {                             // uniform over the warp
                                const RingRow rr = sRows[r0 + g];
                                const double *src = slab + (size_t)rr.localStart * SLAB;
                                double *out = args.values + (size_t)rr.valueStart * 9;
                                const int total = rr.len * 9;
                                int e = lane / 9, c = lane - 9 * e;            // double m of the row = component c of entry e
                                for (int m = lane; m < total; m += 32) {
                                    out[m] = src[e * SLAB + c];
                                    e += 3; c += 5;                            // 32 = 3 x 9 + 5
                                    if (c >= 9) { c -= 9; e++; }
                                }
                            }
                        }
#else
                        if (rowOk) {
                            const RingRow rr = sRows[r];
                            const int len = rr.len;
                            node = rr.node; diagOff = rr.diagOff;             // 0xFFFF never equals a position
                            const double *sp = slab + (size_t)rr.localStart * SLAB + comp;
                            double *out = args.values + (size_t)rr.valueStart * 9 + comp, *op = out;
                            double a = 0.0;
#if MFB_RING_SPLIT_DIAG
                            const int dq = diagOff < len ? diagOff : len;
                            for (int q = 0; q < dq; q++, sp += SLAB, op += 9) { const double v = *sp; a += v; *op = v; }
                            sp += SLAB; op += 9;
                            for (int q = dq + 1; q < len; q++, sp += SLAB, op += 9) { const double v = *sp; a += v; *op = v; }
#else
                            for (int q = 0; q < len; q++, sp += SLAB, op += 9) {
                                if (q != diagOff) { const double v = *sp; a += v; *op = v; }
                            }
#endif
                            diag = 0.0 - a;
                            if (diagOff != 0xFFFF) out[diagOff * 9] = diag;
                        }
#endif
                        if (args.fusePrec) {
                            // prec_init + prec_inversion (src/preconditioner.cc:25-87, src/Fortran/elasclpr.f:19-53) by the
                            // nine lanes that hold the block: Dirichlet rows / columns to identity, then cofactor / determinant
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
                            const bool invert = rowOk && !isInterface && diagOff != 0xFFFF;
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
                            if (rowOk) args.prec[(size_t)(node & kRingNodeMask) * 9 + comp] = out;
                        }
                    }
                }
            }
            // ---- slab k & 1 is drained: tile k + 2 may store into it -----------------------------------------------
            if (kn < nbMine) {
                __syncwarp ();
                if (lane == 0) ring_mbar_arrive (ready + (kn & 1));
            }
            // ---- peer-to-peer halo: this warp's rows of a tile that owns interface nodes are in global memory ----------
            if (k >= 0 && args.intfDone != nullptr && reinterpret_cast<const RingTileHeader*> (head_of (k))->hasInterface) {
                ring_fence_gpu ();
                __syncwarp ();
                if (lane == 0) ring_signal_add (args.intfDone);
            }
#if MFB_RING_STAGE_LDG
            load_coords (k + 3);
#endif
        }
    }
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

int ring_write_out_warps (int operatorID, int threads)
{
    return threads / 32 - ring_job_warps (operatorID == 0 ? 1 : 9, threads);
}

size_t ring_smem_bytes (int operatorID, const DeviceRingPlan &plan)
{
    return ring_smem_layout (operatorID, plan).total;
}

#ifndef MFB_RING_HOST_EMULATION
// CTA shapes: 384 threads x 2 CTAs per SM; 640 (96 registers), 768 (80) x 1; 896 and 1024 x 1 with per-warpgroup
// register counts (setmaxnreg).  One visitor for configure / occupancy / launch.
template <class F>
cudaError_t ring_dispatch (int operatorID, int threads, F &&f)
{
    const bool lap = operatorID == 0;
    switch (threads) {
    case 384:  return lap ? f (ring_assembly_kernel<1, 384, 2>, 384)   : f (ring_assembly_kernel<9, 384, 2>, 384);
    case 640:  return lap ? f (ring_assembly_kernel<1, 640, 1>, 640)   : f (ring_assembly_kernel<9, 640, 1>, 640);
    case 768:  return lap ? f (ring_assembly_kernel<1, 768, 1>, 768)   : f (ring_assembly_kernel<9, 768, 1>, 768);
    case 896:  return lap ? f (ring_assembly_kernel<1, 896, 1>, 896)   : f (ring_assembly_kernel<9, 896, 1>, 896);
    case 1024: return lap ? f (ring_assembly_kernel<1, 1024, 1>, 1024) : f (ring_assembly_kernel<9, 1024, 1>, 1024);
    }
    return cudaErrorInvalidValue;
}

bool ring_threads_supported (int threads) { return threads == 384 || threads == 640 || threads == 768 || threads == 896 || threads == 1024; }

cudaError_t ring_configure (int operatorID)
{
    for (int threads : {384, 640, 768, 896, 1024}) {
        cudaError_t e = ring_dispatch (operatorID, threads, [] (auto kernel, int) { return ring_opt_in (kernel); });
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t ring_ctas_per_sm (int operatorID, int threads, size_t smemBytes, int *ctas)
{
    return ring_dispatch (operatorID, threads, [&] (auto kernel, int t) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor (ctas, kernel, t, smemBytes); });
}

cudaError_t launch_ring (int operatorID, const DeviceRingPlan &plan, int firstTile, int nbTiles, int ctas,
                         int threads, size_t smemBytes, const double *coord, double *values, double *prec,
                         int fusePrec, cudaStream_t stream, unsigned *intfDone)
{
    if (nbTiles <= 0) return cudaSuccess;
    RingArgs args;
    args.plan = plan; args.smem = ring_smem_layout (operatorID, plan);
    args.coord = coord; args.values = values; args.prec = prec;
    args.fusePrec = fusePrec;
    args.firstTile = firstTile; args.lastTile = firstTile + nbTiles;
    args.intfDone = intfDone;
    static const unsigned pollNs = getenv ("MFB_RING_POLL_NS") ? (unsigned)std::max (atoi (getenv ("MFB_RING_POLL_NS")), 0) : kRingPollNs;
    args.pollNs = pollNs;
    const int grid = std::max (1, std::min (ctas, nbTiles));
    return ring_dispatch (operatorID, threads, [&] (auto kernel, int t) {
        kernel<<<grid, t, smemBytes, stream>>> (args);
        return cudaGetLastError ();
    });
}
#endif

}  // namespace mfb
