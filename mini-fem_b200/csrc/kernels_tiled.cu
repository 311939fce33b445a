// TILED path: write-once assembly (+ fused preconditioner) over node tiles.
//
// Each CTA walks tiles of the plan (host/tile_plan.h).  Per tile:
//   0. one bulk copy (TMA, cp.async.bulk + mbarrier) stages the tile's whole plan record
//      into shared memory: row table, node list, element connectivity, contribution codes;
//   1. gather the coordinates of every node the tile's elements reference into shared
//      memory (SoA planes);
//   2. one thread per tile element: 12 gradient coefficients (elem_coef_seq,
//      src/assembly.cc:85-121) into shared memory, SoA planes indexed [local node][element];
//   3. diagonal blocks: 4 lanes per owned row walk the row's incident elements, reduce
//      with two shuffles, keep the block in shared memory and (fused mode) write the
//      preconditioner entry — prec_init + ela_invert_prec for that node
//      (src/preconditioner.cc:52-87, src/Fortran/elasclpr.f:2-56);
//   4. off-diagonal blocks: one lane per CSR entry of the owned rows accumulates
//      A = sum_e c_a d_b^T over the elements shared by the node pair in registers;
//      the entry is K = 1.25 A + tr(A) I, which is what the reference adds element by
//      element (src/assembly.cc:386-409: diagonal c_p d_p*2.25 + the two other
//      products, off-diagonal c_p d_q*1.25).  A warp's 32 blocks are transposed through
//      a 2304-byte shared slab and leave as coalesced 8-byte stores.
// Every CSR entry is written exactly once, by plain stores: no zero-fill, no atomics, and
// a summation order fixed by the plan (bit-reproducible run to run).
#include "kernels.cuh"
#include "device_math.cuh"

namespace mfb {

namespace {

static_assert (sizeof (TileRow) == 16 && sizeof (TileBlobHeader) == 48 && sizeof (TileBatch) == 8,
               "plan records are copied to the device verbatim");

struct TiledArgs {
    DeviceTilePlan plan;
    const double *coord;
    double *values;
    double *prec;
    const int *checkBounds;
    int nbNodes;
    int fusePrec;
    int firstTile, lastTile;      // [firstTile, lastTile)
};

// tileOffset entries: byte offset of the record in the low 48 bits, (head bytes / 16) above
__device__ __forceinline__ uint64_t record_offset (uint64_t packed) { return packed & 0xFFFFFFFFFFFFull; }
__device__ __forceinline__ unsigned record_head_bytes (uint64_t packed) { return (unsigned)(packed >> 48) << 4; }

__device__ __forceinline__ unsigned smem_u32 (const void *p) { return (unsigned)__cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t *bar, unsigned count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx (uint64_t *bar, unsigned bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait (uint64_t *bar, unsigned parity)
{
    asm volatile (
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}

// TMA bulk copy global -> shared, completion counted on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

// Laplacian: lap_pair_slot / lap_diag_slot (host/tile_plan.h) place the 10 dot products of an element;
// the plan's codes already are those slots.

template <int OPDIM, int MINB, int STRIDE>
__global__ void __launch_bounds__(256, MINB)
tiled_assembly_kernel (const TiledArgs args)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const DeviceTilePlan &P = args.plan;
    const int tid = threadIdx.x, nThreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nWarps = nThreads >> 5;
    const int strideE = STRIDE ? STRIDE : P.elemStride;     // compile-time where the plan uses a default cap

    // shared memory: [blob][cX cY cZ][sDiag][scratch = coordinates, later the warp slabs][mbarrier]
    unsigned char *sBlob = smemRaw;
    double *cX = reinterpret_cast<double*> (smemRaw + ((P.maxBlobBytes + 127u) & ~127u));
    double *cY = cX + 4 * strideE;
    double *cZ = cY + 4 * strideE;
    double *sDiag = cZ + 4 * strideE;
    double *scratch = sDiag + P.maxRows * OPDIM;
    const int scratchDoubles = max (3 * P.maxNodesRef, OPDIM == 9 ? nWarps * 288 : 0);
    uint64_t *bar = reinterpret_cast<uint64_t*> (scratch + scratchDoubles);
    double *sX = scratch, *sY = sX + P.maxNodesRef, *sZ = sY + P.maxNodesRef;
    double *slab = scratch + warp * 288;

    if (tid == 0) mbar_init (bar, 1);
    __syncthreads ();

    unsigned parity = 0;
    for (int tile = args.firstTile + blockIdx.x; tile < args.lastTile; tile += gridDim.x) {
        // ---- 0. stage the plan record ------------------------------------------------
        if (tid == 0) {
            const uint64_t off = record_offset (P.tileOffset[tile]);
            const unsigned bytes = (unsigned)(record_offset (P.tileOffset[tile + 1]) - off);
            mbar_expect_tx (bar, bytes);
            bulk_load (sBlob, P.blob + off, bytes, bar);
        }
        mbar_wait (bar, parity);
        parity ^= 1;
        const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (sBlob);
        const int nbRows = hdr.nbRows, nbElems = hdr.nbElems, nbNodesRef = hdr.nbNodesRef;
        const TileRow *sRows = reinterpret_cast<const TileRow*> (sBlob + sizeof (TileBlobHeader));
        const int *tileNodes = reinterpret_cast<const int*> (sBlob + hdr.offNodes);
        const ushort4 *tileElems = reinterpret_cast<const ushort4*> (sBlob + hdr.offElems);
        const uint8_t *entryRow = sBlob + hdr.offEntryRow;
        const uint16_t *laneEntry = reinterpret_cast<const uint16_t*> (sBlob + hdr.offLaneEntry);
        const TileBatch *batches = reinterpret_cast<const TileBatch*> (sBlob + hdr.offBatches);
        const uint16_t *diagCodes = reinterpret_cast<const uint16_t*> (sBlob + hdr.offDiag);
        const uint16_t *pairCodes = reinterpret_cast<const uint16_t*> (sBlob + hdr.offPair);

        // ---- 1. coordinates ------------------------------------------------------------
        for (int n = tid; n < nbNodesRef; n += nThreads) {
            const double *q = args.coord + (size_t)tileNodes[n] * 3;
            sX[n] = __ldg (q); sY[n] = __ldg (q + 1); sZ[n] = __ldg (q + 2);
        }
        if (tid < 64) {                                   // the 16 all-zero slots of each plane (padding codes)
            const int z = (tid >> 4) * strideE + nbElems + (tid & 15);
            cX[z] = 0.0; cY[z] = 0.0; cZ[z] = 0.0;
            if (OPDIM == 1 && tid < 16) cX[lap_pair_slot (1, 0, nbElems + tid, strideE - 4)] = 0.0;
        }
        __syncthreads ();

        // ---- 2. gradient coefficients ----------------------------------------------------
        for (int e = tid; e < nbElems; e += nThreads) {
            const ushort4 ln = tileElems[e];
            if (ln.x == 0xFFFF) continue;                 // hole of the coset numbering
            const int ids[4] = {ln.x, ln.y, ln.z, ln.w};
            double p[12], c[12];
            #pragma unroll
            for (int i = 0; i < 4; i++) { p[3 * i] = sX[ids[i]]; p[3 * i + 1] = sY[ids[i]]; p[3 * i + 2] = sZ[ids[i]]; }
            elem_coef (p, c);
            if (OPDIM == 1) {                             // the 10 dot products (assembly.cc:539-541)
                const int PS = strideE - 4;
                #pragma unroll
                for (int a = 0; a < 4; a++) {
                    #pragma unroll
                    for (int b = a; b < 4; b++) {
                        const double dot = c[3 * a] * c[3 * b] + c[3 * a + 1] * c[3 * b + 1] + c[3 * a + 2] * c[3 * b + 2];
                        cX[a == b ? lap_diag_slot (a, e, PS) : lap_pair_slot (a, b, e, PS)] = dot;
                    }
                }
            }
            else {
                #pragma unroll
                for (int a = 0; a < 4; a++) {
                    cX[a * strideE + e] = c[3 * a]; cY[a * strideE + e] = c[3 * a + 1]; cZ[a * strideE + e] = c[3 * a + 2];
                }
            }
        }
        __syncthreads ();

        // ---- 3. diagonal blocks, 4 lanes per row ------------------------------------------
        for (int r0 = warp * 8; r0 < nbRows; r0 += nWarps * 8) {
            const int r = r0 + (lane >> 2), sub = lane & 3;
            const bool live = r < nbRows;
            const int begin = live ? sRows[r].diagCodeBase : 0;
            const int end   = live ? sRows[r + 1].diagCodeBase : 0;
            double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
            for (int k = begin + sub; k < end; k += 4) {
                const int code = diagCodes[k];
                if (OPDIM == 1) { a00 += cX[code]; continue; }
                const int v = (code & 3) * strideE + (code >> 2);
                const double x = cX[v], y = cY[v], z = cZ[v];
                a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
            }
            #pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                a00 += __shfl_xor_sync (0xffffffffu, a00, off);
                if (OPDIM == 9) {
                    a01 += __shfl_xor_sync (0xffffffffu, a01, off);
                    a02 += __shfl_xor_sync (0xffffffffu, a02, off);
                    a11 += __shfl_xor_sync (0xffffffffu, a11, off);
                    a12 += __shfl_xor_sync (0xffffffffu, a12, off);
                    a22 += __shfl_xor_sync (0xffffffffu, a22, off);
                }
            }
            if (live && sub == 0) {
                if (OPDIM == 1) sDiag[r] = a00;
                else {
                    const double tr = a00 + a11 + a22;
                    sDiag[r * 9 + 0] = 1.25 * a00 + tr; sDiag[r * 9 + 1] = 1.25 * a01; sDiag[r * 9 + 2] = 1.25 * a02;
                    sDiag[r * 9 + 3] = 1.25 * a01; sDiag[r * 9 + 4] = 1.25 * a11 + tr; sDiag[r * 9 + 5] = 1.25 * a12;
                    sDiag[r * 9 + 6] = 1.25 * a02; sDiag[r * 9 + 7] = 1.25 * a12; sDiag[r * 9 + 8] = 1.25 * a22 + tr;
                }
            }
        }
        __syncthreads ();      // sDiag complete; the coordinate planes are dead: scratch becomes the slabs

        // ---- 4. off-diagonal blocks, one lane per CSR entry ----------------------------------
        const int nbBatches = hdr.nbBatches;
        for (int b = warp; b < nbBatches; b += nWarps) {
            const TileBatch tb = batches[b];
            const int q = laneEntry[b * 32 + lane];
            const bool live = q != 0xFFFF;
            const int r = live ? entryRow[q] : 0;
            const TileRow tr = sRows[r];
            const bool isDiag = live && q == tr.diagLocal;
            const int g = tr.valueStart + (q - tr.localStart);         // global CSR entry
            const uint16_t *codes = pairCodes + tb.codeBase + lane;

            double acc[OPDIM];
            #pragma unroll
            for (int k = 0; k < OPDIM; k++) acc[k] = 0.0;
            #pragma unroll 2
            for (int t = 0; t < tb.steps; t++) {
                const int code = codes[t * 32];
                const int e = code >> 4;
                if (OPDIM == 1) { acc[0] += cX[code]; continue; }
                const int va = ((code >> 2) & 3) * strideE + e, vb = (code & 3) * strideE + e;
                const double ax = cX[va], ay = cY[va], az = cZ[va];
                const double bx = cX[vb], by = cY[vb], bz = cZ[vb];
                {
                    acc[0] += ax * bx; acc[1] += ax * by; acc[2] += ax * bz;
                    acc[3] += ay * bx; acc[4] += ay * by; acc[5] += ay * bz;
                    acc[6] += az * bx; acc[7] += az * by; acc[8] += az * bz;
                }
            }

            if (OPDIM == 1) {
                if (live) args.values[g] = isDiag ? sDiag[r] : acc[0];
            }
            else {
                const double trA = acc[0] + acc[4] + acc[8];
                #pragma unroll
                for (int k = 0; k < 9; k++) {
                    double v = 1.25 * acc[k] + ((k == 0 || k == 4 || k == 8) ? trA : 0.0);
                    if (isDiag) v = sDiag[r * 9 + k];
                    slab[lane * 9 + k] = v;
                }
                __syncwarp ();
                // each half-warp holds consecutive entries of one row: two contiguous runs
                const unsigned liveMask = __ballot_sync (0xffffffffu, live);
                const int run0 = __popc (liveMask & 0xffffu) * 9, run1 = __popc (liveMask >> 16) * 9;
                double *out0 = args.values + (size_t)__shfl_sync (0xffffffffu, g, 0) * 9;
                double *out1 = args.values + (size_t)__shfl_sync (0xffffffffu, g, 16) * 9 - 144;
                #pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int m = k * 32 + lane;
                    if (m < 144) { if (m < run0) out0[m] = slab[m]; }
                    else if (m - 144 < run1) out1[m] = slab[m];
                }
                __syncwarp ();
            }
        }
        // ---- 5. fused preconditioner: one thread per owned row (last warps first: they get fewer batches) ----
        if (args.fusePrec) {
            for (int r = nThreads - 1 - tid; r < nbRows; r += nThreads) {
                const int nodeField = sRows[r].node;
                const int node = nodeField & 0x7fffffff;
                const bool isInterface = nodeField < 0;
                if (OPDIM == 1) {
                    const double d = sDiag[r];
                    args.prec[node] = isInterface ? d : 1.0 / d;
                }
                else {
                    double b[9];
                    #pragma unroll
                    for (int q = 0; q < 9; q++) b[q] = sDiag[r * 9 + q];
                    if (!isInterface) {
                        int mx = 0, my = 0, mz = 0;
                        if (args.checkBounds) {
                            mx = __ldg (args.checkBounds + node);
                            my = __ldg (args.checkBounds + (size_t)args.nbNodes + node);
                            mz = __ldg (args.checkBounds + 2 * (size_t)args.nbNodes + node);
                        }
                        mask_block (b, mx, my, mz);
                        if (sRows[r].diagLocal != 0xFFFF) invert3_lu (b);
                    }
                    double *dst = args.prec + (size_t)node * 9;
                    #pragma unroll
                    for (int q = 0; q < 9; q++) dst[q] = b[q];
                }
            }
        }
        __syncthreads ();      // every reader of the blob / coefficients is done before the next tile lands
    }
}


// ------------------------------------------------------------------------------------
// Prefetching variant (the default).  Same five phases, but nothing at the head of a tile waits
// on global memory: the record is split into a HEAD (header, row table, node list, element
// connectivity: what the coefficient phase needs) and a TAIL (lane tables and contribution
// codes: needed from the diagonal pass on).
//   * the head of tile t+1 is fetched by TMA into the other head buffer at the start of tile t;
//   * when tile t enters its off-diagonal pass, that head has landed and every thread issues
//     cp.async copies of tile t+1's node coordinates, which complete during the pass;
//   * the tail of tile t is fetched by TMA at the start of tile t and lands during the
//     coefficient phase.
// Shared memory stays under a third of an SM (three CTAs per SM) because the write-out slab is
// per half-warp (144 doubles per warp).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_f64 (double *dst, const double *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32 (dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all () { asm volatile ("cp.async.wait_all;" ::: "memory"); }

// Bounded wait: a protocol bug must trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait_or_trap (uint64_t *bar, unsigned parity)
{
    unsigned done = 0;
    for (long spin = 0; spin < (1l << 22); spin++) {
        asm volatile (
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap ();
}

template <int OPDIM, int STRIDE>
__global__ void __launch_bounds__(256, 3)
tiled_prefetch_kernel (const TiledArgs args)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const DeviceTilePlan &P = args.plan;
    const int tid = threadIdx.x, nThreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nWarps = nThreads >> 5;
    const int strideE = STRIDE ? STRIDE : P.elemStride;

    // shared memory: [head 0][head 1][tail][cX cY cZ][coordinates][half-warp slabs][3 mbarriers]
    const unsigned headBytes = (P.maxHeadBytes + 127u) & ~127u, tailBytes = (P.maxTailBytes + 127u) & ~127u;
    unsigned char *sHead0 = smemRaw, *sTail = smemRaw + 2 * headBytes;
    double *cX = reinterpret_cast<double*> (sTail + tailBytes);
    double *cY = cX + 4 * strideE;
    double *cZ = cY + 4 * strideE;
    double *sX = cZ + 4 * strideE, *sY = sX + P.maxNodesRef, *sZ = sY + P.maxNodesRef;
    double *slabs = sZ + P.maxNodesRef;
    double *slab = slabs + warp * 144;
    uint64_t *bars = reinterpret_cast<uint64_t*> (slabs + (OPDIM == 9 ? nWarps * 144 : 0));
    uint64_t *headFull = bars, *tailFull = bars + 2;            // headFull[2], tailFull

    if (tid == 0) { mbar_init (headFull, 1); mbar_init (headFull + 1, 1); mbar_init (tailFull, 1); }
    __syncthreads ();

    const int firstTile = args.firstTile + blockIdx.x, tileStep = gridDim.x;
    auto fetch_head = [&] (uint64_t packed, int k) {             // thread 0 only
        const unsigned bytes = record_head_bytes (packed);
        mbar_expect_tx (headFull + (k & 1), bytes);
        bulk_load (sHead0 + (k & 1) * headBytes, P.blob + record_offset (packed), bytes, headFull + (k & 1));
    };
    // thread 0 keeps the packed offsets of this tile and the next in registers and loads the one after
    // that a whole tile ahead, so that issuing the copies never waits on global memory
    uint64_t offCur = 0, offNext = 0;
    if (tid == 0 && firstTile < args.lastTile) {
        offCur = P.tileOffset[firstTile];
        if (firstTile + tileStep < args.lastTile) offNext = P.tileOffset[firstTile + tileStep];
    }
    auto gather_coords = [&] (const unsigned char *head) {       // all threads, asynchronous
        const TileBlobHeader &h = *reinterpret_cast<const TileBlobHeader*> (head);
        const int *nodes = reinterpret_cast<const int*> (head + h.offNodes);
        for (int n = tid; n < h.nbNodesRef; n += nThreads) {
            const double *q = args.coord + (size_t)nodes[n] * 3;
            cp_async_f64 (sX + n, q); cp_async_f64 (sY + n, q + 1); cp_async_f64 (sZ + n, q + 2);
        }
    };

    // prologue: head and coordinates of this CTA's first tile
    if (firstTile < args.lastTile) {
        if (tid == 0) fetch_head (offCur, 0);
        mbar_wait_or_trap (headFull, 0);
        gather_coords (sHead0);
        cp_async_wait_all ();
    }
    __syncthreads ();

    int k = 0;
    for (int tile = firstTile; tile < args.lastTile; tile += tileStep, k++) {
        const unsigned char *sHead = sHead0 + (k & 1) * headBytes;
        const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (sHead);
        const bool hasNext = tile + tileStep < args.lastTile;
        // ---- 0. this tile's tail and the next tile's head start travelling -----------------
        uint64_t offAfter = 0;
        if (tid == 0) {
            if (tile + 2 * tileStep < args.lastTile) offAfter = P.tileOffset[tile + 2 * tileStep];   // used next iteration
            const unsigned bytes = hdr.blobBytes - hdr.offEntryRow;
            mbar_expect_tx (tailFull, bytes);
            bulk_load (sTail, P.blob + record_offset (offCur) + hdr.offEntryRow, bytes, tailFull);
            if (hasNext) fetch_head (offNext, k + 1);
        }
        const int nbRows = hdr.nbRows, nbElems = hdr.nbElems;
        const TileRow *sRows = reinterpret_cast<const TileRow*> (sHead + sizeof (TileBlobHeader));
        const ushort4 *tileElems = reinterpret_cast<const ushort4*> (sHead + hdr.offElems);
        const unsigned tailBase = hdr.offEntryRow;
        const uint8_t *entryRow = sTail;
        const uint16_t *laneEntry = reinterpret_cast<const uint16_t*> (sTail + (hdr.offLaneEntry - tailBase));
        const TileBatch *batches = reinterpret_cast<const TileBatch*> (sTail + (hdr.offBatches - tailBase));
        const uint16_t *diagCodes = reinterpret_cast<const uint16_t*> (sTail + (hdr.offDiag - tailBase));
        const uint16_t *pairCodes = reinterpret_cast<const uint16_t*> (sTail + (hdr.offPair - tailBase));

        // ---- 2. gradient coefficients (coordinates were prefetched) ----------------------------
        if (tid < 64) {                                   // the 16 all-zero slots of each plane (padding codes)
            const int z = (tid >> 4) * strideE + nbElems + (tid & 15);
            cX[z] = 0.0; cY[z] = 0.0; cZ[z] = 0.0;
        }
        for (int e = tid; e < nbElems; e += nThreads) {
            const ushort4 ln = tileElems[e];
            if (ln.x == 0xFFFF) continue;                 // hole of the coset numbering
            const int ids[4] = {ln.x, ln.y, ln.z, ln.w};
            double p[12], c[12];
            #pragma unroll
            for (int i = 0; i < 4; i++) { p[3 * i] = sX[ids[i]]; p[3 * i + 1] = sY[ids[i]]; p[3 * i + 2] = sZ[ids[i]]; }
            elem_coef (p, c);
            if (OPDIM == 1) {                             // the 10 dot products (assembly.cc:539-541)
                const int PS = strideE - 4;
                #pragma unroll
                for (int a = 0; a < 4; a++) {
                    #pragma unroll
                    for (int b = a; b < 4; b++) {
                        const double dot = c[3 * a] * c[3 * b] + c[3 * a + 1] * c[3 * b + 1] + c[3 * a + 2] * c[3 * b + 2];
                        cX[a == b ? lap_diag_slot (a, e, PS) : lap_pair_slot (a, b, e, PS)] = dot;
                    }
                }
            }
            else {
                #pragma unroll
                for (int a = 0; a < 4; a++) {
                    cX[a * strideE + e] = c[3 * a]; cY[a * strideE + e] = c[3 * a + 1]; cZ[a * strideE + e] = c[3 * a + 2];
                }
            }
        }
        __syncthreads ();      // coefficients complete, coordinate planes free
        mbar_wait_or_trap (tailFull, k & 1);

        // ---- 3. diagonal blocks, 4 lanes per row ------------------------------------------------
        for (int r0 = warp * 8; r0 < nbRows; r0 += nWarps * 8) {
            const int r = r0 + (lane >> 2), sub = lane & 3;
            const bool live = r < nbRows;
            const int begin = live ? sRows[r].diagCodeBase : 0;
            const int end   = live ? sRows[r + 1].diagCodeBase : 0;
            double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
            for (int q = begin + sub; q < end; q += 4) {
                const int code = diagCodes[q];
                if (OPDIM == 1) { a00 += cX[code]; continue; }
                const int v = (code & 3) * strideE + (code >> 2);
                const double x = cX[v], y = cY[v], z = cZ[v];
                a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
            }
            #pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                a00 += __shfl_xor_sync (0xffffffffu, a00, off);
                if (OPDIM == 9) {
                    a01 += __shfl_xor_sync (0xffffffffu, a01, off);
                    a02 += __shfl_xor_sync (0xffffffffu, a02, off);
                    a11 += __shfl_xor_sync (0xffffffffu, a11, off);
                    a12 += __shfl_xor_sync (0xffffffffu, a12, off);
                    a22 += __shfl_xor_sync (0xffffffffu, a22, off);
                }
            }
            // every lane of the quad now holds the row's sums: the diagonal entry and the node's
            // preconditioner block (prec_init + prec_inversion, src/preconditioner.cc:25-87,
            // src/Fortran/elasclpr.f:19-53) leave from here, lane `sub` storing components
            // sub, sub + 4 (and 8)
            if (live) {
                const TileRow tr = sRows[r];
                const int node = tr.node & 0x7fffffff;
                const bool isInterface = tr.node < 0, hasDiag = tr.diagLocal != 0xFFFF;
                const size_t gd = (size_t)(tr.valueStart + ((int)tr.diagLocal - (int)tr.localStart));
                if (OPDIM == 1) {
                    if (sub == 0) {
                        if (hasDiag) args.values[gd] = a00;
                        if (args.fusePrec) args.prec[node] = isInterface ? a00 : 1.0 / a00;
                    }
                }
                else {
                    const double tr3 = a00 + a11 + a22;
                    double b[9] = {1.25 * a00 + tr3, 1.25 * a01, 1.25 * a02, 1.25 * a01, 1.25 * a11 + tr3, 1.25 * a12,
                                   1.25 * a02, 1.25 * a12, 1.25 * a22 + tr3};
                    if (hasDiag) {
                        double *dst = args.values + gd * 9;
                        dst[sub] = sub == 0 ? b[0] : sub == 1 ? b[1] : sub == 2 ? b[2] : b[3];
                        dst[sub + 4] = sub == 0 ? b[4] : sub == 1 ? b[5] : sub == 2 ? b[6] : b[7];
                        if (sub == 0) dst[8] = b[8];
                    }
                    if (args.fusePrec) {
                        if (!isInterface) {
                            int mx = 0, my = 0, mz = 0;
                            if (args.checkBounds) {
                                mx = __ldg (args.checkBounds + node);
                                my = __ldg (args.checkBounds + (size_t)args.nbNodes + node);
                                mz = __ldg (args.checkBounds + 2 * (size_t)args.nbNodes + node);
                            }
                            mask_block (b, mx, my, mz);
                            if (hasDiag) invert3_lu (b);
                        }
                        double *dst = args.prec + (size_t)node * 9;
                        dst[sub] = sub == 0 ? b[0] : sub == 1 ? b[1] : sub == 2 ? b[2] : b[3];
                        dst[sub + 4] = sub == 0 ? b[4] : sub == 1 ? b[5] : sub == 2 ? b[6] : b[7];
                        if (sub == 0) dst[8] = b[8];
                    }
                }
            }
        }
        // no barrier: the off-diagonal pass reads nothing the diagonal pass wrote

        // ---- next tile's coordinates start travelling ---------------------------------------------
        if (hasNext) {
            mbar_wait_or_trap (headFull + ((k + 1) & 1), ((k + 1) >> 1) & 1);
            gather_coords (sHead0 + ((k + 1) & 1) * headBytes);
        }

        // ---- 4. off-diagonal blocks, one lane per CSR entry ----------------------------------------
        const int nbBatches = hdr.nbBatches;
        for (int b = nWarps - 1 - warp; b < nbBatches; b += nWarps) {   // the low warps ran the diagonal pass
            const TileBatch tb = batches[b];
            const int q = laneEntry[b * 32 + lane];
            const bool live = q != 0xFFFF;
            const int r = live ? entryRow[q] : 0;
            const TileRow tr = sRows[r];
            const bool isDiag = live && q == tr.diagLocal;
            const int g = tr.valueStart + (q - tr.localStart);         // global CSR entry
            const uint16_t *codes = pairCodes + tb.codeBase + lane;

            double acc[OPDIM];
            #pragma unroll
            for (int i = 0; i < OPDIM; i++) acc[i] = 0.0;
            #pragma unroll 2
            for (int t = 0; t < tb.steps; t++) {
                const int code = codes[t * 32];
                const int e = code >> 4;
                if (OPDIM == 1) { acc[0] += cX[code]; continue; }
                const int va = ((code >> 2) & 3) * strideE + e, vb = (code & 3) * strideE + e;
                const double ax = cX[va], ay = cY[va], az = cZ[va];
                const double bx = cX[vb], by = cY[vb], bz = cZ[vb];
                acc[0] += ax * bx; acc[1 % OPDIM] += ax * by; acc[2 % OPDIM] += ax * bz;
                acc[3 % OPDIM] += ay * bx; acc[4 % OPDIM] += ay * by; acc[5 % OPDIM] += ay * bz;
                acc[6 % OPDIM] += az * bx; acc[7 % OPDIM] += az * by; acc[8 % OPDIM] += az * bz;
            }

            if (OPDIM == 1) {
                if (live && !isDiag) args.values[g] = acc[0];
            }
            else {
                const double trA = acc[0] + acc[4 % OPDIM] + acc[8 % OPDIM];
                double blk[9];
                #pragma unroll
                for (int i = 0; i < 9; i++) {
                    blk[i] = 1.25 * acc[i % OPDIM] + ((i == 0 || i == 4 || i == 8) ? trA : 0.0);
                }
                // each half-warp holds consecutive entries of one row: one contiguous run each,
                // streamed out through a 144-double slab, half-warp after half-warp; the 9 slots of
                // the row's diagonal entry (written by the diagonal pass) are stepped over
                const unsigned liveMask = __ballot_sync (0xffffffffu, live);
                const unsigned diagMask = __ballot_sync (0xffffffffu, isDiag);
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    if ((lane >> 4) == h) {
                        #pragma unroll
                        for (int i = 0; i < 9; i++) slab[(lane & 15) * 9 + i] = blk[i];
                    }
                    __syncwarp ();
                    const int run = __popc ((liveMask >> (16 * h)) & 0xffffu) * 9;
                    const unsigned dm = (diagMask >> (16 * h)) & 0xffffu;
                    const int dlo = dm ? (__ffs (dm) - 1) * 9 : -16;
                    double *out = args.values + (size_t)__shfl_sync (0xffffffffu, g, 16 * h) * 9;
                    #pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const int m = i * 32 + lane;
                        if (m < run && (unsigned)(m - dlo) >= 9u) out[m] = slab[m];
                    }
                    __syncwarp ();
                }
            }
        }

        offCur = offNext; offNext = offAfter;
        cp_async_wait_all ();  // next tile's coordinates are in
        __syncthreads ();      // every reader of this tile's records / coefficients is done
    }
}

// ------------------------------------------------------------------------------------
// Pipelined variant: ONE persistent CTA of 16 warps per SM, warps specialised by role, two
// tiles in flight.  While the row warps run the shared-memory-bound diagonal / off-diagonal
// passes of tile t, the coefficient warps run the FP64-bound element pass of tile t+1 and the
// producer warp has the plan record of tile t+2 in flight (TMA) and gathers the coordinates
// of tile t+1 with cp.async.  Stages are handed over through mbarriers:
//   blobFull[s]   TMA bytes of the record have landed            (producer arms, TMA completes)
//   coordFull[s]  the 32 producer lanes' cp.async copies are done (cp.async.mbarrier.arrive)
//   coefFull[s]   the coefficient warps have written all planes
//   stageFree[s]  every row-warp thread is done with the stage
// ------------------------------------------------------------------------------------
constexpr int kPipeWarps = 24, kCoefWarps = 4, kRowWarps = kPipeWarps - 1 - kCoefWarps;
constexpr int kPipeThreads = kPipeWarps * 32;

__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
    asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}

// Arrive on `bar` once all cp.async copies this thread issued so far have completed.
__device__ __forceinline__ void cp_async_arrive (uint64_t *bar)
{
    asm volatile ("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void cp_async_8 (void *dst, const void *src)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32 (dst)), "l"(src) : "memory");
}

// Bounded wait: a protocol bug must trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait_bounded (uint64_t *bar, unsigned parity)
{
    unsigned done = 0;
    for (long spin = 0; spin < (1l << 22); spin++) {
        asm volatile (
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap ();
}

struct PipeStage {
    unsigned char *blob;
    double *sX, *sY, *sZ, *cX, *cY, *cZ, *sDiag;
    uint64_t *blobFull, *coordFull, *coefFull, *stageFree;
};

template <int OPDIM>
__global__ void __launch_bounds__(kPipeThreads, 1)
tiled_pipeline_kernel (const TiledArgs args)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const DeviceTilePlan &P = args.plan;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strideE = P.elemStride;

    // shared memory: per stage [blob][coordinates][coefficient planes][diagonal blocks], then the
    // row warps' slabs, then 8 mbarriers
    const size_t blobBytes = (P.maxBlobBytes + 127u) & ~127u;
    const size_t stageDoubles = 3 * (size_t)P.maxNodesRef + 12 * (size_t)strideE + (size_t)P.maxRows * OPDIM;
    const size_t stageBytes = blobBytes + ((stageDoubles * 8 + 127) & ~(size_t)127);
    double *slabs = reinterpret_cast<double*> (smemRaw + 2 * stageBytes);
    uint64_t *bars = reinterpret_cast<uint64_t*> (slabs + (OPDIM == 9 ? kRowWarps * 288 : 0));
    auto stage = [&] (int s) {                 // computed, not stored: keeps everything in registers
        PipeStage S;
        S.blob = smemRaw + s * stageBytes;
        S.sX = reinterpret_cast<double*> (S.blob + blobBytes);
        S.sY = S.sX + P.maxNodesRef;
        S.sZ = S.sY + P.maxNodesRef;
        S.cX = S.sZ + P.maxNodesRef;
        S.cY = S.cX + 4 * strideE;
        S.cZ = S.cY + 4 * strideE;
        S.sDiag = S.cZ + 4 * strideE;
        S.blobFull = bars + 4 * s; S.coordFull = bars + 4 * s + 1;
        S.coefFull = bars + 4 * s + 2; S.stageFree = bars + 4 * s + 3;
        return S;
    };
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            const PipeStage S = stage (s);
            mbar_init (S.blobFull, 1);
            mbar_init (S.coordFull, 32);
            mbar_init (S.coefFull, kCoefWarps * 32);
            mbar_init (S.stageFree, kRowWarps * 32);
        }
    }
    __syncthreads ();

    const int firstTile = args.firstTile + blockIdx.x, tileStep = gridDim.x;

    if (warp == 0) {
        // ================= producer: plan records (TMA) and coordinates (cp.async) =========
        int k = 0;
        for (int tile = firstTile; tile < args.lastTile; tile += tileStep, k++) {
            const PipeStage S = stage (k & 1);
            const unsigned use = (unsigned)(k >> 1);
            if (use > 0) mbar_wait_bounded (S.stageFree, (use - 1) & 1);     // previous tenant is gone
            if (lane == 0) {
                const uint64_t off = record_offset (P.tileOffset[tile]);
                const unsigned bytes = (unsigned)(record_offset (P.tileOffset[tile + 1]) - off);
                mbar_expect_tx (S.blobFull, bytes);
                bulk_load (S.blob, P.blob + off, bytes, S.blobFull);
            }
            mbar_wait_bounded (S.blobFull, use & 1);
            const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (S.blob);
            const int *tileNodes = reinterpret_cast<const int*> (S.blob + hdr.offNodes);
            const int nbNodesRef = hdr.nbNodesRef;
            for (int n = lane; n < nbNodesRef; n += 32) {
                const double *q = args.coord + (size_t)tileNodes[n] * 3;
                cp_async_8 (S.sX + n, q); cp_async_8 (S.sY + n, q + 1); cp_async_8 (S.sZ + n, q + 2);
            }
            cp_async_arrive (S.coordFull);
        }
    }
    else if (warp <= kCoefWarps) {
        // ================= coefficient warps: elem_coef of every tile element ==============
        const int ct = tid - 32, nCoef = kCoefWarps * 32;
        int k = 0;
        for (int tile = firstTile; tile < args.lastTile; tile += tileStep, k++) {
            const PipeStage S = stage (k & 1);
            const unsigned use = (unsigned)(k >> 1);
            mbar_wait_bounded (S.blobFull, use & 1);
            mbar_wait_bounded (S.coordFull, use & 1);
            const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (S.blob);
            const ushort4 *tileElems = reinterpret_cast<const ushort4*> (S.blob + hdr.offElems);
            const int nbElems = hdr.nbElems;
            if (ct < 64) {                                // the 16 all-zero slots of each plane (padding codes)
                const int z = (ct >> 4) * strideE + nbElems + (ct & 15);
                S.cX[z] = 0.0; S.cY[z] = 0.0; S.cZ[z] = 0.0;
            }
            for (int e = ct; e < nbElems; e += nCoef) {
                const ushort4 ln = tileElems[e];
                if (ln.x == 0xFFFF) continue;             // hole of the coset numbering
                const int ids[4] = {ln.x, ln.y, ln.z, ln.w};
                double p[12], c[12];
                #pragma unroll
                for (int i = 0; i < 4; i++) { p[3 * i] = S.sX[ids[i]]; p[3 * i + 1] = S.sY[ids[i]]; p[3 * i + 2] = S.sZ[ids[i]]; }
                elem_coef (p, c);
                if (OPDIM == 1) {                         // the 10 dot products (assembly.cc:539-541)
                    const int PS = strideE - 4;
                    #pragma unroll
                    for (int a = 0; a < 4; a++) {
                        #pragma unroll
                        for (int b = a; b < 4; b++) {
                            const double dot = c[3 * a] * c[3 * b] + c[3 * a + 1] * c[3 * b + 1] + c[3 * a + 2] * c[3 * b + 2];
                            S.cX[a == b ? lap_diag_slot (a, e, PS) : lap_pair_slot (a, b, e, PS)] = dot;
                        }
                    }
                }
                else {
                    #pragma unroll
                    for (int a = 0; a < 4; a++) {
                        S.cX[a * strideE + e] = c[3 * a]; S.cY[a * strideE + e] = c[3 * a + 1]; S.cZ[a * strideE + e] = c[3 * a + 2];
                    }
                }
            }
            mbar_arrive (S.coefFull);
        }
    }
    else {
        // ================= row warps: diagonal pass, off-diagonal pass, write-out ============
        const int rw = warp - 1 - kCoefWarps;
        double *slab = slabs + rw * 288;
        int k = 0;
        for (int tile = firstTile; tile < args.lastTile; tile += tileStep, k++) {
            const PipeStage S = stage (k & 1);
            const unsigned use = (unsigned)(k >> 1);
            mbar_wait_bounded (S.blobFull, use & 1);
            mbar_wait_bounded (S.coefFull, use & 1);
            const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (S.blob);
            const int nbRows = hdr.nbRows;
            const TileRow *sRows = reinterpret_cast<const TileRow*> (S.blob + sizeof (TileBlobHeader));
            const uint8_t *entryRow = S.blob + hdr.offEntryRow;
            const uint16_t *laneEntry = reinterpret_cast<const uint16_t*> (S.blob + hdr.offLaneEntry);
            const TileBatch *batches = reinterpret_cast<const TileBatch*> (S.blob + hdr.offBatches);
            const uint16_t *diagCodes = reinterpret_cast<const uint16_t*> (S.blob + hdr.offDiag);
            const uint16_t *pairCodes = reinterpret_cast<const uint16_t*> (S.blob + hdr.offPair);
            const double *cX = S.cX, *cY = S.cY, *cZ = S.cZ;
            double *sDiag = S.sDiag;

            for (int r0 = rw * 8; r0 < nbRows; r0 += kRowWarps * 8) {
                const int r = r0 + (lane >> 2), sub = lane & 3;
                const bool live = r < nbRows;
                const int begin = live ? sRows[r].diagCodeBase : 0;
                const int end   = live ? sRows[r + 1].diagCodeBase : 0;
                double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
                for (int q = begin + sub; q < end; q += 4) {
                    const int code = diagCodes[q];
                    if (OPDIM == 1) { a00 += cX[code]; continue; }
                    const int v = (code & 3) * strideE + (code >> 2);
                    const double x = cX[v], y = cY[v], z = cZ[v];
                    a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
                }
                #pragma unroll
                for (int off = 1; off <= 2; off <<= 1) {
                    a00 += __shfl_xor_sync (0xffffffffu, a00, off);
                    if (OPDIM == 9) {
                        a01 += __shfl_xor_sync (0xffffffffu, a01, off);
                        a02 += __shfl_xor_sync (0xffffffffu, a02, off);
                        a11 += __shfl_xor_sync (0xffffffffu, a11, off);
                        a12 += __shfl_xor_sync (0xffffffffu, a12, off);
                        a22 += __shfl_xor_sync (0xffffffffu, a22, off);
                    }
                }
                if (live && sub == 0) {
                    const int nodeField = sRows[r].node;
                    const int node = nodeField & 0x7fffffff;
                    const bool isInterface = nodeField < 0;
                    const bool hasDiag = sRows[r].diagLocal != 0xFFFF;
                    if (OPDIM == 1) {
                        sDiag[r] = a00;
                        if (args.fusePrec) args.prec[node] = isInterface ? a00 : 1.0 / a00;
                    }
                    else {
                        const double tr = a00 + a11 + a22;
                        double b[9] = {1.25 * a00 + tr, 1.25 * a01, 1.25 * a02,
                                       1.25 * a01, 1.25 * a11 + tr, 1.25 * a12,
                                       1.25 * a02, 1.25 * a12, 1.25 * a22 + tr};
                        #pragma unroll
                        for (int q = 0; q < 9; q++) sDiag[r * 9 + q] = b[q];
                        if (args.fusePrec) {
                            if (!isInterface) {
                                int mx = 0, my = 0, mz = 0;
                                if (args.checkBounds) {
                                    mx = __ldg (args.checkBounds + node);
                                    my = __ldg (args.checkBounds + (size_t)args.nbNodes + node);
                                    mz = __ldg (args.checkBounds + 2 * (size_t)args.nbNodes + node);
                                }
                                mask_block (b, mx, my, mz);
                                if (hasDiag) invert3_lu (b);
                            }
                            double *dst = args.prec + (size_t)node * 9;
                            #pragma unroll
                            for (int q = 0; q < 9; q++) dst[q] = b[q];
                        }
                    }
                }
            }
            // the diagonal blocks of all rows must be in place before any entry lane copies one
            asm volatile ("bar.sync 1, %0;" :: "n"(kRowWarps * 32) : "memory");

            const int nbBatches = hdr.nbBatches;
            for (int b = rw; b < nbBatches; b += kRowWarps) {
                const TileBatch tb = batches[b];
                const int q = laneEntry[b * 32 + lane];
                const bool live = q != 0xFFFF;
                const int r = live ? entryRow[q] : 0;
                const TileRow tr = sRows[r];
                const bool isDiag = live && q == tr.diagLocal;
                const int g = tr.valueStart + (q - tr.localStart);         // global CSR entry
                const uint16_t *codes = pairCodes + tb.codeBase + lane;

                double acc[OPDIM];
                #pragma unroll
                for (int i = 0; i < OPDIM; i++) acc[i] = 0.0;
                #pragma unroll 2
                for (int t = 0; t < tb.steps; t++) {
                    const int code = codes[t * 32];
                    if (OPDIM == 1) { acc[0] += cX[code]; continue; }
                    const int e = code >> 4;
                    const int va = ((code >> 2) & 3) * strideE + e, vb = (code & 3) * strideE + e;
                    const double ax = cX[va], ay = cY[va], az = cZ[va];
                    const double bx = cX[vb], by = cY[vb], bz = cZ[vb];
                    {
                        acc[0] += ax * bx; acc[1] += ax * by; acc[2] += ax * bz;
                        acc[3] += ay * bx; acc[4] += ay * by; acc[5] += ay * bz;
                        acc[6] += az * bx; acc[7] += az * by; acc[8] += az * bz;
                    }
                }
                if (OPDIM == 1) {
                    if (live) args.values[g] = isDiag ? sDiag[r] : acc[0];
                }
                else {
                    const double trA = acc[0] + acc[4] + acc[8];
                    #pragma unroll
                    for (int i = 0; i < 9; i++) {
                        double v = 1.25 * acc[i] + ((i == 0 || i == 4 || i == 8) ? trA : 0.0);
                        if (isDiag) v = sDiag[r * 9 + i];
                        slab[lane * 9 + i] = v;
                    }
                    __syncwarp ();
                    // each half-warp holds consecutive entries of one row: two contiguous runs
                    const unsigned liveMask = __ballot_sync (0xffffffffu, live);
                    const int run0 = __popc (liveMask & 0xffffu) * 9, run1 = __popc (liveMask >> 16) * 9;
                    double *out0 = args.values + (size_t)__shfl_sync (0xffffffffu, g, 0) * 9;
                    double *out1 = args.values + (size_t)__shfl_sync (0xffffffffu, g, 16) * 9 - 144;
                    #pragma unroll
                    for (int i = 0; i < 9; i++) {
                        const int m = i * 32 + lane;
                        if (m < 144) { if (m < run0) out0[m] = slab[m]; }
                        else if (m - 144 < run1) out1[m] = slab[m];
                    }
                    __syncwarp ();
                }
            }
            mbar_arrive (S.stageFree);
        }
    }
}

}  // namespace

size_t tiled_smem_bytes (int operatorID, const DeviceTilePlan &plan, int threads)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    const size_t scratch = std::max<size_t> (3 * (size_t)plan.maxNodesRef, opDim == 9 ? (size_t)(threads / 32) * 288 : 0);
    const size_t doubles = 12 * (size_t)plan.elemStride + (size_t)plan.maxRows * opDim + scratch;
    return (((size_t)plan.maxBlobBytes + 127) & ~(size_t)127) + doubles * sizeof (double) + 16;
}

size_t tiled_pipeline_smem_bytes (int operatorID, const DeviceTilePlan &plan)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    const size_t blobBytes = ((size_t)plan.maxBlobBytes + 127) & ~(size_t)127;
    const size_t stageDoubles = 3 * (size_t)plan.maxNodesRef + 12 * (size_t)plan.elemStride + (size_t)plan.maxRows * opDim;
    const size_t stageBytes = blobBytes + ((stageDoubles * 8 + 127) & ~(size_t)127);
    return 2 * stageBytes + (opDim == 9 ? (size_t)kRowWarps * 288 * 8 : 0) + 8 * sizeof (uint64_t);
}

int tiled_pipeline_threads () { return kPipeThreads; }

size_t tiled_prefetch_smem_bytes (int operatorID, const DeviceTilePlan &plan, int threads)
{
    const int opDim = operatorID == 0 ? 1 : 9;
    const size_t headBytes = ((size_t)plan.maxHeadBytes + 127) & ~(size_t)127, tailBytes = ((size_t)plan.maxTailBytes + 127) & ~(size_t)127;
    const size_t doubles = 12 * (size_t)plan.elemStride + 3 * (size_t)plan.maxNodesRef +
                           (opDim == 9 ? (size_t)(threads / 32) * 144 : 0);
    return 2 * headBytes + tailBytes + doubles * sizeof (double) + 3 * sizeof (uint64_t);
}

// Instantiations: stride 420 = default caps (36 rows / 384 elements, three CTAs per SM),
// 660 = 64 rows / 624 elements (two CTAs per SM), 0 = any other cap (stride read at run time).
constexpr int kStrideSmall = 420, kStrideLarge = 660;

template <class K>
cudaError_t opt_in (K kernel)
{
    // The attribute belongs to the kernel, not to a context: several contexts (one per
    // subdomain) with different tile caps share it, so always opt in to the device maximum.
    return cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

template <int OPDIM>
cudaError_t configure_op ()
{
    cudaError_t e;
    if ((e = opt_in (tiled_assembly_kernel<OPDIM, 3, kStrideSmall>)) != cudaSuccess) return e;
    if ((e = opt_in (tiled_assembly_kernel<OPDIM, 2, kStrideLarge>)) != cudaSuccess) return e;
    if ((e = opt_in (tiled_assembly_kernel<OPDIM, 3, 0>)) != cudaSuccess) return e;
    if ((e = opt_in (tiled_assembly_kernel<OPDIM, 2, 0>)) != cudaSuccess) return e;
    if ((e = opt_in (tiled_prefetch_kernel<OPDIM, kStrideSmall>)) != cudaSuccess) return e;
    if ((e = opt_in (tiled_prefetch_kernel<OPDIM, 0>)) != cudaSuccess) return e;
    return opt_in (tiled_pipeline_kernel<OPDIM>);
}

cudaError_t tiled_configure (int operatorID, size_t smemBytes)
{
    (void)smemBytes;
    return operatorID == 0 ? configure_op<1> () : configure_op<9> ();
}

template <int OPDIM>
void launch_op (const TiledArgs &args, int grid, int threads, size_t smemBytes, cudaStream_t stream, bool prefetch)
{
    if (prefetch) {
        if (args.plan.elemStride == kStrideSmall) tiled_prefetch_kernel<OPDIM, kStrideSmall><<<grid, threads, smemBytes, stream>>> (args);
        else                                      tiled_prefetch_kernel<OPDIM, 0><<<grid, threads, smemBytes, stream>>> (args);
        return;
    }
    if (threads == kPipeThreads) {          // pipelined variant: one 24-warp CTA per SM
        tiled_pipeline_kernel<OPDIM><<<grid, kPipeThreads, smemBytes, stream>>> (args);
        return;
    }
    // three co-resident CTAs per SM when the tile record is small enough (register cap 85)
    const bool three = smemBytes + 1024 <= (size_t)(227 * 1024) / 3 && threads * 3 <= 2048;
    const int stride = args.plan.elemStride;
    if (three && stride == kStrideSmall)       tiled_assembly_kernel<OPDIM, 3, kStrideSmall><<<grid, threads, smemBytes, stream>>> (args);
    else if (!three && stride == kStrideLarge) tiled_assembly_kernel<OPDIM, 2, kStrideLarge><<<grid, threads, smemBytes, stream>>> (args);
    else if (three)                            tiled_assembly_kernel<OPDIM, 3, 0><<<grid, threads, smemBytes, stream>>> (args);
    else                                       tiled_assembly_kernel<OPDIM, 2, 0><<<grid, threads, smemBytes, stream>>> (args);
}

cudaError_t launch_tiled (int operatorID, const DeviceTilePlan &plan, int firstTile, int nbTiles, int ctas,
                          int threads, size_t smemBytes, const double *coord, double *values,
                          double *prec, const int *checkBounds, int nbNodes, int fusePrec,
                          cudaStream_t stream, bool prefetch)
{
    if (nbTiles <= 0) return cudaSuccess;
    TiledArgs args;
    args.plan = plan; args.coord = coord; args.values = values; args.prec = prec;
    args.checkBounds = checkBounds; args.nbNodes = nbNodes; args.fusePrec = fusePrec;
    args.firstTile = firstTile; args.lastTile = firstTile + nbTiles;
    const int grid = std::max (1, std::min (ctas, nbTiles));
    if (operatorID == 0) launch_op<1> (args, grid, threads, smemBytes, stream, prefetch);
    else                 launch_op<9> (args, grid, threads, smemBytes, stream, prefetch);
    return cudaGetLastError ();
}

}  // namespace mfb
