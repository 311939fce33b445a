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

template <int OPDIM, int MINBLOCKS>
__global__ void __launch_bounds__(256, MINBLOCKS)
tiled_assembly_kernel (const TiledArgs args)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const DeviceTilePlan &P = args.plan;
    const int tid = threadIdx.x, nThreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nWarps = nThreads >> 5;
    const int strideE = P.elemStride;

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
            const uint64_t off = P.tileOffset[tile];
            const unsigned bytes = (unsigned)(P.tileOffset[tile + 1] - off);
            mbar_expect_tx (bar, bytes);
            bulk_load (sBlob, P.blob + off, bytes, bar);
        }
        mbar_wait (bar, parity);
        parity ^= 1;
        const TileBlobHeader &hdr = *reinterpret_cast<const TileBlobHeader*> (sBlob);
        const int nbRows = hdr.nbRows, nbElems = hdr.nbElems, nbEntries = hdr.nbEntries, nbNodesRef = hdr.nbNodesRef;
        const TileRow *sRows = reinterpret_cast<const TileRow*> (sBlob + sizeof (TileBlobHeader));
        const int *tileNodes = reinterpret_cast<const int*> (sBlob + hdr.offNodes);
        const ushort4 *tileElems = reinterpret_cast<const ushort4*> (sBlob + hdr.offElems);
        const uint8_t *entryRow = sBlob + hdr.offEntryRow;
        const TileBatch *batches = reinterpret_cast<const TileBatch*> (sBlob + hdr.offBatches);
        const uint16_t *diagCodes = reinterpret_cast<const uint16_t*> (sBlob + hdr.offDiag);
        const uint16_t *pairCodes = reinterpret_cast<const uint16_t*> (sBlob + hdr.offPair);

        // ---- 1. coordinates ------------------------------------------------------------
        for (int n = tid; n < nbNodesRef; n += nThreads) {
            const double *q = args.coord + (size_t)tileNodes[n] * 3;
            sX[n] = __ldg (q); sY[n] = __ldg (q + 1); sZ[n] = __ldg (q + 2);
        }
        if (tid == 0) { cX[nbElems] = 0.0; cY[nbElems] = 0.0; cZ[nbElems] = 0.0; }   // padding slot
        __syncthreads ();

        // ---- 2. gradient coefficients ----------------------------------------------------
        for (int e = tid; e < nbElems; e += nThreads) {
            const ushort4 ln = tileElems[e];
            const int ids[4] = {ln.x, ln.y, ln.z, ln.w};
            double p[12], c[12];
            #pragma unroll
            for (int i = 0; i < 4; i++) { p[3 * i] = sX[ids[i]]; p[3 * i + 1] = sY[ids[i]]; p[3 * i + 2] = sZ[ids[i]]; }
            elem_coef (p, c);
            #pragma unroll
            for (int a = 0; a < 4; a++) {
                cX[a * strideE + e] = c[3 * a]; cY[a * strideE + e] = c[3 * a + 1]; cZ[a * strideE + e] = c[3 * a + 2];
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
                const int v = (code & 3) * strideE + (code >> 2);
                const double x = cX[v], y = cY[v], z = cZ[v];
                if (OPDIM == 1) { a00 += x * x + y * y + z * z; }
                else { a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z; }
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
        __syncthreads ();      // sDiag complete; the coordinate planes are dead: scratch becomes the slabs

        // ---- 4. off-diagonal blocks, one lane per CSR entry ----------------------------------
        const int nbBatches = (nbEntries + 31) >> 5;
        for (int b = warp; b < nbBatches; b += nWarps) {
            const TileBatch tb = batches[b];
            const int q = b * 32 + lane;
            const bool live = q < nbEntries;
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
                const int va = ((code >> 2) & 3) * strideE + e, vb = (code & 3) * strideE + e;
                const double ax = cX[va], ay = cY[va], az = cZ[va];
                const double bx = cX[vb], by = cY[vb], bz = cZ[vb];
                if (OPDIM == 1) {
                    acc[0] += ax * bx + ay * by + az * bz;
                }
                else {
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
                const int liveEntries = min (32, nbEntries - b * 32);
                #pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int m = k * 32 + lane;
                    const int ent = m / 9, comp = m - ent * 9;
                    const int gs = __shfl_sync (0xffffffffu, g, ent);
                    if (ent < liveEntries) args.values[(size_t)gs * 9 + comp] = slab[m];
                }
                __syncwarp ();
            }
        }
        __syncthreads ();      // every reader of the blob / coefficients is done before the next tile lands
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

template <int OPDIM, int MINBLOCKS>
cudaError_t configure_one ()
{
    // The attribute belongs to the kernel, not to a context: several contexts (one per
    // subdomain) with different tile caps share it, so always opt in to the device maximum.
    return cudaFuncSetAttribute (tiled_assembly_kernel<OPDIM, MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t tiled_configure (int operatorID, size_t smemBytes)
{
    (void)smemBytes;
    cudaError_t e;
    if (operatorID == 0) {
        if ((e = configure_one<1, 2> ()) != cudaSuccess) return e;
        return configure_one<1, 3> ();
    }
    if ((e = configure_one<9, 2> ()) != cudaSuccess) return e;
    return configure_one<9, 3> ();
}

cudaError_t launch_tiled (int operatorID, const DeviceTilePlan &plan, int firstTile, int nbTiles, int ctas,
                          int threads, size_t smemBytes, const double *coord, double *values,
                          double *prec, const int *checkBounds, int nbNodes, int fusePrec,
                          cudaStream_t stream)
{
    if (nbTiles <= 0) return cudaSuccess;
    TiledArgs args;
    args.plan = plan; args.coord = coord; args.values = values; args.prec = prec;
    args.checkBounds = checkBounds; args.nbNodes = nbNodes; args.fusePrec = fusePrec;
    args.firstTile = firstTile; args.lastTile = firstTile + nbTiles;
    const int grid = std::max (1, std::min (ctas, nbTiles));
    // three co-resident CTAs per SM when the tile record is small enough (register cap 85)
    const bool three = smemBytes + 1024 <= (227 * 1024) / 3 && threads * 3 <= 2048;
    if (operatorID == 0) {
        if (three) tiled_assembly_kernel<1, 3><<<grid, threads, smemBytes, stream>>> (args);
        else       tiled_assembly_kernel<1, 2><<<grid, threads, smemBytes, stream>>> (args);
    }
    else {
        if (three) tiled_assembly_kernel<9, 3><<<grid, threads, smemBytes, stream>>> (args);
        else       tiled_assembly_kernel<9, 2><<<grid, threads, smemBytes, stream>>> (args);
    }
    return cudaGetLastError ();
}

}  // namespace mfb
