// Element-parallel scatter-add kernels: the ATOMIC path (native FP64 red.global.add)
// and the COLOR path (plain read-modify-write inside one of the reference's colours).
// Both do, per element, what assembly_{lap,ela}_seq do (src/assembly.cc:485-588,
// :332-479) through the precomputed elemToEdge index (the OPTIMIZED variant, :380-412
// / :533-544); they differ only in how a contribution reaches nodeToNodeValue.
//
// Also here: the once-per-context index kernels (elemToEdge on the GPU, diagonal index).
#include "kernels.cuh"
#include "device_math.cuh"

namespace mfb {

namespace {

template <bool ATOMIC>
__device__ __forceinline__ void accumulate (double *addr, double v)
{
    if (ATOMIC) atomicAdd (addr, v);       // RED.E.ADD.F64 (result unused)
    else        *addr += v;                // conflict-free by colouring
}

// One element: what assembly_{lap,ela}_seq do for it (src/assembly.cc:332-479, :485-588).
template <int OPDIM, bool ATOMIC>
__device__ __forceinline__ void scatter_element (const double *__restrict__ coord, const int4 *__restrict__ elemToNode,
                                                 const int4 *__restrict__ elemToEdge, double *__restrict__ values, size_t e)
{
    const int4 nd = __ldg (elemToNode + e);                 // 1-based ids, one 16 B load
    const int ids[4] = {nd.x - 1, nd.y - 1, nd.z - 1, nd.w - 1};
    double p[12], c[12];
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const double *q = coord + (size_t)ids[i] * 3;
        p[3 * i] = __ldg (q); p[3 * i + 1] = __ldg (q + 1); p[3 * i + 2] = __ldg (q + 2);
    }
    elem_coef (p, c);

    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int4 idx4 = __ldg (elemToEdge + e * 4 + j);
        const int idx[4] = {idx4.x, idx4.y, idx4.z, idx4.w};
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (OPDIM == 1) {
                const double v = c[3 * j] * c[3 * k] + c[3 * j + 1] * c[3 * k + 1] + c[3 * j + 2] * c[3 * k + 2];
                accumulate<ATOMIC> (values + idx[k], v);
            }
            else {
                double blk[9];
                ela_block (c + 3 * j, c + 3 * k, blk);
                double *dst = values + (size_t)idx[k] * 9;
                #pragma unroll
                for (int q = 0; q < 9; q++) accumulate<ATOMIC> (dst + q, blk[q]);
            }
        }
    }
}

template <int OPDIM, bool ATOMIC>
__global__ void __launch_bounds__(128)
scatter_elements_kernel (const double *__restrict__ coord, const int4 *__restrict__ elemToNode,
                         const int4 *__restrict__ elemToEdge, double *__restrict__ values,
                         int firstElem, int nbElemInterval)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbElemInterval) return;
    scatter_element<OPDIM, ATOMIC> (coord, elemToNode, elemToEdge, values, (size_t)firstElem + t);
}

// BLOCKCOLOR path (host/mesh_topology.h: build_block_coloring): one CTA per block of a block colour — the blocks of one
// launch share no node —, the block's elements local colour by local colour with a block barrier in between: plain
// read-modify-write without atomics, and a CSR row is updated many times by ONE CTA while it sits in L1 / L2.
constexpr int kBlockColorThreads = 128;

// The element of scatter_element for a CTA that owns its rows: the four CSR entries of a row of the element matrix are
// read together, updated, written together — four round trips to L2 per element instead of sixteen (the plain `+=` of
// scatter_element has to keep the order of its sixteen read-modify-writes: their indices could coincide).  The first
// measurement of this path (9.2 ms per EIB iteration) was bound by exactly that chain.
template <int OPDIM>
__device__ __forceinline__ void scatter_element_rows (const double *__restrict__ coord, const int4 *__restrict__ elemToNode,
                                                      const int4 *__restrict__ elemToEdge, double *values, size_t e)
{
    const int4 nd = __ldg (elemToNode + e);
    const int ids[4] = {nd.x - 1, nd.y - 1, nd.z - 1, nd.w - 1};
    double p[12], c[12];
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const double *q = coord + (size_t)ids[i] * 3;
        p[3 * i] = __ldg (q); p[3 * i + 1] = __ldg (q + 1); p[3 * i + 2] = __ldg (q + 2);
    }
    elem_coef (p, c);
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int4 idx4 = __ldg (elemToEdge + e * 4 + j);
        const int idx[4] = {idx4.x, idx4.y, idx4.z, idx4.w};
        double old[4 * OPDIM];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            #pragma unroll
            for (int q = 0; q < OPDIM; q++) old[k * OPDIM + q] = __ldcg (values + (size_t)idx[k] * OPDIM + q);
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (OPDIM == 1) {
                old[k] += c[3 * j] * c[3 * k] + c[3 * j + 1] * c[3 * k + 1] + c[3 * j + 2] * c[3 * k + 2];
            }
            else {
                double blk[9];
                ela_block (c + 3 * j, c + 3 * k, blk);
                #pragma unroll
                for (int q = 0; q < 9; q++) old[k * OPDIM + q % OPDIM] += blk[q];
            }
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            #pragma unroll
            for (int q = 0; q < OPDIM; q++) __stcg (values + (size_t)idx[k] * OPDIM + q, old[k * OPDIM + q]);
        }
    }
}

template <int OPDIM>
__global__ void __launch_bounds__(kBlockColorThreads, 3)
scatter_blocks_kernel (const double *__restrict__ coord, const int4 *__restrict__ elemToNode,
                       const int4 *__restrict__ elemToEdge, double *values,
                       const int *__restrict__ localIndex, const int *__restrict__ localStart, int firstBlock)
{
    const int b = firstBlock + blockIdx.x;
    const int first = localIndex[b], nbLocal = localIndex[b + 1] - first - 1;
    for (int c = 0; c < nbLocal; c++) {
        const int lo = localStart[first + c], hi = localStart[first + c + 1];
        for (int e = lo + threadIdx.x; e < hi; e += kBlockColorThreads) {
            scatter_element_rows<OPDIM> (coord, elemToNode, elemToEdge, values, (size_t)e);
        }
        __syncthreads ();          // the next local colour adds to entries this one has written
    }
}

// create_elemToEdge (src/matrix.cc:25-52) on the device: one thread per (element, j, k).
__global__ void elem_to_edge_kernel (const int *__restrict__ row, const int *__restrict__ col,
                                     const int *__restrict__ elemToNode, int *__restrict__ elemToEdge,
                                     size_t nbPairs, int *__restrict__ missing)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbPairs) return;
    const size_t e = t >> 4;
    const int j = (t >> 2) & 3, k = t & 3;
    const int n1 = elemToNode[e * 4 + j] - 1, n2 = elemToNode[e * 4 + k];
    int found = -1;
    for (int l = row[n1]; l < row[n1 + 1]; l++) {
        if (col[l] == n2) { found = l; break; }
    }
    if (found < 0) atomicAdd (missing, 1);
    elemToEdge[t] = found;
}

// Position of each node's own column in its row (prec_init's search,
// src/preconditioner.cc:78-79), -1 if the row has none.
__global__ void diag_index_kernel (const int *__restrict__ row, const int *__restrict__ col,
                                   int *__restrict__ diagIndex, int nbNodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbNodes) return;
    int found = -1;
    for (int j = row[i]; j < row[i + 1]; j++) {
        if (col[j] - 1 == i) { found = j; break; }
    }
    diagIndex[i] = found;
}

}  // namespace

cudaError_t launch_scatter (int operatorID, bool atomic, const double *coord, const int *elemToNode,
                            const int *elemToEdge, double *values, int firstElem, int count,
                            cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    const int threads = 128, blocks = (count + threads - 1) / threads;
    const int4 *e2n = reinterpret_cast<const int4*> (elemToNode);
    const int4 *e2e = reinterpret_cast<const int4*> (elemToEdge);
    if (operatorID == 0) {
        if (atomic) scatter_elements_kernel<1, true><<<blocks, threads, 0, stream>>> (coord, e2n, e2e, values, firstElem, count);
        else        scatter_elements_kernel<1, false><<<blocks, threads, 0, stream>>> (coord, e2n, e2e, values, firstElem, count);
    }
    else {
        if (atomic) scatter_elements_kernel<9, true><<<blocks, threads, 0, stream>>> (coord, e2n, e2e, values, firstElem, count);
        else        scatter_elements_kernel<9, false><<<blocks, threads, 0, stream>>> (coord, e2n, e2e, values, firstElem, count);
    }
    return cudaGetLastError ();
}

cudaError_t launch_scatter_blocks (int operatorID, const double *coord, const int *elemToNode, const int *elemToEdge,
                                   double *values, const int *localIndex, const int *localStart, int firstBlock,
                                   int nbBlocks, cudaStream_t stream)
{
    if (nbBlocks <= 0) return cudaSuccess;
    const int4 *e2n = reinterpret_cast<const int4*> (elemToNode);
    const int4 *e2e = reinterpret_cast<const int4*> (elemToEdge);
    if (operatorID == 0) scatter_blocks_kernel<1><<<nbBlocks, kBlockColorThreads, 0, stream>>> (coord, e2n, e2e, values, localIndex, localStart, firstBlock);
    else                 scatter_blocks_kernel<9><<<nbBlocks, kBlockColorThreads, 0, stream>>> (coord, e2n, e2e, values, localIndex, localStart, firstBlock);
    return cudaGetLastError ();
}

cudaError_t launch_elem_to_edge (const int *row, const int *col, const int *elemToNode,
                                 int *elemToEdge, int nbElem, int *missing, cudaStream_t stream)
{
    const size_t pairs = (size_t)nbElem * 16;
    if (pairs == 0) return cudaSuccess;
    const int threads = 256;
    const unsigned blocks = (unsigned)((pairs + threads - 1) / threads);
    elem_to_edge_kernel<<<blocks, threads, 0, stream>>> (row, col, elemToNode, elemToEdge, pairs, missing);
    return cudaGetLastError ();
}

cudaError_t launch_diag_index (const int *row, const int *col, int *diagIndex, int nbNodes,
                               cudaStream_t stream)
{
    if (nbNodes <= 0) return cudaSuccess;
    const int threads = 256, blocks = (nbNodes + threads - 1) / threads;
    diag_index_kernel<<<blocks, threads, 0, stream>>> (row, col, diagIndex, nbNodes);
    return cudaGetLastError ();
}

}  // namespace mfb
