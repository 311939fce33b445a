// Preconditioner and interface kernels:
//   prec_init        src/preconditioner.cc:52-87   (zero + copy of each diagonal block)
//   prec_inversion   src/preconditioner.cc:25-49 + src/Fortran/elasclpr.f:2-56
//   halo pack / add  src/halo.cc:77-80 / :113-116
#include "kernels.cuh"
#include "device_math.cuh"

namespace mfb {

namespace {

template <int OPDIM>
__global__ void prec_init_kernel (double *__restrict__ prec, const double *__restrict__ values,
                                  const int *__restrict__ diagIndex, size_t total)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t node = t / OPDIM;
    const int k = (int)(t - node * OPDIM);
    const int d = diagIndex[node];
    prec[t] = d >= 0 ? values[(size_t)d * OPDIM + k] : 0.0;
}

__global__ void prec_invert_lap_kernel (double *__restrict__ prec, int nbNodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbNodes) prec[i] = 1.0 / prec[i];
}

__device__ __forceinline__ void invert_node_block (double b[9], int node, int nbNodes,
                                                   const int *__restrict__ diagIndex,
                                                   const int *__restrict__ checkBounds)
{
    int mx = 0, my = 0, mz = 0;
    if (checkBounds) {
        mx = checkBounds[node];
        my = checkBounds[(size_t)nbNodes + node];
        mz = checkBounds[2 * (size_t)nbNodes + node];
    }
    mask_block (b, mx, my, mz);
    if (diagIndex[node] >= 0) invert3_lu (b);    // elasclpr.f:29-32: only rows with a diagonal entry
}

// One warp per block of 32 consecutive nodes: the 32 x 72 B of blocks are moved with
// coalesced 8-byte accesses through a per-warp shared-memory slab, each lane inverts
// one node's block in registers.
constexpr int kInvWarps = 4;
__global__ void __launch_bounds__(kInvWarps * 32)
prec_invert_ela_kernel (double *__restrict__ prec, const int *__restrict__ diagIndex,
                        const int *__restrict__ checkBounds, int nbNodes)
{
    __shared__ double slab[kInvWarps][32 * 9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int firstNode = (blockIdx.x * kInvWarps + warp) * 32;
    if (firstNode >= nbNodes) return;
    const int nodesHere = min (32, nbNodes - firstNode);
    double *base = prec + (size_t)firstNode * 9;
    for (int m = lane; m < nodesHere * 9; m += 32) slab[warp][m] = base[m];
    __syncwarp ();
    if (lane < nodesHere) {
        double b[9];
        #pragma unroll
        for (int q = 0; q < 9; q++) b[q] = slab[warp][lane * 9 + q];
        invert_node_block (b, firstNode + lane, nbNodes, diagIndex, checkBounds);
        #pragma unroll
        for (int q = 0; q < 9; q++) slab[warp][lane * 9 + q] = b[q];
    }
    __syncwarp ();
    for (int m = lane; m < nodesHere * 9; m += 32) base[m] = slab[warp][m];
}

// Same arithmetic on a list of 0-based node ids (the interface nodes after the halo sum).
__global__ void prec_invert_list_kernel (double *__restrict__ prec, const int *__restrict__ diagIndex,
                                         const int *__restrict__ checkBounds, int nbNodes,
                                         const int *__restrict__ nodes, int count, int operatorID)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int node = nodes[t];
    if (operatorID == 0) { prec[node] = 1.0 / prec[node]; return; }
    double b[9];
    double *blk = prec + (size_t)node * 9;
    #pragma unroll
    for (int q = 0; q < 9; q++) b[q] = blk[q];
    invert_node_block (b, node, nbNodes, diagIndex, checkBounds);
    #pragma unroll
    for (int q = 0; q < 9; q++) blk[q] = b[q];
}

// halo.cc:77-80: bufferSend[j*dim+k] = prec[(intfNodes[j]-1)*dim+k]
__global__ void halo_pack_kernel (double *__restrict__ sendBuf, const double *__restrict__ prec,
                                  const int *__restrict__ intfNodes, int dim, size_t total)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t j = t / dim;
    const int k = (int)(t - j * dim);
    sendBuf[t] = prec[(size_t)(intfNodes[j] - 1) * dim + k];
}

// halo.cc:113-116 as a gather: a node listed in several interfaces (subdomain edges and
// corners) receives its additions from one thread, in increasing interface position —
// the order of the reference's serial (REF) loop — so the sum is deterministic.
__global__ void halo_add_kernel (double *__restrict__ prec, const double *__restrict__ recvBuf,
                                 const int *__restrict__ uniqNodes, const int *__restrict__ slotIndex,
                                 const int *__restrict__ slots, int dim, size_t total)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t u = t / dim;
    const int k = (int)(t - u * dim);
    double *dst = prec + (size_t)uniqNodes[u] * dim + k;
    double acc = *dst;
    for (int s = slotIndex[u]; s < slotIndex[u + 1]; s++) acc += recvBuf[(size_t)slots[s] * dim + k];
    *dst = acc;
}

inline unsigned blocks_for (size_t total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

cudaError_t launch_prec_init (int operatorDim, double *prec, const double *values,
                              const int *diagIndex, int nbNodes, cudaStream_t stream)
{
    const size_t total = (size_t)nbNodes * operatorDim;
    if (total == 0) return cudaSuccess;
    if (operatorDim == 1) prec_init_kernel<1><<<blocks_for (total, 256), 256, 0, stream>>> (prec, values, diagIndex, total);
    else                  prec_init_kernel<9><<<blocks_for (total, 256), 256, 0, stream>>> (prec, values, diagIndex, total);
    return cudaGetLastError ();
}

cudaError_t launch_prec_inversion (int operatorID, double *prec, const int *diagIndex,
                                   const int *checkBounds, int nbNodes, cudaStream_t stream)
{
    if (nbNodes <= 0) return cudaSuccess;
    if (operatorID == 0) {
        prec_invert_lap_kernel<<<blocks_for (nbNodes, 256), 256, 0, stream>>> (prec, nbNodes);
    }
    else {
        const int nodesPerBlock = kInvWarps * 32;
        prec_invert_ela_kernel<<<blocks_for (nbNodes, nodesPerBlock), kInvWarps * 32, 0, stream>>> (
            prec, diagIndex, checkBounds, nbNodes);
    }
    return cudaGetLastError ();
}

cudaError_t launch_prec_inversion_list (int operatorID, double *prec, const int *diagIndex,
                                        const int *checkBounds, int nbNodes, const int *nodes,
                                        int count, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    prec_invert_list_kernel<<<blocks_for (count, 128), 128, 0, stream>>> (prec, diagIndex, checkBounds,
                                                                         nbNodes, nodes, count, operatorID);
    return cudaGetLastError ();
}

cudaError_t launch_halo_pack (double *sendBuf, const double *prec, const int *intfNodes, int dim,
                              int nbIntfNodes, cudaStream_t stream)
{
    const size_t total = (size_t)nbIntfNodes * dim;
    if (total == 0) return cudaSuccess;
    halo_pack_kernel<<<blocks_for (total, 256), 256, 0, stream>>> (sendBuf, prec, intfNodes, dim, total);
    return cudaGetLastError ();
}

cudaError_t launch_halo_add (double *prec, const double *recvBuf, const int *uniqNodes,
                             const int *slotIndex, const int *slots, int dim, int nbUniq,
                             cudaStream_t stream)
{
    const size_t total = (size_t)nbUniq * dim;
    if (total == 0) return cudaSuccess;
    halo_add_kernel<<<blocks_for (total, 256), 256, 0, stream>>> (prec, recvBuf, uniqNodes, slotIndex, slots, dim, total);
    return cudaGetLastError ();
}

// compute_double_norm (src/FEM.cc:48-56) on the device: sqrt of the sum of squares, summed in
// two levels whose order depends only on the (fixed) launch shape, so the result is
// reproducible run to run.
namespace {
constexpr int kNormBlocks = 592, kNormThreads = 256;

__global__ void __launch_bounds__(kNormThreads)
sum_squares_kernel (const double *__restrict__ x, int64_t n, double *partials)
{
    __shared__ double warpSums[kNormThreads / 32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kNormThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kNormThreads) {
        const double v = x[i];
        acc += v * v;
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync (0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = acc;
    __syncthreads ();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int w = 0; w < kNormThreads / 32; w++) total += warpSums[w];
        partials[blockIdx.x] = total;
    }
}

__global__ void finish_norm_kernel (const double *partials, double *out)
{
    if (threadIdx.x != 0) return;
    double total = 0.0;
    for (int b = 0; b < kNormBlocks; b++) total += partials[b];
    *out = sqrt (total);
}
}  // namespace

int double_norm_scratch_doubles () { return kNormBlocks; }

cudaError_t launch_double_norm (const double *x, int64_t n, double *partials, double *out, cudaStream_t stream)
{
    sum_squares_kernel<<<kNormBlocks, kNormThreads, 0, stream>>> (x, n, partials);
    cudaError_t e = cudaGetLastError ();
    if (e != cudaSuccess) return e;
    finish_norm_kernel<<<1, 32, 0, stream>>> (partials, out);
    return cudaGetLastError ();
}

}  // namespace mfb
