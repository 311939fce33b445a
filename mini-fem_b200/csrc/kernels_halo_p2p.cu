// Peer-to-peer halo exchange of the fused RING iteration: ONE small kernel next to the assembly kernel replaces
// pack -> ncclGroup{Recv, Send} -> add -> interface inversion (four launches and the NCCL proxy).
//
// The reference's interface sum (MPI_halo_exchange, src/halo.cc:39-122) packs prec of the interface nodes, exchanges
// the buffers with every neighbour and adds what arrives; its GASPI variant (src/halo.cc:127-226) writes straight
// into the neighbour's receive segment and notifies it.  This kernel is the latter over NVLink / NVSwitch peer memory
// (cudaIpc windows, or plain pointers when both subdomains live in one process):
//   1. wait until the assembly kernel, which runs at the same time on the rest of the device and takes the tiles
//      that own interface nodes first, has written the raw diagonal blocks of all of them (a counter the write-out
//      warps add to, ring_assembly_kernel);
//   2. every interface node's block is stored straight into the neighbour's receive window (halo.cc:77-80 and the
//      transfer in one step); system-scope fence; the last CTA to finish raises the epoch flag of this subdomain in
//      every neighbour's window (write + notify);
//   3. wait for the epoch flag of every neighbour;
//   4. one thread per distinct interface node adds the received blocks in increasing interface position — the order
//      of the reference's serial loop (halo.cc:113-116), so the sum is deterministic — masks and inverts the block
//      (prec_inversion, src/preconditioner.cc:25-49, src/Fortran/elasclpr.f) and writes prec.
// Receive buffers are double-buffered by epoch parity: a neighbour can be one iteration ahead, never two (its next
// exchange needs this subdomain's flag of that epoch).  Every wait is bounded; a timeout sets *status and returns.
#include "kernels.cuh"
#include "device_math.cuh"

namespace mfb {

namespace {

constexpr int kP2PThreads = 512;
constexpr long long kP2PSpinCycles = 6000000000ll;     // ~3 s at 1.9 GHz

__device__ __forceinline__ unsigned ld_acquire_gpu (const unsigned *p)
{
    unsigned v;
    asm volatile ("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys (const unsigned *p)
{
    unsigned v;
    asm volatile ("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys (unsigned *p, unsigned v)
{
    asm volatile ("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add_acq_rel_gpu (unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile ("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// true when `*word >= target` (wrap-safe) was seen before the spin budget ran out
template <bool SYSTEM>
__device__ __forceinline__ bool spin_until (const unsigned *word, unsigned target)
{
    const long long t0 = clock64 ();
    for (;;) {
        const unsigned v = SYSTEM ? ld_acquire_sys (word) : ld_acquire_gpu (word);
        if ((int)(v - target) >= 0) return true;
        if (clock64 () - t0 > kP2PSpinCycles) return false;
        __nanosleep (200);
    }
}

__device__ __forceinline__ void invert_interface_block (double b[9], int node, const HaloP2PArgs &a)
{
    int mx = 0, my = 0, mz = 0;
    if (a.checkBounds) {
        mx = a.checkBounds[node];
        my = a.checkBounds[(size_t)a.nbNodes + node];
        mz = a.checkBounds[2 * (size_t)a.nbNodes + node];
    }
    mask_block (b, mx, my, mz);
    if (a.diagIndex[node] >= 0) invert3_lu (b);          // elasclpr.f:29-32: only rows with a diagonal entry
}

template <int DIM>
__global__ void __launch_bounds__(kP2PThreads)
halo_p2p_kernel (const HaloP2PArgs a)
{
    __shared__ int sOk;
    HaloP2PState *st = a.state;
    const int tid = threadIdx.x;
    // the epoch of this exchange: read before anyone can advance it (the last CTA does, after every CTA has arrived
    // at `finished`, i.e. has read it)
    const unsigned epoch = ld_acquire_gpu (&st->epoch) + 1u;
    const unsigned parity = epoch & 1u;

    // ---- 1. the interface tiles of the assembly kernel are written -------------------------------------------------
    if (tid == 0) sOk = spin_until<false> (&st->intfDone, a.intfTarget) ? 1 : 0;
    __syncthreads ();
    if (!sOk && tid == 0) *(volatile unsigned*)a.status = 1u;

    // ---- 2. pack straight into the neighbours' windows --------------------------------------------------------------
    const size_t gtid = (size_t)blockIdx.x * kP2PThreads + tid, gsize = (size_t)gridDim.x * kP2PThreads;
    for (int i = 0; i < a.nbIntf; i++) {
        const int begin = a.intfIndex[i], end = a.intfIndex[i + 1];
        double *dst = a.peerRecv[2 * i + parity];
        const size_t total = (size_t)(end - begin) * DIM;
        for (size_t t = gtid; t < total; t += gsize) {
            const size_t j = t / DIM;
            const int k = (int)(t - j * DIM);
            dst[t] = __ldcg (a.prec + (size_t)(a.intfNodes[begin + j] - 1) * DIM + k);
        }
    }
    __threadfence_system ();
    __syncthreads ();
    if (tid == 0) {
        const unsigned old = atom_add_acq_rel_gpu (&st->packDone, 1u);
        if (old == gridDim.x - 1) {
            __threadfence_system ();
            for (int i = 0; i < a.nbIntf; i++) st_release_sys (a.peerFlag[i], epoch);      // write + notify
        }
    }

    // ---- 3. every neighbour's blocks have arrived -------------------------------------------------------------------
    if (tid == 0) sOk = 1;
    __syncthreads ();
    if (tid < a.nbIntf && !spin_until<true> (a.localFlags + tid, epoch)) { sOk = 0; *(volatile unsigned*)a.status = 2u; }
    __syncthreads ();

    // ---- 4. deterministic sum + inversion of the interface blocks --------------------------------------------------
    const double *recv = a.localRecv[parity];
    for (size_t u = gtid; u < (size_t)a.nbUniq; u += gsize) {
        const int node = a.uniqNodes[u];
        double b[DIM];
        double *blk = a.prec + (size_t)node * DIM;
        #pragma unroll
        for (int k = 0; k < DIM; k++) b[k] = __ldcg (blk + k);
        for (int s = a.slotIndex[u]; s < a.slotIndex[u + 1]; s++) {
            const double *src = recv + (size_t)a.slots[s] * DIM;
            #pragma unroll
            for (int k = 0; k < DIM; k++) b[k] += __ldcg (src + k);
        }
        if (DIM == 1) b[0] = 1.0 / b[0];
        else invert_interface_block (b, node, a);
        #pragma unroll
        for (int k = 0; k < DIM; k++) blk[k] = b[k];
    }

    // ---- the last CTA out re-arms the counters for the next iteration ---------------------------------------------
    __syncthreads ();
    if (tid == 0) {
        const unsigned old = atom_add_acq_rel_gpu (&st->finished, 1u);
        if (old == gridDim.x - 1) {
            st->intfDone = 0; st->packDone = 0; st->finished = 0;
            __threadfence ();
            st->epoch = epoch;
        }
    }
}

}  // namespace

cudaError_t launch_halo_p2p (const HaloP2PArgs &args, int operatorDim, int ctas, cudaStream_t stream)
{
    if (ctas < 1) ctas = 1;
    if (operatorDim == 1) halo_p2p_kernel<1><<<ctas, kP2PThreads, 0, stream>>> (args);
    else                  halo_p2p_kernel<9><<<ctas, kP2PThreads, 0, stream>>> (args);
    return cudaGetLastError ();
}

}  // namespace mfb
