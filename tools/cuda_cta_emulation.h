// TEST AID: runs a CUDA kernel's source on the host, one std::thread per CUDA thread of a CTA, so
// that the control flow, the indexing and — under -fsanitize=thread — the synchronisation
// protocol of csrc/kernels_ring.cu can be checked without a GPU (tools/ring_kernel_host.cc).
// What it emulates: threadIdx / blockIdx / blockDim / gridDim, dynamic shared memory,
// __syncthreads / __syncwarp (real barriers), warp shuffles (exchange through a per-warp buffer
// between two warp barriers), mbarriers with transaction counts, TMA bulk copies and cp.async.
// The asynchronous copies are performed AT ISSUE — the earliest moment the hardware could touch
// the destination — so a reader that has not been ordered before the issue shows up as a data
// race, and a reader that does not wait for the mbarrier / the block barrier is unordered with the
// copy and shows up as well.  What it cannot show: alignment rules, proxy fences, bank conflicts.
#ifndef MFB_CUDA_CTA_EMULATION_H
#define MFB_CUDA_CTA_EMULATION_H

#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>      // dim3, double2, make_double2; __global__ / __device__ expand to nothing under g++

namespace cta_emu {

struct MbarState { uint64_t *addr; unsigned expected, pending; uint64_t bytes; };
struct NamedBarrier { int id; pthread_barrier_t *barrier; };
struct Cta {
    int threads = 0;
    std::mutex mbarMutex;
    std::vector<MbarState> mbars;
    std::vector<NamedBarrier> named;
    pthread_barrier_t blockBarrier;
    std::vector<pthread_barrier_t> warpBarrier;
    std::vector<double> xchg;                 // 32 doubles per warp
    std::vector<unsigned char> smem;
};

struct Tls {
    uint3 threadIdx, blockIdx;
    dim3 blockDim, gridDim;
    Cta *cta = nullptr;
};
inline thread_local Tls tls;

inline void syncthreads () { pthread_barrier_wait (&tls.cta->blockBarrier); }
inline void syncwarp () { pthread_barrier_wait (&tls.cta->warpBarrier[tls.threadIdx.x >> 5]); }

inline double shfl (double v, int srcLane)
{
    Cta &c = *tls.cta;
    const int warp = (int)(tls.threadIdx.x >> 5), lane = (int)(tls.threadIdx.x & 31);
    double *x = c.xchg.data () + (size_t)warp * 32;
    x[lane] = v;
    pthread_barrier_wait (&c.warpBarrier[warp]);
    const double r = (srcLane >= 0 && srcLane < 32) ? x[srcLane] : v;
    pthread_barrier_wait (&c.warpBarrier[warp]);
    return r;
}

// mbarriers: the 8-byte object in shared memory only holds the number of completed phases; expected arrivals,
// pending arrivals and pending transaction bytes live in a side table of the CTA, under its mutex (which also
// gives ThreadSanitizer the happens-before edges of release / acquire).
inline MbarState &mbar_state (uint64_t *bar)
{
    for (MbarState &m : tls.cta->mbars) if (m.addr == bar) return m;
    fprintf (stderr, "cta_emu: mbarrier used before init\n"); abort ();
}
inline void mbar_complete_if_done (MbarState &m)
{
    if (m.pending == 0 && m.bytes == 0) { __atomic_store_n (m.addr, *m.addr + 1, __ATOMIC_RELEASE); m.pending = m.expected; }
}
inline void mbar_init (uint64_t *bar, unsigned count = 1)
{
    std::lock_guard<std::mutex> lock (tls.cta->mbarMutex);
    tls.cta->mbars.push_back ({bar, count, count, 0});
    __atomic_store_n (bar, 0ull, __ATOMIC_RELEASE);
}
inline void mbar_arrive (uint64_t *bar)
{
    std::lock_guard<std::mutex> lock (tls.cta->mbarMutex);
    MbarState &m = mbar_state (bar);
    if (m.pending == 0) { fprintf (stderr, "cta_emu: more arrivals than expected\n"); abort (); }
    m.pending--;
    mbar_complete_if_done (m);
}
inline void mbar_expect_tx (uint64_t *bar, unsigned bytes)      // arrive + expect `bytes`
{
    std::lock_guard<std::mutex> lock (tls.cta->mbarMutex);
    MbarState &m = mbar_state (bar);
    if (m.pending == 0) { fprintf (stderr, "cta_emu: more arrivals than expected\n"); abort (); }
    m.pending--;
    m.bytes += bytes;
    mbar_complete_if_done (m);
}
inline void bulk_load (void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15) || (bytes & 15)) { fprintf (stderr, "cta_emu: bulk copy not 16-byte aligned\n"); abort (); }
    memcpy (dst, src, bytes);
    std::lock_guard<std::mutex> lock (tls.cta->mbarMutex);
    MbarState &m = mbar_state (bar);
    if (m.bytes < bytes) { fprintf (stderr, "cta_emu: more bytes copied than expected\n"); abort (); }
    m.bytes -= bytes;
    mbar_complete_if_done (m);
}
// bulk copy shared -> global, performed at issue (the earliest moment the TMA unit could read the source)
inline void bulk_store (void *dst, const void *src, unsigned bytes)
{
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15) || (bytes & 15)) { fprintf (stderr, "cta_emu: bulk store not 16-byte aligned\n"); abort (); }
    memcpy (dst, src, bytes);
}
inline bool any_sync (bool pred)
{
    Cta &c = *tls.cta;
    const int warp = (int)(tls.threadIdx.x >> 5), lane = (int)(tls.threadIdx.x & 31);
    double *x = c.xchg.data () + (size_t)warp * 32;
    x[lane] = pred ? 1.0 : 0.0;
    pthread_barrier_wait (&c.warpBarrier[warp]);
    bool r = false;
    for (int l = 0; l < 32; l++) r = r || x[l] != 0.0;
    pthread_barrier_wait (&c.warpBarrier[warp]);
    return r;
}
inline void mbar_wait (uint64_t *bar, unsigned parity)          // returns once the phase of that parity has completed
{
    for (long spin = 0; spin < (1l << 26); spin++) {
        const uint64_t v = __atomic_load_n (bar, __ATOMIC_ACQUIRE);
        if (((uint32_t)v & 1u) != parity) return;
        std::this_thread::yield ();
    }
    fprintf (stderr, "cta_emu: mbarrier wait timed out (block %u thread %u)\n", tls.blockIdx.x, tls.threadIdx.x);
    abort ();
}
// named barrier (bar.sync id, count): one pthread barrier per (id, count), created on first use
inline void bar_sync (int id, int count)
{
    Cta &c = *tls.cta;
    pthread_barrier_t *b = nullptr;
    {
        std::lock_guard<std::mutex> lock (c.mbarMutex);
        for (auto &nb : c.named) if (nb.id == id) b = nb.barrier;
        if (!b) {
            b = new pthread_barrier_t;
            pthread_barrier_init (b, nullptr, (unsigned)count);
            c.named.push_back ({id, b});
        }
    }
    pthread_barrier_wait (b);
}

// Runs kernel() for every thread of every CTA of the grid; CTAs one after the other.
inline void launch (int grid, int threads, size_t smemBytes, const std::function<void ()> &kernel)
{
    for (int b = 0; b < grid; b++) {
        Cta cta;
        cta.threads = threads;
        pthread_barrier_init (&cta.blockBarrier, nullptr, (unsigned)threads);
        cta.warpBarrier.resize ((size_t)threads / 32);
        for (auto &wb : cta.warpBarrier) pthread_barrier_init (&wb, nullptr, 32);
        cta.xchg.assign ((size_t)threads, 0.0);
        cta.smem.assign (smemBytes + 128, 0xCD);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) {
            pool.emplace_back ([&, t] () {
                tls.threadIdx = {(unsigned)t, 0, 0}; tls.blockIdx = {(unsigned)b, 0, 0};
                tls.blockDim = dim3 ((unsigned)threads); tls.gridDim = dim3 ((unsigned)grid);
                tls.cta = &cta;
                kernel ();
            });
        }
        for (auto &th : pool) th.join ();
        pthread_barrier_destroy (&cta.blockBarrier);
        for (auto &wb : cta.warpBarrier) pthread_barrier_destroy (&wb);
    }
}

inline unsigned char *dynamic_smem ()          // 128-byte aligned like the kernel assumes
{
    return reinterpret_cast<unsigned char*> (((uintptr_t)tls.cta->smem.data () + 127) & ~(uintptr_t)127);
}

}  // namespace cta_emu

#define threadIdx (cta_emu::tls.threadIdx)
#define blockIdx (cta_emu::tls.blockIdx)
#define blockDim (cta_emu::tls.blockDim)
#define gridDim (cta_emu::tls.gridDim)
#define __launch_bounds__(...)
#define __syncthreads() cta_emu::syncthreads ()
#define __syncwarp() cta_emu::syncwarp ()
#define __shfl_xor_sync(mask, v, off) cta_emu::shfl ((v), (int)(threadIdx.x & 31) ^ (off))
#define __shfl_down_sync(mask, v, delta) cta_emu::shfl ((v), (int)(threadIdx.x & 31) + (delta))
#define __shfl_sync(mask, v, src) cta_emu::shfl ((v), (src) & 31)
#define __any_sync(mask, pred) cta_emu::any_sync (pred)
#define __ldg(p) (*(p))
#define __trap() abort ()
using std::max;
using std::min;

#endif
