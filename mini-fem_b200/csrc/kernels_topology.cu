// GPU construction of the layouts the assembly path consumes:
//
//   device_node_to_elem    DC_create_nodeToElem (call sites main.cc:247, coloring.cc:90)
//   device_build_csr       create_nodeToNode    (src/matrix.cc:55-91)
//   device_color_elements  coloring_creation    (src/coloring.cc:46-109)
//
// Same outputs as the reference's serial host code, bit for bit (column order is "first seen
// while sweeping the node's elements in increasing id", colours are greedy first-fit in element
// order), reached with parallel algorithms:
//   * node -> element lists: atomic counting sort, then every node sorts its short list;
//   * CSR: one thread per node collects its neighbours in first-seen order into a scratch row of
//     3*degree+1 slots (an element brings at most 3 new nodes), a prefix sum of the counts gives
//     nodeToNodeRow, a second kernel compacts the scratch rows into nodeToNodeColumn;
//   * colouring: an element can take its colour as soon as every lower-numbered element around its
//     4 nodes has one.  Node lists are sorted, so "ready" means the element is the next uncoloured
//     one in all 4 lists; ready elements never share a node, so a whole front is coloured in
//     parallel and the result is the sequential first-fit's.  One persistent CTA walks the fronts
//     (a 100^3 Kuhn mesh has ~2800 of them, ~2000 elements wide).
#include "kernels.cuh"

#include <cstdio>

namespace mfb {

namespace {

constexpr int kScanThreads = 1024, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

// Exclusive prefix sum of one tile per block; blockSums[b] = sum of the tile.
__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel (const int *in, int *out, int *blockSums, int64_t n)
{
    __shared__ int warpSums[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)tid * kScanItems;
    int v[kScanItems], total = 0;
    #pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = base + i < n ? in[base + i] : 0;
        total += v[i];
    }
    int incl = total;
    #pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync (0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads ();
    if (warp == 0) {
        int w = warpSums[lane];
        #pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync (0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        warpSums[lane] = w;                       // inclusive over warps
    }
    __syncthreads ();
    int run = (warp ? warpSums[warp - 1] : 0) + incl - total;
    #pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (tid == kScanThreads - 1 && blockSums) blockSums[blockIdx.x] = warpSums[31];
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel (int *out, const int *blockOffsets, int64_t n)
{
    const int add = blockOffsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    #pragma unroll
    for (int i = 0; i < kScanItems; i++) if (base + i < n) out[base + i] += add;
}

// In-place exclusive scan of n ints (recursive over the block sums).
cudaError_t exclusive_scan (int *data, int64_t n, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int64_t blocks = (n + kScanTile - 1) / kScanTile;
    if (blocks == 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, stream>>> (data, data, nullptr, n);
        return cudaGetLastError ();
    }
    int *sums = nullptr;
    cudaError_t e = cudaMalloc (&sums, sizeof (int) * (size_t)blocks);
    if (e != cudaSuccess) return e;
    scan_tile_kernel<<<(unsigned)blocks, kScanThreads, 0, stream>>> (data, data, sums, n);
    e = cudaGetLastError ();
    if (e == cudaSuccess) e = exclusive_scan (sums, blocks, stream);
    if (e == cudaSuccess) {
        scan_add_kernel<<<(unsigned)blocks, kScanThreads, 0, stream>>> (data, sums, n);
        e = cudaGetLastError ();
    }
    cudaError_t e2 = cudaStreamSynchronize (stream);
    cudaFree (sums);
    return e != cudaSuccess ? e : e2;
}

// ---------------------------------------------------------------- node -> elements

// flags[0] = number of node ids outside [1, nbNodes]
__global__ void incidence_count_kernel (const int *__restrict__ elemToNode, int64_t nbInc, int nbNodes,
                                        int *count, int *flags)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbInc) return;
    const int node = elemToNode[k] - 1;
    if (node < 0 || node >= nbNodes) { atomicAdd (flags, 1); return; }
    atomicAdd (count + node, 1);
}

__global__ void incidence_fill_kernel (const int *__restrict__ elemToNode, int64_t nbInc,
                                       const int *__restrict__ index, int *cursor, int *value)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbInc) return;
    const int node = elemToNode[k] - 1;
    value[index[node] + atomicAdd (cursor + node, 1)] = (int)(k >> 2);
}

// Each node's list in increasing element id (the atomics fill it in arrival order).
__global__ void sort_lists_kernel (const int *__restrict__ index, int *value, int nbNodes)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nbNodes) return;
    const int begin = index[n], end = index[n + 1];
    for (int i = begin + 1; i < end; i++) {
        const int v = value[i];
        int j = i - 1;
        while (j >= begin && value[j] > v) { value[j + 1] = value[j]; j--; }
        value[j + 1] = v;
    }
}

// ---------------------------------------------------------------- CSR

// scratch row of node i starts at 3*index[i] + i and holds 3*degree + 1 slots
__global__ void neighbour_sweep_kernel (const int *__restrict__ elemToNode, const int *__restrict__ index,
                                        const int *__restrict__ value, int nbNodes, int *scratch, int *count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbNodes) return;
    const int begin = index[i], end = index[i + 1];
    int *mine = scratch + 3 * (int64_t)begin + i;
    int seen = 0;
    for (int p = begin; p < end; p++) {
        const int4 nodes = *reinterpret_cast<const int4*> (elemToNode + (size_t)value[p] * 4);
        const int cand[4] = {nodes.x, nodes.y, nodes.z, nodes.w};
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            bool known = false;
            for (int s = 0; s < seen; s++) known |= mine[s] == cand[k];
            if (!known) mine[seen++] = cand[k];
        }
    }
    count[i] = seen;
}

__global__ void compact_rows_kernel (const int *__restrict__ index, const int *__restrict__ row,
                                     const int *__restrict__ scratch, int nbNodes, int *col)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbNodes) return;
    const int *mine = scratch + 3 * (int64_t)index[i] + i;
    const int begin = row[i], n = row[i + 1] - begin;
    for (int s = 0; s < n; s++) col[begin + s] = mine[s];
}

// ---------------------------------------------------------------- colouring

struct ColorState {
    const int *elemToNode, *index, *value;
    int *ptr;             // per node: how many leading elements of its list are coloured
    unsigned *usedAt;     // per node: 4 words = the 128 colour bits taken around it
    int *readyCnt;        // per element: in how many of its 4 node lists it is the next one
    int *colorPart;
    int *queue[2];        // fronts
    int *counters;        // [0], [1] = sizes of the two queues; [2] = error (no colour left);
                          // [3] = highest colour; [4] = elements coloured
};

// An element whose node lists hold it twice (repeated node id) would never become ready:
// flags[1] counts them.
__global__ void color_seed_kernel (ColorState s, int nbNodes)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nbNodes) return;
    const int begin = s.index[n];
    if (begin == s.index[n + 1]) return;
    const int c = s.value[begin];
    if (atomicAdd (s.readyCnt + c, 1) == 3) s.queue[0][atomicAdd (s.counters, 1)] = c;
}

__global__ void __launch_bounds__(1024)
color_fronts_kernel (ColorState s)
{
    __shared__ int frontSize;
    const int lane = threadIdx.x & 31;
    int cur = 0, highest = 0;
    for (;;) {
        if (threadIdx.x == 0) { frontSize = *(volatile int*)(s.counters + cur); s.counters[cur ^ 1] = 0; }
        __syncthreads ();
        const int n = frontSize;
        if (n == 0) break;
        int *front = s.queue[cur], *next = s.queue[cur ^ 1];
        for (int w0 = 0; w0 < n; w0 += blockDim.x) {            // whole warps stay in the loop (ballots below)
            const int w = w0 + threadIdx.x;
            bool ready[4] = {false, false, false, false};
            int follower[4] = {-1, -1, -1, -1};
            if (w < n) {
                const int e = __ldcg (front + w);
                const int4 nd = *reinterpret_cast<const int4*> (s.elemToNode + (size_t)e * 4);
                const int nodes[4] = {nd.x - 1, nd.y - 1, nd.z - 1, nd.w - 1};
                // everything this element needs from its 4 nodes, requested together (L2 loads: the
                // previous front wrote these words from other threads)
                uint4 used[4];
                int done[4], begin[4], end[4];
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    used[k] = __ldcg (reinterpret_cast<const uint4*> (s.usedAt) + nodes[k]);
                    done[k] = __ldcg (s.ptr + nodes[k]) + 1;
                    begin[k] = s.index[nodes[k]];
                    end[k] = s.index[nodes[k] + 1];
                }
                const unsigned taken[4] = {used[0].x | used[1].x | used[2].x | used[3].x, used[0].y | used[1].y | used[2].y | used[3].y,
                                           used[0].z | used[1].z | used[2].z | used[3].z, used[0].w | used[1].w | used[2].w | used[3].w};
                int color = -1;
                #pragma unroll
                for (int q = 3; q >= 0; q--) if (~taken[q]) color = 32 * q + __ffs (~taken[q]) - 1;
                if (color < 0) { atomicAdd (s.counters + 2, 1); color = 0; }
                s.colorPart[e] = color;
                if (color > highest) highest = color;
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    // nobody else touches this node in this front
                    unsigned *word = s.usedAt + (size_t)nodes[k] * 4 + (color >> 5);
                    const unsigned old = (color >> 5) == 0 ? used[k].x : (color >> 5) == 1 ? used[k].y : (color >> 5) == 2 ? used[k].z : used[k].w;
                    __stcg (word, old | (1u << (color & 31)));
                    __stcg (s.ptr + nodes[k], done[k]);
                    follower[k] = begin[k] + done[k] < end[k] ? s.value[begin[k] + done[k]] : -1;
                }
                #pragma unroll
                for (int k = 0; k < 4; k++) ready[k] = follower[k] >= 0 && atomicAdd (s.readyCnt + follower[k], 1) == 3;
            }
            // one queue reservation per warp
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned pushers = __ballot_sync (0xffffffffu, ready[k]);
                if (!pushers) continue;
                int base = 0;
                if (lane == __ffs (pushers) - 1) base = atomicAdd (s.counters + (cur ^ 1), __popc (pushers));
                base = __shfl_sync (0xffffffffu, base, __ffs (pushers) - 1);
                if (ready[k]) __stcg (next + base + __popc (pushers & ((1u << lane) - 1)), follower[k]);
            }
        }
        if (threadIdx.x == 0) s.counters[4] += n;
        __syncthreads ();
        cur ^= 1;
    }
    atomicMax (s.counters + 3, highest);
}

// colorPerm = stable counting sort by colour (DC_create_permutation as used at coloring.cc:107):
// histogram per chunk of elements, scan over (colour, chunk), then ranks inside each chunk.
constexpr int kPermChunk = 1024;

__global__ void __launch_bounds__(256)
perm_histogram_kernel (const int *__restrict__ colorPart, int nbElem, int nbChunks, int *hist /* [128][nbChunks] */)
{
    __shared__ int local[128];
    const int chunk = blockIdx.x;
    if (threadIdx.x < 128) local[threadIdx.x] = 0;
    __syncthreads ();
    for (int i = threadIdx.x; i < kPermChunk; i += blockDim.x) {
        const int e = chunk * kPermChunk + i;
        if (e < nbElem) atomicAdd (local + colorPart[e], 1);
    }
    __syncthreads ();
    if (threadIdx.x < 128) hist[(size_t)threadIdx.x * nbChunks + chunk] = local[threadIdx.x];
}

// One warp per chunk walks it 32 elements at a time, in order: rank = earlier equal colours.
__global__ void __launch_bounds__(32)
perm_rank_kernel (const int *__restrict__ colorPart, int nbElem, int nbChunks, const int *__restrict__ hist,
                  int *colorPerm)
{
    __shared__ int base[128];
    const int chunk = blockIdx.x, lane = threadIdx.x;
    for (int c = lane; c < 128; c += 32) base[c] = hist[(size_t)c * nbChunks + chunk];
    __syncwarp ();
    for (int i = 0; i < kPermChunk; i += 32) {
        const int e = chunk * kPermChunk + i + lane;
        const bool live = e < nbElem;
        const int color = live ? colorPart[e] : -1 - lane;            // idle lanes match nobody
        const unsigned same = __match_any_sync (0xffffffffu, color);
        const int before = __popc (same & ((1u << lane) - 1));
        if (live) colorPerm[e] = base[color] + before;
        __syncwarp ();
        if (live && before == __popc (same) - 1) base[color] += before + 1;   // last lane of the group
        __syncwarp ();
    }
}

__global__ void color_offsets_kernel (const int *__restrict__ hist, int nbChunks, int nbColors, int *colorToElem)
{
    const int c = threadIdx.x;
    if (c < nbColors) colorToElem[c] = hist[(size_t)c * nbChunks];
}

struct DeviceBuffer {
    void *p = nullptr;
    ~DeviceBuffer () { if (p) cudaFree (p); }
    cudaError_t alloc (size_t bytes) { return cudaMalloc (&p, bytes ? bytes : 1); }
    template <class T> T *as () { return static_cast<T*> (p); }
};

#define TOPO_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

unsigned blocks_for (int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

cudaError_t device_node_to_elem (const int *dElemToNode, int nbElem, int nbNodes, int *dIndex, int *dValue,
                                 int *badIds, cudaStream_t stream)
{
    *badIds = 0;
    const int64_t nbInc = (int64_t)nbElem * 4;
    DeviceBuffer cursor, flags;
    TOPO_TRY (cursor.alloc (sizeof (int) * (size_t)(nbNodes + 1)));
    TOPO_TRY (flags.alloc (sizeof (int)));
    TOPO_TRY (cudaMemsetAsync (dIndex, 0, sizeof (int) * (size_t)(nbNodes + 1), stream));
    TOPO_TRY (cudaMemsetAsync (cursor.p, 0, sizeof (int) * (size_t)(nbNodes + 1), stream));
    TOPO_TRY (cudaMemsetAsync (flags.p, 0, sizeof (int), stream));
    if (nbInc > 0) {
        incidence_count_kernel<<<blocks_for (nbInc, 256), 256, 0, stream>>> (dElemToNode, nbInc, nbNodes, dIndex, flags.as<int> ());
        TOPO_TRY (cudaGetLastError ());
    }
    TOPO_TRY (cudaMemcpyAsync (badIds, flags.p, sizeof (int), cudaMemcpyDeviceToHost, stream));
    TOPO_TRY (cudaStreamSynchronize (stream));
    if (*badIds) return cudaSuccess;
    TOPO_TRY (exclusive_scan (dIndex, (int64_t)nbNodes + 1, stream));
    if (nbInc > 0) {
        incidence_fill_kernel<<<blocks_for (nbInc, 256), 256, 0, stream>>> (dElemToNode, nbInc, dIndex, cursor.as<int> (), dValue);
        TOPO_TRY (cudaGetLastError ());
        sort_lists_kernel<<<blocks_for (nbNodes, 128), 128, 0, stream>>> (dIndex, dValue, nbNodes);
        TOPO_TRY (cudaGetLastError ());
    }
    return cudaStreamSynchronize (stream);
}

cudaError_t device_build_csr (const int *dElemToNode, const int *dIndex, const int *dValue, int nbElem,
                              int nbNodes, int *dRow, int **dColOut, int64_t *nbEdges, cudaStream_t stream)
{
    *dColOut = nullptr;
    *nbEdges = 0;
    DeviceBuffer scratch;
    TOPO_TRY (scratch.alloc (sizeof (int) * ((size_t)nbElem * 12 + (size_t)nbNodes)));
    TOPO_TRY (cudaMemsetAsync (dRow, 0, sizeof (int) * (size_t)(nbNodes + 1), stream));
    if (nbNodes > 0) {
        neighbour_sweep_kernel<<<blocks_for (nbNodes, 128), 128, 0, stream>>> (dElemToNode, dIndex, dValue, nbNodes,
                                                                                 scratch.as<int> (), dRow);
        TOPO_TRY (cudaGetLastError ());
    }
    // the counts are bounded by 3*4*nbElem + nbNodes; the reference keeps them in an int (IO.cc:77)
    TOPO_TRY (exclusive_scan (dRow, (int64_t)nbNodes + 1, stream));
    int total = 0;
    TOPO_TRY (cudaMemcpyAsync (&total, dRow + nbNodes, sizeof (int), cudaMemcpyDeviceToHost, stream));
    TOPO_TRY (cudaStreamSynchronize (stream));
    if (total < 0) return cudaErrorInvalidValue;
    int *dCol = nullptr;
    TOPO_TRY (cudaMalloc (&dCol, sizeof (int) * (size_t)(total ? total : 1)));
    if (nbNodes > 0) {
        compact_rows_kernel<<<blocks_for (nbNodes, 128), 128, 0, stream>>> (dIndex, dRow, scratch.as<int> (), nbNodes, dCol);
        cudaError_t e = cudaGetLastError ();
        if (e == cudaSuccess) e = cudaStreamSynchronize (stream);
        if (e != cudaSuccess) { cudaFree (dCol); return e; }
    }
    *dColOut = dCol;
    *nbEdges = total;
    return cudaSuccess;
}

cudaError_t device_color_elements (const int *dElemToNode, const int *dIndex, const int *dValue, int nbElem,
                                   int nbNodes, int *dColorPart, int *dColorPerm, int *colorToElem /* host, 129 */,
                                   int *nbColors, cudaStream_t stream)
{
    *nbColors = 0;
    for (int c = 0; c <= 128; c++) colorToElem[c] = 0;
    if (nbElem == 0) return cudaSuccess;
    DeviceBuffer ptr, usedAt, readyCnt, q0, q1, counters, hist, dOffsets;
    TOPO_TRY (ptr.alloc (sizeof (int) * (size_t)nbNodes));
    TOPO_TRY (usedAt.alloc (sizeof (unsigned) * 4 * (size_t)nbNodes));
    TOPO_TRY (readyCnt.alloc (sizeof (int) * (size_t)nbElem));
    TOPO_TRY (q0.alloc (sizeof (int) * (size_t)nbElem));
    TOPO_TRY (q1.alloc (sizeof (int) * (size_t)nbElem));
    TOPO_TRY (counters.alloc (sizeof (int) * 8));
    TOPO_TRY (cudaMemsetAsync (ptr.p, 0, sizeof (int) * (size_t)nbNodes, stream));
    TOPO_TRY (cudaMemsetAsync (usedAt.p, 0, sizeof (unsigned) * 4 * (size_t)nbNodes, stream));
    TOPO_TRY (cudaMemsetAsync (readyCnt.p, 0, sizeof (int) * (size_t)nbElem, stream));
    TOPO_TRY (cudaMemsetAsync (counters.p, 0, sizeof (int) * 8, stream));
    ColorState s;
    s.elemToNode = dElemToNode; s.index = dIndex; s.value = dValue;
    s.ptr = ptr.as<int> (); s.usedAt = usedAt.as<unsigned> (); s.readyCnt = readyCnt.as<int> ();
    s.colorPart = dColorPart; s.queue[0] = q0.as<int> (); s.queue[1] = q1.as<int> (); s.counters = counters.as<int> ();
    color_seed_kernel<<<blocks_for (nbNodes, 256), 256, 0, stream>>> (s, nbNodes);
    TOPO_TRY (cudaGetLastError ());
    color_fronts_kernel<<<1, 1024, 0, stream>>> (s);
    TOPO_TRY (cudaGetLastError ());
    int host[8];
    TOPO_TRY (cudaMemcpyAsync (host, counters.p, sizeof (host), cudaMemcpyDeviceToHost, stream));
    TOPO_TRY (cudaStreamSynchronize (stream));
    if (host[2] > 0) { *nbColors = -1; return cudaSuccess; }          // more than 128 colours (coloring.cc:66-69)
    if (host[4] != nbElem) { *nbColors = -2; return cudaSuccess; }    // an element lists a node twice
    const int colors = host[3] + 1;

    const int nbChunks = (nbElem + kPermChunk - 1) / kPermChunk;
    TOPO_TRY (hist.alloc (sizeof (int) * 128 * (size_t)nbChunks + sizeof (int)));
    perm_histogram_kernel<<<nbChunks, 256, 0, stream>>> (dColorPart, nbElem, nbChunks, hist.as<int> ());
    TOPO_TRY (cudaGetLastError ());
    TOPO_TRY (cudaMemsetAsync (hist.as<int> () + 128 * (size_t)nbChunks, 0, sizeof (int), stream));
    TOPO_TRY (exclusive_scan (hist.as<int> (), 128 * (int64_t)nbChunks + 1, stream));
    perm_rank_kernel<<<nbChunks, 32, 0, stream>>> (dColorPart, nbElem, nbChunks, hist.as<int> (), dColorPerm);
    TOPO_TRY (cudaGetLastError ());
    TOPO_TRY (dOffsets.alloc (sizeof (int) * 129));
    color_offsets_kernel<<<1, 128, 0, stream>>> (hist.as<int> (), nbChunks, colors, dOffsets.as<int> ());
    TOPO_TRY (cudaGetLastError ());
    TOPO_TRY (cudaMemcpyAsync (colorToElem, dOffsets.p, sizeof (int) * (size_t)colors, cudaMemcpyDeviceToHost, stream));
    TOPO_TRY (cudaStreamSynchronize (stream));
    colorToElem[colors] = nbElem;
    *nbColors = colors;
    return cudaSuccess;
}

}  // namespace mfb
