// What HBM3e sustains for the traffic MIX of the assembly kernel: an EIB elasticity iteration reads 0.24 GB and
// writes 1.11 GB (values + prec), while MEASURED_PEAKS.json's 6548 GB/s is a copy (half reads, half writes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu
// Kernels: fill (stores only, 8 / 16 bytes per lane), copy (1 : 1), mix (1 byte read per `ratio` bytes written),
// and the write-out's own store shape (three 72-byte runs per warp instruction, 27 live lanes).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void fill8 (double *dst, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = 1.0; }
__global__ void fill16 (double2 *dst, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = make_double2 (1.0, 2.0); }
__global__ void copy16 (double2 *dst, const double2 *src, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i]; }
// every thread writes `ratio` double2 for each one it reads
__global__ void mix16 (double2 *dst, const double2 *src, size_t nRead, int ratio)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRead; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = src[i];
        for (int r = 0; r < ratio; r++) dst[(size_t)r * nRead + i] = v;
    }
}
// the write-out's store: a warp writes three runs of 72 bytes (rows of 15 entries: 1080 bytes each, consecutive rows)
__global__ void fill_rows72 (double *dst, size_t nRows, int entries)
{
    const int lane = threadIdx.x & 31, grp = lane / 10, comp = lane - 10 * grp;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    if (grp >= 3 || comp >= 9) return;
    for (size_t r0 = warp * 3; r0 < nRows; r0 += nWarps * 3) {
        const size_t r = r0 + grp;
        if (r >= nRows) continue;
        double *out = dst + r * entries * 9 + comp;
        for (int q = 0; q < entries; q++) out[q * 9] = 1.0 + q;
    }
}

// the coalesced write-out: a warp writes a row (entries x 9 consecutive doubles) 32 doubles per instruction, three rows
// one after the other
__global__ void fill_rows_coalesced (double *dst, size_t nRows, int entries)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int total = entries * 9;
    for (size_t r0 = warp * 3; r0 < nRows; r0 += nWarps * 3) {
        for (int g = 0; g < 3 && r0 + g < nRows; g++) {
            double *out = dst + (r0 + g) * total;
            for (int m = lane; m < total; m += 32) out[m] = 1.0 + m;
        }
    }
}

template <class F> float time_ms (F &&launch, int reps)
{
    cudaEvent_t a, b; cudaEventCreate (&a); cudaEventCreate (&b);
    for (int i = 0; i < 3; i++) launch ();
    std::vector<float> ms;
    for (int i = 0; i < reps; i++) { cudaEventRecord (a); launch (); cudaEventRecord (b); cudaEventSynchronize (b); float t; cudaEventElapsedTime (&t, a, b); ms.push_back (t); }
    std::sort (ms.begin (), ms.end ());
    return ms[ms.size () / 2];
}

int main ()
{
    const size_t writeBytes = 1113000000ull, readBytes = 239000000ull;     // ncu: dram bytes of one EIB ela iteration
    double *dst, *src;
    cudaMalloc (&dst, writeBytes + 4096); cudaMalloc (&src, writeBytes + 4096);
    cudaMemset (src, 0, writeBytes);
    const int grid = 148 * 8, block = 256;
    const size_t n8 = writeBytes / 8, n16 = writeBytes / 16;
    float t;
    t = time_ms ([&] { fill8<<<grid, block>>> (dst, n8); }, 20);
    printf ("fill, 8-byte stores:   %.4f ms  %.0f GB/s written\n", t, writeBytes / t / 1e6);
    t = time_ms ([&] { fill16<<<grid, block>>> ((double2*)dst, n16); }, 20);
    printf ("fill, 16-byte stores:  %.4f ms  %.0f GB/s written\n", t, writeBytes / t / 1e6);
    t = time_ms ([&] { cudaMemsetAsync (dst, 0, writeBytes); }, 20);
    printf ("cudaMemsetAsync:       %.4f ms  %.0f GB/s written\n", t, writeBytes / t / 1e6);
    t = time_ms ([&] { copy16<<<grid, block>>> ((double2*)dst, (const double2*)src, n16); }, 20);
    printf ("copy 1:1:              %.4f ms  %.0f GB/s read + written\n", t, 2.0 * writeBytes / t / 1e6);
    const size_t nRead = readBytes / 16;
    const int ratio = (int)(writeBytes / readBytes);       // 4
    t = time_ms ([&] { mix16<<<grid, block>>> ((double2*)dst, (const double2*)src, nRead, ratio); }, 20);
    printf ("mix 1 read : %d written: %.4f ms  %.0f GB/s read + written (%.3f GB)\n", ratio, t, (double)(nRead * 16 * (1 + ratio)) / t / 1e6, nRead * 16.0 * (1 + ratio) / 1e9);
    const size_t nRows = writeBytes / (15 * 72);
    t = time_ms ([&] { fill_rows72<<<148 * 4, 352>>> (dst, nRows, 15); }, 20);
    printf ("write-out store shape (3 x 72 B per warp instruction, rows of 15 entries): %.4f ms  %.0f GB/s written\n", t, (double)nRows * 15 * 72 / t / 1e6);
    t = time_ms ([&] { fill_rows72<<<148 * 16, 256>>> (dst, nRows, 15); }, 20);
    printf ("  the same with 148 x 16 CTAs of 256 threads: %.4f ms  %.0f GB/s written\n", t, (double)nRows * 15 * 72 / t / 1e6);
    t = time_ms ([&] { fill_rows_coalesced<<<148 * 4, 352>>> (dst, nRows, 15); }, 20);
    printf ("coalesced rows (32 consecutive doubles per warp instruction, 135 per row): %.4f ms  %.0f GB/s written\n", t, (double)nRows * 15 * 72 / t / 1e6);
    t = time_ms ([&] { fill_rows_coalesced<<<148 * 16, 256>>> (dst, nRows, 15); }, 20);
    printf ("  the same with 148 x 16 CTAs of 256 threads: %.4f ms  %.0f GB/s written\n", t, (double)nRows * 15 * 72 / t / 1e6);
    return 0;
}
