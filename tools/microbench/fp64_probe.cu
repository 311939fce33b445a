// Microbenchmarks behind the RING kernel's cost model (B200): FP64 issue rate and latency, MUFU.RCP64H seed
// accuracy, shared-memory load latency.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_tp (double *out, int iters, double a, double b)
{
    double x[ILP];
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; i++) {
        #pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = fma (x[k], a, b);
    }
    double s = 0;
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_lat (double *out, long long *cycles, int iters, double a, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64 ();
    for (int i = 0; i < iters; i++) x = fma (x, a, b);
    long long t1 = clock64 ();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void lds_lat (double *out, long long *cycles, int iters)
{
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i * 17 + 5) & 1023;
    __syncthreads ();
    int p = threadIdx.x;
    long long t0 = clock64 ();
    for (int i = 0; i < iters; i++) p = next[p];
    long long t1 = clock64 ();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void rcp_seed (const double *x, double *seed, double *cubic, double *full, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i]));
    seed[i] = r;
    double e = fma (-x[i], r, 1.0);
    e = fma (e, e, e);
    double r3 = fma (r, e, r);
    cubic[i] = r3;
    e = fma (-x[i], r3, 1.0);
    full[i] = fma (r3, e, r3);
}

// FP64 chain interleaved with integer / FP32 work: does the FP64 pipe co-issue with the rest?
__global__ void mix_tp (double *out, int iters, double a, double b)
{
    double x[4];
    int y[4];
    for (int k = 0; k < 4; k++) { x[k] = threadIdx.x * 1e-3 + k; y[k] = threadIdx.x + k; }
    for (int i = 0; i < iters; i++) {
        #pragma unroll
        for (int k = 0; k < 4; k++) { x[k] = fma (x[k], a, b); y[k] = y[k] * 3 + i; }
    }
    double s = 0;
    for (int k = 0; k < 4; k++) s += x[k] + y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main ()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties (&prop, 0);
    int clk = 0;
    cudaDeviceGetAttribute (&clk, cudaDevAttrClockRate, 0);
    printf ("%s: %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, clk);
    double *out; long long *cyc;
    cudaMalloc (&out, sizeof (double) * 148 * 2048 * 4);
    cudaMallocManaged (&cyc, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0); cudaEventCreate (&e1);
    const int iters = 20000;
    auto time_tp = [&] (auto kernel, int ilp, int threads, int blocksPerSM, const char *name, double flopsPerIter) {
        kernel<<<prop.multiProcessorCount * blocksPerSM, threads>>> (out, 100, 1.0000001, 1e-9);
        cudaDeviceSynchronize ();
        cudaEventRecord (e0);
        kernel<<<prop.multiProcessorCount * blocksPerSM, threads>>> (out, iters, 1.0000001, 1e-9);
        cudaEventRecord (e1);
        cudaEventSynchronize (e1);
        float ms; cudaEventElapsedTime (&ms, e0, e1);
        const double fma = (double)prop.multiProcessorCount * blocksPerSM * threads * iters * flopsPerIter;
        printf ("%-28s threads/SM %4d ilp %d: %.3f ms  %.2f T DFMA/s = %.1f TFLOP/s; per SM per clock (at %.0f MHz): %.1f DFMA\n", name,
                threads * blocksPerSM, ilp, ms, fma / ms / 1e9, 2 * fma / ms / 1e9, clk / 1e3, fma / (ms * 1e-3) / prop.multiProcessorCount / (clk * 1e3));
    };
    time_tp (dfma_tp<8>, 8, 256, 4, "DFMA throughput", 8);
    time_tp (dfma_tp<8>, 8, 256, 3, "DFMA throughput", 8);
    time_tp (dfma_tp<4>, 4, 256, 3, "DFMA throughput", 4);
    time_tp (dfma_tp<2>, 2, 256, 3, "DFMA throughput", 2);
    time_tp (dfma_tp<1>, 1, 256, 3, "DFMA throughput", 1);
    time_tp (dfma_tp<1>, 1, 256, 8, "DFMA throughput", 1);
    time_tp (mix_tp, 4, 256, 3, "DFMA + IMAD mix", 4);
    dfma_lat<<<1, 32>>> (out, cyc, 10000, 1.0000001, 1e-9);
    cudaDeviceSynchronize ();
    printf ("DFMA dependent latency: %.2f cycles\n", *cyc / 10000.0);
    lds_lat<<<1, 32>>> (out, cyc, 10000);
    cudaDeviceSynchronize ();
    printf ("LDS dependent latency (pointer chase incl. address calc): %.2f cycles\n", *cyc / 10000.0);
    // reciprocal seed accuracy
    const int n = 1 << 20;
    std::vector<double> hx (n), hs (n), hc (n), hf (n);
    uint64_t state = 88172645463325252ull;
    for (int i = 0; i < n; i++) {
        state ^= state << 13; state ^= state >> 7; state ^= state << 17;
        const double u = (state >> 11) * (1.0 / 9007199254740992.0);
        hx[i] = ldexp (1.0 + u, (int)(state % 120) - 60) * ((state >> 3) & 1 ? 1 : -1);
    }
    double *dx, *ds, *dc, *df;
    cudaMalloc (&dx, 8 * n); cudaMalloc (&ds, 8 * n); cudaMalloc (&dc, 8 * n); cudaMalloc (&df, 8 * n);
    cudaMemcpy (dx, hx.data (), 8 * n, cudaMemcpyHostToDevice);
    rcp_seed<<<n / 256, 256>>> (dx, ds, dc, df, n);
    cudaMemcpy (hs.data (), ds, 8 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy (hc.data (), dc, 8 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy (hf.data (), df, 8 * n, cudaMemcpyDeviceToHost);
    double es = 0, ec = 0, ef = 0;
    for (int i = 0; i < n; i++) {
        const long double t = 1.0L / hx[i];
        es = fmax (es, (double)fabsl ((hs[i] - t) / t)); ec = fmax (ec, (double)fabsl ((hc[i] - t) / t)); ef = fmax (ef, (double)fabsl ((hf[i] - t) / t));
    }
    printf ("rcp.approx.ftz.f64 max relative error: seed %.3e (%.1f bits), after the cubic step %.3e, after cubic + Newton %.3e\n", es, -log2 (es), ec, ef);
    return 0;
}
