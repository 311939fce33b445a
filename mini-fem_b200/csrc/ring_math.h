// Arithmetic of the RING path (host/ring_plan.h), shared verbatim by the sm_100a kernel
// (csrc/kernels_ring.cu) and by the host replay the tests run (tools/ring_replay.cc).
#ifndef MFB_RING_MATH_H
#define MFB_RING_MATH_H

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MFB_RM __host__ __device__ __forceinline__
#else
#define MFB_RM inline
#endif

#ifndef MFB_RING_FAST_RCP
#define MFB_RING_FAST_RCP 1
#endif

namespace mfb {

// 1 / x for a normal, finite, non-zero x.  Device: MUFU.RCP64H seed (rcp.approx.ftz.f64: it looks at the high
// word of x only; measured on B200 over 2^20 random arguments: relative error <= 9.9e-7, 19.9 bits —
// tools/microbench/fp64_probe.cu) refined by one cubic step r (1 + e + e^2): the error is the cube of the
// seed's, 1e-18 before rounding (measured: 1.1e-16, the same as with a further Newton step, which nvcc's
// own 1.0 / x adds together with range checks and a slow path).  3 DFMA.  The host replay starts from a
// seed cut to 20 bits so that the tests exercise the same iteration.
MFB_RM double ring_rcp (double x)
{
#if MFB_RING_FAST_RCP
    double r;
#if defined(__CUDA_ARCH__)
    asm ("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
    r = 1.0 / x;
    uint64_t bits;
    memcpy (&bits, &r, 8);
    bits &= 0xFFFFFFFF00000000ull;      // sign, exponent, upper 20 mantissa bits
    memcpy (&r, &bits, 8);
#endif
    double e = fma (-x, r, 1.0);
    e = fma (e, e, e);
    return fma (r, e, r);
#else
    return 1.0 / x;
#endif
}

// One element (i, j, p, q) of the ring around the mesh edge (i, j).  d = x_j - x_i, u = x_p - x_i,
// w = x_q - x_i.  Face normals n_j = u x w (opposite j) and n_i = (u - d) x (w - d)
// = n_j + (w - u) x d (opposite i); det = n_j . d = n_i . d = 6 V (signed).  The gradients the
// reference computes in elem_coef_seq (src/assembly.cc:85-121) are grad_j = n_j / det and
// grad_i = -n_i / det, so the element adds grad_i grad_j^T = -(n_i n_j^T) / det^2 to A_ij;
// the order of p and q does not matter (both normals and det change sign).
// OPDIM 9: acc = the 3x3 sum A_ij (row component from node i); OPDIM 1: acc[0] = sum of
// grad_i . grad_j, the Laplacian entry (src/assembly.cc:539-541).
// live = false: a padding step of the regular loop; the contribution is discarded by selecting a zero
// weight (the operands are finite coordinates differences; a degenerate u, w only poisons r, which is dropped).
template <int OPDIM>
MFB_RM void ring_accumulate (const double d[3], const double u[3], const double w[3], double acc[OPDIM], bool live = true)
{
    const double njx = u[1] * w[2] - u[2] * w[1];
    const double njy = u[2] * w[0] - u[0] * w[2];
    const double njz = u[0] * w[1] - u[1] * w[0];
    const double ex = w[0] - u[0], ey = w[1] - u[1], ez = w[2] - u[2];
    // n_i = n_j + e x d, two fused multiply-adds per component
    const double nix = fma (-ez, d[1], fma (ey, d[2], njx));
    const double niy = fma (-ex, d[2], fma (ez, d[0], njy));
    const double niz = fma (-ey, d[0], fma (ex, d[1], njz));
    const double det = njx * d[0] + njy * d[1] + njz * d[2];
    const double rr = ring_rcp (-(det * det));          // -1 / det^2: the sign rides on the operand
    const double r = live ? rr : 0.0;
    if (OPDIM == 1) {
        acc[0] += r * (nix * njx + niy * njy + niz * njz);
    }
    else {
        const double mx = r * nix, my = r * niy, mz = r * niz;
        acc[0] += mx * njx; acc[1 % OPDIM] += mx * njy; acc[2 % OPDIM] += mx * njz;
        acc[3 % OPDIM] += my * njx; acc[4 % OPDIM] += my * njy; acc[5 % OPDIM] += my * njz;
        acc[6 % OPDIM] += mz * njx; acc[7 % OPDIM] += mz * njy; acc[8 % OPDIM] += mz * njz;
    }
}

// CSR block of the elasticity operator from A = sum grad_i grad_j^T (src/assembly.cc:386-409):
// K = 1.25 A + tr(A) I, row-major.
MFB_RM void ring_block (const double acc[9], double k[9])
{
    const double tr = acc[0] + acc[4] + acc[8];
    k[0] = 1.25 * acc[0] + tr; k[1] = 1.25 * acc[1]; k[2] = 1.25 * acc[2];
    k[3] = 1.25 * acc[3]; k[4] = 1.25 * acc[4] + tr; k[5] = 1.25 * acc[5];
    k[6] = 1.25 * acc[6]; k[7] = 1.25 * acc[7]; k[8] = 1.25 * acc[8] + tr;
}

// position of component k = 3a + b in the transposed block
MFB_RM int ring_transposed (int k) { return 3 * (k % 3) + k / 3; }

}  // namespace mfb

#endif
