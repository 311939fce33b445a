// Per-element and per-node arithmetic shared by every kernel of the path.
#ifndef MFB_DEVICE_MATH_CUH
#define MFB_DEVICE_MATH_CUH

#include <cmath>

// __host__ too: tools/ring_replay.cc (a test aid) runs mask_block / invert3_lu on the host
#if defined(__CUDACC__)
#define MFB_DM __host__ __device__ __forceinline__
#else
#define MFB_DM inline
#endif

namespace mfb {

// Gradient coefficients of one P1 tetrahedron — what elem_coef_seq computes
// (src/assembly.cc:85-121): edge vectors from node 3 to nodes 0, 2, 1, their cross
// products, row 3 = minus the sum of rows 0..2, all scaled by 1/vol with
// vol = a . row0 (no |vol|/6 weight; the sign of vol cancels in every product).
// p = 4 nodes x (x,y,z); c = 4 rows x 3.
MFB_DM void elem_coef (const double p[12], double c[12])
{
    const double xa = p[0] - p[9],  xb = p[6] - p[9],  xc = p[3] - p[9];
    const double ya = p[1] - p[10], yb = p[7] - p[10], yc = p[4] - p[10];
    const double za = p[2] - p[11], zb = p[8] - p[11], zc = p[5] - p[11];
    c[0] = yb * zc - yc * zb;  c[1] = zb * xc - zc * xb;  c[2] = xb * yc - xc * yb;
    c[3] = ya * zb - yb * za;  c[4] = za * xb - zb * xa;  c[5] = xa * yb - xb * ya;
    c[6] = yc * za - ya * zc;  c[7] = zc * xa - za * xc;  c[8] = xc * ya - xa * yc;
    c[9]  = -(c[0] + c[3] + c[6]);
    c[10] = -(c[1] + c[4] + c[7]);
    c[11] = -(c[2] + c[5] + c[8]);
    const double vol = xa * c[0] + ya * c[1] + za * c[2];
    const double inv = 1.0 / vol;
    #pragma unroll
    for (int k = 0; k < 12; k++) c[k] *= inv;
}

// One elasticity node-pair block (src/assembly.cc:386-409), row-major 3x3:
// K = (a.b) I + 1.25 a b^T, i.e. diagonal a_p b_p * 2.25 + the two other products.
MFB_DM void ela_block (const double a[3], const double b[3], double k[9])
{
    const double p00 = a[0] * b[0], p11 = a[1] * b[1], p22 = a[2] * b[2];
    k[0] = p00 * 2.25 + p11 + p22;
    k[1] = a[0] * b[1] * 1.25;
    k[2] = a[0] * b[2] * 1.25;
    k[3] = a[1] * b[0] * 1.25;
    k[4] = p00 + p11 * 2.25 + p22;
    k[5] = a[1] * b[2] * 1.25;
    k[6] = a[2] * b[0] * 1.25;
    k[7] = a[2] * b[1] * 1.25;
    k[8] = p00 + p11 + p22 * 2.25;
}

MFB_DM void swap2 (double &x, double &y) { double t = x; x = y; y = t; }

// What ela_invert_prec does to one node's 3x3 block (src/Fortran/elasclpr.f:19-53):
// Dirichlet components get their row and column zeroed and a unit diagonal, then the
// block is inverted by LU with partial pivoting (DGETRF) and DGETRI's back-substitution.
// The Fortran views the C block column-major (a(ki,kj) = blk[3*kj+ki]); the same view
// is kept so that pivoting picks the same rows.  Registers only: every index is static.
MFB_DM void mask_block (double b[9], int mx, int my, int mz)
{
    if (mx) { b[0] = 1.0; b[1] = b[2] = b[3] = b[6] = 0.0; }
    if (my) { b[4] = 1.0; b[1] = b[3] = b[5] = b[7] = 0.0; }
    if (mz) { b[8] = 1.0; b[2] = b[5] = b[6] = b[7] = 0.0; }
}

MFB_DM void invert3_lu (double b[9])
{
    // rows of the column-major view: r_i = (a(i,0), a(i,1), a(i,2)) = (b[i], b[3+i], b[6+i])
    double a00 = b[0], a01 = b[3], a02 = b[6];
    double a10 = b[1], a11 = b[4], a12 = b[7];
    double a20 = b[2], a21 = b[5], a22 = b[8];
    // column 0: first maximum of |a(i,0)|
    int p0 = 0;
    double best = fabs (a00);
    if (fabs (a10) > best) { best = fabs (a10); p0 = 1; }
    if (fabs (a20) > best) { p0 = 2; }
    if (p0 == 1) { swap2 (a00, a10); swap2 (a01, a11); swap2 (a02, a12); }
    if (p0 == 2) { swap2 (a00, a20); swap2 (a01, a21); swap2 (a02, a22); }
    double r = 1.0 / a00;
    a10 *= r; a20 *= r;
    a11 -= a10 * a01; a12 -= a10 * a02;
    a21 -= a20 * a01; a22 -= a20 * a02;
    // column 1
    int p1 = 1;
    if (fabs (a21) > fabs (a11)) p1 = 2;
    if (p1 == 2) { swap2 (a10, a20); swap2 (a11, a21); swap2 (a12, a22); }
    r = 1.0 / a11;
    a21 *= r;
    a22 -= a21 * a12;
    // inv(U) (DTRTI2, upper, non-unit)
    double u00 = 1.0 / a00;
    double u11 = 1.0 / a11;
    double u01 = -u11 * (a01 * u00);
    double u22 = 1.0 / a22;
    double t0 = a02, t1 = a12;          // column 2 above the diagonal, times inv(U)(0:2,0:2)
    t0 = u00 * t0 + u01 * t1;
    t1 = u11 * t1;
    double u02 = -u22 * t0, u12 = -u22 * t1;
    // inv(A) * L = inv(U): columns 2, 1, 0 (L unit lower: l10 = a10, l20 = a20, l21 = a21)
    double x02 = u02, x12 = u12, x22 = u22;
    double x01 = u01 - x02 * a21, x11 = u11 - x12 * a21, x21 = -x22 * a21;
    double x00 = u00 - x01 * a10 - x02 * a20;
    double x10 = -x11 * a10 - x12 * a20;
    double x20 = -x21 * a10 - x22 * a20;
    // column interchanges, j = 1 then j = 0
    if (p1 == 2) { swap2 (x01, x02); swap2 (x11, x12); swap2 (x21, x22); }
    if (p0 == 1) { swap2 (x00, x01); swap2 (x10, x11); swap2 (x20, x21); }
    if (p0 == 2) { swap2 (x00, x02); swap2 (x10, x12); swap2 (x20, x22); }
    b[0] = x00; b[3] = x01; b[6] = x02;
    b[1] = x10; b[4] = x11; b[7] = x12;
    b[2] = x20; b[5] = x21; b[8] = x22;
}

// The same block — masked like mask_block, inverted like ela_invert_prec — the way the RING kernel's write-out
// forms it: nine lanes hold one component each, so instead of an LU with row exchanges every lane builds the
// cofactor it needs from four other components (fetched by shuffles) and divides by the determinant (expansion
// along row 0; the reciprocal by `rcp`).  For the 3x3 diagonal blocks of the elasticity operator (symmetric
// positive definite, condition number of a few units) this agrees with LAPACK's result to a few ulp of the
// largest entry.  Returns false when the determinant is zero or not finite: the caller then runs invert3_lu,
// whose infinities and NaNs are the ones DGETRF / DGETRI produce.
// m = masked block, row-major; component c = 3a + b of the inverse is cofactor (b, a) / det.
MFB_DM int adj_src (int r, int s) { return 3 * (r % 3) + s % 3; }

template <class Rcp>
MFB_DM bool invert3_adj (const double m[9], double inv[9], Rcp rcp)
{
    double cof[9];
    for (int c = 0; c < 9; c++) {
        const int a = c / 3, b = c % 3;            // cofactor of row b, column a
        const double t = m[adj_src (b + 1, a + 2)] * m[adj_src (b + 2, a + 1)];
        cof[c] = fma (m[adj_src (b + 1, a + 1)], m[adj_src (b + 2, a + 2)], -t);
    }
    // cofactors of row 0 sit at c = 0, 3, 6 (b = 0, a = 0..2)
    const double det = (m[0] * cof[0] + m[1] * cof[3]) + m[2] * cof[6];
    if (!(fabs (det) > 0.0) || !(fabs (det) < 1.0e300)) return false;
    const double r = rcp (det);
    for (int c = 0; c < 9; c++) inv[c] = cof[c] * r;
    return true;
}

}  // namespace mfb

#endif
