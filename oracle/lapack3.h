/* TEST INFRASTRUCTURE — restatement of the two LAPACK routines the reference calls
 * on every 3x3 preconditioner block (src/Fortran/elasclpr.f:39 DGETRF, :44 DGETRI).
 *
 * The reference links them from MKL (build/CMakeLists.txt:30 "-mkl:sequential",
 * version unpinned); MKL is not in this image, so the published reference-LAPACK
 * algorithm is restated: DGETF2 (unblocked right-looking LU, partial pivoting,
 * first-maximum IDAMAX, reciprocal scaling) and DGETRI's unblocked path (DTRTI2 for
 * inv(U), then inv(A)*L = inv(U) column by column, then the column interchanges).
 * PARITY UNPINNED against MKL itself; the survey measured a restatement of this kind
 * against scipy's LAPACK at 4.4e-16 block-wise.
 *
 * Column-major, leading dimension n, 0-based pivots.  Plain C; included by
 * oracle/minifem_oracle.c and oracle/shim/fortran_shim.cc only.
 */
#ifndef MINIFEM_ORACLE_LAPACK3_H
#define MINIFEM_ORACLE_LAPACK3_H

#include <math.h>
#include <float.h>

#define L3_A(i, j) a[(j) * n + (i)]

/* DGETF2.  Returns LAPACK's info (0, or k+1 for the first exactly-zero pivot). */
static int l3_getrf (int n, double *a, int *ipiv)
{
    int info = 0;
    const double sfmin = DBL_MIN;
    for (int j = 0; j < n; j++) {
        int p = j;
        double best = fabs (L3_A (j, j));
        for (int i = j + 1; i < n; i++) {
            if (fabs (L3_A (i, j)) > best) { best = fabs (L3_A (i, j)); p = i; }
        }
        ipiv[j] = p;
        if (L3_A (p, j) != 0.0) {
            if (p != j) {
                for (int c = 0; c < n; c++) {
                    double t = L3_A (j, c); L3_A (j, c) = L3_A (p, c); L3_A (p, c) = t;
                }
            }
            if (fabs (L3_A (j, j)) >= sfmin) {
                double r = 1.0 / L3_A (j, j);
                for (int i = j + 1; i < n; i++) L3_A (i, j) *= r;
            }
            else {
                for (int i = j + 1; i < n; i++) L3_A (i, j) /= L3_A (j, j);
            }
        }
        else if (info == 0) {
            info = j + 1;
        }
        for (int c = j + 1; c < n; c++) {
            for (int i = j + 1; i < n; i++) {
                L3_A (i, c) -= L3_A (i, j) * L3_A (j, c);
            }
        }
    }
    return info;
}

/* DGETRI (unblocked path; n <= 8).  a holds the factors from l3_getrf. */
static int l3_getri (int n, double *a, const int *ipiv)
{
    double work[8];
    /* DTRTI2, upper, non-unit: a <- inv(U) in place (singular U -> info, no work) */
    for (int j = 0; j < n; j++) if (L3_A (j, j) == 0.0) return j + 1;
    for (int j = 0; j < n; j++) {
        L3_A (j, j) = 1.0 / L3_A (j, j);
        double ajj = -L3_A (j, j);
        /* DTRMV upper, no-transpose, non-unit on the leading j x j block */
        for (int c = 0; c < j; c++) {
            if (L3_A (c, j) != 0.0) {
                double t = L3_A (c, j);
                for (int i = 0; i < c; i++) L3_A (i, j) += t * L3_A (i, c);
                L3_A (c, j) *= L3_A (c, c);
            }
        }
        for (int i = 0; i < j; i++) L3_A (i, j) *= ajj;
    }
    /* inv(A) * L = inv(U), last column first */
    for (int j = n - 1; j >= 0; j--) {
        for (int i = j + 1; i < n; i++) { work[i] = L3_A (i, j); L3_A (i, j) = 0.0; }
        for (int c = j + 1; c < n; c++) {          /* DGEMV: a(:,j) -= a(:,c) * work[c] */
            double t = -work[c];
            if (t != 0.0) for (int i = 0; i < n; i++) L3_A (i, j) += t * L3_A (i, c);
        }
    }
    /* undo the row interchanges as column interchanges */
    for (int j = n - 2; j >= 0; j--) {
        int p = ipiv[j];
        if (p != j) {
            for (int i = 0; i < n; i++) {
                double t = L3_A (i, j); L3_A (i, j) = L3_A (i, p); L3_A (i, p) = t;
            }
        }
    }
    return 0;
}

#undef L3_A
#endif
