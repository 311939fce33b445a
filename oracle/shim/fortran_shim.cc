// TEST INFRASTRUCTURE — C++ restatement of the three Fortran routines the reference
// links (no gfortran in this image), with the symbols and argument lists the C++
// side declares (main.cc:36-39, preconditioner.h:20-23).  Used only by the
// oracle/_ref build of the reference's own sources.
//
//   dqmrd4_          src/Fortran/qdmrd4.f:1-23
//   e_essbcm_        src/Fortran/e_cgmelissa.F:1-51 (codes: DefMesh_jlog.h:10-20)
//   ela_invert_prec_ src/Fortran/elasclpr.f:2-56    (DGETRF/DGETRI -> ../lapack3.h)
#include <cstdio>
#include "../lapack3.h"

enum { JLSYMX = 52, JLSYMY = 53, JLSYMZ = 54 };

extern "C" void dqmrd4_ (int *nnt, int *jlog, int *nnd, int *nlim, int *ier)
{
    int capacity = *nnd, found = 0;
    *ier = 0;
    for (int i = 1; i <= *nnt; i++) {
        if (jlog[i - 1] != 0) {
            found++;
            if (found <= capacity) nlim[found - 1] = i;
        }
    }
    *nnd = found;
    if (found > capacity) *ier = -1;
}

extern "C" void e_essbcm_ (int *ndim, int *nbp, int *nnd, int *nlim, int *jlog,
                           int *markf)
{
    const int N = *nbp, D = *ndim;
    for (int k = 0; k < N * D; k++) markf[k] = 0;
    // markf(node, comp) is column-major: markf[(comp-1)*N + (node-1)]
    for (int q = 0; q < *nnd; q++) {
        int node = nlim[q], code = jlog[node - 1];
        if (D == 2) {
            if (code == JLSYMX)      markf[0 * N + node - 1] = 1;
            else if (code == JLSYMZ) markf[1 * N + node - 1] = 1;
            else { markf[node - 1] = 1; markf[N + node - 1] = 1; }
        }
        else if (D == 3) {
            if (code == JLSYMX)      markf[0 * N + node - 1] = 1;
            else if (code == JLSYMY) markf[1 * N + node - 1] = 1;
            else if (code == JLSYMZ) markf[2 * N + node - 1] = 1;
            else {
                markf[node - 1] = 1; markf[N + node - 1] = 1; markf[2 * N + node - 1] = 1;
            }
        }
    }
}

extern "C" void ela_invert_prec_ (int *nspa, int *nnt, int *innv, int *nnv,
                                  double *prec, int *ier, int *markf, int *cur)
{
    const int D = *nspa, N = *nnt, i = *cur;
    // prec(ki,kj,i) column-major -> blk[(kj-1)*D + (ki-1)]
    double *blk = prec + (size_t)(i - 1) * D * D;
    for (int ki = 0; ki < D; ki++) {
        if (markf[ki * N + (i - 1)] != 0) {
            for (int kj = 0; kj < D; kj++) { blk[kj * D + ki] = 0.0; blk[ki * D + kj] = 0.0; }
            blk[ki * D + ki] = 1.0;
        }
    }
    bool hasDiag = false;
    for (int m = innv[i - 1] + 1; m <= innv[i]; m++) {
        if (nnv[m - 1] == i) { hasDiag = true; break; }
    }
    if (!hasDiag) return;
    double a[9];
    int ipiv[3];
    for (int k = 0; k < D * D; k++) a[k] = blk[k];
    int info = l3_getrf (D, a, ipiv);
    if (info != 0) { printf (" !!! info[DGETRF]= %d\n", info); *ier = info; }
    info = l3_getri (D, a, ipiv);
    if (info != 0) { printf (" !!! info[DGETRI]= %d\n", info); *ier = info; }
    for (int k = 0; k < D * D; k++) blk[k] = a[k];
}
