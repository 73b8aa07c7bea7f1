/*
 * mkl_shim.c -- Gustavson CSR x CSR product behind the mkl_dcsrmultcsr name
 * (TEST INFRASTRUCTURE; see mkl_spblas.h).  One-based indices in and out,
 * request 0 (compute everything), output rows left in first-touch order
 * (the reference passes sort=7, "no sorting").
 */
#include "mkl_spblas.h"
#include <stdlib.h>
#include <string.h>

void mkl_dcsrmultcsr(const char *trans, const MKL_INT *request, const MKL_INT *sort, const MKL_INT *m,
                     const MKL_INT *n, const MKL_INT *k, double *a, MKL_INT *ja, MKL_INT *ia, double *b,
                     MKL_INT *jb, MKL_INT *ib, double *c, MKL_INT *jc, MKL_INT *ic, const MKL_INT *nzmax,
                     MKL_INT *info) {
    (void)trans; (void)request; (void)sort; (void)n;
    const int M = *m, K = *k;
    int *slot = (int *)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1)); /* column -> position in c, or -1 */
    memset(slot, 0xff, sizeof(int) * (size_t)(K > 0 ? K : 1));
    long nz = 0;
    *info = 0;
    ic[0] = 1;
    for (int i = 0; i < M; ++i) {
        const long row_start = nz;
        for (int p = ia[i] - 1; p < ia[i + 1] - 1; ++p) {
            const int r = ja[p] - 1;
            const double av = a[p];
            for (int q = ib[r] - 1; q < ib[r + 1] - 1; ++q) {
                const int col = jb[q] - 1;
                if (slot[col] < 0) {
                    if (nz >= *nzmax) { *info = i + 1; goto done; }
                    slot[col] = (int)nz;
                    jc[nz] = col + 1;
                    c[nz] = av * b[q];
                    ++nz;
                } else {
                    c[slot[col]] += av * b[q];
                }
            }
        }
        for (long t = row_start; t < nz; ++t) slot[jc[t] - 1] = -1;
        ic[i + 1] = (MKL_INT)(nz + 1);
    }
done:
    free(slot);
}
