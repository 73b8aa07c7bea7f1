/*
 * mkl_spblas.h -- stand-in for the one Intel MKL entry point the reference's
 * AMG *setup* calls (TEST INFRASTRUCTURE, not product code).
 *
 * Reference call site: /root/reference/src/saena_object_setup_matmat.cpp:214-218
 * (mkl_dcsrmultcsr, request=0, sort=7, one-based 3-array CSR).  MKL is an
 * un-vendored, un-pinned dependency of the reference ($MKLROOT,
 * CMakeLists.txt:106-128) and is absent from this image; mkl_shim.c restates
 * the published contract of the routine (C = A*B, row-by-row Gustavson
 * product, one-based indices, info=0 on success).  Any exact CSR product is
 * equivalent up to summation order inside one output entry.
 */
#ifndef SAENA_B200_ORACLE_MKL_SPBLAS_H
#define SAENA_B200_ORACLE_MKL_SPBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MKL_INT;
void mkl_dcsrmultcsr(const char *trans, const MKL_INT *request, const MKL_INT *sort, const MKL_INT *m,
                     const MKL_INT *n, const MKL_INT *k, double *a, MKL_INT *ja, MKL_INT *ia, double *b,
                     MKL_INT *jb, MKL_INT *ib, double *c, MKL_INT *jc, MKL_INT *ic, const MKL_INT *nzmax,
                     MKL_INT *info);
#ifdef __cplusplus
}
#endif
#endif
