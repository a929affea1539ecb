/*
 * Minimal cblas.h for building the reference sources in oracle/_ref.
 *
 * TEST INFRASTRUCTURE ONLY. The reference's M>1 linears and its causal conv
 * call cblas_sgemm (voxtral_kernels.c:56,73,91,322 in /root/reference); this
 * container has an OpenBLAS binary (bundled inside the scipy wheel) but no
 * header, so this file declares the one entry point the reference uses and
 * maps it onto the scipy-prefixed symbol of that library.
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H

enum CBLAS_ORDER     { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

#ifdef ORACLE_SCIPY_OPENBLAS
#define cblas_sgemm scipy_cblas_sgemm
#endif

void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int m, int n, int k, float alpha, const float *a, int lda,
                 const float *b, int ldb, float beta, float *c, int ldc);

#endif
