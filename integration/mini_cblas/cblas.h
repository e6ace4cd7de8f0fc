/* integration/mini_cblas/cblas.h -- the two BLAS entry points the reference's CPU code calls when built with HAVE_CBLAS
 * (lib/ccv_algebra.c, lib/nnc/cmd/blas/cpu_sys/_ccv_nnc_gemm_cpu_sys.c), for the drop-in proof build only: the reference's cnnp layer
 * autotunes every CPU backend of a command, CPU_OPT's GEMM path included, and that path asserts without a BLAS.  A plain triple loop
 * (integration/mini_cblas/mini_cblas.c) -- test infrastructure for running ccv_cnnp models on the CPU, never on a measured path. */
#ifndef MINI_CBLAS_H
#define MINI_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE transa, const enum CBLAS_TRANSPOSE transb, const int m, const int n, const int k,
	const float alpha, const float* a, const int lda, const float* b, const int ldb, const float beta, float* c, const int ldc);
void cblas_dgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE transa, const enum CBLAS_TRANSPOSE transb, const int m, const int n, const int k,
	const double alpha, const double* a, const int lda, const double* b, const int ldb, const double beta, double* c, const int ldc);
#ifdef __cplusplus
}
#endif
#endif
