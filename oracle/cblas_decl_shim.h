/* oracle/cblas_decl_shim.h -- TEST INFRASTRUCTURE. Declaration-only stand-in for <cblas.h>:
 * /root/reference/lib/nnc/cmd/blas/cpu_sys/_ccv_nnc_gemm_cpu_sys.c:13-24 names cblas_sgemm and the
 * Cblas* enumerators outside its HAVE_CBLAS guard (:67). Without HAVE_CBLAS the callers are unused
 * static-inline functions, so nothing here is ever linked or called. */
#ifndef ORACLE_CBLAS_DECL_SHIM_H
#define ORACLE_CBLAS_DECL_SHIM_H
enum { CblasRowMajor = 101, CblasColMajor = 102 };
enum { CblasNoTrans = 111, CblasTrans = 112 };
void cblas_sgemm(int order, int transa, int transb, int m, int n, int k, float alpha, const float* a, int lda,
	const float* b, int ldb, float beta, float* c, int ldc);
#endif
