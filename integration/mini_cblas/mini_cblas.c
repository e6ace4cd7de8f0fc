/* integration/mini_cblas/mini_cblas.c -- see cblas.h.  C = alpha op(A) op(B) + beta C for both storage orders, as the BLAS defines it. */
#include "cblas.h"

#define GEMM_BODY(T) \
	/* a column-major product is the row-major product of the transposed problem: C^T = op(B)^T op(A)^T */ \
	if (order == CblasColMajor) \
	{ \
		const enum CBLAS_TRANSPOSE tt = transa; const T* pt = a; const int lt = lda, mt = m; \
		transa = transb, a = b, lda = ldb, m = n; \
		transb = tt, b = pt, ldb = lt, n = mt; \
	} \
	int i, j, p; \
	for (i = 0; i < m; i++) \
		for (j = 0; j < n; j++) \
		{ \
			T acc = 0; \
			for (p = 0; p < k; p++) \
				acc += (transa == CblasNoTrans ? a[(long)i * lda + p] : a[(long)p * lda + i]) * (transb == CblasNoTrans ? b[(long)p * ldb + j] : b[(long)j * ldb + p]); \
			c[(long)i * ldc + j] = alpha * acc + (beta == 0 ? 0 : beta * c[(long)i * ldc + j]); \
		}

void cblas_sgemm(const enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, const int k,
	const float alpha, const float* a, int lda, const float* b, int ldb, const float beta, float* c, const int ldc)
{
	GEMM_BODY(float)
}

void cblas_dgemm(const enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, const int k,
	const double alpha, const double* a, int lda, const double* b, int ldb, const double beta, double* c, const int ldc)
{
	GEMM_BODY(double)
}
