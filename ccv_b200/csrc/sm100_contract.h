// sm100_contract.h -- internal C++ interface between the command layer (sm100_backend.cu) and the tensor-core
// contraction launchers (sm100_contract.cu).  All pointers are device pointers; shapes in elements.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace sm100 {

struct ConvGeom {
	// NHWC activations a[N, H, W, C] (element strides an, ah, aw; channel stride 1),
	// filters w[K, R, S, C] contiguous, outputs b[N, P, Q, K] (element strides bn, bh, bw).
	int N, H, W, C;
	int K, R, S;
	int P, Q;
	int stride_h, stride_w;
	int pad_h0, pad_h1, pad_w0, pad_w1; // border begin / end
	int dil_h, dil_w;
	long long an, ah, aw;
	long long bn, bh, bw;
};

// returns 0 on success, 1 if the shape/alignment cannot use the TMA+tcgen05 path (caller falls back to FFMA), <0 on CUDA error.
// C[M, N] (+)= op(A) * op(B) + bias; A is [M, K] (lda) or, if trans_a, [K, M]; B is [K, N] (ldb) or, if trans_b, [N, K].
int gemm_tf32(cudaStream_t stream, int M, int N, int K, const float* a, long long lda, int trans_a, const float* b, long long ldb, int trans_b, float* c, long long ldc, const float* bias, int accumulate);
int conv_fprop_tf32(cudaStream_t stream, const ConvGeom& g, const float* a, const float* w, const float* bias, float* b);
int conv_dgrad_tf32(cudaStream_t stream, const ConvGeom& g, const float* grad_b, const float* w, float* grad_a);
int conv_wgrad_tf32(cudaStream_t stream, const ConvGeom& g, const float* grad_b, const float* a, float* grad_w, int accumulate);

// CUDA-core fp32 versions of the same contractions: any stride/alignment, groups, exact fp32 products.
int gemm_ffma(cudaStream_t stream, int M, int N, int K, const float* a, long long a_rs, long long a_cs, const float* b, long long b_rs, long long b_cs, float* c, long long ldc, const float* bias, int accumulate);
int conv_fprop_ffma(cudaStream_t stream, const ConvGeom& g, int groups, const float* a, const float* w, const float* bias, float* b);
int conv_dgrad_ffma(cudaStream_t stream, const ConvGeom& g, int groups, const float* grad_b, const float* w, float* grad_a);
int conv_wgrad_ffma(cudaStream_t stream, const ConvGeom& g, int groups, const float* grad_b, const float* a, float* grad_w, int accumulate);

// bookkeeping shared by every launcher in the backend
void count_launch(int n = 1);
unsigned long long launch_count();
void set_last_error(const char* what, cudaError_t err);
const char* last_error();

} // namespace sm100
