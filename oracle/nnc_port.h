/* oracle/nnc_port.h -- TEST INFRASTRUCTURE ONLY (see nnc_port.c). Dense row-major fp32 arrays, NHWC activations. */
#ifndef ORACLE_NNC_PORT_H
#define ORACLE_NNC_PORT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int port_num_threads(void);
/* b[M,N] = bias[N] + op(a) op(w); a is [M,K] (or [K,M] if ta), w is [K,N] (or [N,K] if tb) */
void port_gemm_forw(const float* a, const float* w, const float* bias, float* b, int M, int N, int K, int ta, int tb);
/* h[M,K] (layout of a), dw (layout of w), dbias[N]; NULL outputs are skipped; accumulate = CCV_NNC_ACCUMULATE_OUTPUT */
void port_gemm_back(const float* g, const float* a, const float* w, float* h, float* dw, float* dbias, int M, int N, int K, int ta, int tb, int accumulate);
typedef struct {
	int N, H, W, C, K, R, S, P, Q, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups;
} port_conv_t;
void port_conv_forw(const port_conv_t* c, const float* a, const float* w, const float* bias, float* b);
void port_conv_back(const port_conv_t* c, const float* g, const float* a, const float* w, float* h, float* dw, float* dbias, int accumulate);
/* batch norm over [rows, C] (NHWC per-channel) */
void port_bnorm_forw_train(const float* x, const float* scale, const float* bias, float* mean, float* var, float* y, float* saved_mean, float* saved_inv_std, size_t rows, int C, float epsilon, float momentum);
void port_bnorm_back(const float* g, const float* x, const float* scale, const float* saved_mean, const float* saved_inv_std, float* h, float* dscale, float* dbias, size_t rows, int C);
void port_relu_forw(const float* a, float* b, size_t n);
void port_relu_back(const float* g, const float* b, float* h, size_t n);
typedef struct {
	int N, H, W, C, P, Q, R, S, stride_h, stride_w, pad_h, pad_w;
} port_pool_t;
void port_max_pool_forw(const port_pool_t* p, const float* a, float* b);
void port_max_pool_back(const port_pool_t* p, const float* g, const float* a, const float* b, float* h);
void port_avg_pool_forw(const port_pool_t* p, const float* a, float* b);
void port_avg_pool_back(const port_pool_t* p, const float* g, float* h);
void port_softmax_forw(const float* a, float* b, int batch, int count);
void port_softmax_back(const float* g, const float* b, float* h, int batch, int count);
void port_cce_forw(const float* a, const int* label, float* c, int batch, int count, float trim0, float trim1);
void port_cce_back(const float* g, const float* a, const int* label, float* h, int batch, int count, float trim0, float trim1);
void port_sgd(const float* g, const float* a, const float* m, float* b, float* n, size_t count, int nesterov, float rate, float scale, float decay, float momentum, float dampening);
void port_float_to_half(const float* f, uint16_t* h, size_t n);
/* scaled dot product attention forward, packed [B, S, H, D] tensors, optional additive mask [Sq, Sk] (cpu_ref.c:88-183) */
void port_sdpa_forw(const float* q, const float* k, const float* v, const float* mask, float* o, int B, int Sq, int Sk, int H, int Hk, int D, int Dv, float scale, int is_causal);
/* layer norm (rms = 0) / rms norm (rms = 1) forward over the trailing `inner` elements of `rows` rows */
void port_row_norm_forw(const float* x, const float* scale, const float* bias, float* y, float* saved_mean, float* saved_inv_std, int rows, int inner, float epsilon, int rms);
#ifdef __cplusplus
}
#endif
#endif
