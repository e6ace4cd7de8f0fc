/* oracle/nnc_port.c -- TEST INFRASTRUCTURE ONLY: a plain-C restatement of the CCV_NNC_BACKEND_CPU_REF algorithms on the
 * hot path, written from the reference's behaviour (citations relative to /root/reference/lib/nnc/cmd unless noted).
 *
 * Pinned (tests/test_oracle.py, run in the build container): against the reference's literal known-answer vectors
 * (test/unit/nnc/gemm.tests.c) and against the compiled reference itself (oracle/_ref/libccv_ref.so) on seeded inputs.
 * It is never linked into, imported by or executed from the product (ccv_b200/): only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it (when oracle/_ref/libccv_ref.so is unavailable).
 *
 * Loop orders and accumulation types follow the reference so that results are the same to rounding: fp32 scalar
 * accumulators in (r, s, c) order for convolution (convolution/ccv_nnc_conv_cpu_ref.c:93-100), sequential-k fp32 for
 * GEMM (blas/ccv_nnc_gemm_cpu_ref.c:28-33), double max/sum for softmax (softmax/ccv_nnc_softmax_cpu_ref.c:27-36). */
#include "nnc_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int port_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

#define A_AT(i, k) (ta ? a[(size_t)(k) * M + (i)] : a[(size_t)(i) * K + (k)])
#define W_AT(k, j) (tb ? w[(size_t)(j) * K + (k)] : w[(size_t)(k) * N + (j)])

/* blas/ccv_nnc_gemm_cpu_ref.c:13-108 (_ccv_nnc_gbmm_and_bias / _ccv_nnc_gbmm) */
void port_gemm_forw(const float* a, const float* w, const float* bias, float* b, int M, int N, int K, int ta, int tb)
{
	int i;
#pragma omp parallel for schedule(dynamic)
	for (i = 0; i < M; i++)
	{
		int j, k;
		for (j = 0; j < N; j++)
		{
			float v = bias ? bias[j] : 0;
			for (k = 0; k < K; k++)
				v += A_AT(i, k) * W_AT(k, j);
			b[(size_t)i * N + j] = v;
		}
	}
}

/* blas/ccv_nnc_gemm_cpu_ref.c:186-448: dbias = column sums of g; dw = a^T g; h = g w^T */
void port_gemm_back(const float* g, const float* a, const float* w, float* h, float* dw, float* dbias, int M, int N, int K, int ta, int tb, int accumulate)
{
	int i, j, k;
	if (dbias)
	{
		if (!accumulate)
			memset(dbias, 0, sizeof(float) * N);
		for (i = 0; i < M; i++)
			for (j = 0; j < N; j++)
				dbias[j] += g[(size_t)i * N + j];
	}
	if (dw)
	{
		if (!accumulate)
			memset(dw, 0, sizeof(float) * (size_t)K * N);
#pragma omp parallel for schedule(dynamic) private(i, j)
		for (k = 0; k < K; k++)
			for (j = 0; j < N; j++)
			{
				float v = 0;
				for (i = 0; i < M; i++)
					v += A_AT(i, k) * g[(size_t)i * N + j];
				if (tb)
					dw[(size_t)j * K + k] += v;
				else
					dw[(size_t)k * N + j] += v;
			}
	}
	if (h)
	{
		if (!accumulate)
			memset(h, 0, sizeof(float) * (size_t)M * K);
#pragma omp parallel for schedule(dynamic) private(j, k)
		for (i = 0; i < M; i++)
			for (k = 0; k < K; k++)
			{
				float v = 0;
				for (j = 0; j < N; j++)
					v += g[(size_t)i * N + j] * W_AT(k, j);
				if (ta)
					h[(size_t)k * M + i] += v;
				else
					h[(size_t)i * K + k] += v;
			}
	}
}

/* convolution/ccv_nnc_conv_cpu_ref.c:47-108 (NHWC): the window is clipped to the input (SET_BORDER_OFFSET_SIZE_FOR,
 * ../ccv_nnc_internal.h:209-213), accumulation is a float scalar in (r, s, c) order starting from the bias. */
void port_conv_forw(const port_conv_t* c, const float* a, const float* w, const float* bias, float* b)
{
	const int cg = c->C / c->groups, kg = c->K / c->groups;
	int idx;
#pragma omp parallel for schedule(dynamic)
	for (idx = 0; idx < c->N * c->K; idx++)
	{
		const int n = idx / c->K, k = idx % c->K, gi = k / kg;
		int p, q, r, s, ch;
		for (p = 0; p < c->P; p++)
			for (q = 0; q < c->Q; q++)
			{
				float v = bias ? bias[k] : 0;
				for (r = 0; r < c->R; r++)
				{
					const int y = p * c->stride_h - c->pad_h + r * c->dil_h;
					if (y < 0 || y >= c->H)
						continue;
					for (s = 0; s < c->S; s++)
					{
						const int x = q * c->stride_w - c->pad_w + s * c->dil_w;
						if (x < 0 || x >= c->W)
							continue;
						const float* ap = a + (((size_t)n * c->H + y) * c->W + x) * c->C + gi * cg;
						const float* wp = w + (((size_t)k * c->R + r) * c->S + s) * cg;
						for (ch = 0; ch < cg; ch++)
							v += wp[ch] * ap[ch];
					}
				}
				b[(((size_t)n * c->P + p) * c->Q + q) * c->K + k] = v;
			}
	}
}

/* convolution/ccv_nnc_conv_cpu_ref.c:174-345: dw / dbias zeroed unless accumulating (:186-192); h is always zeroed and
 * scatter-added (:286,:333); gradients equal to 0 are skipped by the reference (:254,:325), which changes nothing. */
void port_conv_back(const port_conv_t* c, const float* g, const float* a, const float* w, float* h, float* dw, float* dbias, int accumulate)
{
	const int cg = c->C / c->groups, kg = c->K / c->groups;
	int k;
	if (dw || dbias)
	{
		if (!accumulate)
		{
			if (dw)
				memset(dw, 0, sizeof(float) * (size_t)c->K * c->R * c->S * cg);
			if (dbias)
				memset(dbias, 0, sizeof(float) * c->K);
		}
#pragma omp parallel for schedule(dynamic)
		for (k = 0; k < c->K; k++)
		{
			const int gi = k / kg;
			int n, p, q, r, s, ch;
			for (n = 0; n < c->N; n++)
				for (p = 0; p < c->P; p++)
					for (q = 0; q < c->Q; q++)
					{
						const float v = g[(((size_t)n * c->P + p) * c->Q + q) * c->K + k];
						if (dbias)
							dbias[k] += v;
						if (!dw)
							continue;
						for (r = 0; r < c->R; r++)
						{
							const int y = p * c->stride_h - c->pad_h + r * c->dil_h;
							if (y < 0 || y >= c->H)
								continue;
							for (s = 0; s < c->S; s++)
							{
								const int x = q * c->stride_w - c->pad_w + s * c->dil_w;
								if (x < 0 || x >= c->W)
									continue;
								const float* ap = a + (((size_t)n * c->H + y) * c->W + x) * c->C + gi * cg;
								float* wp = dw + (((size_t)k * c->R + r) * c->S + s) * cg;
								for (ch = 0; ch < cg; ch++)
									wp[ch] += v * ap[ch];
							}
						}
					}
		}
	}
	if (h)
	{
		int n;
		memset(h, 0, sizeof(float) * (size_t)c->N * c->H * c->W * c->C);
#pragma omp parallel for schedule(dynamic)
		for (n = 0; n < c->N; n++)
		{
			int kk, p, q, r, s, ch;
			for (kk = 0; kk < c->K; kk++)
			{
				const int gi = kk / kg;
				for (p = 0; p < c->P; p++)
					for (q = 0; q < c->Q; q++)
					{
						const float v = g[(((size_t)n * c->P + p) * c->Q + q) * c->K + kk];
						for (r = 0; r < c->R; r++)
						{
							const int y = p * c->stride_h - c->pad_h + r * c->dil_h;
							if (y < 0 || y >= c->H)
								continue;
							for (s = 0; s < c->S; s++)
							{
								const int x = q * c->stride_w - c->pad_w + s * c->dil_w;
								if (x < 0 || x >= c->W)
									continue;
								float* hp = h + (((size_t)n * c->H + y) * c->W + x) * c->C + gi * cg;
								const float* wp = w + (((size_t)kk * c->R + r) * c->S + s) * cg;
								for (ch = 0; ch < cg; ch++)
									hp[ch] += v * wp[ch];
							}
						}
					}
			}
		}
	}
}

/* norm/ccv_nnc_batch_norm_cpu_ref.c:47-250: mean, biased variance of (x - mean), running = momentum * running +
 * (1 - momentum) * batch, inv_std = 1 / sqrt(var + eps), y = (x - mean) * inv_std * scale + bias */
void port_bnorm_forw_train(const float* x, const float* scale, const float* bias, float* mean, float* var, float* y, float* saved_mean, float* saved_inv_std, size_t rows, int C, float epsilon, float momentum)
{
	const float inv = 1. / rows;
	size_t r;
	int c;
	for (c = 0; c < C; c++)
		saved_mean[c] = 0, saved_inv_std[c] = 0;
	for (r = 0; r < rows; r++)
		for (c = 0; c < C; c++)
			saved_mean[c] += x[r * C + c];
	for (c = 0; c < C; c++)
	{
		saved_mean[c] *= inv;
		mean[c] = momentum * mean[c] + (1. - momentum) * saved_mean[c];
	}
	for (r = 0; r < rows; r++)
		for (c = 0; c < C; c++)
		{
			const float d = x[r * C + c] - saved_mean[c];
			saved_inv_std[c] += d * d;
		}
	for (c = 0; c < C; c++)
	{
		saved_inv_std[c] *= inv;
		var[c] = momentum * var[c] + (1. - momentum) * saved_inv_std[c];
		saved_inv_std[c] = 1. / sqrtf(saved_inv_std[c] + epsilon);
	}
	for (r = 0; r < rows; r++)
		for (c = 0; c < C; c++)
			y[r * C + c] = (x[r * C + c] - saved_mean[c]) * saved_inv_std[c] * scale[c] + bias[c];
}

/* norm/ccv_nnc_batch_norm_cpu_ref.c:312-470 */
void port_bnorm_back(const float* g, const float* x, const float* scale, const float* saved_mean, const float* saved_inv_std, float* h, float* dscale, float* dbias, size_t rows, int C)
{
	size_t r;
	int c;
	for (c = 0; c < C; c++)
		dscale[c] = 0, dbias[c] = 0;
	for (r = 0; r < rows; r++)
		for (c = 0; c < C; c++)
		{
			const float xh = (x[r * C + c] - saved_mean[c]) * saved_inv_std[c];
			dbias[c] += g[r * C + c];
			dscale[c] += xh * g[r * C + c];
		}
	if (!h)
		return;
	for (r = 0; r < rows; r++)
		for (c = 0; c < C; c++)
		{
			const float xh = (x[r * C + c] - saved_mean[c]) * saved_inv_std[c];
			const float sisb = scale[c] * saved_inv_std[c] / rows;
			h[r * C + c] = sisb * (rows * g[r * C + c] - dbias[c] - xh * dscale[c]);
		}
}

/* relu/ccv_nnc_relu_cpu_ref.c:13-55 */
void port_relu_forw(const float* a, float* b, size_t n)
{
	size_t i;
	for (i = 0; i < n; i++)
		b[i] = a[i] > 0 ? a[i] : 0;
}

void port_relu_back(const float* g, const float* b, float* h, size_t n)
{
	size_t i;
	for (i = 0; i < n; i++)
		h[i] = b[i] > 0 ? g[i] : 0;
}

#define POOL_WINDOW(p_, q_) \
	const int y0 = (p_) * p->stride_h - p->pad_h < 0 ? 0 : (p_) * p->stride_h - p->pad_h; \
	const int y1 = (p_) * p->stride_h - p->pad_h + p->R > p->H ? p->H : (p_) * p->stride_h - p->pad_h + p->R; \
	const int x0 = (q_) * p->stride_w - p->pad_w < 0 ? 0 : (q_) * p->stride_w - p->pad_w; \
	const int x1 = (q_) * p->stride_w - p->pad_w + p->S > p->W ? p->W : (q_) * p->stride_w - p->pad_w + p->S

/* pool/ccv_nnc_max_pool_cpu_ref.c:13-59, applied to every image (the reference itself only walks image 0) */
void port_max_pool_forw(const port_pool_t* p, const float* a, float* b)
{
	int n, pp, q, c, y, x;
	for (n = 0; n < p->N; n++)
		for (pp = 0; pp < p->P; pp++)
			for (q = 0; q < p->Q; q++)
			{
				POOL_WINDOW(pp, q);
				for (c = 0; c < p->C; c++)
				{
					float v = a[(((size_t)n * p->H + y0) * p->W + x0) * p->C + c];
					for (y = y0; y < y1; y++)
						for (x = x0; x < x1; x++)
							if (a[(((size_t)n * p->H + y) * p->W + x) * p->C + c] > v)
								v = a[(((size_t)n * p->H + y) * p->W + x) * p->C + c];
					b[(((size_t)n * p->P + pp) * p->Q + q) * p->C + c] = v;
				}
			}
}

/* pool/ccv_nnc_max_pool_cpu_ref.c:61-139: every position equal to the window maximum receives the gradient */
void port_max_pool_back(const port_pool_t* p, const float* g, const float* a, const float* b, float* h)
{
	int n, pp, q, c, y, x;
	memset(h, 0, sizeof(float) * (size_t)p->N * p->H * p->W * p->C);
	for (n = 0; n < p->N; n++)
		for (pp = 0; pp < p->P; pp++)
			for (q = 0; q < p->Q; q++)
			{
				POOL_WINDOW(pp, q);
				for (c = 0; c < p->C; c++)
				{
					const size_t o = (((size_t)n * p->P + pp) * p->Q + q) * p->C + c;
					for (y = y0; y < y1; y++)
						for (x = x0; x < x1; x++)
							if (a[(((size_t)n * p->H + y) * p->W + x) * p->C + c] == b[o])
								h[(((size_t)n * p->H + y) * p->W + x) * p->C + c] += g[o];
				}
			}
}

/* pool/ccv_nnc_avg_pool_cpu_ref.c:13-58: divides by the clipped window size */
void port_avg_pool_forw(const port_pool_t* p, const float* a, float* b)
{
	int n, pp, q, c, y, x;
	for (n = 0; n < p->N; n++)
		for (pp = 0; pp < p->P; pp++)
			for (q = 0; q < p->Q; q++)
			{
				POOL_WINDOW(pp, q);
				for (c = 0; c < p->C; c++)
				{
					float v = 0;
					for (y = y0; y < y1; y++)
						for (x = x0; x < x1; x++)
							v += a[(((size_t)n * p->H + y) * p->W + x) * p->C + c];
					b[(((size_t)n * p->P + pp) * p->Q + q) * p->C + c] = v / ((y1 - y0) * (x1 - x0));
				}
			}
}

/* pool/ccv_nnc_avg_pool_cpu_ref.c:60-110 */
void port_avg_pool_back(const port_pool_t* p, const float* g, float* h)
{
	int n, pp, q, c, y, x;
	memset(h, 0, sizeof(float) * (size_t)p->N * p->H * p->W * p->C);
	for (n = 0; n < p->N; n++)
		for (pp = 0; pp < p->P; pp++)
			for (q = 0; q < p->Q; q++)
			{
				POOL_WINDOW(pp, q);
				for (c = 0; c < p->C; c++)
				{
					const float u = g[(((size_t)n * p->P + pp) * p->Q + q) * p->C + c] / ((y1 - y0) * (x1 - x0));
					for (y = y0; y < y1; y++)
						for (x = x0; x < x1; x++)
							h[(((size_t)n * p->H + y) * p->W + x) * p->C + c] += u;
				}
			}
}

/* softmax/ccv_nnc_softmax_cpu_ref.c:13-40 */
void port_softmax_forw(const float* a, float* b, int batch, int count)
{
	int i, j;
	for (i = 0; i < batch; i++)
	{
		const float* ap = a + (size_t)i * count;
		float* bp = b + (size_t)i * count;
		double maxval = ap[0], sumval = 0;
		for (j = 1; j < count; j++)
			if (ap[j] > maxval)
				maxval = ap[j];
		for (j = 0; j < count; j++)
			sumval += (bp[j] = expf(ap[j] - maxval));
		sumval = 1.0 / sumval;
		for (j = 0; j < count; j++)
			bp[j] *= sumval;
	}
}

/* softmax/ccv_nnc_softmax_cpu_ref.c:42-75 */
void port_softmax_back(const float* g, const float* b, float* h, int batch, int count)
{
	int i, j;
	for (i = 0; i < batch; i++)
	{
		float sumval = 0;
		for (j = 0; j < count; j++)
			sumval += g[(size_t)i * count + j] * b[(size_t)i * count + j];
		for (j = 0; j < count; j++)
			h[(size_t)i * count + j] = (g[(size_t)i * count + j] - sumval) * b[(size_t)i * count + j];
	}
}

/* loss/ccv_nnc_categorical_crossentropy_cpu_ref.c:74-103 (int32 labels) */
void port_cce_forw(const float* a, const int* label, float* c, int batch, int count, float trim0, float trim1)
{
	int i, j;
	for (i = 0; i < batch; i++)
	{
		const float* ap = a + (size_t)i * count;
		if (trim0 == 0 && trim1 == 1)
		{
			c[i] = -logf(ap[label[i]]);
			continue;
		}
		float p = 0;
		for (j = 0; j < count; j++)
			p += -(j == label[i] ? trim1 : trim0) * logf(ap[j]);
		c[i] = p;
	}
}

/* loss/ccv_nnc_categorical_crossentropy_cpu_ref.c:171-203: h = -g * t / a; g == NULL means 1 */
void port_cce_back(const float* g, const float* a, const int* label, float* h, int batch, int count, float trim0, float trim1)
{
	int i, j;
	for (i = 0; i < batch; i++)
	{
		const float gp = g ? g[i] : 1;
		for (j = 0; j < count; j++)
		{
			const float t = j == label[i] ? trim1 : trim0;
			h[(size_t)i * count + j] = t == 0 ? 0 : -gp * t / a[(size_t)i * count + j];
		}
	}
}

/* sgd/ccv_nnc_sgd_cpu_ref.c:16-126 */
void port_sgd(const float* g, const float* a, const float* m, float* b, float* n, size_t count, int nesterov, float rate, float scale, float decay, float momentum, float dampening)
{
	const float inv_dampening = 1 - dampening;
	size_t i;
	for (i = 0; i < count; i++)
	{
		if (nesterov)
		{
			float grad = scale * g[i];
			const float mom = n[i] = momentum * m[i] + grad + decay * a[i];
			grad += momentum * mom;
			b[i] = a[i] - rate * grad;
		} else {
			const float mom = n[i] = momentum * m[i] + inv_dampening * (scale * g[i] + decay * a[i]);
			b[i] = a[i] - rate * mom;
		}
	}
}

/* /root/reference/lib/ccv_util.c:1434-1440 and the base/shift tables above it (van der Zijp's method), as arithmetic:
 * truncating mantissa, flush below 2^-24, overflow to infinity, NaN keeps its top bits. */
void port_float_to_half(const float* f, uint16_t* h, size_t n)
{
	size_t i;
	for (i = 0; i < n; i++)
	{
		uint32_t u;
		memcpy(&u, f + i, 4);
		const uint32_t sign = (u >> 16) & 0x8000u;
		const int e = (int)((u >> 23) & 0xff) - 127;
		const uint32_t m = u & 0x007fffffu;
		if (e < -24)
			h[i] = (uint16_t)sign;
		else if (e < -14)
			h[i] = (uint16_t)(sign | ((0x0400u >> (-e - 14)) + (m >> (-e - 1))));
		else if (e <= 15)
			h[i] = (uint16_t)(sign | (((uint32_t)(e + 15) << 10) + (m >> 13)));
		else if (e < 128)
			h[i] = (uint16_t)(sign | 0x7c00u);
		else
			h[i] = (uint16_t)(sign | (0x7c00u + (m >> 13)));
	}
}

/* ------------------------------------------------------------------------------------------------ attention / row norms
 * Restated from /root/reference/lib/nnc/cmd/scaled_dot_product_attention/ccv_nnc_scaled_dot_product_attention_cpu_ref.c:88-183.
 * Packed tensors q [B, Sq, H, D], k [B, Sk, Hk, D], v [B, Sk, Hk, Dv], o [B, Sq, H, Dv]; query head h reads key / value head
 * h / (H / Hk) (:99,113-114); logits = scale * q.k (+ mask[Sq, Sk] when given, :127); causal keeps keys [0, x - Sq + Sk + 1) of
 * query x (:147) -- the masked tail contributes exactly 0; softmax with max and sum kept in double, exponent in float (:150-171);
 * output = sum_y p[y] * v[y] accumulated in key order (:172-180).  `scratch` holds Sk floats per call. */
static void port_sdpa_row(const float* q, const float* k, const float* v, const float* mask, float* o, float* p, int Sk, int D, int Dv, long long k_stride, long long v_stride, float scale, int visible)
{
	int y, d;
	for (y = 0; y < Sk; y++)
	{
		float dot = 0;
		for (d = 0; d < D; d++)
			dot += q[d] * k[(long long)y * k_stride + d];
		p[y] = scale * dot + (mask ? mask[y] : 0);
	}
	if (visible >= 0 && visible < Sk)
		memset(p + visible, 0, sizeof(float) * (size_t)(Sk - visible));
	if (visible > 0)
	{
		double mx = p[0], sum = 0;
		for (y = 1; y < visible; y++)
			if (p[y] > mx)
				mx = p[y];
		for (y = 0; y < visible; y++)
			sum += (p[y] = expf(p[y] - mx));
		sum = 1.0 / sum;
		for (y = 0; y < visible; y++)
			p[y] *= sum;
	}
	for (d = 0; d < Dv; d++)
		o[d] = 0;
	for (y = 0; y < Sk; y++)
		for (d = 0; d < Dv; d++)
			o[d] += p[y] * v[(long long)y * v_stride + d];
}

void port_sdpa_forw(const float* q, const float* k, const float* v, const float* mask, float* o, int B, int Sq, int Sk, int H, int Hk, int D, int Dv, float scale, int is_causal)
{
	const int ratio = H / Hk;
	const long long total = (long long)B * H * Sq;
	long long r;
#pragma omp parallel for schedule(static)
	for (r = 0; r < total; r++)
	{
		const int x = (int)(r % Sq), h = (int)((r / Sq) % H), b = (int)(r / ((long long)Sq * H));
		float* const p = (float*)malloc(sizeof(float) * (size_t)Sk);
		int visible = Sk;
		if (is_causal)
		{
			visible = x - Sq + Sk + 1;
			if (visible < 0)
				visible = 0;
			if (visible > Sk)
				visible = Sk;
		}
		port_sdpa_row(q + (((long long)b * Sq + x) * H + h) * D, k + ((long long)b * Sk * Hk + h / ratio) * D, v + ((long long)b * Sk * Hk + h / ratio) * Dv,
			mask ? mask + (long long)x * Sk : 0, o + (((long long)b * Sq + x) * H + h) * Dv, p, Sk, D, Dv, (long long)Hk * D, (long long)Hk * Dv, scale, visible);
		free(p);
	}
}

/* /root/reference/lib/nnc/cmd/norm/ccv_nnc_layer_norm_cpu_ref.c:16-190 for statistics over the trailing `inner` elements of each of
 * `rows` rows: mean = sum / n, inv_std = 1 / sqrtf(sum((x - mean)^2) / n + epsilon) (the epsilon sits inside the square root, :105),
 * y = (x - mean) * inv_std [* scale + bias].  rms = 1: /root/reference/lib/nnc/cmd/norm/ccv_nnc_rmsnorm_cpu_ref.c:16-130, no mean,
 * inv_std = 1 / sqrtf(sum(x^2) / n + epsilon), y = x * inv_std * scale. */
void port_row_norm_forw(const float* x, const float* scale, const float* bias, float* y, float* saved_mean, float* saved_inv_std, int rows, int inner, float epsilon, int rms)
{
	int r;
	const float inv_n = 1. / inner;
	for (r = 0; r < rows; r++)
	{
		const float* const xr = x + (size_t)r * inner;
		float* const yr = y + (size_t)r * inner;
		float mean = 0, var = 0;
		int i;
		if (!rms)
		{
			for (i = 0; i < inner; i++)
				mean += xr[i];
			mean = mean * inv_n;
		}
		for (i = 0; i < inner; i++)
		{
			const float w = xr[i] - mean;
			var += w * w;
		}
		const float inv_std = 1. / sqrtf(var * inv_n + epsilon);
		if (saved_mean)
			saved_mean[r] = mean;
		if (saved_inv_std)
			saved_inv_std[r] = inv_std;
		for (i = 0; i < inner; i++)
		{
			float t = (xr[i] - mean) * inv_std;
			if (scale)
				t = t * scale[i];
			if (bias)
				t += bias[i];
			yr[i] = t;
		}
	}
}
