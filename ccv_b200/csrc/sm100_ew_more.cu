// sm100_ew_more.cu -- the remaining element-wise / reduction commands the graphs around the hot path reach (VERDICT
// r01 "missing" 8): SIGMOID, TANH, LEAKY_RELU, EWDIV, EWEXP, EWLOG, EWSQRT, CLAMP (forward / backward), REDUCE_SUM / MEAN / MAX / MIN /
// NORM2 (forward / backward) and MASKED_FILL (forward / backward).  All HBM-bound: grid-stride kernels with 16-byte accesses on
// packed tensors, fp32 arithmetic, fp32 / bf16 / fp16 tensors; reductions are fixed-order (one warp or one block per output
// element, shuffle tree), so results are bit-reproducible.  Semantics (paths relative to /root/reference/lib/nnc/cmd):
//   sigmoid/ccv_nnc_sigmoid_cpu_ref.c:14-78   b = 1 / (1 + exp(-a));  backward (g, -, b): h = g b (1 - b)
//   tanh/ccv_nnc_tanh_cpu_ref.c:14-63         b = tanh(a);            backward (g, -, b): h = g (1 - b^2)
//   leaky_relu/ccv_nnc_leaky_relu_cpu_ref.c   b = a > 0 ? a : slope a; backward (g, -, b): h = b >= 0 ? g : slope g
//   ew/ccv_nnc_ew_cpu_ref.c:501-971           EWDIV c = a / b (a == NULL: 1 / b); backward (g, a, b, c): ha = g / b, hb = -g c / b
//   ew/ccv_nnc_ew_cpu_ref.c:974-1196          EWEXP / EWLOG / EWSQRT; backward h = g b | g / a | 0.5 g / b
//   ew/ccv_nnc_ew_cpu_ref.c:1198-1500         CLAMP to [min, max] (NaN = open side); backward (g, -, b): 0 where b sits on a bound
//   reduce/ccv_nnc_reduce_{sum,mean,max,min,norm2}_cpu_ref.c   the output's unit dimensions are the reduced axes
//   util/ccv_nnc_util_cpu_ref.c:1300-1481     MASKED_FILL c = mask == p ? q : a (mask int32 or fp32, broadcast); backward q = 0
// A missing gradient (inputs[0] == NULL) reads as all ones, as in the reference.
#include "../../include/ccv_nnc_sm100.h"
#include "sm100_contract.h"
#include "sm100_elem.cuh"
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

using namespace sm100;

namespace {

int g_sms = 0;
int sms()
{
	if (!g_sms)
	{
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
		if (g_sms <= 0)
			g_sms = 148;
	}
	return g_sms;
}
int grid_for(size_t work, int threads)
{
	size_t blocks = (work + threads - 1) / threads;
	const size_t cap = (size_t)sms() * 8;
	if (blocks > cap)
		blocks = cap;
	return blocks < 1 ? 1 : (int)blocks;
}
int check(const char* what)
{
	count_launch();
	const cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
	{
		set_last_error(what, e);
		return -1;
	}
	return 0;
}
inline cudaStream_t stream_of(ccv_nnc_stream_context_t* const sc) { return (cudaStream_t)ccv_nnc_stream_context_get_stream(sc); }
inline int kind_of(const ccv_nnc_tensor_t* const t)
{
	switch (CCV_GET_DATA_TYPE(t->info.datatype))
	{
		case CCV_32F: return 0;
		case CCV_16BF: return 1;
		case CCV_16F: return 2;
	}
	return -1;
}
int nd_of(const ccv_nnc_tensor_t* const t)
{
	int i;
	for (i = 0; i < CCV_NNC_MAX_DIM_ALLOC && t->info.dim[i] > 0; i++) {}
	return i;
}
size_t count_of(const ccv_nnc_tensor_t* const t)
{
	size_t n = 1;
	for (int i = 0; i < CCV_NNC_MAX_DIM_ALLOC && t->info.dim[i] > 0; i++)
		n *= (size_t)t->info.dim[i];
	return n;
}
bool same_dims(const ccv_nnc_tensor_t* const a, const ccv_nnc_tensor_t* const b)
{
	for (int i = 0; i < CCV_NNC_MAX_DIM_ALLOC; i++)
	{
		if (a->info.dim[i] != b->info.dim[i])
			return false;
		if (a->info.dim[i] == 0)
			break;
	}
	return true;
}
// right-aligned 4-d shape and strides (ccv_nnc_tensor_view_get_dim / _get_stride); false when the tensor has more than 4 axes
bool dims4(const ccv_nnc_tensor_t* const t, int dim[4], int stride[4])
{
	const int nd = nd_of(t);
	if (nd > 4 || nd < 1)
		return false;
	const int off = 4 - nd;
	int packed = 1;
	for (int i = 3; i >= 0; i--)
	{
		if (i < off)
			dim[i] = 1, stride[i] = 0;
		else {
			dim[i] = t->info.dim[i - off];
			stride[i] = CCV_IS_TENSOR_VIEW(t) ? ((const ccv_nnc_tensor_view_t*)t)->stride[i - off] : packed;
			packed *= dim[i];
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ element-wise family
enum { OP_SIGMOID, OP_TANH, OP_LEAKY, OP_EXP, OP_LOG, OP_SQRT, OP_CLAMP, OP_DIV,
	OP_SIGMOID_B, OP_TANH_B, OP_LEAKY_B, OP_EXP_B, OP_LOG_B, OP_SQRT_B, OP_CLAMP_B, OP_DIV_BA, OP_DIV_BB };
struct EwArg {
	float p0, p1; // slope | (min, max)
	int has_x, has_y; // optional first / second operand present
};
// x = first operand (a, or the gradient g: 1 when absent), y = second operand (b / the saved output), z = third (c of EWDIV)
template <int OP>
__device__ __forceinline__ float ew_op(const float x, const float y, const float z, const EwArg& k)
{
	switch (OP)
	{
		case OP_SIGMOID: return 1.f / (1.f + expf(-x));
		case OP_TANH: return tanhf(x);
		case OP_LEAKY: return x > 0 ? x : x * k.p0;
		case OP_EXP: return expf(x);
		case OP_LOG: return logf(x);
		case OP_SQRT: return sqrtf(x);
		case OP_CLAMP: {
			float v = x;
			if (!isnan(k.p1))
				v = fminf(v, k.p1);
			if (!isnan(k.p0))
				v = fmaxf(v, k.p0);
			return v;
		}
		case OP_DIV: return x / y;
		case OP_SIGMOID_B: return x * y * (1.f - y);
		case OP_TANH_B: return x * (1.f - y * y);
		case OP_LEAKY_B: return y >= 0 ? x : k.p0 * x;
		case OP_EXP_B: return x * y;
		case OP_LOG_B: return x / y;
		case OP_SQRT_B: return 0.5f * x / y;
		case OP_CLAMP_B: return ((!isnan(k.p1) && y >= k.p1) || (!isnan(k.p0) && y <= k.p0)) ? 0.f : x;
		case OP_DIV_BA: return x / y;
		case OP_DIV_BB: return -x * z / y;
	}
	return 0.f;
}
template <typename T, int OP>
__global__ void __launch_bounds__(256) ew_kernel(const T* __restrict__ xp, const T* __restrict__ yp, const T* __restrict__ zp, T* __restrict__ out, const size_t n, const int vec, const EwArg k)
{
	constexpr int W = Vec16<T>::W;
	const size_t nw = vec ? n / W : 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nw; i += (size_t)gridDim.x * blockDim.x)
	{
		float x[W], y[W], z[W], o[W];
		if (xp)
			ldv(xp + i * W, x);
		if (yp)
			ldv(yp + i * W, y);
		if (zp)
			ldv(zp + i * W, z);
#pragma unroll
		for (int j = 0; j < W; j++)
			o[j] = ew_op<OP>(xp ? x[j] : 1.f, yp ? y[j] : 0.f, zp ? z[j] : 0.f, k);
		stv(out + i * W, o);
	}
	for (size_t i = nw * W + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		stf(out + i, ew_op<OP>(xp ? ldf(xp + i) : 1.f, yp ? ldf(yp + i) : 0.f, zp ? ldf(zp + i) : 0.f, k));
}
template <typename T, int OP>
int run_ew_t(cudaStream_t s, const void* x, const void* y, const void* z, void* out, const size_t n, const EwArg& k)
{
	if (n == 0)
		return 0;
	const int vec = aligned_v16((const T*)out) && (!x || aligned_v16((const T*)x)) && (!y || aligned_v16((const T*)y)) && (!z || aligned_v16((const T*)z));
	ew_kernel<T, OP><<<grid_for(vec ? n / Vec16<T>::W + 1 : n, 256), 256, 0, s>>>((const T*)x, (const T*)y, (const T*)z, (T*)out, n, vec, k);
	return check("element-wise");
}
template <int OP>
int run_ew(cudaStream_t s, const int kind, const void* x, const void* y, const void* z, void* out, const size_t n, const EwArg& k)
{
	if (kind == 0)
		return run_ew_t<float, OP>(s, x, y, z, out, n, k);
	if (kind == 1)
		return run_ew_t<__nv_bfloat16, OP>(s, x, y, z, out, n, k);
	return run_ew_t<__half, OP>(s, x, y, z, out, n, k);
}
// packed tensor of `kind` with the shape of `like`; NULL allowed when `optional`
bool ok_operand(const ccv_nnc_tensor_t* const t, const ccv_nnc_tensor_t* const like, const int kind)
{
	return t && kind_of(t) == kind && CCV_IS_TENSOR_CONTIGUOUS(t) && same_dims(t, like);
}
EwArg arg_of(const ccv_nnc_cmd_t& cmd, const int op)
{
	EwArg k;
	memset(&k, 0, sizeof(k));
	if (op == OP_LEAKY || op == OP_LEAKY_B)
		k.p0 = cmd.info.leaky_relu.negative_slope;
	if (op == OP_CLAMP || op == OP_CLAMP_B)
		k.p0 = cmd.info.clamp.min, k.p1 = cmd.info.clamp.max;
	return k;
}
// forward: inputs (a) -> outputs (b)
template <int OP>
int exec_unary_forw(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 1 || output_size < 1 || !inputs[0] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(outputs[0]);
	if (kind < 0 || !CCV_IS_TENSOR_CONTIGUOUS(outputs[0]) || !ok_operand(inputs[0], outputs[0], kind))
		return CCV_NNC_EXEC_INVALID;
	return run_ew<OP>(stream_of(stream_context), kind, inputs[0]->data.u8, 0, 0, outputs[0]->data.u8, count_of(outputs[0]), arg_of(cmd, OP)) ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
// backward: inputs (g, a, b) -> outputs (h); WHICH = 1: the derivative reads the forward input a, 2: the forward output b
template <int OP, int WHICH>
int exec_unary_back(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size <= WHICH || output_size < 1 || !inputs[WHICH] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(outputs[0]);
	if (kind < 0 || !CCV_IS_TENSOR_CONTIGUOUS(outputs[0]) || !ok_operand(inputs[WHICH], outputs[0], kind) || (inputs[0] && !ok_operand(inputs[0], outputs[0], kind)))
		return CCV_NNC_EXEC_INVALID;
	return run_ew<OP>(stream_of(stream_context), kind, inputs[0] ? inputs[0]->data.u8 : 0, inputs[WHICH]->data.u8, 0, outputs[0]->data.u8, count_of(outputs[0]), arg_of(cmd, OP)) ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
// EWDIV forward: inputs (a | NULL, b) -> c
int exec_ewdiv_forw(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 2 || output_size < 1 || !inputs[1] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(outputs[0]);
	if (kind < 0 || !CCV_IS_TENSOR_CONTIGUOUS(outputs[0]) || !ok_operand(inputs[1], outputs[0], kind) || (inputs[0] && !ok_operand(inputs[0], outputs[0], kind)))
		return CCV_NNC_EXEC_INVALID;
	return run_ew<OP_DIV>(stream_of(stream_context), kind, inputs[0] ? inputs[0]->data.u8 : 0, inputs[1]->data.u8, 0, outputs[0]->data.u8, count_of(outputs[0]), arg_of(cmd, OP_DIV)) ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
// EWDIV backward: inputs (g | NULL, a, b, c) -> outputs (ha | NULL, hb | NULL)
int exec_ewdiv_back(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 3 || output_size < 1 || !inputs[2])
		return CCV_NNC_EXEC_INVALID;
	ccv_nnc_tensor_t* const ha = outputs[0];
	ccv_nnc_tensor_t* const hb = output_size > 1 ? outputs[1] : 0;
	const ccv_nnc_tensor_t* const b = inputs[2];
	const int kind = kind_of(b);
	if (kind < 0 || !CCV_IS_TENSOR_CONTIGUOUS(b) || (inputs[0] && !ok_operand(inputs[0], b, kind)))
		return CCV_NNC_EXEC_INVALID;
	cudaStream_t s = stream_of(stream_context);
	const EwArg k = arg_of(cmd, OP_DIV);
	const size_t n = count_of(b);
	if (ha)
	{
		if (!ok_operand(ha, b, kind))
			return CCV_NNC_EXEC_INVALID;
		if (run_ew<OP_DIV_BA>(s, kind, inputs[0] ? inputs[0]->data.u8 : 0, b->data.u8, 0, ha->data.u8, n, k))
			return CCV_NNC_EXEC_INVALID;
	}
	if (hb)
	{
		if (input_size < 4 || !inputs[3] || !ok_operand(inputs[3], b, kind) || !ok_operand(hb, b, kind))
			return CCV_NNC_EXEC_INVALID;
		if (run_ew<OP_DIV_BB>(s, kind, inputs[0] ? inputs[0]->data.u8 : 0, b->data.u8, inputs[3]->data.u8, hb->data.u8, n, k))
			return CCV_NNC_EXEC_INVALID;
	}
	return CCV_NNC_EXEC_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ reductions
// a [d0, d1, d2, d3] -> b whose unit dimensions are the reduced axes.  The 4-d index space is split into the kept axes (one
// output element each) and the reduced axes (walked in a fixed order by one warp or one block, combined by a shuffle tree).
enum { RED_SUM, RED_MEAN, RED_MAX, RED_MIN, RED_NORM2 };
struct Red4 {
	int kd[4], rd[4]; // kept / reduced extent per axis (1 where the axis is of the other kind)
	long long as[4], bs[4]; // element strides of a and b
	long long rcount;
};
template <int MODE> __device__ __forceinline__ float red_identity() { return MODE == RED_MAX ? -INFINITY : MODE == RED_MIN ? INFINITY : 0.f; }
template <int MODE> __device__ __forceinline__ float red_combine(const float x, const float y) { return MODE == RED_MAX ? fmaxf(x, y) : MODE == RED_MIN ? fminf(x, y) : x + y; }
template <int MODE> __device__ __forceinline__ float red_map(const float v) { return MODE == RED_NORM2 ? v * v : v; }
template <typename T, int MODE, int THREADS>
__global__ void __launch_bounds__(256) reduce_kernel(const T* __restrict__ a, T* __restrict__ b, const Red4 r, const size_t outputs, const float scale)
{
	// one group of THREADS (a warp or the block) per output element
	constexpr int GROUPS = 256 / THREADS;
	__shared__ float part[8];
	const int lane = threadIdx.x % THREADS, group = threadIdx.x / THREADS;
	for (size_t o = (size_t)blockIdx.x * GROUPS + group; o < outputs; o += (size_t)gridDim.x * GROUPS)
	{
		size_t t = o;
		const int k3 = (int)(t % r.kd[3]); t /= r.kd[3];
		const int k2 = (int)(t % r.kd[2]); t /= r.kd[2];
		const int k1 = (int)(t % r.kd[1]); t /= r.kd[1];
		const int k0 = (int)t;
		const T* const ap = a + k0 * r.as[0] + k1 * r.as[1] + k2 * r.as[2] + k3 * r.as[3];
		float acc = red_identity<MODE>();
		for (long long j = lane; j < r.rcount; j += THREADS)
		{
			long long u = j;
			const int j3 = (int)(u % r.rd[3]); u /= r.rd[3];
			const int j2 = (int)(u % r.rd[2]); u /= r.rd[2];
			const int j1 = (int)(u % r.rd[1]); u /= r.rd[1];
			const int j0 = (int)u;
			acc = red_combine<MODE>(acc, red_map<MODE>(ldf(ap + j0 * r.as[0] + j1 * r.as[1] + j2 * r.as[2] + j3 * r.as[3])));
		}
#pragma unroll
		for (int off = 16; off > 0; off >>= 1)
			acc = red_combine<MODE>(acc, __shfl_xor_sync(0xffffffffu, acc, off));
		if (THREADS > 32)
		{
			if ((threadIdx.x & 31) == 0)
				part[threadIdx.x >> 5] = acc;
			__syncthreads();
			acc = part[0];
#pragma unroll
			for (int w = 1; w < THREADS / 32; w++)
				acc = red_combine<MODE>(acc, part[w]);
			__syncthreads();
		}
		if (lane == 0)
		{
			if (MODE == RED_MEAN)
				acc *= scale;
			if (MODE == RED_NORM2)
				acc = sqrtf(acc);
			stf(b + k0 * r.bs[0] + k1 * r.bs[1] + k2 * r.bs[2] + k3 * r.bs[3], acc);
		}
	}
}
// backward of every reduction as one broadcast kernel: h[i] = f(g[o(i)], a[i], b[o(i)])
template <typename T, int MODE>
__global__ void __launch_bounds__(256) reduce_back_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ h, const int d1, const int d2, const int d3, const size_t n,
	const long long hs0, const long long hs1, const long long hs2, const long long hs3, const long long as0, const long long as1, const long long as2, const long long as3,
	const long long gs0, const long long gs1, const long long gs2, const long long gs3, const long long bs0, const long long bs1, const long long bs2, const long long bs3, const float scale)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
	{
		size_t t = i;
		const int i3 = (int)(t % d3); t /= d3;
		const int i2 = (int)(t % d2); t /= d2;
		const int i1 = (int)(t % d1); t /= d1;
		const int i0 = (int)t;
		const float gv = g ? ldf(g + i0 * gs0 + i1 * gs1 + i2 * gs2 + i3 * gs3) : 1.f;
		float v;
		if (MODE == RED_SUM)
			v = gv;
		else if (MODE == RED_MEAN)
			v = gv * scale;
		else {
			const float av = ldf(a + i0 * as0 + i1 * as1 + i2 * as2 + i3 * as3), bv = ldf(b + i0 * bs0 + i1 * bs1 + i2 * bs2 + i3 * bs3);
			v = MODE == RED_NORM2 ? gv * av / bv : (av == bv ? gv : 0.f);
		}
		stf(h + i0 * hs0 + i1 * hs1 + i2 * hs2 + i3 * hs3, v);
	}
}
template <typename T, int MODE>
int run_reduce_t(cudaStream_t s, const void* a, void* b, const Red4& r, const size_t outputs, const float scale)
{
	if (r.rcount <= 64)
		reduce_kernel<T, MODE, 32><<<grid_for(outputs * 32, 256), 256, 0, s>>>((const T*)a, (T*)b, r, outputs, scale);
	else
		reduce_kernel<T, MODE, 256><<<(int)(outputs < (size_t)sms() * 8 ? outputs : (size_t)sms() * 8), 256, 0, s>>>((const T*)a, (T*)b, r, outputs, scale);
	return check("reduce");
}
// does `small` broadcast onto `big` (every axis equal or 1)?
bool broadcast_onto(const int small[4], const int big[4])
{
	for (int i = 0; i < 4; i++)
		if (small[i] != big[i] && small[i] != 1)
			return false;
	return true;
}
template <int MODE>
int exec_reduce_forw(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 1 || output_size < 1 || !inputs[0] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(inputs[0]);
	int ad[4], as[4], bd[4], bs[4];
	if (kind < 0 || kind_of(outputs[0]) != kind || !dims4(inputs[0], ad, as) || !dims4(outputs[0], bd, bs) || !broadcast_onto(bd, ad))
		return CCV_NNC_EXEC_INVALID;
	Red4 r;
	size_t outs = 1;
	r.rcount = 1;
	for (int i = 0; i < 4; i++)
	{
		const int reduced = bd[i] == 1 && ad[i] != 1;
		r.kd[i] = reduced ? 1 : ad[i], r.rd[i] = reduced ? ad[i] : 1;
		r.as[i] = as[i], r.bs[i] = bd[i] == 1 ? 0 : bs[i];
		outs *= (size_t)r.kd[i], r.rcount *= r.rd[i];
	}
	if (outs == 0 || r.rcount == 0)
		return CCV_NNC_EXEC_SUCCESS;
	const float scale = 1.f / (float)r.rcount; // count(b) / count(a), reduce_mean_cpu_ref.c:36
	cudaStream_t s = stream_of(stream_context);
	int rc;
	if (kind == 0)
		rc = run_reduce_t<float, MODE>(s, inputs[0]->data.u8, outputs[0]->data.u8, r, outs, scale);
	else if (kind == 1)
		rc = run_reduce_t<__nv_bfloat16, MODE>(s, inputs[0]->data.u8, outputs[0]->data.u8, r, outs, scale);
	else
		rc = run_reduce_t<__half, MODE>(s, inputs[0]->data.u8, outputs[0]->data.u8, r, outs, scale);
	return rc ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
// inputs (g | NULL, a, b) -> outputs (h); SUM / MEAN only read g
template <int MODE>
int exec_reduce_back(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (output_size < 1 || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	ccv_nnc_tensor_t* const h = outputs[0];
	const ccv_nnc_tensor_t* const g = input_size > 0 ? inputs[0] : 0;
	const int needs_ab = MODE == RED_MAX || MODE == RED_MIN || MODE == RED_NORM2;
	const ccv_nnc_tensor_t* const a = needs_ab && input_size > 1 ? inputs[1] : 0;
	const ccv_nnc_tensor_t* const b = needs_ab && input_size > 2 ? inputs[2] : 0;
	if (needs_ab && (!a || !b))
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(h);
	int hd[4], hs[4], gd[4] = { 1, 1, 1, 1 }, gs[4] = { 0, 0, 0, 0 }, ad[4], as[4] = { 0, 0, 0, 0 }, bd[4] = { 1, 1, 1, 1 }, bs[4] = { 0, 0, 0, 0 };
	if (kind < 0 || !dims4(h, hd, hs))
		return CCV_NNC_EXEC_INVALID;
	if (g && (kind_of(g) != kind || !dims4(g, gd, gs) || !broadcast_onto(gd, hd)))
		return CCV_NNC_EXEC_INVALID;
	if (needs_ab && (kind_of(a) != kind || kind_of(b) != kind || !dims4(a, ad, as) || !dims4(b, bd, bs) || memcmp(ad, hd, sizeof(ad)) != 0 || !broadcast_onto(bd, hd)))
		return CCV_NNC_EXEC_INVALID;
	const size_t n = (size_t)hd[0] * hd[1] * hd[2] * hd[3];
	if (n == 0)
		return CCV_NNC_EXEC_SUCCESS;
	size_t gcount = 1;
	for (int i = 0; i < 4; i++)
	{
		gcount *= (size_t)gd[i];
		if (gd[i] == 1)
			gs[i] = 0;
		if (bd[i] == 1)
			bs[i] = 0;
	}
	// the mean's weight: with no gradient the reference uses 1 / (extent of the axes named by the command); with one, count(g) / count(h)
	float scale = 1.f;
	if (MODE == RED_MEAN)
	{
		if (g)
			scale = (float)gcount / (float)n;
		else {
			long long dims = 1;
			const int nd = nd_of(h);
			for (int i = 0; i < cmd.info.reduce.count; i++)
				if (cmd.info.reduce.axis[i] >= 0 && cmd.info.reduce.axis[i] < nd)
					dims *= h->info.dim[cmd.info.reduce.axis[i]];
			scale = 1.f / (float)dims;
		}
	}
	cudaStream_t s = stream_of(stream_context);
	const int grid = grid_for(n, 256);
#define LAUNCH_RB(T) reduce_back_kernel<T, MODE><<<grid, 256, 0, s>>>(g ? (const T*)g->data.u8 : 0, a ? (const T*)a->data.u8 : 0, b ? (const T*)b->data.u8 : 0, (T*)h->data.u8, hd[1], hd[2], hd[3], n, \
		hs[0], hs[1], hs[2], hs[3], as[0], as[1], as[2], as[3], gs[0], gs[1], gs[2], gs[3], bs[0], bs[1], bs[2], bs[3], scale)
	if (kind == 0)
		LAUNCH_RB(float);
	else if (kind == 1)
		LAUNCH_RB(__nv_bfloat16);
	else
		LAUNCH_RB(__half);
#undef LAUNCH_RB
	return check("reduce backward") ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ masked fill
// c = mask == p ? q : a, a and mask broadcast onto c
template <typename T, typename M>
__global__ void __launch_bounds__(256) masked_fill_kernel(const T* __restrict__ a, const M* __restrict__ mask, T* __restrict__ c, const int d1, const int d2, const int d3, const size_t n,
	const long long as0, const long long as1, const long long as2, const long long as3, const long long ms0, const long long ms1, const long long ms2, const long long ms3,
	const long long cs0, const long long cs1, const long long cs2, const long long cs3, const M p, const float q)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
	{
		size_t t = i;
		const int i3 = (int)(t % d3); t /= d3;
		const int i2 = (int)(t % d2); t /= d2;
		const int i1 = (int)(t % d1); t /= d1;
		const int i0 = (int)t;
		const M m = mask[i0 * ms0 + i1 * ms1 + i2 * ms2 + i3 * ms3];
		stf(c + i0 * cs0 + i1 * cs1 + i2 * cs2 + i3 * cs3, m == p ? q : ldf(a + i0 * as0 + i1 * as1 + i2 * as2 + i3 * as3));
	}
}
int masked_fill(cudaStream_t s, const float p, const float q, const ccv_nnc_tensor_t* const a, const ccv_nnc_tensor_t* const mask, ccv_nnc_tensor_t* const c)
{
	const int kind = kind_of(c);
	int ad[4], as[4], md[4], ms[4], cd[4], cs[4];
	const int mtype = CCV_GET_DATA_TYPE(mask->info.datatype);
	if (kind < 0 || kind_of(a) != kind || (mtype != CCV_32F && mtype != CCV_32S) || !dims4(a, ad, as) || !dims4(mask, md, ms) || !dims4(c, cd, cs) || !broadcast_onto(ad, cd) || !broadcast_onto(md, cd))
		return CCV_NNC_EXEC_INVALID;
	for (int i = 0; i < 4; i++)
	{
		if (ad[i] == 1)
			as[i] = 0;
		if (md[i] == 1)
			ms[i] = 0;
	}
	const size_t n = (size_t)cd[0] * cd[1] * cd[2] * cd[3];
	if (n == 0)
		return CCV_NNC_EXEC_SUCCESS;
	const int grid = grid_for(n, 256);
#define LAUNCH_MF(T, M, pv) masked_fill_kernel<T, M><<<grid, 256, 0, s>>>((const T*)a->data.u8, (const M*)mask->data.u8, (T*)c->data.u8, cd[1], cd[2], cd[3], n, \
		as[0], as[1], as[2], as[3], ms[0], ms[1], ms[2], ms[3], cs[0], cs[1], cs[2], cs[3], pv, q)
	if (mtype == CCV_32S)
	{
		const int pi = (int)(p + 0.5f); // util_cpu_ref.c:1467
		if (kind == 0) LAUNCH_MF(float, int, pi); else if (kind == 1) LAUNCH_MF(__nv_bfloat16, int, pi); else LAUNCH_MF(__half, int, pi);
	} else {
		if (kind == 0) LAUNCH_MF(float, float, p); else if (kind == 1) LAUNCH_MF(__nv_bfloat16, float, p); else LAUNCH_MF(__half, float, p);
	}
#undef LAUNCH_MF
	return check("masked fill") ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
int exec_masked_fill_forw(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 2 || output_size < 1 || !inputs[0] || !inputs[1] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	return masked_fill(stream_of(stream_context), cmd.info.blas.a[0], cmd.info.blas.a[1], inputs[0], inputs[1], outputs[0]);
}
// inputs (g, a, mask) -> outputs (h): the gradient passes where the mask did not fill
int exec_masked_fill_back(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 3 || output_size < 1 || !inputs[0] || !inputs[2] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	return masked_fill(stream_of(stream_context), cmd.info.blas.a[0], 0.f, inputs[0], inputs[2], outputs[0]);
}

// ------------------------------------------------------------------------------------------------ random fill
// RANDOM_UNIFORM / RANDOM_NORMAL (rand/gpu/ccv_nnc_rand_{uniform,normal}_gpu_ref.cu:11-66): the reference draws a 32-bit seed from
// the stream context's generator (ccv_nnc_stream_context_genrand_uint32) and fills the tensor from a counter-based Philox
// 4x32-10 stream.  Same contract here: seed from the stream context, Philox 4x32-10 written out below (Salmon et al., SC'11:
// public algorithm) with the element-quad index as the counter, so a fill is reproducible for a given seed and independent of
// the launch shape.  What parameter initialisation (lib/nnc/ccv_cnnp_model_addons.c:996,1158) needs is the distribution, which the
// tests check; the reference's own sample sequence (cuRAND's counter layout) is not part of its API.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], const uint32_t k0, const uint32_t k1)
{
	const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
	const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
	const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
	c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(const uint64_t counter, const uint32_t seed, uint32_t (&out)[4])
{
	uint32_t c[4] = { (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u };
	uint32_t k0 = seed, k1 = 0x5eed5eedu;
#pragma unroll
	for (int r = 0; r < 10; r++)
	{
		philox_round(c, k0, k1);
		k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
	}
	out[0] = c[0], out[1] = c[1], out[2] = c[2], out[3] = c[3];
}
// (0, 1]: 2^-32 (x + 1) rounded into float never returns 0 (the reference's curand_uniform is open at 0, closed at 1)
__device__ __forceinline__ float u01(const uint32_t x) { return fmaxf((float)x * 2.3283064365386963e-10f + 2.3283064365386963e-10f * 0.5f, 1.1754944e-38f); }
template <typename T, int NORMAL>
__global__ void __launch_bounds__(256) random_kernel(T* __restrict__ a, const size_t n, const uint32_t seed, const float p0, const float p1)
{
	const size_t quads = (n + 3) / 4;
	for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x)
	{
		uint32_t r[4];
		philox4x32_10(q, seed, r);
		float v[4];
		if (NORMAL)
		{
			// Box-Muller on two pairs: p0 = std, p1 = mean
#pragma unroll
			for (int j = 0; j < 2; j++)
			{
				const float radius = sqrtf(-2.f * logf(u01(r[2 * j]))), angle = 6.283185307179586f * u01(r[2 * j + 1]);
				v[2 * j] = radius * cosf(angle) * p0 + p1, v[2 * j + 1] = radius * sinf(angle) * p0 + p1;
			}
		} else {
			// p0 = lower, p1 = upper: r u + (1 - r) l, as the reference
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				const float u = u01(r[j]);
				v[j] = u * p1 + (1.f - u) * p0;
			}
		}
#pragma unroll
		for (int j = 0; j < 4; j++)
			if (q * 4 + j < n)
				stf(a + q * 4 + j, v[j]);
	}
}
template <int NORMAL>
int exec_random(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (output_size < 1 || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	ccv_nnc_tensor_t* const a = outputs[0];
	const int kind = kind_of(a);
	if (kind < 0 || !CCV_IS_TENSOR_CONTIGUOUS(a))
		return CCV_NNC_EXEC_INVALID;
	const size_t n = count_of(a);
	if (n == 0)
		return CCV_NNC_EXEC_SUCCESS;
	const uint32_t seed = ccv_nnc_stream_context_genrand_uint32(stream_context);
	cudaStream_t s = stream_of(stream_context);
	const int grid = grid_for((n + 3) / 4, 256);
	const float p0 = cmd.info.blas.a[0], p1 = cmd.info.blas.a[1];
	if (kind == 0)
		random_kernel<float, NORMAL><<<grid, 256, 0, s>>>((float*)a->data.u8, n, seed, p0, p1);
	else if (kind == 1)
		random_kernel<__nv_bfloat16, NORMAL><<<grid, 256, 0, s>>>((__nv_bfloat16*)a->data.u8, n, seed, p0, p1);
	else
		random_kernel<__half, NORMAL><<<grid, 256, 0, s>>>((__half*)a->data.u8, n, seed, p0, p1);
	return check("random fill") ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ dropout
// dropout/ccv_nnc_dropout_cpu_ref.c:16-215: b = mask ? 0 : a / (1 - p), mask[i] = (u_i <= p) as one byte per element in the reserved
// second output (dropout/ccv_nnc_dropout.c:21-44 sizes it in 128-byte lines, never below one byte per element); `entirety` = one
// decision for the whole tensor, kept as an int32 in mask[0].  Backward (g, -, -, -, mask): h = mask ? 0 : g / (1 - p).  The
// uniforms are the Philox stream of the random-fill commands above, keyed by the stream context's seed: as for those, the contract
// is the distribution and the forward / backward consistency of the mask (there is no bit-level parity to define
// against the reference's dSFMT / cuDNN generator states).
template <typename T>
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const T* __restrict__ a, T* __restrict__ b, uint8_t* __restrict__ mask, const size_t n, const uint32_t seed, const float p, const float inv_p, const int entirety)
{
	if (entirety)
	{
		uint32_t r[4];
		philox4x32_10(0, seed, r);
		const int drop = u01(r[0]) <= p;
		if (blockIdx.x == 0 && threadIdx.x == 0)
			*reinterpret_cast<int32_t*>(mask) = drop;
		for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
			stf(b + i, drop ? 0.f : ldf(a + i) * inv_p);
		return;
	}
	const size_t quads = (n + 3) / 4;
	for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x)
	{
		uint32_t r[4];
		philox4x32_10(q, seed, r);
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			const size_t i = q * 4 + j;
			if (i < n)
			{
				const int drop = u01(r[j]) <= p;
				mask[i] = (uint8_t)drop;
				stf(b + i, drop ? 0.f : ldf(a + i) * inv_p);
			}
		}
	}
}
template <typename T>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const T* __restrict__ g, const uint8_t* __restrict__ mask, T* __restrict__ h, const size_t n, const float inv_p, const int entirety)
{
	const int all = entirety ? *reinterpret_cast<const int32_t*>(mask) : 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		stf(h + i, (entirety ? all : mask[i]) ? 0.f : ldf(g + i) * inv_p);
}
size_t bytes_of(const ccv_nnc_tensor_t* const t)
{
	const int dt = CCV_GET_DATA_TYPE(t->info.datatype);
	return count_of(t) * (dt == CCV_64F || dt == CCV_64S ? 8 : dt == CCV_16F || dt == CCV_16BF ? 2 : dt == CCV_8U ? 1 : 4);
}
int exec_dropout_forw(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 1 || output_size < 2 || !inputs[0] || !outputs[0] || !outputs[1])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(inputs[0]), entirety = cmd.info.dropout.entirety;
	if (kind < 0 || !ok_operand(outputs[0], inputs[0], kind) || !CCV_IS_TENSOR_CONTIGUOUS(inputs[0]) || !CCV_IS_TENSOR_CONTIGUOUS(outputs[1]))
		return CCV_NNC_EXEC_INVALID;
	const size_t n = count_of(inputs[0]);
	if (bytes_of(outputs[1]) < (entirety ? sizeof(int32_t) : n))
		return CCV_NNC_EXEC_INVALID;
	if (n == 0)
		return CCV_NNC_EXEC_SUCCESS;
	const float p = cmd.info.dropout.p, inv_p = 1.f / (1.f - p);
	const uint32_t seed = ccv_nnc_stream_context_genrand_uint32(stream_context);
	cudaStream_t s = stream_of(stream_context);
	const int grid = grid_for(entirety ? n : (n + 3) / 4, 256);
	uint8_t* const mask = outputs[1]->data.u8;
	if (kind == 0)
		dropout_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)inputs[0]->data.u8, (float*)outputs[0]->data.u8, mask, n, seed, p, inv_p, entirety);
	else if (kind == 1)
		dropout_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)inputs[0]->data.u8, (__nv_bfloat16*)outputs[0]->data.u8, mask, n, seed, p, inv_p, entirety);
	else
		dropout_fwd_kernel<__half><<<grid, 256, 0, s>>>((const __half*)inputs[0]->data.u8, (__half*)outputs[0]->data.u8, mask, n, seed, p, inv_p, entirety);
	return check("dropout") ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}
int exec_dropout_back(const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context)
{
	if (input_size < 5 || output_size < 1 || !inputs[0] || !inputs[4] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int kind = kind_of(inputs[0]), entirety = cmd.info.dropout.entirety;
	if (kind < 0 || !ok_operand(outputs[0], inputs[0], kind) || !CCV_IS_TENSOR_CONTIGUOUS(inputs[0]) || !CCV_IS_TENSOR_CONTIGUOUS(inputs[4]))
		return CCV_NNC_EXEC_INVALID;
	const size_t n = count_of(inputs[0]);
	if (bytes_of(inputs[4]) < (entirety ? sizeof(int32_t) : n))
		return CCV_NNC_EXEC_INVALID;
	if (n == 0)
		return CCV_NNC_EXEC_SUCCESS;
	const float inv_p = 1.f / (1.f - cmd.info.dropout.p);
	cudaStream_t s = stream_of(stream_context);
	const int grid = grid_for(n, 256);
	const uint8_t* const mask = inputs[4]->data.u8;
	if (kind == 0)
		dropout_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)inputs[0]->data.u8, mask, (float*)outputs[0]->data.u8, n, inv_p, entirety);
	else if (kind == 1)
		dropout_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)inputs[0]->data.u8, mask, (__nv_bfloat16*)outputs[0]->data.u8, n, inv_p, entirety);
	else
		dropout_bwd_kernel<__half><<<grid, 256, 0, s>>>((const __half*)inputs[0]->data.u8, mask, (__half*)outputs[0]->data.u8, n, inv_p, entirety);
	return check("dropout backward") ? CCV_NNC_EXEC_INVALID : CCV_NNC_EXEC_SUCCESS;
}

void fill(ccv_nnc_cmd_backend_registry_t* const registry, const int datatypes, const ccv_nnc_cmd_exec_f exec)
{
	registry->tensor_formats = CCV_TENSOR_FORMAT_NCHW | CCV_TENSOR_FORMAT_NHWC | CCV_TENSOR_FORMAT_CHWN;
	registry->tensor_datatypes = datatypes;
	registry->tensor_memory = CCV_TENSOR_GPU_MEMORY;
	registry->algorithms = 1;
	registry->exec = exec;
	registry->autotune = 0;
	registry->aux = 0;
}

} // namespace

#define F3 (CCV_32F | CCV_16F | CCV_16BF)
#define REGISTER_SM100(cmd) extern "C" void _register_command_ ## cmd ## _backend_CCV_NNC_BACKEND_GPU_SM100(ccv_nnc_cmd_backend_registry_t* const registry)
REGISTER_SM100(CCV_NNC_SIGMOID_FORWARD) { fill(registry, F3, exec_unary_forw<OP_SIGMOID>); }
REGISTER_SM100(CCV_NNC_SIGMOID_BACKWARD) { fill(registry, F3, exec_unary_back<OP_SIGMOID_B, 2>); }
REGISTER_SM100(CCV_NNC_TANH_FORWARD) { fill(registry, F3, exec_unary_forw<OP_TANH>); }
REGISTER_SM100(CCV_NNC_TANH_BACKWARD) { fill(registry, F3, exec_unary_back<OP_TANH_B, 2>); }
REGISTER_SM100(CCV_NNC_LEAKY_RELU_FORWARD) { fill(registry, F3, exec_unary_forw<OP_LEAKY>); }
REGISTER_SM100(CCV_NNC_LEAKY_RELU_BACKWARD) { fill(registry, F3, exec_unary_back<OP_LEAKY_B, 2>); }
REGISTER_SM100(CCV_NNC_EWEXP_FORWARD) { fill(registry, F3, exec_unary_forw<OP_EXP>); }
REGISTER_SM100(CCV_NNC_EWEXP_BACKWARD) { fill(registry, F3, exec_unary_back<OP_EXP_B, 2>); }
REGISTER_SM100(CCV_NNC_EWLOG_FORWARD) { fill(registry, F3, exec_unary_forw<OP_LOG>); }
REGISTER_SM100(CCV_NNC_EWLOG_BACKWARD) { fill(registry, F3, exec_unary_back<OP_LOG_B, 1>); }
REGISTER_SM100(CCV_NNC_EWSQRT_FORWARD) { fill(registry, F3, exec_unary_forw<OP_SQRT>); }
REGISTER_SM100(CCV_NNC_EWSQRT_BACKWARD) { fill(registry, F3, exec_unary_back<OP_SQRT_B, 2>); }
REGISTER_SM100(CCV_NNC_CLAMP_FORWARD) { fill(registry, F3, exec_unary_forw<OP_CLAMP>); }
REGISTER_SM100(CCV_NNC_CLAMP_BACKWARD) { fill(registry, F3, exec_unary_back<OP_CLAMP_B, 2>); }
REGISTER_SM100(CCV_NNC_EWDIV_FORWARD) { fill(registry, F3, exec_ewdiv_forw); }
REGISTER_SM100(CCV_NNC_EWDIV_BACKWARD) { fill(registry, F3, exec_ewdiv_back); }
REGISTER_SM100(CCV_NNC_REDUCE_SUM_FORWARD) { fill(registry, F3, exec_reduce_forw<RED_SUM>); }
REGISTER_SM100(CCV_NNC_REDUCE_SUM_BACKWARD) { fill(registry, F3, exec_reduce_back<RED_SUM>); }
REGISTER_SM100(CCV_NNC_REDUCE_MEAN_FORWARD) { fill(registry, F3, exec_reduce_forw<RED_MEAN>); }
REGISTER_SM100(CCV_NNC_REDUCE_MEAN_BACKWARD) { fill(registry, F3, exec_reduce_back<RED_MEAN>); }
REGISTER_SM100(CCV_NNC_REDUCE_MAX_FORWARD) { fill(registry, F3, exec_reduce_forw<RED_MAX>); }
REGISTER_SM100(CCV_NNC_REDUCE_MAX_BACKWARD) { fill(registry, F3, exec_reduce_back<RED_MAX>); }
REGISTER_SM100(CCV_NNC_REDUCE_MIN_FORWARD) { fill(registry, F3, exec_reduce_forw<RED_MIN>); }
REGISTER_SM100(CCV_NNC_REDUCE_MIN_BACKWARD) { fill(registry, F3, exec_reduce_back<RED_MIN>); }
REGISTER_SM100(CCV_NNC_REDUCE_NORM2_FORWARD) { fill(registry, F3, exec_reduce_forw<RED_NORM2>); }
REGISTER_SM100(CCV_NNC_REDUCE_NORM2_BACKWARD) { fill(registry, F3, exec_reduce_back<RED_NORM2>); }
REGISTER_SM100(CCV_NNC_MASKED_FILL_FORWARD) { fill(registry, F3 | CCV_32S, exec_masked_fill_forw); }
REGISTER_SM100(CCV_NNC_MASKED_FILL_BACKWARD) { fill(registry, F3 | CCV_32S, exec_masked_fill_back); }
REGISTER_SM100(CCV_NNC_RANDOM_UNIFORM_FORWARD) { fill(registry, F3, exec_random<0>); }
REGISTER_SM100(CCV_NNC_RANDOM_UNIFORM_BACKWARD) { fill(registry, F3, exec_random<0>); }
REGISTER_SM100(CCV_NNC_RANDOM_NORMAL_FORWARD) { fill(registry, F3, exec_random<1>); }
REGISTER_SM100(CCV_NNC_RANDOM_NORMAL_BACKWARD) { fill(registry, F3, exec_random<1>); }
REGISTER_SM100(CCV_NNC_DROPOUT_FORWARD) { fill(registry, F3, exec_dropout_forw); }
REGISTER_SM100(CCV_NNC_DROPOUT_BACKWARD) { fill(registry, F3, exec_dropout_back); }
