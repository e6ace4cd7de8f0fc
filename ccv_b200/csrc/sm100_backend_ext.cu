// sm100_backend_ext.cu -- command layer for attention, layer / rms norm, upsample and the gradient allreduce.
#include "../../include/ccv_nnc_sm100.h"
#include "sm100_contract.h"
#include "sm100_ew.h"
#include <cuda_runtime.h>
#include <string.h>

using namespace sm100;

namespace {

inline cudaStream_t stream_of(ccv_nnc_stream_context_t* const stream_context) { return (cudaStream_t)ccv_nnc_stream_context_get_stream(stream_context); }

int nd_of(const ccv_nnc_tensor_t* const t)
{
	int i;
	for (i = 0; i < CCV_NNC_MAX_DIM_ALLOC; i++)
		if (t->info.dim[i] == 0)
			return i;
	return CCV_NNC_MAX_DIM_ALLOC;
}

size_t count_of(const ccv_nnc_tensor_t* const t)
{
	size_t c = 1;
	for (int i = 0; i < nd_of(t); i++)
		c *= (size_t)t->info.dim[i];
	return c;
}

bool packed_f32(const ccv_nnc_tensor_t* const t)
{
	return CCV_GET_DATA_TYPE(t->info.datatype) == CCV_32F && CCV_IS_TENSOR_CONTIGUOUS(t);
}

// the normalised axes must be the trailing ones: x is then [rows, inner]
bool rows_inner(const ccv_nnc_tensor_t* const x, const int* const axis, const int axis_count, int& rows, int& inner)
{
	if (!packed_f32(x))
		return false;
	const int nd = nd_of(x);
	if (axis_count < 1 || axis_count > nd)
		return false;
	bool reduced[CCV_NNC_MAX_DIM_ALLOC] = { false };
	for (int i = 0; i < axis_count; i++)
	{
		if (axis[i] < 0 || axis[i] >= nd)
			return false;
		reduced[axis[i]] = true;
	}
	long long r = 1, in = 1;
	bool seen_reduced = false;
	for (int i = 0; i < nd; i++)
	{
		if (reduced[i])
			seen_reduced = true, in *= x->info.dim[i];
		else {
			if (seen_reduced && x->info.dim[i] != 1)
				return false; // a kept axis after a reduced one: not a trailing reduction
			r *= x->info.dim[i];
		}
	}
	if (r > 0x7fffffff || in > 0x7fffffff)
		return false;
	rows = (int)r, inner = (int)in;
	return true;
}

} // namespace

#define SM100_EXEC_ARGS const ccv_nnc_cmd_t cmd, const ccv_nnc_hint_t hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, ccv_nnc_stream_context_t* const stream_context

extern "C" {

int ccv_nnc_sm100_exec_sdpa_forw(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
int ccv_nnc_sm100_exec_sdpa_back(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }
// ---- layer norm / rms norm: statistics over the trailing axes (saved_mean dims [d0, .., 1, .., 1]) -----------------
// norm/ccv_nnc_layer_norm_cpu_ref.c:16-190: inputs (x, [scale, bias]) -> outputs (y, [saved_mean, saved_inv_std])
int ccv_nnc_sm100_exec_lnorm_forw(SM100_EXEC_ARGS)
{
	if (input_size < 1 || output_size < 1 || !inputs[0] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	const int affine = cmd.info.lnorm.elementwise_affine;
	if (affine && (input_size < 3 || !inputs[1] || !inputs[2]))
		return CCV_NNC_EXEC_INVALID;
	int rows, inner;
	if (!rows_inner(inputs[0], cmd.info.lnorm.axis, cmd.info.lnorm.count, rows, inner) || !packed_f32(outputs[0]) || count_of(outputs[0]) != (size_t)rows * inner)
		return CCV_NNC_EXEC_INVALID;
	if (affine && (!packed_f32(inputs[1]) || !packed_f32(inputs[2]) || count_of(inputs[1]) != (size_t)inner || count_of(inputs[2]) != (size_t)inner))
		return CCV_NNC_EXEC_INVALID;
	float* sm = output_size > 1 && outputs[1] ? outputs[1]->data.f32 : 0;
	float* sis = output_size > 2 && outputs[2] ? outputs[2]->data.f32 : 0;
	if ((sm && count_of(outputs[1]) != (size_t)rows) || (sis && count_of(outputs[2]) != (size_t)rows))
		return CCV_NNC_EXEC_INVALID;
	if (layer_norm_fwd_f32(stream_of(stream_context), inputs[0]->data.f32, affine ? inputs[1]->data.f32 : 0, affine ? inputs[2]->data.f32 : 0, outputs[0]->data.f32, sm, sis, rows, inner, cmd.info.lnorm.epsilon))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}

// norm/ccv_nnc_layer_norm_cpu_ref.c:192-420: inputs[0] = g, [3] = x, [4] = scale (affine), [7 | 5] = saved_mean, [8 | 6] = saved_inv_std
int ccv_nnc_sm100_exec_lnorm_back(SM100_EXEC_ARGS)
{
	const int affine = cmd.info.lnorm.elementwise_affine;
	const int mi = affine ? 7 : 5, si = affine ? 8 : 6;
	if (input_size <= si || output_size < 1 || !inputs[0] || !inputs[3] || !inputs[mi] || !inputs[si] || (affine && !inputs[4]))
		return CCV_NNC_EXEC_INVALID;
	int rows, inner;
	if (!rows_inner(inputs[3], cmd.info.lnorm.axis, cmd.info.lnorm.count, rows, inner) || !packed_f32(inputs[0]) || count_of(inputs[0]) != (size_t)rows * inner)
		return CCV_NNC_EXEC_INVALID;
	if (count_of(inputs[mi]) != (size_t)rows || count_of(inputs[si]) != (size_t)rows)
		return CCV_NNC_EXEC_INVALID;
	float* h = outputs[0] ? outputs[0]->data.f32 : 0;
	float* ds = output_size > 1 && outputs[1] ? outputs[1]->data.f32 : 0;
	float* db = output_size > 2 && outputs[2] ? outputs[2]->data.f32 : 0;
	if (layer_norm_bwd_f32(stream_of(stream_context), inputs[0]->data.f32, inputs[3]->data.f32, affine ? inputs[4]->data.f32 : 0, inputs[mi]->data.f32, inputs[si]->data.f32, h, ds, db, rows, inner, 0))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}

// norm/ccv_nnc_rmsnorm_cpu_ref.c:16-130: inputs (x, scale) -> outputs (y, saved_inv_std)
int ccv_nnc_sm100_exec_rmsnorm_forw(SM100_EXEC_ARGS)
{
	if (input_size < 2 || output_size < 1 || !inputs[0] || !inputs[1] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	int rows, inner;
	if (!rows_inner(inputs[0], cmd.info.rmsnorm.axis, cmd.info.rmsnorm.count, rows, inner) || !packed_f32(outputs[0]) || !packed_f32(inputs[1]) || count_of(inputs[1]) != (size_t)inner)
		return CCV_NNC_EXEC_INVALID;
	float* sis = output_size > 1 && outputs[1] ? outputs[1]->data.f32 : 0;
	if (sis && count_of(outputs[1]) != (size_t)rows)
		return CCV_NNC_EXEC_INVALID;
	if (rmsnorm_fwd_f32(stream_of(stream_context), inputs[0]->data.f32, inputs[1]->data.f32, outputs[0]->data.f32, sis, rows, inner, cmd.info.rmsnorm.epsilon))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}

// norm/ccv_nnc_rmsnorm_cpu_ref.c:132-330: inputs[0] = g, [2] = x, [3] = scale, [5] = saved_inv_std -> (h, dscale)
int ccv_nnc_sm100_exec_rmsnorm_back(SM100_EXEC_ARGS)
{
	if (input_size < 6 || output_size < 1 || !inputs[0] || !inputs[2] || !inputs[3] || !inputs[5])
		return CCV_NNC_EXEC_INVALID;
	int rows, inner;
	if (!rows_inner(inputs[2], cmd.info.rmsnorm.axis, cmd.info.rmsnorm.count, rows, inner) || !packed_f32(inputs[0]) || count_of(inputs[5]) != (size_t)rows || count_of(inputs[3]) != (size_t)inner)
		return CCV_NNC_EXEC_INVALID;
	float* h = outputs[0] ? outputs[0]->data.f32 : 0;
	float* ds = output_size > 1 && outputs[1] ? outputs[1]->data.f32 : 0;
	if (rmsnorm_bwd_f32(stream_of(stream_context), inputs[0]->data.f32, inputs[2]->data.f32, inputs[3]->data.f32, inputs[5]->data.f32, h, ds, rows, inner, 0))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}

// upsample/ccv_nnc_upsample_cpu_ref.c:16-509: nearest / bilinear, align_corners, NHWC and NCHW
static bool upsample_geom(const ccv_nnc_tensor_t* const a, const ccv_nnc_tensor_t* const b, int& N, int& H, int& W, int& C, int& OH, int& OW, int& nchw)
{
	if (!packed_f32(a) || !packed_f32(b) || a->info.format != b->info.format)
		return false;
	int ad[4], bd[4];
	const int and_ = nd_of(a), bnd = nd_of(b);
	if (and_ > 4 || bnd != and_ || and_ < 3)
		return false;
	for (int i = 0; i < 4; i++)
		ad[i] = i < 4 - and_ ? 1 : a->info.dim[i - (4 - and_)], bd[i] = i < 4 - bnd ? 1 : b->info.dim[i - (4 - bnd)];
	nchw = a->info.format == CCV_TENSOR_FORMAT_NCHW;
	N = ad[0];
	if (nchw)
		C = ad[1], H = ad[2], W = ad[3], OH = bd[2], OW = bd[3];
	else
		H = ad[1], W = ad[2], C = ad[3], OH = bd[1], OW = bd[2];
	return bd[0] == N && (nchw ? bd[1] : bd[3]) == C && OH >= H && OW >= W;
}

int ccv_nnc_sm100_exec_upsample_forw(SM100_EXEC_ARGS)
{
	if (input_size < 1 || output_size < 1 || !inputs[0] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	int N, H, W, C, OH, OW, nchw;
	if (!upsample_geom(inputs[0], outputs[0], N, H, W, C, OH, OW, nchw))
		return CCV_NNC_EXEC_INVALID;
	if (upsample_fwd_f32(stream_of(stream_context), inputs[0]->data.f32, outputs[0]->data.f32, N, H, W, C, OH, OW, cmd.info.upsample.type, cmd.info.upsample.align_corners, nchw))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}

int ccv_nnc_sm100_exec_upsample_back(SM100_EXEC_ARGS)
{
	if (input_size < 1 || output_size < 1 || !inputs[0] || !outputs[0])
		return CCV_NNC_EXEC_INVALID;
	int N, H, W, C, OH, OW, nchw;
	if (!upsample_geom(outputs[0], inputs[0], N, H, W, C, OH, OW, nchw))
		return CCV_NNC_EXEC_INVALID;
	if (upsample_bwd_f32(stream_of(stream_context), inputs[0]->data.f32, outputs[0]->data.f32, N, H, W, C, OH, OW, cmd.info.upsample.type, cmd.info.upsample.align_corners, nchw))
		return CCV_NNC_EXEC_INVALID;
	return CCV_NNC_EXEC_SUCCESS;
}
int ccv_nnc_sm100_exec_allreduce(SM100_EXEC_ARGS) { return CCV_NNC_EXEC_NO_KERNEL; }

}
