/* A plain C translation unit (gcc, -std=c99) compiled against include/ccv_nnc_sm100.h and linked with
 * libccv_nnc_sm100.so: the drop-in boundary exercised the way the reference's own tests exercise libccv -- by-value
 * ccv_nnc_cmd_t / ccv_nnc_hint_t through ccv_nnc_cmd_exec, tensors from ccv_nnc_tensor_new, host <-> device movement with
 * CMD_DATA_TRANSFER (the protocol of test/int/nnc/cublas.tests.c:1155-1750) -- on the literal known-answer GEMM cases of
 * test/unit/nnc/gemm.tests.c:13-200 with the backend set to CCV_NNC_BACKEND_GPU_SM100.
 *
 *   gemm_literals            run every case on device 0, exit code = number of failures
 *   gemm_literals --no-gpu   only the host-side checks (struct sizes, registration, refusals): what a CPU box can run
 */
#include "ccv_nnc_sm100.h"
#include <math.h>
#include <stdio.h>
#include <string.h>

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static ccv_nnc_tensor_param_t params(const int memory, const int d0, const int d1, const int d2)
{
	ccv_nnc_tensor_param_t p;
	memset(&p, 0, sizeof(p));
	p.type = memory; /* device 0 */
	p.format = CCV_TENSOR_FORMAT_NHWC;
	p.datatype = CCV_32F;
	p.dim[0] = d0, p.dim[1] = d1, p.dim[2] = d2;
	return p;
}

static ccv_nnc_cmd_t gemm_cmd(const uint32_t which, const int ta0, const int ta1, const int tb0, const int tb1)
{
	ccv_nnc_cmd_param_t info;
	memset(&info, 0, sizeof(info));
	info.size.dim[0] = info.size.dim[1] = info.size.dim[2] = 1; /* what CMD_GEMM_FORWARD() fills in (ccv_nnc_cmd_easy.h) */
	info.blas.a[0] = info.blas.a[1] = 1;
	info.blas.transpose_a[0] = ta0, info.blas.transpose_a[1] = ta1;
	info.blas.transpose_b[0] = tb0, info.blas.transpose_b[1] = tb1;
	ccv_nnc_cmd_t cmd = ccv_nnc_cmd(which, 0, info, 0);
	cmd.backend = CCV_NNC_BACKEND_GPU_SM100;
	return cmd;
}

static ccv_nnc_cmd_t transfer_cmd(void)
{
	ccv_nnc_cmd_param_t info;
	memset(&info, 0, sizeof(info));
	return ccv_nnc_cmd(CCV_NNC_DATA_TRANSFER_FORWARD, 0, info, 0);
}

static ccv_nnc_hint_t no_hint(void)
{
	ccv_nnc_hint_t h;
	memset(&h, 0, sizeof(h));
	return h;
}

/* host array -> fresh device tensor, through the command API */
static ccv_nnc_tensor_t* to_device(const float* const host, const int d0, const int d1, const int d2)
{
	ccv_nnc_tensor_t* const h = ccv_nnc_tensor_new(host, params(CCV_TENSOR_CPU_MEMORY, d0, d1, d2), 0);
	ccv_nnc_tensor_t* const g = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, d0, d1, d2), 0);
	ccv_nnc_tensor_t* ins[1] = { h };
	ccv_nnc_tensor_t* outs[1] = { g };
	CHECK(ccv_nnc_cmd_exec(transfer_cmd(), no_hint(), 0, ins, 1, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "host -> device transfer");
	ccv_nnc_tensor_free(h);
	return g;
}

static void expect(ccv_nnc_tensor_t* const g, const float* const want, const int count, const int d0, const int d1, const int d2, const char* const what)
{
	float got[64];
	ccv_nnc_tensor_t* const h = ccv_nnc_tensor_new(got, params(CCV_TENSOR_CPU_MEMORY, d0, d1, d2), 0);
	ccv_nnc_tensor_t* ins[1] = { g };
	ccv_nnc_tensor_t* outs[1] = { h };
	int i;
	memset(got, 0, sizeof(got));
	CHECK(ccv_nnc_cmd_exec(transfer_cmd(), no_hint(), 0, ins, 1, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "device -> host transfer");
	for (i = 0; i < count; i++)
		CHECK(fabsf(got[i] - want[i]) <= 1e-5f * fmaxf(1.f, fabsf(want[i])), "%s: element %d is %g, expected %g", what, i, got[i], want[i]);
	ccv_nnc_tensor_free(h);
}

static void host_checks(void)
{
	ccv_nnc_cmd_backend_registry_t r;
	/* by-value ABI sizes of the reference (lib/nnc/ccv_nnc.h:111-306) */
	CHECK(sizeof(ccv_nnc_cmd_t) == 152, "sizeof(ccv_nnc_cmd_t) = %zu", sizeof(ccv_nnc_cmd_t));
	CHECK(sizeof(ccv_nnc_cmd_param_t) == 120, "sizeof(ccv_nnc_cmd_param_t) = %zu", sizeof(ccv_nnc_cmd_param_t));
	CHECK(sizeof(ccv_nnc_hint_t) == 144, "sizeof(ccv_nnc_hint_t) = %zu", sizeof(ccv_nnc_hint_t));
	/* the registration symbol REGISTER_COMMAND_BACKEND expands to (lib/nnc/ccv_nnc_internal.h:196-202) fills the record */
	memset(&r, 0, sizeof(r));
	_register_command_CCV_NNC_GEMM_FORWARD_backend_CCV_NNC_BACKEND_GPU_SM100(&r);
	CHECK(r.exec != 0 && (r.tensor_datatypes & CCV_32F) && (r.tensor_memory & CCV_TENSOR_GPU_MEMORY) && r.algorithms >= 1, "GEMM_FORWARD registry record");
	ccv_nnc_init();
	CHECK(ccv_nnc_cmd_ok(CCV_NNC_GEMM_FORWARD, CCV_NNC_BACKEND_GPU_SM100) == 1, "ccv_nnc_cmd_ok(GEMM_FORWARD, GPU_SM100)");
	{
		/* host tensors handed to a device backend: refused, never computed on the CPU */
		float ap[4] = { 1, 2, 3, 4 }, cp[4];
		ccv_nnc_tensor_t* const a = ccv_nnc_tensor_new(ap, params(CCV_TENSOR_CPU_MEMORY, 2, 2, 0), 0);
		ccv_nnc_tensor_t* const c = ccv_nnc_tensor_new(cp, params(CCV_TENSOR_CPU_MEMORY, 2, 2, 0), 0);
		ccv_nnc_tensor_t* ins[2] = { a, a };
		ccv_nnc_tensor_t* outs[1] = { c };
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_FORWARD, 0, 0, 0, 0), no_hint(), 0, ins, 2, outs, 1, 0) == CCV_NNC_EXEC_NO_KERNEL, "host tensors must be refused with NO_KERNEL");
		ccv_nnc_tensor_free(a);
		ccv_nnc_tensor_free(c);
	}
}

int main(int argc, char** argv)
{
	host_checks();
	if (argc > 1 && strcmp(argv[1], "--no-gpu") == 0)
	{
		printf("gemm_literals (host checks only): %d failure(s)\n", failures);
		return failures;
	}
	if (ccv_nnc_device_count(CCV_STREAM_CONTEXT_GPU) <= 0)
	{
		fprintf(stderr, "no CUDA device: this backend has no CPU path\n");
		return 100;
	}
	{
		const float ap[] = { 1, 2, 3, 4, 5, 6, 7, 8 };
		const float bp[] = { 7, 8, 9, 10, 11, 12 };
		const float btp[] = { 7, 10, 8, 11, 9, 12 };
		const float atp[] = { 1, 3, 5, 7, 2, 4, 6, 8 };
		const float biasp[] = { -1, 0, 1 };
		const float want[] = { 1 * 7 + 2 * 10, 1 * 8 + 2 * 11, 1 * 9 + 2 * 12, 3 * 7 + 4 * 10, 3 * 8 + 4 * 11, 3 * 9 + 4 * 12, 5 * 7 + 6 * 10, 5 * 8 + 6 * 11, 5 * 9 + 6 * 12, 7 * 7 + 8 * 10, 7 * 8 + 8 * 11, 7 * 9 + 8 * 12 };
		float want_bias[12];
		int i;
		ccv_nnc_tensor_t *a, *b, *c, *bias;
		ccv_nnc_tensor_t* ins[3];
		ccv_nnc_tensor_t* outs[3];
		for (i = 0; i < 12; i++)
			want_bias[i] = want[i] + biasp[i % 3];
		/* gemm.tests.c:13-40: [4x2] * [2x3] */
		a = to_device(ap, 4, 2, 0), b = to_device(bp, 2, 3, 0), c = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, 4, 3, 0), 0);
		ins[0] = a, ins[1] = b, outs[0] = c;
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_FORWARD, 0, 0, 0, 0), no_hint(), 0, ins, 2, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "GEMM NN");
		expect(c, want, 12, 4, 3, 0, "GEMM NN");
		/* :117-150: + bias [-1, 0, 1] */
		bias = to_device(biasp, 3, 0, 0);
		ins[2] = bias;
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_FORWARD, 0, 0, 0, 0), no_hint(), 0, ins, 3, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "GEMM NN + bias");
		expect(c, want_bias, 12, 4, 3, 0, "GEMM NN + bias");
		ccv_nnc_tensor_free(b);
		/* :66-93: b transposed, TRANSPOSE(0, 1) */
		b = to_device(btp, 3, 2, 0);
		ins[1] = b;
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_FORWARD, 0, 0, 0, 1), no_hint(), 0, ins, 2, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "GEMM NT");
		expect(c, want, 12, 4, 3, 0, "GEMM NT");
		ccv_nnc_tensor_free(a);
		ccv_nnc_tensor_free(c);
		/* :95-115: a [1, 2, 4] transposed over axes (1, 2), b transposed: batched form, c [1, 4, 3] */
		a = to_device(atp, 1, 2, 4), c = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, 1, 4, 3), 0);
		ins[0] = a, outs[0] = c;
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_FORWARD, 1, 2, 0, 1), no_hint(), 0, ins, 2, outs, 1, 0) == CCV_NNC_EXEC_SUCCESS, "GEMM TT");
		expect(c, want, 12, 1, 4, 3, "GEMM TT");
		ccv_nnc_tensor_free(a);
		ccv_nnc_tensor_free(b);
		ccv_nnc_tensor_free(c);
		ccv_nnc_tensor_free(bias);
	}
	{
		/* backward (gemm.tests.c, "backward gemm with no transpose"): g [4x3], a [4x2], w [2x3] -> h = g w^T, dw = a^T g, dbias = sum_rows g */
		const float gp[] = { 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12 };
		const float ap[] = { 13, 14, 15, 16, 17, 18, 19, 20 };
		const float wp[] = { 21, 22, 23, 24, 25, 26 };
		float h_want[8], dw_want[6], db_want[3];
		int i, j, k;
		ccv_nnc_tensor_t *g, *a, *w, *h, *dw, *db;
		ccv_nnc_tensor_t* ins[3];
		ccv_nnc_tensor_t* outs[3];
		for (i = 0; i < 4; i++)
			for (k = 0; k < 2; k++)
			{
				h_want[i * 2 + k] = 0;
				for (j = 0; j < 3; j++)
					h_want[i * 2 + k] += gp[i * 3 + j] * wp[k * 3 + j];
			}
		for (k = 0; k < 2; k++)
			for (j = 0; j < 3; j++)
			{
				dw_want[k * 3 + j] = 0;
				for (i = 0; i < 4; i++)
					dw_want[k * 3 + j] += ap[i * 2 + k] * gp[i * 3 + j];
			}
		for (j = 0; j < 3; j++)
			db_want[j] = gp[j] + gp[3 + j] + gp[6 + j] + gp[9 + j];
		g = to_device(gp, 4, 3, 0), a = to_device(ap, 4, 2, 0), w = to_device(wp, 2, 3, 0);
		h = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, 4, 2, 0), 0), dw = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, 2, 3, 0), 0), db = ccv_nnc_tensor_new(0, params(CCV_TENSOR_GPU_MEMORY, 3, 0, 0), 0);
		ins[0] = g, ins[1] = a, ins[2] = w, outs[0] = h, outs[1] = dw, outs[2] = db;
		CHECK(ccv_nnc_cmd_exec(gemm_cmd(CCV_NNC_GEMM_BACKWARD, 0, 0, 0, 0), no_hint(), 0, ins, 3, outs, 3, 0) == CCV_NNC_EXEC_SUCCESS, "GEMM backward");
		expect(h, h_want, 8, 4, 2, 0, "GEMM backward h");
		expect(dw, dw_want, 6, 2, 3, 0, "GEMM backward dw");
		expect(db, db_want, 3, 3, 0, 0, "GEMM backward dbias");
		ccv_nnc_tensor_free(g), ccv_nnc_tensor_free(a), ccv_nnc_tensor_free(w), ccv_nnc_tensor_free(h), ccv_nnc_tensor_free(dw), ccv_nnc_tensor_free(db);
	}
	printf("gemm_literals: %d failure(s)\n", failures);
	return failures;
}
