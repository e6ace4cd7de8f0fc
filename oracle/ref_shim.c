/* oracle/ref_shim.c -- TEST INFRASTRUCTURE (part of the parity oracle; never linked into the product).
 *
 * Pointer-argument entry points over the UNMODIFIED reference library so that Python/ctypes (tests,
 * bench.py's cpu_baseline leg) can drive CCV_NNC_BACKEND_CPU_REF without passing ccv's large structs by
 * value (ccv_nnc_cmd_t is 152 bytes, ccv_nnc_hint_t 144; lib/nnc/ccv_nnc.h:279-306).  Compiled against the
 * reference's own headers in /root/reference/lib by oracle/Makefile into oracle/_ref/libccv_ref.so.
 */
#include "ccv.h"
#include "nnc/ccv_nnc.h"
#include "nnc/ccv_nnc_easy.h"
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void ref_nnc_init(void)
{
	ccv_nnc_init(); /* lib/nnc/ccv_nnc_cmd.c:27 */
}

int ref_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* explicit thread count for the timed CPU_REF runs: torchrun exports OMP_NUM_THREADS=1 to its workers, which silently turned
 * the reference arm into a one-thread measurement */
void ref_set_num_threads(const int n)
{
#ifdef _OPENMP
	if (n > 0)
		omp_set_num_threads(n);
#else
	(void)n;
#endif
}

/* sizes the boundary header must reproduce */
void ref_abi_sizes(int* const out)
{
	out[0] = (int)sizeof(ccv_nnc_cmd_t);
	out[1] = (int)sizeof(ccv_nnc_hint_t);
	out[2] = (int)sizeof(ccv_nnc_cmd_param_t);
	out[3] = (int)sizeof(ccv_nnc_tensor_param_t);
	out[4] = (int)sizeof(ccv_nnc_tensor_t);
	out[5] = (int)sizeof(ccv_nnc_tensor_view_t);
	out[6] = (int)offsetof(ccv_nnc_cmd_t, info);
	out[7] = (int)offsetof(ccv_nnc_tensor_t, info);
	out[8] = (int)offsetof(ccv_nnc_tensor_view_t, stride);
	out[9] = (int)offsetof(ccv_nnc_tensor_t, data);
}

ccv_nnc_tensor_t* ref_tensor_new(void* const ptr, const ccv_nnc_tensor_param_t* const params)
{
	return ccv_nnc_tensor_new(ptr, *params, 0); /* lib/nnc/ccv_nnc_tensor.c */
}

void ref_tensor_free(ccv_nnc_tensor_t* const t)
{
	ccv_nnc_tensor_free(t);
}

ccv_nnc_tensor_view_t* ref_tensor_view_new(const ccv_nnc_tensor_t* const t, const ccv_nnc_tensor_param_t* const params, const int* const ofs, const int* const stride)
{
	return ccv_nnc_tensor_view_new(t, *params, ofs, stride);
}

void ref_tensor_view_free(ccv_nnc_tensor_view_t* const tv)
{
	ccv_nnc_tensor_view_free(tv);
}

void ref_hint_auto(const ccv_nnc_cmd_param_t* const info, const ccv_nnc_tensor_param_t* const a, const ccv_nnc_tensor_param_t* const b, ccv_nnc_hint_t* const hint)
{
	*hint = ccv_nnc_hint_auto(*info, *a, *b); /* lib/nnc/ccv_nnc_cmd.c */
}

void ref_hint_tensor_auto(const uint32_t cmd, const ccv_nnc_cmd_param_t* const info, const ccv_nnc_tensor_param_t* const inputs, const int input_size, const ccv_nnc_hint_t* const hint, ccv_nnc_tensor_param_t* const outputs, const int output_size)
{
	ccv_nnc_cmd_t c = ccv_nnc_cmd(cmd, 0, *info, 0);
	ccv_nnc_hint_tensor_auto(c, inputs, input_size, *hint, outputs, output_size);
}

/* ccv_nnc_cmd_exec (lib/nnc/ccv_nnc_cmd.c:651) with cmd.backend forced (CPU_REF for the oracle). */
int ref_cmd_exec(const uint32_t cmd, const uint32_t backend, const int algorithm, const ccv_nnc_cmd_param_t* const info, const ccv_nnc_hint_t* const hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size)
{
	ccv_nnc_cmd_t c = ccv_nnc_cmd(cmd, 0, *info, 0);
	c.backend = backend;
	c.algorithm = algorithm;
	return ccv_nnc_cmd_exec(c, *hint, flags, inputs, input_size, outputs, output_size, 0);
}

/* wall-clock timing helper used by bench.py's cpu_baseline: runs the command `reps` times, returns best seconds */
double ref_cmd_time(const uint32_t cmd, const uint32_t backend, const ccv_nnc_cmd_param_t* const info, const ccv_nnc_hint_t* const hint, const int flags, ccv_nnc_tensor_t* const* const inputs, const int input_size, ccv_nnc_tensor_t* const* const outputs, const int output_size, const int reps)
{
	ccv_nnc_cmd_t c = ccv_nnc_cmd(cmd, 0, *info, 0);
	c.backend = backend;
	double best = 1e30;
	int i;
	for (i = 0; i < reps; i++)
	{
		const uint64_t t0 = ccv_nnc_cmd_mono_time();
		const int status = ccv_nnc_cmd_exec(c, *hint, flags, inputs, input_size, outputs, output_size, 0);
		const uint64_t t1 = ccv_nnc_cmd_mono_time();
		if (status != 0)
			return -1;
		const double s = (double)(t1 - t0) * 1e-9;
		if (s < best)
			best = s;
	}
	return best;
}

/* f32 -> f16 / f16 -> f32 helpers of the reference (lib/ccv_util.c:1434-1440: truncating table method) */
void ref_float_to_half(const float* const f, uint16_t* const h, const size_t n)
{
	ccv_float_to_half_precision((float*)f, h, n);
}

void ref_half_to_float(const uint16_t* const h, float* const f, const size_t n)
{
	ccv_half_precision_to_float((uint16_t*)h, f, n);
}
