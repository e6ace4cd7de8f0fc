// sm100_fmha.cuh -- helpers shared by the 16-bit flash-attention kernels (sm100_fmha.cu forward, sm100_fmha_bwd.cu backward):
// 16-bit packing, the sm_100 packed fp32 instructions, and the 4-D TMA tensor map over a [B, S, H, D] tensor.
#pragma once
#include "sm100_ptx.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>
#include <stdlib.h>

namespace sm100 {
namespace {

__device__ __forceinline__ uint32_t pack2(const float a, const float b, const int is_bf16)
{
	if (is_bf16)
	{
		const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
		return *reinterpret_cast<const uint32_t*>(&v);
	}
	const __half2 v = __floats2half2_rn(a, b);
	return *reinterpret_cast<const uint32_t*>(&v);
}

// sm_100 packed / 3-input fp32 instructions: FMNMX3 halves the row-max pass, FFMA2 halves the scale-and-shift in front of the exp2
// and the running-output update (the softmax warps are issue-limited: profiles/r01_ncu_fmha_bf16_config5.txt)
__device__ __forceinline__ float max3(const float a, const float b, const float c)
{
	float d;
	asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ void fma2(float& d0, float& d1, const float a0, const float a1, const float b0, const float b1, const float c0, const float c1)
{
	asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\tfma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
		: "=f"(d0), "=f"(d1)
		: "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

typedef CUresult (*encode_tiled_f)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_tiled_f& encode_fn()
{
	static encode_tiled_f fn = 0;
	return fn;
}

inline bool encode_init()
{
	static std::once_flag once;
	std::call_once(once, []() {
		void* fn = 0;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
			encode_fn() = (encode_tiled_f)fn;
	});
	return encode_fn() != 0;
}

// 16-bit [B, S, H, D] tensor with element strides (sb, ss, sh), D contiguous: box = {64 d, 1 head, box_rows rows, 1 batch}
inline bool make_map_bshd(CUtensorMap* map, const void* ptr, int B, int S, int H, int D, long long sb, long long ss, long long sh, int is_bf16, int box_rows)
{
	if ((((uintptr_t)ptr) & 15) || ((sb * 2) & 15) || ((ss * 2) & 15) || ((sh * 2) & 15))
		return false;
	// a stride of 0 is not encodable; extents of 1 never advance, any positive multiple of 16 bytes will do
	if (H == 1 && sh <= 0)
		sh = D;
	if (B == 1 && sb <= 0)
		sb = (long long)S * ss;
	cuuint64_t dims[4] = { (cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B };
	cuuint64_t strides[3] = { (cuuint64_t)sh * 2, (cuuint64_t)ss * 2, (cuuint64_t)sb * 2 };
	cuuint32_t box[4] = { 64, 1, (cuuint32_t)box_rows, 1 };
	cuuint32_t estr[4] = { 1, 1, 1, 1 };
	return encode_fn()(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int env_int(const char* name, int dflt)
{
	const char* e = getenv(name);
	return e ? atoi(e) : dflt;
}

} // namespace
} // namespace sm100
