// sm100_umma_gemm.cuh -- the one tensor-core contraction kernel of the backend.
//
//   D[M, N] (+)= A[M, Kdim] * B[Kdim, N]        fp32 in HBM, TF32 tcgen05.mma, fp32 accumulators in TMEM
//
// GEMM fwd/bwd, convolution fprop / dgrad / wgrad (implicit GEMM, NHWC) are all this kernel with a different way of
// bringing the A and B tiles into shared memory (OperandMode).  Data path per CTA (one 128 x BN output tile,
// optional split-K along gridDim.z):
//
//   warp 0   TMA producer: cp.async.bulk.tensor (tile mode or im2col mode) -> 128B-swizzled smem stages, mbarrier tx
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 8, kind::tf32), tcgen05.commit
//   warp 2-5 epilogue: tcgen05.ld 32x32b -> registers -> (+bias, +C) -> 128-bit global stores (split-K: into this split's scratch slice)
//
// Two CTAs are resident per SM (3 stages x 32 KB at BN = 128), so one CTA's epilogue overlaps the other's main loop.
#pragma once
#include "sm100_ptx.cuh"

namespace sm100 {

enum OperandMode {
	OP_K2D = 0,    // K-major 2-D tensor [rows, K]: one box {32 k, rows} per stage
	OP_MN2D = 1,   // MN-major 2-D tensor [K, rows]: rows/32 boxes {32 mn, BLOCK_K k} per stage
	OP_IM2COL = 2, // NHWC tensor through im2col-mode TMA.  As A: pixels are rows (K-major, channels = K).
	               //                                       As B: pixels are K   (MN-major, channels = N).
};

constexpr int UMMA_BLOCK_M = 128;
constexpr int UMMA_BLOCK_K = 32; // fp32 elements per stage along K = one 128-byte swizzle span
constexpr int UMMA_MAX_TAPS = 64;

struct UmmaRowMap {
	// output row m -> element offset.  mode 0: m * ld.  mode 1: m = (n, i, j) over a (Pc x Qc) grid per image:
	// n * n_stride + i * h_stride + j * w_stride (used by strided dgrad, which writes every other pixel).
	int mode;
	int Pc, Qc;
	long long ld, n_stride, h_stride, w_stride;
};

struct UmmaGemmParams {
	int M, N;
	int k_iters;        // BLOCK_K iterations over the whole reduction (all taps x chunks)
	int chunks_per_tap; // k-iteration it -> tap = it / chunks_per_tap, chunk = it % chunks_per_tap
	int splits;         // split-K factor; gridDim.z = grid_taps * splits.  Split s writes its partial tile to out + s * split_out_stride
	                    // (a scratch slice per split, plain stores); splitk_reduce_kernel then adds the slices in a fixed order, so the
	                    // result does not depend on CTA arrival order (no red.global.add anywhere in the contraction kernels)
	long long split_out_stride;
	int grid_taps;      // > 1 only for wgrad: the filter tap is a grid dimension
	long long grid_tap_out_stride; // output column offset per grid tap
	// im2col geometry of the operand that uses OP_IM2COL: base pixel p -> (w, h, n) TMA coordinates
	int P, Q;
	int stride_h, stride_w;
	int base_h, base_w;
	unsigned short tap_off_h[UMMA_MAX_TAPS], tap_off_w[UMMA_MAX_TAPS];
	int tap_b_col[UMMA_MAX_TAPS]; // B's column (or row-block) origin for this tap
	// epilogue
	float* out;
	const float* bias;
	int accumulate; // 1: D += existing output (CCV_NNC_ACCUMULATE_OUTPUT)
	float alpha;    // scales the product (1 for the reference's commands)
	UmmaRowMap rowmap;
	// optional per-column statistics of the output (batch-norm forward fused into the producing convolution; persistent kernel,
	// TMA-store epilogue only).  Four planes of stats_rows x N floats: count, shift k, sum(v - k), sum((v - k)^2).  Row
	// blockIdx.x * 4 + quarter belongs to one epilogue warp quarter of one CTA and every (row, column) slot to exactly one thread,
	// which updates it with plain loads / stores tile after tile: no atomics, a fixed summation order, and a per-slot shift (the
	// first value seen) that keeps the one-pass variance well conditioned.  The count plane must be zero on entry.
	float* stats;
	int stats_rows;
	int tma_store; // persistent kernel: the epilogue writes 32 x 32 chunks with TMA tile stores through the output tensor map
	// Operand / output format of the PERSISTENT kernel (the one-tile and taps kernels are fp32-only and ignore these).  Operand
	// tiles are always 128-byte-span swizzled rows, so a stage holds bk = 128 / sizeof(element) elements along K: 32 fp32
	// (kind::tf32, 4 MMAs of K = 8) or 64 bf16 / fp16 (kind::f16, 4 MMAs of K = 16); all byte sizes are the same.
	int kind16;            // 0: fp32 operands through kind::tf32; 1: 16-bit operands through kind::f16 (idesc names bf16 / fp16); the
	                       // kernel itself is compiled per operand size (template K16), this records which one the host chose
	int out_kind;          // OUT16 kernels: 1 = bf16, 2 = fp16 output elements (0: fp32 output)
	const void* bias16;    // OUT16 kernels: bias in the output's 16-bit type (p.bias, fp32, wins when both are given)
	uint32_t idesc;
	// smem descriptor fields for MN-major operands (layout SWIZZLE_128B_BASE32B: 128 B of MN x 4 k-rows per atom)
	uint32_t mn_lbo, mn_sbo, mn_layout;
};

template <int BN, int STAGES>
struct UmmaSmem {
	static constexpr int A_BYTES = UMMA_BLOCK_M * UMMA_BLOCK_K * 4;
	static constexpr int B_BYTES = BN * UMMA_BLOCK_K * 4;
	static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
	static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
	static constexpr int TOTAL = BAR_OFFSET + 256 + 1024; // barriers + slack for the 1024-byte alignment
};

template <int AMODE, int BMODE, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const UmmaGemmParams p)
{
	using S = UmmaSmem<BN, STAGES>;
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint64_t* full_bar = (uint64_t*)(smem + S::BAR_OFFSET);
	uint64_t* empty_bar = full_bar + STAGES;
	uint64_t* tmem_full_bar = empty_bar + STAGES;
	uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int m0 = blockIdx.x * UMMA_BLOCK_M;
	const int n0 = blockIdx.y * BN;
	const int split = blockIdx.z % p.splits;
	const int gtap = blockIdx.z / p.splits;
	// this CTA's slice of the reduction
	const int per = (p.k_iters + p.splits - 1) / p.splits;
	const int it_begin = split * per;
	const int it_end = min(p.k_iters, it_begin + per);
	const int n_it = it_end - it_begin;

	if (warp == 0 && lane == 0)
	{
		tma_prefetch_desc(&tmA);
		tma_prefetch_desc(&tmB);
		for (int s = 0; s < STAGES; s++)
		{
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
		}
		mbar_init(tmem_full_bar, 1);
		fence_mbar_init();
	}
	if (warp == 1)
	{
		tmem_alloc(tmem_slot, BN);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0)
	{
		// ------------------------------------------------------------------ TMA producer
		if (n_it > 0 && elect_one())
		{
			// im2col base coordinates of this tile's first row (A side); fixed for the whole loop
			int a_w = 0, a_h = 0, a_n = 0;
			if (AMODE == OP_IM2COL)
			{
				const int q = m0 % p.Q;
				const int t = m0 / p.Q;
				a_w = q * p.stride_w + p.base_w;
				a_h = (t % p.P) * p.stride_h + p.base_h;
				a_n = t / p.P;
			}
			int stage = 0;
			uint32_t phase = 0;
			for (int it = it_begin; it < it_end; it++)
			{
				const int tap = it / p.chunks_per_tap;
				const int chunk = it - tap * p.chunks_per_tap;
				mbar_wait(&empty_bar[stage], phase ^ 1);
				uint8_t* sA = smem + stage * S::STAGE_BYTES;
				uint8_t* sB = sA + S::A_BYTES;
				mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
				if (AMODE == OP_K2D)
					tma_load_2d(sA, &tmA, &full_bar[stage], chunk * UMMA_BLOCK_K, m0);
				else if (AMODE == OP_MN2D) {
#pragma unroll
					for (int j = 0; j < UMMA_BLOCK_M / 32; j++)
						tma_load_2d(sA + j * 4096, &tmA, &full_bar[stage], m0 + 32 * j, it * UMMA_BLOCK_K);
				} else
					tma_load_im2col_4d(sA, &tmA, &full_bar[stage], chunk * UMMA_BLOCK_K, a_w, a_h, a_n, p.tap_off_w[tap], p.tap_off_h[tap]);
				if (BMODE == OP_K2D)
					tma_load_2d(sB, &tmB, &full_bar[stage], p.tap_b_col[tap] + chunk * UMMA_BLOCK_K, n0);
				else if (BMODE == OP_MN2D) {
#pragma unroll
					for (int j = 0; j < BN / 32; j++)
						tma_load_2d(sB + j * 4096, &tmB, &full_bar[stage], p.tap_b_col[tap] + n0 + 32 * j, chunk * UMMA_BLOCK_K);
				} else {
					// pixels are the reduction: this iteration covers base pixels [it * 32, it * 32 + 32)
					const int pix = it * UMMA_BLOCK_K;
					const int q = pix % p.Q;
					const int t = pix / p.Q;
					const int b_w = q * p.stride_w + p.base_w;
					const int b_h = (t % p.P) * p.stride_h + p.base_h;
					const int b_n = t / p.P;
#pragma unroll
					for (int j = 0; j < BN / 32; j++)
						tma_load_im2col_4d(sB + j * 4096, &tmB, &full_bar[stage], n0 + 32 * j, b_w, b_h, b_n, p.tap_off_w[gtap], p.tap_off_h[gtap]);
				}
				if (++stage == STAGES) { stage = 0; phase ^= 1; }
			}
		}
	} else if (warp == 1) {
		// ------------------------------------------------------------------ MMA issuer: one elected thread, operands in uniform registers
		// (sm100_umma_persistent.cuh explains why not `lane == 0`)
		if (n_it > 0 && elect_one())
		{
			const uint32_t smem_base = smem_u32(smem);
			// K-major (SWIZZLE_128B): 8 fp32 = 32 bytes along the swizzled row per MMA; SBO = 8 rows x 128 B.
			// MN-major tf32 (SWIZZLE_128B_BASE32B is the only legal layout): 8 k-rows x 128 B = 1024 bytes per MMA;
			// LBO = stride between 32-wide MN groups, SBO = stride between 4-row k atoms.  The start-address field counts 16-byte units.
			const uint64_t a_desc0 = (AMODE == OP_MN2D) ? umma_smem_desc(smem_base, p.mn_lbo, p.mn_sbo, p.mn_layout) : umma_smem_desc(smem_base, 16, 1024, 2);
			const uint64_t b_desc0 = (BMODE == OP_K2D) ? umma_smem_desc(smem_base + S::A_BYTES, 16, 1024, 2) : umma_smem_desc(smem_base + S::A_BYTES, p.mn_lbo, p.mn_sbo, p.mn_layout);
			constexpr uint32_t A_STEP = (AMODE == OP_MN2D ? 1024 : 32) >> 4, B_STEP = (BMODE == OP_K2D ? 32 : 1024) >> 4;
			const uint32_t idesc = p.idesc;
			int stage = 0;
			uint32_t phase = 0;
			for (int it = 0; it < n_it; it++)
			{
				mbar_wait(&full_bar[stage], phase);
				tc_fence_after();
				const uint64_t da0 = a_desc0 + (uint32_t)stage * (uint32_t)(S::STAGE_BYTES >> 4);
				const uint64_t db0 = b_desc0 + (uint32_t)stage * (uint32_t)(S::STAGE_BYTES >> 4);
#pragma unroll
				for (int k = 0; k < UMMA_BLOCK_K / 8; k++)
					umma_tf32(tmem_base, da0 + k * A_STEP, db0 + k * B_STEP, idesc, (it > 0 || k > 0) ? 1u : 0u);
				umma_commit(&empty_bar[stage]); // frees the smem stage once these MMAs have read it
				if (it == n_it - 1)
					umma_commit(tmem_full_bar);
				if (++stage == STAGES) { stage = 0; phase ^= 1; }
			}
		}
	} else {
		// ------------------------------------------------------------------ epilogue (warps 2..5)
		const int quarter = warp & 3; // TMEM lanes [32 * quarter, +32) are the ones this warp may read
		const int row = m0 + quarter * 32 + lane;
		if (n_it > 0)
		{
			mbar_wait(tmem_full_bar, 0);
			tc_fence_after();
		}
		long long row_off;
		if (p.rowmap.mode == 0)
			row_off = (long long)row * p.rowmap.ld;
		else {
			const int pq = p.rowmap.Pc * p.rowmap.Qc;
			const int n = row / pq;
			const int rem = row - n * pq;
			const int i = rem / p.rowmap.Qc;
			const int j = rem - i * p.rowmap.Qc;
			row_off = n * p.rowmap.n_stride + i * p.rowmap.h_stride + j * p.rowmap.w_stride;
		}
		float* const orow = p.out + row_off + (long long)gtap * p.grid_tap_out_stride + (long long)split * p.split_out_stride;
		const bool row_ok = row < p.M;
		const bool vec_ok = ((((uintptr_t)orow) & 15) == 0);
		const bool add_bias = p.bias != 0 && split == 0;
#pragma unroll 1
		for (int c = 0; c < BN / 32; c++)
		{
			uint32_t r[32];
			if (n_it > 0)
			{
				tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
				tmem_ld_wait();
			} else {
#pragma unroll
				for (int i = 0; i < 32; i++)
					r[i] = 0;
			}
			const int col0 = n0 + c * 32;
			if (!row_ok || col0 >= p.N)
				continue;
			float v[32];
#pragma unroll
			for (int i = 0; i < 32; i++)
				v[i] = __uint_as_float(r[i]) * p.alpha;
			if (add_bias)
			{
#pragma unroll
				for (int i = 0; i < 32; i++)
					if (col0 + i < p.N)
						v[i] += __ldg(p.bias + col0 + i);
			}
			float* const o = orow + col0;
			if (col0 + 32 <= p.N && vec_ok)
			{
				if (p.accumulate)
				{
#pragma unroll
					for (int i = 0; i < 32; i += 4)
					{
						const float4 e = *reinterpret_cast<const float4*>(o + i);
						v[i] += e.x, v[i + 1] += e.y, v[i + 2] += e.z, v[i + 3] += e.w;
					}
				}
#pragma unroll
				for (int i = 0; i < 32; i += 4)
					*reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
			} else {
				for (int i = 0; i < 32; i++)
					if (col0 + i < p.N)
						o[i] = p.accumulate ? o[i] + v[i] : v[i];
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1)
		tmem_dealloc(tmem_base, BN);
}

} // namespace sm100
