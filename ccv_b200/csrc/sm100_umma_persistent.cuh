// sm100_umma_persistent.cuh -- persistent form of the contraction kernel (same operand modes and parameters as
// sm100_umma_gemm.cuh): one CTA per SM loops over output tiles, so that
//   * the TMA producer runs ahead across tile boundaries (no pipeline refill, no per-tile barrier init / TMEM alloc),
//   * accumulators are double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the MMAs of tile i + 1,
//   * the epilogue handles each 32 x 32 accumulator chunk (tcgen05.ld, lane = row) in one of two ways:
//       - plain "write the tile" (dense row-major output, no split-K / accumulate): the chunk is staged (+bias) in shared memory
//         in the 128-byte-swizzled layout of the output tensor map and written by ONE TMA tile store; two buffers per warp,
//         cp.async.bulk.wait_group.read keeps one store in flight (profiles/r01_ncu_expand_1x1_persistent_8epiwarps.txt: the
//         row-store variant below was instruction-issue-bound);
//       - otherwise (strided rows, accumulate, split-K scratch slices): transposed through a pitch-36 scratch tile and written
//         with 128-bit stores in which every 128-byte line is touched whole (4 rows x 128 B per warp instruction);
//     the TMA-store form can also fold per-column shifted sums of what it writes into per-(CTA, warp quarter) slots
//     (batch-norm statistics, p.stats): plain loads / stores by the one thread that owns the slot -- deterministic.
// Roles: warp 0 TMA producer, warp 1 TMEM allocator + tcgen05.mma issuer, warps 2.. epilogue (EPIW = 4: one warp per TMEM
// lane quarter; 8: two, each taking every other 32-column chunk -- with a single warp per scheduler the dependent
// LDTM -> STS -> LDS -> STG chain was latency-bound, profiles/r01_ncu_gemm_n256k64_persistent_4epiwarps.txt).
#pragma once
#include "sm100_umma_gemm.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace sm100 {

// 16-bit output helpers (OUT16 kernels): kind 1 = bf16, 2 = fp16, round to nearest even
__device__ __forceinline__ uint32_t pack16x2(const float a, const float b, const int kind)
{
	if (kind == 1)
	{
		const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
		return *reinterpret_cast<const uint32_t*>(&v);
	}
	const __half2 v = __floats2half2_rn(a, b);
	return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float cvt16(const uint16_t x, const int kind)
{
	return kind == 1 ? __uint_as_float((uint32_t)x << 16) : __half2float(__ushort_as_half(x));
}

// X3 = 1 selects the error-compensated "3xTF32" form (CCV_NNC_SM100_ALGO_3XTF32): the operands arrive as raw fp32
// (tensor map data type FLOAT32, no rounding), four extra warps split every staged element into hi = x with the low 13
// mantissa bits cleared (exactly representable in TF32) and lo = x - hi (exact in fp32), hi rewritten in place and lo into a
// second buffer of the stage, and the issuer runs three MMAs per k-step -- lo*hi + hi*lo + hi*hi (lo*lo ~ 2^-22 is dropped)
// -- into the same fp32 TMEM accumulator.  The product then carries ~2^-21 relative error per term instead of TF32's 2^-11:
// fp32-grade results (the reference's fp32 GEMM is CUBLAS_COMPUTE_32F, lib/nnc/gpu/ccv_nnc_compat.cu:786-803) at a third of
// the tensor rate.
template <int BN, int STAGES, int EPIW, int X3 = 0>
struct UmmaPersistentSmem {
	static constexpr int A_BYTES = UMMA_BLOCK_M * UMMA_BLOCK_K * 4;
	static constexpr int B_BYTES = BN * UMMA_BLOCK_K * 4;
	static constexpr int RAW_BYTES = A_BYTES + B_BYTES;        // what the TMA fills per stage
	static constexpr int STAGE_BYTES = RAW_BYTES * (X3 ? 2 : 1); // X3: [A hi][B hi][A lo][B lo]
	static constexpr int XFORM_WARPS = X3 ? 4 : 0;
	static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
	static constexpr int EPI_PITCH = 36;                       // floats per scratch row: 16-byte aligned, bank-shifted
	static constexpr int EPI_WARPS = EPIW; // 4: one warp per TMEM lane quarter; 8: two, each taking every other 32-column chunk
	static constexpr int EPI_WARP_BYTES = 8192;                // per epilogue warp: two 4 KB chunk buffers (TMA-store path) / one pitch-36 chunk
	static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
	static constexpr int THREADS = 64 + EPI_WARPS * 32 + XFORM_WARPS * 32;
	static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
	static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

// K16 = 1: 16-bit (bf16 / fp16) operands through kind::f16 -- 64 elements per 128-byte operand row, K = 16 per MMA, ordinary
// 128-byte swizzle for MN-major tiles; K16 = 0: fp32 operands through kind::tf32 (32 elements per row, K = 8 per MMA).  A
// compile-time switch so that the producer's box loops and the issuer's descriptor arithmetic stay fully unrolled constants.
// OUT16 = 1: the output tensor (and an optional 16-bit bias) is bf16 / fp16 (p.out_kind); accumulation stays fp32 in TMEM.
template <int AMODE, int BMODE, int BN, int STAGES, int EPIW, int X3 = 0, int K16 = 0, int OUT16 = 0>
__global__ void __launch_bounds__(64 + EPIW * 32 + (X3 ? 128 : 0), 1) umma_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const UmmaGemmParams p)
{
	using S = UmmaPersistentSmem<BN, STAGES, EPIW, X3>;
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint64_t* full_bar = (uint64_t*)(smem + S::BAR_OFFSET);
	uint64_t* empty_bar = full_bar + STAGES;
	uint64_t* tmem_full_bar = empty_bar + STAGES; // [2]
	uint64_t* tmem_empty_bar = tmem_full_bar + 2; // [2]
	uint64_t* xform_bar = tmem_empty_bar + 2;     // [STAGES] (X3): the hi / lo split of the stage is in place
	uint32_t* tmem_slot = (uint32_t*)(xform_bar + STAGES);

	constexpr int BK = K16 ? 64 : 32;                  // K elements per stage = MN elements per MN-major box: one 128-byte span
	constexpr int MN_BOX_BYTES = BK * 128;             // an MN-major box: BK k-rows of 128 bytes
	constexpr int MN_STEP = (K16 ? 16 : 8) * 128;      // start-address advance per MMA of an MN-major operand (UMMA_K k-rows)
	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int tiles_m = (p.M + UMMA_BLOCK_M - 1) / UMMA_BLOCK_M;
	const int tiles_n = (p.N + BN - 1) / BN;
	const int total_tiles = tiles_m * tiles_n * p.grid_taps * p.splits;
	const int per = (p.k_iters + p.splits - 1) / p.splits;

	if (warp == 0 && lane == 0)
	{
		tma_prefetch_desc(&tmA);
		tma_prefetch_desc(&tmB);
		for (int s = 0; s < STAGES; s++)
		{
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
			mbar_init(&xform_bar[s], X3 ? S::XFORM_WARPS : 1);
		}
		for (int a = 0; a < 2; a++)
		{
			mbar_init(&tmem_full_bar[a], 1);
			mbar_init(&tmem_empty_bar[a], S::EPI_WARPS); // one arrival per epilogue warp
		}
		fence_mbar_init();
	}
	if (warp == 1)
	{
		tmem_alloc(tmem_slot, 2 * BN);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0)
	{
		// ------------------------------------------------------------------ TMA producer (one elected thread, see the MMA issuer)
		if (elect_one())
		{
			int stage = 0;
			uint32_t phase = 0;
			for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
			{
				const int n_blk = tile % tiles_n;
				const int rest = tile / tiles_n;
				const int m_blk = rest % tiles_m;
				const int zz = rest / tiles_m;
				const int split = zz % p.splits, gtap = zz / p.splits;
				const int m0 = m_blk * UMMA_BLOCK_M, n0 = n_blk * BN;
				const int it_begin = split * per;
				const int it_end = min(p.k_iters, it_begin + per);
				int a_w = 0, a_h = 0, a_n = 0;
				if (AMODE == OP_IM2COL)
				{
					const int q = m0 % p.Q;
					const int t = m0 / p.Q;
					a_w = q * p.stride_w + p.base_w;
					a_h = (t % p.P) * p.stride_h + p.base_h;
					a_n = t / p.P;
				}
				for (int it = it_begin; it < it_end; it++)
				{
					const int tap = it / p.chunks_per_tap;
					const int chunk = it - tap * p.chunks_per_tap;
					mbar_wait(&empty_bar[stage], phase ^ 1);
					uint8_t* sA = smem + stage * S::STAGE_BYTES;
					uint8_t* sB = sA + S::A_BYTES;
					mbar_expect_tx(&full_bar[stage], S::RAW_BYTES);
					if (AMODE == OP_K2D)
						tma_load_2d(sA, &tmA, &full_bar[stage], chunk * BK, m0);
					else if (AMODE == OP_MN2D) {
#pragma unroll
						for (int j = 0; j < UMMA_BLOCK_M / BK; j++)
							tma_load_2d(sA + j * MN_BOX_BYTES, &tmA, &full_bar[stage], m0 + BK * j, it * BK);
					} else
						tma_load_im2col_4d(sA, &tmA, &full_bar[stage], chunk * BK, a_w, a_h, a_n, p.tap_off_w[tap], p.tap_off_h[tap]);
					if (BMODE == OP_K2D)
						tma_load_2d(sB, &tmB, &full_bar[stage], p.tap_b_col[tap] + chunk * BK, n0);
					else if (BMODE == OP_MN2D) {
#pragma unroll
						for (int j = 0; j < BN / BK; j++)
							tma_load_2d(sB + j * MN_BOX_BYTES, &tmB, &full_bar[stage], p.tap_b_col[tap] + n0 + BK * j, chunk * BK);
					} else {
						const int pix = it * BK;
						const int q = pix % p.Q;
						const int t = pix / p.Q;
						const int b_w = q * p.stride_w + p.base_w;
						const int b_h = (t % p.P) * p.stride_h + p.base_h;
						const int b_n = t / p.P;
#pragma unroll
						for (int j = 0; j < BN / BK; j++)
							tma_load_im2col_4d(sB + j * MN_BOX_BYTES, &tmB, &full_bar[stage], n0 + BK * j, b_w, b_h, b_n, p.tap_off_w[gtap], p.tap_off_h[gtap]);
					}
					if (++stage == STAGES) { stage = 0; phase ^= 1; }
				}
			}
		}
	} else if (warp == 1) {
		// ------------------------------------------------------------------ MMA issuer
		// One elected thread runs the whole role.  `elect_one()` (not a lane test) and descriptors as base + constant keep every operand in
		// uniform registers: with `if (lane == 0)` around each issue ptxas wraps every tcgen05.mma in an elect / broadcast / branch loop and
		// rebuilds the descriptors -- about 12 dependent instructions per MMA, more than a 128 x 64 MMA takes to execute.
		if (elect_one())
		{
			const uint32_t smem_base = smem_u32(smem);
			const uint64_t a_desc0 = (AMODE == OP_MN2D) ? umma_smem_desc(smem_base, p.mn_lbo, p.mn_sbo, p.mn_layout) : umma_smem_desc(smem_base, 16, 1024, 2);
			const uint64_t b_desc0 = (BMODE == OP_K2D) ? umma_smem_desc(smem_base + S::A_BYTES, 16, 1024, 2) : umma_smem_desc(smem_base + S::A_BYTES, p.mn_lbo, p.mn_sbo, p.mn_layout);
			// one MMA covers 32 bytes of K (8 fp32 / 16 16-bit): K-major operands advance 32 B inside the swizzled row, MN-major ones by
			// UMMA_K k-rows of 128 B; the start-address field counts 16-byte units
			constexpr uint32_t A_STEP = (AMODE == OP_MN2D ? MN_STEP : 32) >> 4, B_STEP = (BMODE == OP_K2D ? 32 : MN_STEP) >> 4;
			const uint32_t idesc = p.idesc;
			int stage = 0;
			uint32_t phase = 0;
			int t = 0;
			for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t++)
			{
				const int zz = (tile / tiles_n) / tiles_m;
				const int split = zz % p.splits;
				const int it_begin = split * per;
				const int n_it = min(p.k_iters, it_begin + per) - it_begin;
				const int acc = t & 1;
				const uint32_t acc_phase = (t >> 1) & 1;
				mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1); // the epilogue has drained this accumulator
				tc_fence_after();
				const uint32_t tmem_d = tmem_base + acc * BN;
				for (int it = 0; it < n_it; it++)
				{
					mbar_wait(X3 ? &xform_bar[stage] : &full_bar[stage], phase);
					tc_fence_after();
					const uint64_t da0 = a_desc0 + (uint32_t)stage * (uint32_t)(S::STAGE_BYTES >> 4);
					const uint64_t db0 = b_desc0 + (uint32_t)stage * (uint32_t)(S::STAGE_BYTES >> 4);
#pragma unroll
					for (int k = 0; k < UMMA_BLOCK_K / 8; k++)
					{
						const uint64_t da = da0 + k * A_STEP;
						const uint64_t db = db0 + k * B_STEP;
						if (K16)
							umma_f16(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
						else if (X3)
						{
							// the lo tiles sit RAW_BYTES behind their hi tiles (same layout): small terms first, then hi * hi
							constexpr uint64_t LO = (uint64_t)(S::RAW_BYTES >> 4);
							umma_tf32(tmem_d, da + LO, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
							umma_tf32(tmem_d, da, db + LO, idesc, 1u);
							umma_tf32(tmem_d, da, db, idesc, 1u);
						} else
							umma_tf32(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
					}
					umma_commit(&empty_bar[stage]);
					if (++stage == STAGES) { stage = 0; phase ^= 1; }
				}
				umma_commit(&tmem_full_bar[acc]); // fires once every MMA of this tile has completed
			}
		}
	} else if (X3 && warp >= 2 + EPIW) {
		// ------------------------------------------------------------------ hi / lo split of every staged operand element (X3)
		const int tid = threadIdx.x - (2 + EPIW) * 32; // 0..127
		int stage = 0;
		uint32_t phase = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
		{
			const int zz = (tile / tiles_n) / tiles_m;
			const int split = zz % p.splits;
			const int it_begin = split * per;
			const int n_it = min(p.k_iters, it_begin + per) - it_begin;
			for (int it = 0; it < n_it; it++)
			{
				mbar_wait(&full_bar[stage], phase); // the TMA has filled the raw tiles
				float4* const raw = reinterpret_cast<float4*>(smem + stage * S::STAGE_BYTES);
				float4* const lo = reinterpret_cast<float4*>(smem + stage * S::STAGE_BYTES + S::RAW_BYTES);
#pragma unroll 4
				for (int i = tid; i < S::RAW_BYTES / 16; i += S::XFORM_WARPS * 32)
				{
					const float4 x = raw[i];
					float4 h, l;
					h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u), h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
					h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u), h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
					// x - hi is exact; an infinity keeps lo = 0 so that inf stays inf instead of turning into inf - inf
					l.x = fabsf(x.x) < __int_as_float(0x7f800000) ? x.x - h.x : 0.f, l.y = fabsf(x.y) < __int_as_float(0x7f800000) ? x.y - h.y : 0.f;
					l.z = fabsf(x.z) < __int_as_float(0x7f800000) ? x.z - h.z : 0.f, l.w = fabsf(x.w) < __int_as_float(0x7f800000) ? x.w - h.w : 0.f;
					raw[i] = h;
					lo[i] = l;
				}
				fence_proxy_async(); // the tensor core reads shared memory through the async proxy
				__syncwarp();
				if (lane == 0)
					mbar_arrive(&xform_bar[stage]);
				if (++stage == STAGES) { stage = 0; phase ^= 1; }
			}
		}
	} else {
		// ------------------------------------------------------------------ epilogue (warps 2..9)
		const int quarter = warp & 3;        // TMEM lanes [32 * quarter, +32) are the ones this warp may read
		const int half = (warp - 2) >> 2;    // which of the EPIW / 4 warps of this quarter: chunks half, half + EPIW / 4, ...
		float* const scratch = reinterpret_cast<float*>(smem + S::EPI_OFFSET + (warp - 2) * S::EPI_WARP_BYTES);
		const bool tma_store = p.tma_store != 0;
		int chunk_no = 0;
		const int sub_row = lane >> 3;       // 0..3: row within a group of 4 rows
		const int sub_col = (lane & 7) * 4;  // 0..28: first of this lane's 4 columns
		const bool accumulate = p.accumulate != 0;
		const float* const bias = p.bias;
		const uint16_t* const bias16 = OUT16 && !p.bias ? reinterpret_cast<const uint16_t*>(p.bias16) : 0;
		const int N = p.N;
		// statistics slots of this warp quarter: plane 0 of row blockIdx.x * 4 + quarter (count); planes k, s1, s2 follow at stats_plane.
		// The launcher makes gridDim.x a multiple of tiles_n, so every tile of this CTA covers the same columns: the shifted sums of
		// the warp's chunks live in (thread-local) arrays for the whole kernel and are stored once at the end.
		float* const stats = p.stats && p.tma_store ? p.stats + ((size_t)blockIdx.x * 4 + quarter) * p.N : 0;
		const size_t stats_plane = (size_t)p.stats_rows * p.N;
		constexpr int CHUNKS_PER_WARP = (BN / 32 + EPIW / 4 - 1) / (EPIW / 4);
		float st_k[CHUNKS_PER_WARP], st_s1[CHUNKS_PER_WARP], st_s2[CHUNKS_PER_WARP];
		float st_n = 0.f;
		int st_n0 = 0;
#pragma unroll
		for (int i = 0; i < CHUNKS_PER_WARP; i++)
			st_k[i] = st_s1[i] = st_s2[i] = 0.f;
		int t = 0;
		for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, t++)
		{
			const int n_blk = tile % tiles_n;
			const int rest = tile / tiles_n;
			const int m_blk = rest % tiles_m;
			const int zz = rest / tiles_m;
			const int split = zz % p.splits, gtap = zz / p.splits;
			const int m0 = m_blk * UMMA_BLOCK_M, n0 = n_blk * BN;
			const int it_begin = split * per;
			const int n_it = min(p.k_iters, it_begin + per) - it_begin;
			const int acc = t & 1;
			const uint32_t acc_phase = (t >> 1) & 1;
			const bool add_bias = bias != 0 && split == 0;
			// output row pointers of the 8 rows this lane writes (rows sub_row + 4 * i of the warp's 32-row band)
			constexpr int OSZ = OUT16 ? 2 : 4; // bytes per output element
			char* orow[8];
#pragma unroll
			for (int i = 0; i < 8; i++)
			{
				const int row = m0 + quarter * 32 + sub_row + 4 * i;
				long long off;
				if (p.rowmap.mode == 0)
					off = (long long)row * p.rowmap.ld;
				else {
					const int pq = p.rowmap.Pc * p.rowmap.Qc;
					const int n = row / pq;
					const int rem = row - n * pq;
					const int ii = rem / p.rowmap.Qc;
					const int jj = rem - ii * p.rowmap.Qc;
					off = n * p.rowmap.n_stride + ii * p.rowmap.h_stride + jj * p.rowmap.w_stride;
				}
				orow[i] = row < p.M ? reinterpret_cast<char*>(p.out) + (off + (long long)gtap * p.grid_tap_out_stride + (long long)split * p.split_out_stride) * OSZ : 0;
			}
			mbar_wait(&tmem_full_bar[acc], acc_phase);
			tc_fence_after();
#pragma unroll 1
			for (int c = half; c < BN / 32; c += EPIW / 4)
			{
				uint32_t r[32];
				if (n_it > 0)
				{
					tmem_ld_32x32(tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
					tmem_ld_wait();
				} else {
#pragma unroll
					for (int i = 0; i < 32; i++)
						r[i] = 0;
				}
				if (tma_store)
				{
					// lane = row.  The 32 x 32 chunk goes (+bias) into one of this warp's two 4 KB buffers in the 128-byte-swizzled
					// layout the output tensor map expects and leaves as ONE TMA tile store: ~60 instructions per chunk instead of
					// ~400 (LDS / address checks / STG per row: the epilogue was issue-bound, profiles/r01_ncu_expand_1x1_*.txt)
					const int col0 = n0 + c * 32;
					if (col0 >= N)
						continue; // a chunk wholly past the last column: nothing to store.  It must not take a buffer either -- the
						          // two-buffer rotation below relies on exactly one bulk group being committed per buffer use
					float* const buf = scratch + (chunk_no & 1) * 1024;
					chunk_no++;
					if (lane == 0)
						bulk_wait_group_read<1>(); // the store issued two chunks ago has finished reading this buffer
					__syncwarp();
					float v[32];
#pragma unroll
					for (int i = 0; i < 32; i++)
						v[i] = __uint_as_float(r[i]);
					if (add_bias)
					{
#pragma unroll
						for (int i = 0; i < 32; i += 4)
							if (col0 + i < N)
							{
								const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
								v[i] += b4.x, v[i + 1] += b4.y, v[i + 2] += b4.z, v[i + 3] += b4.w;
							}
					} else if (OUT16 && bias16 && split == 0) {
#pragma unroll
						for (int i = 0; i < 32; i += 8)
							if (col0 + i < N)
							{
								const uint4 b8 = __ldg(reinterpret_cast<const uint4*>(bias16 + col0 + i));
								const uint32_t w[4] = { b8.x, b8.y, b8.z, b8.w };
#pragma unroll
								for (int j = 0; j < 4; j++)
									v[i + 2 * j] += cvt16((uint16_t)(w[j] & 0xffffu), p.out_kind), v[i + 2 * j + 1] += cvt16((uint16_t)(w[j] >> 16), p.out_kind);
							}
					}
					if (!OUT16)
					{
#pragma unroll
						for (int j = 0; j < 8; j++)
							*reinterpret_cast<float4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
					} else {
						// 32 rows x 64 bytes in the 64-byte-swizzled layout of the 16-bit output map: 16-byte unit u of row r sits at
						// unit u ^ ((r >> 1) & 3) (8 consecutive lanes then cover all 32 banks: conflict-free 128-bit stores)
						char* const b8 = reinterpret_cast<char*>(buf) + lane * 64;
#pragma unroll
						for (int j = 0; j < 4; j++)
							*reinterpret_cast<uint4*>(b8 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pack16x2(v[8 * j], v[8 * j + 1], p.out_kind), pack16x2(v[8 * j + 2], v[8 * j + 3], p.out_kind),
								pack16x2(v[8 * j + 4], v[8 * j + 5], p.out_kind), pack16x2(v[8 * j + 6], v[8 * j + 7], p.out_kind));
					}
					fence_proxy_async();
					__syncwarp();
					const int valid = min(32, p.M - (m0 + quarter * 32)); // rows of this chunk inside the tensor (the store clips the rest)
					if (stats && valid > 0)
					{
						// lane = column: fold the staged rows of this column into this thread's running shifted sums (all lanes read the
						// same row: no bank conflict).  k is the first value the thread ever saw for the column.  16-bit outputs are
						// read back as stored (rounded): the statistics describe the tensor the batch norm will normalise.
						const int ci = (c - half) / (EPIW / 4);
						const uint16_t* const buf16 = reinterpret_cast<const uint16_t*>(buf);
						auto staged = [&](const int rr) -> float {
							if (!OUT16)
								return buf[rr * 32 + ((((lane >> 2) ^ (rr & 7))) << 2) + (lane & 3)];
							return cvt16(buf16[rr * 32 + ((((lane >> 3) ^ ((rr >> 1) & 3))) << 3) + (lane & 7)], p.out_kind);
						};
						const float k = st_n > 0.f ? st_k[ci] : staged(0);
						float s1 = 0.f, s2 = 0.f;
						if (valid == 32)
						{
							// the common case, branch-free: two independent accumulator pairs halve the dependent-add chain
							float t1 = 0.f, t2 = 0.f;
#pragma unroll
							for (int rr = 0; rr < 32; rr += 2)
							{
								const float d0 = staged(rr) - k, d1 = staged(rr + 1) - k;
								s1 += d0, t1 += d1;
								s2 = fmaf(d0, d0, s2), t2 = fmaf(d1, d1, t2);
							}
							s1 += t1, s2 += t2;
						} else {
							for (int rr = 0; rr < valid; rr++)
							{
								const float d = staged(rr) - k;
								s1 += d;
								s2 = fmaf(d, d, s2);
							}
						}
						st_k[ci] = k, st_s1[ci] += s1, st_s2[ci] += s2;
					}
					if (lane == 0)
					{
						tma_store_2d(&tmC, buf, col0, m0 + quarter * 32);
						bulk_commit_group();
					}
					continue;
				}
				// lane = row: park the row in the scratch tile, then re-read it as (4 rows x 8 lanes x 16 bytes)
				__syncwarp();
#pragma unroll
				for (int i = 0; i < 32; i += 4)
					*reinterpret_cast<float4*>(scratch + lane * S::EPI_PITCH + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
				__syncwarp();
				const int col = n0 + c * 32 + sub_col;
				if (col < N)
				{
				float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
				const bool full4 = col + 4 <= N;
				if (add_bias)
				{
					bias4.x = __ldg(bias + col);
					bias4.y = col + 1 < N ? __ldg(bias + col + 1) : 0.f;
					bias4.z = col + 2 < N ? __ldg(bias + col + 2) : 0.f;
					bias4.w = col + 3 < N ? __ldg(bias + col + 3) : 0.f;
				} else if (OUT16 && bias16 && split == 0) {
					bias4.x = cvt16(__ldg(bias16 + col), p.out_kind);
					bias4.y = col + 1 < N ? cvt16(__ldg(bias16 + col + 1), p.out_kind) : 0.f;
					bias4.z = col + 2 < N ? cvt16(__ldg(bias16 + col + 2), p.out_kind) : 0.f;
					bias4.w = col + 3 < N ? cvt16(__ldg(bias16 + col + 3), p.out_kind) : 0.f;
				}
#pragma unroll
				for (int i = 0; i < 8; i++)
				{
					if (!orow[i])
						continue;
					float4 v = *reinterpret_cast<const float4*>(scratch + (sub_row + 4 * i) * S::EPI_PITCH + sub_col);
					v.x += bias4.x, v.y += bias4.y, v.z += bias4.z, v.w += bias4.w;
					char* const ob = orow[i] + (long long)col * OSZ;
					if (!OUT16)
					{
						float* const o = reinterpret_cast<float*>(ob);
						if (full4 && ((((uintptr_t)o) & 15) == 0))
						{
							if (accumulate)
							{
								const float4 e = *reinterpret_cast<const float4*>(o);
								v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
							}
							*reinterpret_cast<float4*>(o) = v;
						} else {
							const float vv[4] = { v.x, v.y, v.z, v.w };
							for (int j = 0; j < 4; j++)
								if (col + j < N)
									o[j] = accumulate ? o[j] + vv[j] : vv[j];
						}
					} else {
						uint16_t* const o = reinterpret_cast<uint16_t*>(ob);
						if (full4 && ((((uintptr_t)o) & 7) == 0))
						{
							if (accumulate)
							{
								const uint2 e = *reinterpret_cast<const uint2*>(o);
								v.x += cvt16((uint16_t)(e.x & 0xffffu), p.out_kind), v.y += cvt16((uint16_t)(e.x >> 16), p.out_kind);
								v.z += cvt16((uint16_t)(e.y & 0xffffu), p.out_kind), v.w += cvt16((uint16_t)(e.y >> 16), p.out_kind);
							}
							*reinterpret_cast<uint2*>(o) = make_uint2(pack16x2(v.x, v.y, p.out_kind), pack16x2(v.z, v.w, p.out_kind));
						} else {
							const float vv[4] = { v.x, v.y, v.z, v.w };
							for (int j = 0; j < 4; j++)
								if (col + j < N)
									o[j] = (uint16_t)(pack16x2(accumulate ? cvt16(o[j], p.out_kind) + vv[j] : vv[j], 0.f, p.out_kind) & 0xffffu);
						}
					}
				}
				}
			}
			if (stats)
			{
				const int valid = min(32, p.M - (m0 + quarter * 32));
				if (valid > 0)
					st_n += (float)valid, st_n0 = n0;
			}
			// all tcgen05.ld of this tile are complete (tmem_ld_wait): hand the accumulator back to the MMA warp
			tc_fence_before();
			__syncwarp();
			if (lane == 0)
				mbar_arrive(&tmem_empty_bar[acc]);
		}
		if (stats && st_n > 0.f)
		{
#pragma unroll
			for (int ci = 0; ci < CHUNKS_PER_WARP; ci++)
			{
				const int col = st_n0 + (half + ci * (EPIW / 4)) * 32 + lane;
				if (half + ci * (EPIW / 4) < BN / 32 && col < N)
				{
					float* const slot = stats + col;
					slot[0] = st_n, slot[stats_plane] = st_k[ci], slot[2 * stats_plane] = st_s1[ci], slot[3 * stats_plane] = st_s2[ci];
				}
			}
		}
		if (tma_store && lane == 0)
			bulk_wait_group<0>(); // every tile store of this warp has been written out before the CTA exits
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1)
		tmem_dealloc(tmem_base, 2 * BN);
}

} // namespace sm100
