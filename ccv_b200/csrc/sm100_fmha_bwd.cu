// sm100_fmha_bwd.cu -- SCALED_DOT_PRODUCT_ATTENTION backward for 16-bit tensors (bf16 / fp16), head dimension <= 128 (multiples of 8;
// the tiles are always 128 features wide, missing features are zero-filled by the TMA unit and never stored): a fused
// flash-attention backward on the tcgen05 tensor cores, deterministic (no atomics, every sum in one fixed order).
// Semantics: scaled_dot_product_attention/ccv_nnc_scaled_dot_product_attention_cpu_ref.c:259-479 (dq, dk, dv of
// O = softmax(scale * Q K^T [causal, bottom-right aligned]) V, GQA gradients summed over the query heads of a key head); the reference's
// GPU form is gpu/ccv_nnc_scaled_dot_product_attention_flash_attn.cu:221-451 (recomputation from O and the saved log-sum-exp,
// :410-440 its deterministic split accumulator, :444-449 the GQA sum).
//
//   P  = exp(scale * Q K^T - LSE)            recomputed per tile from the saved log-sum-exp
//   dV = P^T dO
//   dP = dO V^T,   dS = P o (dP - delta),    delta_i = sum_d dO_id O_id
//   dQ = scale * dS K,   dK = scale * dS^T Q
//
// Three launches:
//   fmha_bwd_prep_kernel    (lse * log2 e, delta) pairs per query row, padded to whole tiles (padding rows: lse = +inf -> P = 0)
//   fmha_bwd_kernel<0>      one CTA per 128 keys of one (batch, KEY head): K, V tiles resident in shared memory, Q_i / dO_i blocks of
//                           64 queries streamed for every query head of the group; dV, dK accumulate in TMEM over the whole loop
//   fmha_bwd_kernel<1>      one CTA per 128 queries of one (batch, query head): Q, dO resident, K_j / V_j blocks of 64 keys streamed;
//                           dQ accumulates in TMEM
// Both recompute S and dP (7 GEMMs instead of 5) so that no gradient is ever accumulated across CTAs: that is what makes the result
// run-to-run identical without the reference's per-split fp32 dq_accum buffers.
//
// One kernel body serves both (MODE 0 = dK/dV, 1 = dQ).  R1 / R2 = the resident 128-row tiles, T1 / T2 = the streamed 64-row blocks:
//   MODE 0: R1 = K, R2 = V, T1 = Q_i, T2 = dO_i:  X = R1 T1^T = S^T,  Y = R2 T2^T = dP^T,  acc1 += P^T T2 (dV),  acc2 += dS^T T1 (dK)
//   MODE 1: R1 = Q, R2 = dO, T1 = K_j, T2 = V_j:  X = R1 T1^T = S,    Y = R2 T2^T = dP,    acc2 += dS T1 (dQ)
// 320 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-9 = the element-wise stage (two warps per TMEM lane
// quarter, 32 of the 64 columns each).  X, Y from TMEM -> P, dS as bf16 / fp16 pairs written straight back into TMEM (over the
// columns just read), where the accumulating MMAs take them as their A operand: an MMA whose two operands come from shared memory
// is bound by the operand fetch, not by the tensor core (profiles/r02_ncu_fmha_bwd.txt).  T1 / T2 are the B operand twice --
// K-major for X / Y, MN-major for the accumulation.  The streamed blocks sit in a ring of 4 (MODE 0) / 7 (MODE 1) stages: a stage is only free once the accumulation of its
// block has completed, and the next X / Y is issued one block ahead, so two stages would expose the whole TMA latency every block.
// MODE 1 also keeps its resident tiles (Q, dout) in tensor memory, so every one of its MMAs reads only the streamed block from
// shared memory.  TMEM: X[2] 2 x 64 + Y[2] 2 x 64 + dV 128 (MODE 0) | Q 64 + dout 64 (MODE 1) + dK | dQ 128 = 512 columns.
// smem: R 64 KB + T 4 x 32 KB = 192 KB (MODE 0), T 7 x 32 KB = 224 KB (MODE 1).
#include "sm100_contract.h"
#include "sm100_fmha.cuh"
#include <string.h>

namespace sm100 {

namespace {

constexpr int FB_R = 128;                     // resident rows per CTA
constexpr int FB_T = 64;                      // streamed rows per block
constexpr int FB_D = 128;                     // head dimension
constexpr int FB_R_ATOM = FB_R * 128;         // 16 KB: 128 rows x 64 features
constexpr int FB_R_BYTES = 2 * FB_R_ATOM;     // 32 KB
constexpr int FB_T_ATOM = FB_T * 128;         // 8 KB
constexpr int FB_T_BYTES = 2 * FB_T_ATOM;     // 16 KB
constexpr int FB_STAT_BYTES = FB_T * 8;       // (-lse2, -delta) per streamed query, as float4 (-l0, -l1, -d0, -d1) per query pair

struct FmhaBwdParams {
	int H, Hk, Sq, Sk, Sq_r;
	int D; // actual head dimension (<= 128, a multiple of 8): tiles are 128 wide, missing features are zero-filled and never stored
	int causal;
	int is_bf16;
	float scale, scale_log2;
	const float4* stat;          // [B, H, Sq_r / 2] (-lse2_0, -lse2_1, -delta_0, -delta_1) per query pair, lse2 = lse * log2 e
	void* out1;                  // MODE 0: dV
	void* out2;                  // MODE 0: dK, MODE 1: dQ
	long long o1_b, o1_s, o1_h;  // element strides
	long long o2_b, o2_s, o2_h;
	uint32_t idesc_xy, idesc_acc;
	// MODE 1: the resident tiles (Q, dout) are read straight from global memory into tensor memory
	const void* r1;
	const void* r2;
	long long r1_b, r1_s, r1_h, r2_b, r2_s, r2_h;
};

template <int MODE>
struct FmhaBwdSmem {
	static constexpr int NST = MODE == 0 ? 4 : 7; // stages of the streamed ring
	static constexpr int R1_OFF = 0;              // MODE 0 only: MODE 1 keeps its resident tiles in tensor memory
	static constexpr int R2_OFF = R1_OFF + FB_R_BYTES;
	static constexpr int T_OFF = MODE == 0 ? R2_OFF + FB_R_BYTES : 0; // stage s: T1 at + s * 2 * FB_T_BYTES, T2 right behind it
	static constexpr int STAT_OFF = T_OFF + NST * 2 * FB_T_BYTES;
	static constexpr int BAR_OFF = STAT_OFF + (MODE == 0 ? NST * FB_STAT_BYTES : 0);
	static constexpr int TOTAL = BAR_OFF + 256 + 1024;
	static_assert(TOTAL <= 232448, "227 KB of shared memory per CTA");
};

__device__ __forceinline__ float widen16(const uint16_t u, const int is_bf16)
{
	return is_bf16 ? __uint_as_float((uint32_t)u << 16) : __half2float(__ushort_as_half(u));
}

// (-lse * log2 e, -delta) per query row: 16 lanes per row, 8 features each
__global__ void __launch_bounds__(256) fmha_bwd_prep_kernel(const uint16_t* __restrict__ dout, const uint16_t* __restrict__ out, const float* __restrict__ lse, float* __restrict__ stat, int B, int H, int Sq, int Sq_r, int D, int is_bf16,
	long long do_b, long long do_s, long long do_h, long long o_b, long long o_s, long long o_h)
{
	const long long row = (long long)blockIdx.x * 16 + (threadIdx.x >> 4);
	const int sub = threadIdx.x & 15;
	const long long rows = (long long)B * H * Sq_r;
	if (row >= rows)
		return;
	const int q = (int)(row % Sq_r);
	const int h = (int)((row / Sq_r) % H);
	const int b = (int)(row / ((long long)Sq_r * H));
	float acc = 0.f;
	if (q < Sq && sub * 8 < D)
	{
		const uint4 a = *reinterpret_cast<const uint4*>(dout + b * do_b + (long long)q * do_s + h * do_h + sub * 8);
		const uint4 c = *reinterpret_cast<const uint4*>(out + b * o_b + (long long)q * o_s + h * o_h + sub * 8);
		const uint32_t aw[4] = { a.x, a.y, a.z, a.w }, cw[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
		for (int i = 0; i < 4; i++)
		{
			acc = fmaf(widen16((uint16_t)(aw[i] & 0xffff), is_bf16), widen16((uint16_t)(cw[i] & 0xffff), is_bf16), acc);
			acc = fmaf(widen16((uint16_t)(aw[i] >> 16), is_bf16), widen16((uint16_t)(cw[i] >> 16), is_bf16), acc);
		}
	}
	// fixed-order tree over the 16 lanes of the row
#pragma unroll
	for (int m = 8; m >= 1; m >>= 1)
		acc += __shfl_xor_sync(0xffffffffu, acc, m);
	if (sub == 0)
	{
		float l2 = INFINITY; // padding rows and fully masked rows (lse = -inf): P = exp2(x - inf) = 0
		if (q < Sq)
		{
			const float l = lse[((long long)b * H + h) * Sq + q];
			if (l != -INFINITY)
				l2 = l * 1.4426950408889634f;
		}
		// query pair (2j, 2j + 1) -> (-l, -l', -d, -d'): Sq_r is even, so a pair never straddles two heads
		float* const dst = stat + (row >> 1) * 4 + (row & 1);
		dst[0] = -l2;
		dst[2] = q < Sq ? -acc : 0.f;
	}
}

template <int MODE>
__global__ void __launch_bounds__(320, 1) fmha_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2, const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2, const FmhaBwdParams p)
{
	typedef FmhaBwdSmem<MODE> L;
	constexpr int NST = L::NST;
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint64_t* bars = (uint64_t*)(smem + L::BAR_OFF);
	uint64_t* r_full = bars;
	uint64_t* t_full = bars + 1;          // [NST] TMA -> MMA (and the element-wise warps: the stat pairs ride on the same barrier)
	uint64_t* t_empty = t_full + NST;     // [NST] MMA (accumulation done) -> TMA
	uint64_t* x_full = t_empty + NST;     // [2] MMA -> element-wise: X and Y of this block are in TMEM
	uint64_t* p_full = x_full + 2;        // [2] element-wise -> MMA (8 arrivals): P / dS written back into TMEM
	uint64_t* acc_full = p_full + 2;
	uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r0 = blockIdx.x * FB_R;
	const int hr = blockIdx.y, b = blockIdx.z; // MODE 0: key head, MODE 1: query head
	const int G = p.H / p.Hk;
	const int shift = p.Sk - p.Sq; // query i sees keys <= i + shift
	// streamed blocks this tile needs
	int blk0 = 0, n_blk;
	if (MODE == 0)
	{
		const int nq = (p.Sq + FB_T - 1) / FB_T;
		if (p.causal)
			blk0 = min(nq, max(r0 - shift, 0) / FB_T); // queries below r0 - shift see none of this tile's keys
		n_blk = nq - blk0;
	} else {
		const int kv_end = p.causal ? min(p.Sk, max(r0 + FB_R + shift, 0)) : p.Sk;
		n_blk = (kv_end + FB_T - 1) / FB_T;
	}
	const int n_it = MODE == 0 ? n_blk * G : n_blk;

	if (warp == 0 && lane == 0)
	{
		tma_prefetch_desc(&tmR1);
		tma_prefetch_desc(&tmR2);
		tma_prefetch_desc(&tmT1);
		tma_prefetch_desc(&tmT2);
		mbar_init(r_full, MODE == 0 ? 1 : 8);
		for (int s = 0; s < NST; s++)
			mbar_init(&t_full[s], 1), mbar_init(&t_empty[s], 1);
		for (int s = 0; s < 2; s++)
		{
			mbar_init(&x_full[s], 1), mbar_init(&p_full[s], 8);
		}
		mbar_init(acc_full, 1);
		fence_mbar_init();
	}
	if (warp == 1)
	{
		tmem_alloc(tmem_slot, 512);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t tmem_x = tmem_base;        // X[s] at + 64 s
	const uint32_t tmem_y = tmem_base + 128;  // Y[s] at + 64 s
	const uint32_t tmem_a1 = tmem_base + 256; // dV (MODE 0)
	const uint32_t tmem_a2 = tmem_base + 384; // dK | dQ
	const uint32_t tmem_r1 = tmem_base + 256; // MODE 1: Q and dout as 16-bit pairs, 64 columns each
	const uint32_t tmem_r2 = tmem_base + 320;

	if (warp == 0)
	{
		if (n_it > 0 && elect_one())
		{
			if (MODE == 0)
			{
				mbar_expect_tx(r_full, 2 * FB_R_BYTES);
				tma_load_4d(smem + L::R1_OFF, &tmR1, r_full, 0, hr, r0, b);
				tma_load_4d(smem + L::R1_OFF + FB_R_ATOM, &tmR1, r_full, 64, hr, r0, b);
				tma_load_4d(smem + L::R2_OFF, &tmR2, r_full, 0, hr, r0, b);
				tma_load_4d(smem + L::R2_OFF + FB_R_ATOM, &tmR2, r_full, 64, hr, r0, b);
			}
			int st = 0;
			uint32_t ph = 0;
			for (int it = 0; it < n_it; it++)
			{
				const int ht = MODE == 0 ? hr * G + it / n_blk : hr / G;
				const int t0 = (MODE == 0 ? blk0 + it % n_blk : it) * FB_T;
				uint8_t* const t1 = smem + L::T_OFF + st * 2 * FB_T_BYTES;
				uint8_t* const t2 = t1 + FB_T_BYTES;
				mbar_wait(&t_empty[st], ph ^ 1);
				mbar_expect_tx(&t_full[st], 2 * FB_T_BYTES + (MODE == 0 ? FB_STAT_BYTES : 0));
				tma_load_4d(t1, &tmT1, &t_full[st], 0, ht, t0, b);
				tma_load_4d(t1 + FB_T_ATOM, &tmT1, &t_full[st], 64, ht, t0, b);
				tma_load_4d(t2, &tmT2, &t_full[st], 0, ht, t0, b);
				tma_load_4d(t2 + FB_T_ATOM, &tmT2, &t_full[st], 64, ht, t0, b);
				if (MODE == 0)
					bulk_load_1d(smem + L::STAT_OFF + st * FB_STAT_BYTES, p.stat + (((long long)b * p.H + ht) * p.Sq_r + t0) / 2, FB_STAT_BYTES, &t_full[st]);
				if (++st == NST)
					st = 0, ph ^= 1;
			}
		}
	} else if (warp == 1) {
		// One elected thread issues every MMA of the CTA.  `elect_one()` rather than `lane == 0`, and descriptors as base + constant:
		// with a lane test ptxas cannot prove the operands uniform and wraps every tcgen05.mma in an elect / broadcast / branch loop plus the
		// descriptor arithmetic -- about 12 dependent instructions per MMA, which made the issue rate, not the tensor core, the bound of the
		// first version of this kernel (profiles/r02_ncu_fmha_bwd.txt).
		if (n_it > 0 && elect_one())
		{
			const uint64_t r1_desc = umma_smem_desc(smem_u32(smem + L::R1_OFF), 16, 1024, 2);
			const uint64_t r2_desc = umma_smem_desc(smem_u32(smem + L::R2_OFF), 16, 1024, 2);
			const uint64_t t_kmajor = umma_smem_desc(smem_u32(smem + L::T_OFF), 16, 1024, 2);       // a streamed block as the K-major B operand of X / Y
			const uint64_t t_mnmajor = umma_smem_desc(smem_u32(smem + L::T_OFF), FB_T_ATOM, 1024, 2); // and as the MN-major B operand of the accumulation
			constexpr uint32_t STAGE16 = (2 * FB_T_BYTES) >> 4, T2_16 = FB_T_BYTES >> 4; // descriptor address units (16 bytes)
			int xst = 0, ast = 0; // ring stage of the next X / Y block and of the next accumulation
			uint32_t xph = 0;
			// X = R1 T1^T, Y = R2 T2^T into buffer it & 1: both operands K-major, 8 steps of 16 features over the two 64-wide atoms
			auto issue_xy = [&](const int it) {
				const int s = it & 1;
				// X[s] / Y[s] also hold P / dS of block it - 2 (below): the accumulation that reads them was issued before this point, and the
				// tensor core executes its MMAs in issue order, so no barrier is needed before overwriting them
				mbar_wait(&t_full[xst], xph);
				tc_fence_after();
				const uint64_t t1 = t_kmajor + (uint32_t)xst * STAGE16;
				const uint64_t t2 = t1 + T2_16;
#pragma unroll
				for (int k = 0; k < FB_D / 16; k++)
				{
					const uint32_t ro = ((uint32_t)(k >> 2) * FB_R_ATOM + (uint32_t)(k & 3) * 32) >> 4;
					const uint32_t to = ((uint32_t)(k >> 2) * FB_T_ATOM + (uint32_t)(k & 3) * 32) >> 4;
					if (MODE == 0)
						umma_f16(tmem_x + s * 64, r1_desc + ro, t1 + to, p.idesc_xy, k > 0 ? 1u : 0u);
					else // the resident operand from tensor memory: 8 columns per 16 features
						umma_f16_ts(tmem_x + s * 64, tmem_r1 + k * 8, t1 + to, p.idesc_xy, k > 0 ? 1u : 0u);
				}
#pragma unroll
				for (int k = 0; k < FB_D / 16; k++)
				{
					const uint32_t ro = ((uint32_t)(k >> 2) * FB_R_ATOM + (uint32_t)(k & 3) * 32) >> 4;
					const uint32_t to = ((uint32_t)(k >> 2) * FB_T_ATOM + (uint32_t)(k & 3) * 32) >> 4;
					if (MODE == 0)
						umma_f16(tmem_y + s * 64, r2_desc + ro, t2 + to, p.idesc_xy, k > 0 ? 1u : 0u);
					else
						umma_f16_ts(tmem_y + s * 64, tmem_r2 + k * 8, t2 + to, p.idesc_xy, k > 0 ? 1u : 0u);
				}
				umma_commit(&x_full[s]);
				if (++xst == NST)
					xst = 0, xph ^= 1;
			};
			mbar_wait(r_full, 0);
			tc_fence_after();
			issue_xy(0);
			for (int it = 0; it < n_it; it++)
			{
				if (it + 1 < n_it)
					issue_xy(it + 1); // one block ahead: overlaps the element-wise stage of block it
				const int s = it & 1;
				mbar_wait(&p_full[s], (uint32_t)(it >> 1) & 1);
				tc_fence_after();
				const uint64_t t1 = t_mnmajor + (uint32_t)ast * STAGE16;
				const uint64_t t2 = t1 + T2_16;
				// A = P / dS from tensor memory (written over the first 16 columns of each 32-column half of X[s] / Y[s]: 8 columns per
				// 16 rows of the streamed block), B = the streamed block as an MN-major operand: two 64-wide feature atoms 8 KB apart
				// (LBO), 8-row groups 1 KB apart (SBO), one MMA consumes 16 rows = 2 KB
#pragma unroll
				for (int k = 0; k < FB_T / 16; k++)
				{
					const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
					const uint32_t a_col = (uint32_t)s * 64 + (uint32_t)(k >> 1) * 32 + (uint32_t)(k & 1) * 8;
					if (MODE == 0)
						umma_f16_ts(tmem_a1, tmem_x + a_col, t2 + k * (2048 >> 4), p.idesc_acc, acc);
					umma_f16_ts(tmem_a2, tmem_y + a_col, t1 + k * (2048 >> 4), p.idesc_acc, acc);
				}
				umma_commit(&t_empty[ast]);
				if (it + 1 == n_it)
					umma_commit(acc_full);
				if (++ast == NST)
					ast = 0;
			}
		}
	} else {
		// ------------------------------------------------------------------ element-wise stage and epilogue (warps 2..9)
		const int quarter = warp & 3;     // TMEM lanes [32 * quarter, +32) are the ones this warp may touch
		const int half = (warp - 2) >> 2; // columns [32 * half, +32) of a 64-wide block; [64 * half, +64) of the accumulators
		const int set = half;
		const int row = quarter * 32 + lane;
		const int rr = r0 + row; // MODE 0: key index, MODE 1: query index
		const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
		const int c0 = half * 32;
		float nl = -INFINITY, nd = 0.f; // MODE 1: -lse2 and -delta of this thread's query
		if (MODE == 1)
		{
			const float4 sv = p.stat[(((long long)b * p.H + hr) * p.Sq_r + rr) >> 1]; // Sq_r is a whole number of tiles
			nl = (rr & 1) ? sv.y : sv.x, nd = (rr & 1) ? sv.w : sv.z;
		}
		if (MODE == 1 && n_it > 0)
		{
			// the resident tiles: this thread's row of Q and of dout, 64 of the 128 features each, global -> registers -> tensor memory
			// (two 16-bit elements per column in memory order = the K-major A operand layout); rows past the end are zero
#pragma unroll
			for (int a = 0; a < 2; a++)
			{
				uint32_t r[32];
				if (rr < p.Sq)
				{
					const uint16_t* const src = a == 0 ? (const uint16_t*)p.r1 + b * p.r1_b + (long long)rr * p.r1_s + hr * p.r1_h + half * 64 : (const uint16_t*)p.r2 + b * p.r2_b + (long long)rr * p.r2_s + hr * p.r2_h + half * 64;
#pragma unroll
					for (int i = 0; i < 8; i++)
					{
						const uint4 v = half * 64 + i * 8 < p.D ? __ldg(reinterpret_cast<const uint4*>(src) + i) : make_uint4(0, 0, 0, 0);
						r[i * 4] = v.x, r[i * 4 + 1] = v.y, r[i * 4 + 2] = v.z, r[i * 4 + 3] = v.w;
					}
				} else {
#pragma unroll
					for (int i = 0; i < 32; i++)
						r[i] = 0;
				}
				tmem_st_32x32((a == 0 ? tmem_r1 : tmem_r2) + half * 32 + lane_sel, r);
			}
			tmem_st_wait();
			tc_fence_before();
			__syncwarp();
			if (lane == 0)
				mbar_arrive(r_full);
		}
		int st = 0;
		uint32_t tph = 0;
		for (int it = 0; it < n_it; it++)
		{
			const int s = it & 1;
			const uint32_t ph = (uint32_t)(it >> 1) & 1;
			const int t0 = (MODE == 0 ? blk0 + it % n_blk : it) * FB_T;
			// columns [cmin, cmax) of this row are visible
			int cmin = 0, cmax = FB_T;
			if (MODE == 0)
			{
				if (p.causal)
					cmin = rr - shift - t0; // key rr is seen by queries >= rr - shift
			} else
				cmax = (p.causal ? min(p.Sk, rr + shift + 1) : p.Sk) - t0;
			const bool edge = cmin > c0 || cmax < c0 + 32;
			mbar_wait(&x_full[s], ph);
			if (MODE == 0)
				mbar_wait(&t_full[st], tph); // the stat pairs of this block (the MMA warp has observed this phase already)
			tc_fence_after();
			uint32_t xr[32], yr[32];
			tmem_ld_32x32(tmem_x + s * 64 + c0 + lane_sel, xr);
			tmem_ld_32x32(tmem_y + s * 64 + c0 + lane_sel, yr);
			tmem_ld_wait();
			const float4* const sp = (const float4*)(smem + L::STAT_OFF + st * FB_STAT_BYTES) + (c0 >> 1);
			uint32_t pk[16], dk[16];
#pragma unroll
			for (int i = 0; i < 32; i += 2)
			{
				float l0 = nl, l1 = nl, d0 = nd, d1 = nd;
				if (MODE == 0)
				{
					const float4 sv = sp[i >> 1]; // (-l, -l', -d, -d') of two queries, the same address for the whole warp
					l0 = sv.x, l1 = sv.y, d0 = sv.z, d1 = sv.w;
				}
				float t0f, t1f, u0, u1, g0, g1;
				fma2(t0f, t1f, __uint_as_float(xr[i]), __uint_as_float(xr[i + 1]), p.scale_log2, p.scale_log2, l0, l1);
				float e0 = ex2_approx(t0f);
				float e1 = ex2_approx(t1f);
				if (edge)
				{
					if (c0 + i < cmin || c0 + i >= cmax)
						e0 = 0.f;
					if (c0 + i + 1 < cmin || c0 + i + 1 >= cmax)
						e1 = 0.f;
				}
				fma2(u0, u1, __uint_as_float(yr[i]), __uint_as_float(yr[i + 1]), 1.f, 1.f, d0, d1);
				fma2(g0, g1, e0, e1, u0, u1, 0.f, 0.f);
				if (MODE == 0)
					pk[i >> 1] = pack2(e0, e1, p.is_bf16);
				dk[i >> 1] = pack2(g0, g1, p.is_bf16);
			}
			// P / dS go back into tensor memory as the A operand of the accumulating MMAs (16-bit pairs per column, K-major): over the first
			// 16 columns of this thread's own 32 of X[s] / Y[s], which it has just read -- no shared-memory round trip
			if (MODE == 0)
				tmem_st_32x16(tmem_x + s * 64 + c0 + lane_sel, pk);
			tmem_st_32x16(tmem_y + s * 64 + c0 + lane_sel, dk);
			tmem_st_wait();
			tc_fence_before();
			__syncwarp();
			if (lane == 0)
				mbar_arrive(&p_full[s]);
			if (++st == NST)
				st = 0, tph ^= 1;
		}
		// epilogue: the accumulators -> 16-bit gradients (this thread: one row, 64 of the 128 features)
		if (n_it > 0)
		{
			mbar_wait(acc_full, 0);
			tc_fence_after();
		}
		const int limit = MODE == 0 ? p.Sk : p.Sq;
#pragma unroll
		for (int a = MODE == 0 ? 0 : 1; a < 2; a++)
		{
			uint32_t r[64];
			if (n_it > 0)
			{
				const uint32_t ta = (a == 0 ? tmem_a1 : tmem_a2) + lane_sel + set * 64;
				tmem_ld_32x32(ta, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
				tmem_ld_32x32(ta + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
				tmem_ld_wait();
			} else {
#pragma unroll
				for (int i = 0; i < 64; i++)
					r[i] = 0;
			}
			if (rr < limit)
			{
				const float f = a == 0 ? 1.f : p.scale;
				uint16_t* const orow = a == 0 ? (uint16_t*)p.out1 + b * p.o1_b + (long long)rr * p.o1_s + hr * p.o1_h + set * 64 : (uint16_t*)p.out2 + b * p.o2_b + (long long)rr * p.o2_s + hr * p.o2_h + set * 64;
#pragma unroll
				for (int i = 0; i < 64; i += 8)
				{
					const uint4 v = make_uint4(pack2(__uint_as_float(r[i]) * f, __uint_as_float(r[i + 1]) * f, p.is_bf16), pack2(__uint_as_float(r[i + 2]) * f, __uint_as_float(r[i + 3]) * f, p.is_bf16),
						pack2(__uint_as_float(r[i + 4]) * f, __uint_as_float(r[i + 5]) * f, p.is_bf16), pack2(__uint_as_float(r[i + 6]) * f, __uint_as_float(r[i + 7]) * f, p.is_bf16));
					if (set * 64 + i < p.D)
						*reinterpret_cast<uint4*>(orow + i) = v;
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1)
		tmem_dealloc(tmem_base, 512);
}

template <int MODE>
int launch_bwd(cudaStream_t stream, const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& t1, const CUtensorMap& t2, const FmhaBwdParams& p, int tiles, int heads, int B)
{
	static bool configured_on[64]; // the attribute is per device
	int dev = 0;
	cudaGetDevice(&dev);
	bool& configured = configured_on[dev & 63];
	if (!configured)
	{
		const cudaError_t e = cudaFuncSetAttribute(fmha_bwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaBwdSmem<MODE>::TOTAL);
		if (e != cudaSuccess)
		{
			set_last_error("cudaFuncSetAttribute(fmha_bwd_kernel)", e);
			return -1;
		}
		configured = true;
	}
	fmha_bwd_kernel<MODE><<<dim3(tiles, heads, B), 320, FmhaBwdSmem<MODE>::TOTAL, stream>>>(r1, r2, t1, t2, p);
	count_launch();
	const cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
	{
		set_last_error("fmha_bwd_kernel launch", e);
		return -1;
	}
	return 0;
}

} // namespace

size_t sdpa_backward_f16_workspace_bytes(const SdpaGeom& g, int need_forward)
{
	const size_t sq_r = (size_t)(g.Sq + FB_R - 1) / FB_R * FB_R;
	size_t bytes = (size_t)g.B * g.H * sq_r * 2 * sizeof(float) + 256;
	if (need_forward) // O (16-bit, packed [B, Sq, H, D]) and the log-sum-exp [B, H, Sq]
		bytes += (size_t)g.B * g.Sq * g.H * g.Dv * 2 + 256 + (size_t)g.B * g.H * g.Sq * sizeof(float) + 256;
	return bytes;
}

// g: q / k / v strides, o_* = the strides of dout; dg: the strides of dq / dk / dv in its q_* / k_* / v_* fields.  `out` (strides oo_*)
// and `lse` are the forward's saved outputs; when either is NULL both are recomputed into the workspace with the forward kernel.
// returns 0 on success, 1 when the shape is outside these kernels (D = Dv <= 128 and a multiple of 8, 16-byte aligned strides), < 0 on CUDA errors
int sdpa_backward_f16(cudaStream_t stream, const SdpaGeom& g, const SdpaGeom& dg, int is_bf16, const void* dout, const void* q, const void* k, const void* v, const void* out, long long oo_b, long long oo_s, long long oo_h,
	const float* lse, void* dq, void* dk, void* dv, void* workspace)
{
	if (g.D > FB_D || g.D < 8 || (g.D & 7) || g.Dv != g.D || g.B <= 0 || g.H <= 0 || g.Hk <= 0 || g.H % g.Hk != 0 || g.Sq <= 0 || g.Sk <= 0 || !encode_init())
		return 1;
	for (const void* ptr : { (const void*)dq, (const void*)dk, (const void*)dv, dout })
		if (((uintptr_t)ptr) & 15)
			return 1;
	for (const long long st : { dg.q_b, dg.q_s, dg.q_h, dg.k_b, dg.k_s, dg.k_h, dg.v_b, dg.v_s, dg.v_h, g.o_b, g.o_s, g.o_h })
		if (st & 7)
			return 1;
	CUtensorMap tmQr, tmQt, tmKr, tmKt, tmVr, tmVt, tmGr, tmGt;
	if (!make_map_bshd(&tmQr, q, g.B, g.Sq, g.H, g.D, g.q_b, g.q_s, g.q_h, is_bf16, FB_R) || !make_map_bshd(&tmQt, q, g.B, g.Sq, g.H, g.D, g.q_b, g.q_s, g.q_h, is_bf16, FB_T) ||
		!make_map_bshd(&tmKr, k, g.B, g.Sk, g.Hk, g.D, g.k_b, g.k_s, g.k_h, is_bf16, FB_R) || !make_map_bshd(&tmKt, k, g.B, g.Sk, g.Hk, g.D, g.k_b, g.k_s, g.k_h, is_bf16, FB_T) ||
		!make_map_bshd(&tmVr, v, g.B, g.Sk, g.Hk, g.Dv, g.v_b, g.v_s, g.v_h, is_bf16, FB_R) || !make_map_bshd(&tmVt, v, g.B, g.Sk, g.Hk, g.Dv, g.v_b, g.v_s, g.v_h, is_bf16, FB_T) ||
		!make_map_bshd(&tmGr, dout, g.B, g.Sq, g.H, g.Dv, g.o_b, g.o_s, g.o_h, is_bf16, FB_R) || !make_map_bshd(&tmGt, dout, g.B, g.Sq, g.H, g.Dv, g.o_b, g.o_s, g.o_h, is_bf16, FB_T))
		return 1;
	const int sq_r = (g.Sq + FB_R - 1) / FB_R * FB_R;
	uint8_t* ws = (uint8_t*)workspace;
	float* const stat = (float*)ws;
	ws += (((size_t)g.B * g.H * sq_r * 2 * sizeof(float)) + 255) & ~(size_t)255;
	if (!out || !lse)
	{
		void* const o_ws = ws;
		ws += (((size_t)g.B * g.Sq * g.H * g.Dv * 2) + 255) & ~(size_t)255;
		float* const lse_ws = (float*)ws;
		SdpaGeom fg = g;
		fg.o_h = g.Dv, fg.o_s = (long long)g.H * g.Dv, fg.o_b = (long long)g.Sq * g.H * g.Dv;
		const int rc = sdpa_forward_f16(stream, fg, is_bf16, q, k, v, o_ws, lse_ws);
		if (rc != 0)
			return rc;
		out = o_ws, lse = lse_ws;
		oo_b = fg.o_b, oo_s = fg.o_s, oo_h = fg.o_h;
	}
	if ((((uintptr_t)out) & 15) || (oo_b & 7) || (oo_s & 7) || (oo_h & 7))
		return 1;
	const long long rows = (long long)g.B * g.H * sq_r;
	fmha_bwd_prep_kernel<<<(unsigned)((rows + 15) / 16), 256, 0, stream>>>((const uint16_t*)dout, (const uint16_t*)out, lse, stat, g.B, g.H, g.Sq, sq_r, g.D, is_bf16, g.o_b, g.o_s, g.o_h, oo_b, oo_s, oo_h);
	count_launch();
	FmhaBwdParams p;
	memset(&p, 0, sizeof(p));
	p.H = g.H, p.Hk = g.Hk, p.Sq = g.Sq, p.Sk = g.Sk, p.Sq_r = sq_r, p.D = g.D;
	p.causal = g.is_causal, p.is_bf16 = is_bf16;
	p.scale = g.scale, p.scale_log2 = g.scale * 1.4426950408889634f;
	p.stat = (const float4*)stat;
	p.idesc_xy = umma_instr_desc(is_bf16 ? 1 : 0, 0, 0, FB_R, FB_T);
	p.idesc_acc = umma_instr_desc(is_bf16 ? 1 : 0, 0, 1, FB_R, FB_D);
	// dK, dV: one CTA per 128 keys of a (batch, key head)
	p.out1 = dv, p.o1_b = dg.v_b, p.o1_s = dg.v_s, p.o1_h = dg.v_h;
	p.out2 = dk, p.o2_b = dg.k_b, p.o2_s = dg.k_s, p.o2_h = dg.k_h;
	int rc = launch_bwd<0>(stream, tmKr, tmVr, tmQt, tmGt, p, (g.Sk + FB_R - 1) / FB_R, g.Hk, g.B);
	if (rc != 0)
		return rc;
	// dQ: one CTA per 128 queries of a (batch, query head)
	p.out1 = 0;
	p.out2 = dq, p.o2_b = dg.q_b, p.o2_s = dg.q_s, p.o2_h = dg.q_h;
	p.r1 = q, p.r1_b = g.q_b, p.r1_s = g.q_s, p.r1_h = g.q_h;
	p.r2 = dout, p.r2_b = g.o_b, p.r2_s = g.o_s, p.r2_h = g.o_h;
	return launch_bwd<1>(stream, tmQr, tmGr, tmKt, tmVt, p, (g.Sq + FB_R - 1) / FB_R, g.H, g.B);
}

} // namespace sm100
