// sm100_umma_wgrad.cuh -- filter gradient of convolutions with FEW output filters (K <= 64), e.g. the 3x3 convolutions of
// the ResNet stem (32 -> 32, 32 -> 64 at 112 x 112) and of layer 1 (64 -> 64 at 56 x 56).
//
// The generic kernel (sm100_umma_gemm.cuh) computes dW^T per filter tap as D[K, C] = dY^T[K, pixels] * X_tap[pixels, C]:
// with K = 32 or 64 only a quarter / half of the 128 UMMA rows carry work.  Here the operands swap sides and several
// taps share one MMA:
//
//   D[(tap, c), k] = sum_pixels X_tap[pixel, c] * dY[pixel, k]        M = taps_per_tile * C (<= 128), N = K
//
//   A (M side)  taps_per_tile im2col-mode TMA boxes {32 channels, 32 pixels} per 32 input channels, each with its own
//               filter-tap offset, laid side by side as the 32-wide MN atoms of one MN-major 128 x 32 tile
//   B (N side)  dY[pixels, K] as an MN-major operand: K / 32 plain 2-D boxes {32 filters, 32 pixels}
//
// grid = (M tiles over the R*S*C filter positions, split-K over the pixels); the epilogue stores its tile into
// out[k, tap, c] + split * split_out_stride (for a fixed k the 32 lanes of a warp hit 32 consecutive floats): dW itself when
// there is one split and nothing to accumulate, otherwise one scratch slice per split that splitk_reduce_kernel adds up in a
// fixed order (deterministic: no red.global.add).
#pragma once
#include "sm100_umma_gemm.cuh"

namespace sm100 {

struct WgradTapsParams {
	int C, K;           // input channels (multiple of 32, <= 128), filters (<= BN)
	int taps;           // R * S
	int taps_per_tile;  // taps that share one 128-row M tile (taps_per_tile * C <= 128)
	int k_iters;        // 32-pixel blocks over N * P * Q
	int splits;         // gridDim.y
	int P, Q, stride_h, stride_w, base_h, base_w;
	unsigned short tap_off_h[UMMA_MAX_TAPS], tap_off_w[UMMA_MAX_TAPS];
	float* out;         // dW [K, R, S, C], or the first scratch slice of the same layout
	long long split_out_stride;
	long long rsc;
	uint32_t idesc, mn_lbo, mn_sbo, mn_layout;
};

template <int BN, int STAGES>
struct WgradTapsSmem {
	static constexpr int A_BYTES = UMMA_BLOCK_M * UMMA_BLOCK_K * 4;
	static constexpr int B_BYTES = BN * UMMA_BLOCK_K * 4;
	static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
	static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
	static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 2) umma_wgrad_taps_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgradTapsParams p)
{
	using S = WgradTapsSmem<BN, STAGES>;
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint64_t* full_bar = (uint64_t*)(smem + S::BAR_OFFSET);
	uint64_t* empty_bar = full_bar + STAGES;
	uint64_t* tmem_full_bar = empty_bar + STAGES;
	uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int tap0 = blockIdx.x * p.taps_per_tile;
	const int nt = min(p.taps_per_tile, p.taps - tap0); // taps of this tile
	const int c32 = p.C >> 5;
	const int per = (p.k_iters + p.splits - 1) / p.splits;
	const int it_begin = blockIdx.y * per;
	const int it_end = min(p.k_iters, it_begin + per);
	const int n_it = it_end - it_begin;
	if (n_it <= 0)
		return;

	if (warp == 0 && lane == 0)
	{
		tma_prefetch_desc(&tmX);
		tma_prefetch_desc(&tmG);
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
		if (elect_one())
		{
			int stage = 0;
			uint32_t phase = 0;
			const uint32_t tx_bytes = (uint32_t)(nt * c32 + BN / 32) * 4096u;
			for (int it = it_begin; it < it_end; it++)
			{
				mbar_wait(&empty_bar[stage], phase ^ 1);
				uint8_t* sA = smem + stage * S::STAGE_BYTES;
				uint8_t* sB = sA + S::A_BYTES;
				mbar_expect_tx(&full_bar[stage], tx_bytes);
				const int pix = it * UMMA_BLOCK_K;
				const int q = pix % p.Q;
				const int t = pix / p.Q;
				const int b_w = q * p.stride_w + p.base_w;
				const int b_h = (t % p.P) * p.stride_h + p.base_h;
				const int b_n = t / p.P;
#pragma unroll
				for (int j = 0; j < BN / 32; j++)
					tma_load_2d(sB + j * 4096, &tmG, &full_bar[stage], 32 * j, pix);
				for (int tl = 0; tl < nt; tl++)
					for (int j = 0; j < c32; j++)
						tma_load_im2col_4d(sA + (tl * c32 + j) * 4096, &tmX, &full_bar[stage], 32 * j, b_w, b_h, b_n, p.tap_off_w[tap0 + tl], p.tap_off_h[tap0 + tl]);
				if (++stage == STAGES) { stage = 0; phase ^= 1; }
			}
		}
	} else if (warp == 1) {
		// one elected thread issues the MMAs, operands in uniform registers (see sm100_umma_persistent.cuh)
		if (elect_one())
		{
			const uint32_t smem_base = smem_u32(smem);
			const uint64_t a_desc0 = umma_smem_desc(smem_base, p.mn_lbo, p.mn_sbo, p.mn_layout);
			const uint64_t b_desc0 = umma_smem_desc(smem_base + S::A_BYTES, p.mn_lbo, p.mn_sbo, p.mn_layout);
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
					umma_tf32(tmem_base, da0 + k * (1024 >> 4), db0 + k * (1024 >> 4), idesc, (it > 0 || k > 0) ? 1u : 0u);
				umma_commit(&empty_bar[stage]);
				if (it == n_it - 1)
					umma_commit(tmem_full_bar);
				if (++stage == STAGES) { stage = 0; phase ^= 1; }
			}
		}
	} else {
		const int quarter = warp & 3;
		const int m = quarter * 32 + lane; // (local tap, channel)
		mbar_wait(tmem_full_bar, 0);
		tc_fence_after();
		const bool row_ok = m < nt * p.C;
		float* const o = p.out + (long long)blockIdx.y * p.split_out_stride + (long long)tap0 * p.C + m;
#pragma unroll 1
		for (int c = 0; c < BN / 32; c++)
		{
			uint32_t r[32];
			tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
			tmem_ld_wait();
			if (row_ok)
			{
#pragma unroll
				for (int i = 0; i < 32; i++)
					if (c * 32 + i < p.K)
						o[(long long)(c * 32 + i) * p.rsc] = __uint_as_float(r[i]);
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1)
		tmem_dealloc(tmem_base, BN);
}

} // namespace sm100
