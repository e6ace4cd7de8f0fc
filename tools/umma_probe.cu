// tools/umma_probe.cu -- stand-alone GPU probe for the contraction kernels (not part of the library).
// Usage: umma_probe <test-id>; each id runs in its own process so that a faulting variant cannot poison the rest.
#include "../ccv_b200/csrc/sm100_contract.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using namespace sm100;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static uint32_t rng_state = 12345;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 65536.f; }
static void fill(std::vector<float>& v, float lo, float hi) { for (auto& x : v) x = lo + (hi - lo) * frand(); }
static float* dev(const std::vector<float>& v) { float* d; CK(cudaMalloc(&d, v.size() * 4 + 16)); CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice)); return d; }
static std::vector<float> host(const float* d, size_t n) { std::vector<float> v(n); CK(cudaMemcpy(v.data(), d, n * 4, cudaMemcpyDeviceToHost)); return v; }

static int report(const char* name, const std::vector<float>& got, const std::vector<double>& ref, double tol)
{
	double max_err = 0, max_ref = 0, sum_err = 0; size_t worst = 0;
	for (size_t i = 0; i < ref.size(); i++) {
		double e = fabs((double)got[i] - ref[i]);
		if (!(e == e)) e = 1e30;
		if (e > max_err) max_err = e, worst = i;
		if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
		sum_err += e;
	}
	const int ok = max_err <= tol * (max_ref > 0 ? max_ref : 1);
	printf("%-44s max_err=%.3e (rel to max|ref| %.3e) mean_err=%.3e max|ref|=%.3e worst@%zu got=%g ref=%g %s\n", name, max_err, max_err / (max_ref > 0 ? max_ref : 1), sum_err / ref.size(), max_ref, worst, got[worst], ref[worst], ok ? "PASS" : "FAIL");
	return ok;
}

static int test_gemm(int M, int N, int K, int ta, int tb, int with_bias, int ffma, float lo, float hi)
{
	std::vector<float> a((size_t)M * K), b((size_t)K * N), bias(N), c((size_t)M * N, -7.f);
	fill(a, lo, hi); fill(b, lo, hi); fill(bias, 0, 1);
	// logical A[m][k], B[k][n]; storage per transposition
	std::vector<float> as(a.size()), bs(b.size());
	for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) as[ta ? (size_t)k * M + m : (size_t)m * K + k] = a[(size_t)m * K + k];
	for (int k = 0; k < K; k++) for (int n = 0; n < N; n++) bs[tb ? (size_t)n * K + k : (size_t)k * N + n] = b[(size_t)k * N + n];
	float *da = dev(as), *db = dev(bs), *dbias = dev(bias), *dc = dev(c);
	int rc;
	if (ffma)
		rc = gemm_ffma(0, M, N, K, da, ta ? 1 : K, ta ? M : 1, db, tb ? 1 : N, tb ? K : 1, dc, N, with_bias ? dbias : 0, 0);
	else
		rc = gemm_tf32(0, M, N, K, da, ta ? M : K, ta, db, tb ? K : N, tb, dc, N, with_bias ? dbias : 0, 0);
	CK(cudaDeviceSynchronize());
	std::vector<float> got = host(dc, c.size());
	std::vector<double> ref((size_t)M * N);
	for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double s = with_bias ? bias[n] : 0; for (int k = 0; k < K; k++) s += (double)a[(size_t)m * K + k] * b[(size_t)k * N + n]; ref[(size_t)m * N + n] = s; }
	char name[128]; snprintf(name, sizeof(name), "gemm%s M%d N%d K%d ta%d tb%d bias%d [%g,%g] rc=%d", ffma ? "_ffma" : "_tf32", M, N, K, ta, tb, with_bias, lo, hi, rc);
	return report(name, got, ref, ffma ? 1e-5 : 2e-3);
}

struct ConvCase { int N, H, W, C, K, R, S, st, pad, dil; };
static ConvGeom geom(const ConvCase& c)
{
	ConvGeom g; memset(&g, 0, sizeof(g));
	g.N = c.N, g.H = c.H, g.W = c.W, g.C = c.C, g.K = c.K, g.R = c.R, g.S = c.S;
	g.stride_h = g.stride_w = c.st; g.pad_h0 = g.pad_h1 = g.pad_w0 = g.pad_w1 = c.pad; g.dil_h = g.dil_w = c.dil;
	g.P = (c.H + 2 * c.pad - ((c.R - 1) * c.dil + 1)) / c.st + 1; g.Q = (c.W + 2 * c.pad - ((c.S - 1) * c.dil + 1)) / c.st + 1;
	g.aw = c.C, g.ah = (long long)c.W * c.C, g.an = (long long)c.H * c.W * c.C;
	g.bw = c.K, g.bh = (long long)g.Q * c.K, g.bn = (long long)g.P * g.Q * c.K;
	return g;
}

// mode 0 fprop, 1 dgrad, 2 wgrad
static int test_conv(const ConvCase& cc, int mode, int ffma, float lo, float hi)
{
	ConvGeom g = geom(cc);
	const size_t na = (size_t)g.N * g.H * g.W * g.C, nw = (size_t)g.K * g.R * g.S * g.C, nb = (size_t)g.N * g.P * g.Q * g.K;
	std::vector<float> a(na), w(nw), gb(nb), bias(g.K);
	fill(a, lo, hi); fill(w, lo, hi); fill(gb, lo, hi); fill(bias, 0, 1);
	for (auto& x : w) x /= (g.C * g.R * g.S);
	float *da = dev(a), *dw = dev(w), *dgb = dev(gb), *dbias = dev(bias);
	std::vector<float> got; std::vector<double> ref; int rc;
	if (mode == 0) {
		std::vector<float> o(nb, -7.f); float* dout = dev(o);
		rc = ffma ? conv_fprop_ffma(0, g, 1, da, dw, dbias, dout) : conv_fprop_tf32(0, g, da, dw, dbias, dout);
		CK(cudaDeviceSynchronize()); got = host(dout, nb); ref.assign(nb, 0);
		for (int n = 0; n < g.N; n++) for (int p = 0; p < g.P; p++) for (int q = 0; q < g.Q; q++) for (int k = 0; k < g.K; k++) {
			double s = bias[k];
			for (int r = 0; r < g.R; r++) for (int ss = 0; ss < g.S; ss++) { int h = p * cc.st - cc.pad + r * cc.dil, x = q * cc.st - cc.pad + ss * cc.dil; if (h < 0 || h >= g.H || x < 0 || x >= g.W) continue;
				for (int c = 0; c < g.C; c++) s += (double)a[((size_t)(n * g.H + h) * g.W + x) * g.C + c] * w[((size_t)(k * g.R + r) * g.S + ss) * g.C + c]; }
			ref[((size_t)(n * g.P + p) * g.Q + q) * g.K + k] = s; }
	} else if (mode == 1) {
		std::vector<float> o(na, -7.f); float* dout = dev(o);
		rc = ffma ? conv_dgrad_ffma(0, g, 1, dgb, dw, dout) : conv_dgrad_tf32(0, g, dgb, dw, dout);
		CK(cudaDeviceSynchronize()); got = host(dout, na); ref.assign(na, 0);
		for (int n = 0; n < g.N; n++) for (int p = 0; p < g.P; p++) for (int q = 0; q < g.Q; q++) for (int k = 0; k < g.K; k++) {
			double gv = gb[((size_t)(n * g.P + p) * g.Q + q) * g.K + k];
			for (int r = 0; r < g.R; r++) for (int ss = 0; ss < g.S; ss++) { int h = p * cc.st - cc.pad + r * cc.dil, x = q * cc.st - cc.pad + ss * cc.dil; if (h < 0 || h >= g.H || x < 0 || x >= g.W) continue;
				for (int c = 0; c < g.C; c++) ref[((size_t)(n * g.H + h) * g.W + x) * g.C + c] += gv * w[((size_t)(k * g.R + r) * g.S + ss) * g.C + c]; } }
	} else {
		std::vector<float> o(nw, -7.f); float* dout = dev(o);
		rc = ffma ? conv_wgrad_ffma(0, g, 1, dgb, da, dout, 0) : conv_wgrad_tf32(0, g, dgb, da, dout, 0);
		CK(cudaDeviceSynchronize()); got = host(dout, nw); ref.assign(nw, 0);
		for (int n = 0; n < g.N; n++) for (int p = 0; p < g.P; p++) for (int q = 0; q < g.Q; q++) for (int k = 0; k < g.K; k++) {
			double gv = gb[((size_t)(n * g.P + p) * g.Q + q) * g.K + k];
			for (int r = 0; r < g.R; r++) for (int ss = 0; ss < g.S; ss++) { int h = p * cc.st - cc.pad + r * cc.dil, x = q * cc.st - cc.pad + ss * cc.dil; if (h < 0 || h >= g.H || x < 0 || x >= g.W) continue;
				for (int c = 0; c < g.C; c++) ref[((size_t)(k * g.R + r) * g.S + ss) * g.C + c] += gv * a[((size_t)(n * g.H + h) * g.W + x) * g.C + c]; } }
	}
	char name[160]; snprintf(name, sizeof(name), "conv%s %s N%d H%d W%d C%d K%d R%d st%d pad%d dil%d rc=%d", ffma ? "_ffma" : "_tf32", mode == 0 ? "fprop" : mode == 1 ? "dgrad" : "wgrad", cc.N, cc.H, cc.W, cc.C, cc.K, cc.R, cc.st, cc.pad, cc.dil, rc);
	return report(name, got, ref, ffma ? 1e-5 : 2e-3);
}

static void rounding_probe()
{
	// out[0][0] = sum_k A[0][k] * B[k][0] with A[0][0] = x, B[0][0] = 1: what does the tensor path do to x's low mantissa bits?
	const int M = 128, N = 64, K = 32;
	const float xs[] = { 1.f + ldexpf(1.f, -11), 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -12), 1.f + ldexpf(1.f, -10) + ldexpf(1.f, -11), 1.f + ldexpf(1.f, -10) - ldexpf(1.f, -20), 1.f + ldexpf(1.f, -12), 3.1415926f };
	for (float x : xs) {
		std::vector<float> a((size_t)M * K, 0.f), b((size_t)N * K, 0.f), c((size_t)M * N, 0.f);
		a[0] = x; b[0] = 1.f;
		float *da = dev(a), *db = dev(b), *dc = dev(c);
		int rc = gemm_tf32(0, M, N, K, da, K, 0, db, K, 1, dc, N, 0, 0);
		CK(cudaDeviceSynchronize());
		std::vector<float> got = host(dc, c.size());
		printf("rounding probe: x=%.10f (0x%08x) -> %.10f (0x%08x) rc=%d\n", x, *(uint32_t*)&x, got[0], *(uint32_t*)&got[0], rc);
	}
}

static void time_gemm(int M, int N, int K, int ta, int tb)
{
	std::vector<float> a((size_t)M * K), b((size_t)K * N); fill(a, -1, 1); fill(b, -1, 1);
	float *da = dev(a), *db = dev(b), *dc; CK(cudaMalloc(&dc, (size_t)M * N * 4));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 3; i++) gemm_tf32(0, M, N, K, da, ta ? M : K, ta, db, tb ? K : N, tb, dc, N, 0, 0);
	CK(cudaDeviceSynchronize());
	const int reps = 20; cudaEventRecord(e0);
	for (int i = 0; i < reps; i++) gemm_tf32(0, M, N, K, da, ta ? M : K, ta, db, tb ? K : N, tb, dc, N, 0, 0);
	cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
	printf("time gemm_tf32 M%d N%d K%d ta%d tb%d: %.3f ms  %.1f TFLOP/s\n", M, N, K, ta, tb, ms, 2.0 * M * N * K / ms * 1e-9);
}

static void time_conv(const ConvCase& cc, int mode)
{
	ConvGeom g = geom(cc);
	const size_t na = (size_t)g.N * g.H * g.W * g.C, nw = (size_t)g.K * g.R * g.S * g.C, nb = (size_t)g.N * g.P * g.Q * g.K;
	std::vector<float> a(na), w(nw), gb(nb); fill(a, 0, 1); fill(w, 0, 1); fill(gb, 0, 1);
	float *da = dev(a), *dw = dev(w), *dgb = dev(gb); float *oa, *ow, *ob; CK(cudaMalloc(&oa, na * 4)); CK(cudaMalloc(&ow, nw * 4)); CK(cudaMalloc(&ob, nb * 4));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	auto run = [&]() { if (mode == 0) conv_fprop_tf32(0, g, da, dw, 0, ob); else if (mode == 1) conv_dgrad_tf32(0, g, dgb, dw, oa); else conv_wgrad_tf32(0, g, dgb, da, ow, 0); };
	for (int i = 0; i < 3; i++) run();
	CK(cudaDeviceSynchronize());
	const int reps = 20; cudaEventRecord(e0); for (int i = 0; i < reps; i++) run(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
	const double flop = 2.0 * g.N * g.P * g.Q * g.K * g.C * g.R * g.S;
	printf("time conv_tf32 %s N%d H%d C%d K%d R%d st%d: %.3f ms  %.1f TFLOP/s\n", mode == 0 ? "fprop" : mode == 1 ? "dgrad" : "wgrad", cc.N, cc.H, cc.C, cc.K, cc.R, cc.st, ms, flop / ms * 1e-9);
}

int main(int argc, char** argv)
{
	const int id = argc > 1 ? atoi(argv[1]) : 0;
	int ok = 1;
	const ConvCase small_s1 = { 2, 12, 12, 32, 64, 3, 3, 1, 1, 1 }, small_s2 = { 2, 12, 12, 64, 64, 3, 3, 2, 1, 1 }, odd = { 3, 13, 9, 96, 160, 3, 3, 1, 1, 1 }, odd_s2 = { 3, 13, 9, 96, 160, 3, 3, 2, 1, 1 };
	const ConvCase pw = { 2, 14, 14, 64, 256, 1, 1, 1, 0, 1 }, dil2 = { 2, 15, 15, 32, 64, 3, 3, 1, 2, 2 }, cfg2 = { 64, 56, 56, 64, 64, 3, 3, 1, 1, 1 }, r50a = { 32, 28, 28, 128, 128, 3, 3, 1, 1, 1 }, k5 = { 1, 11, 11, 32, 32, 5, 5, 1, 2, 1 };
	const ConvCase stem = { 2, 32, 32, 3, 32, 3, 3, 2, 1, 1 };
	switch (id) {
	case 0: ok &= test_gemm(256, 256, 128, 0, 1, 0, 1, -1, 1); ok &= test_gemm(100, 72, 50, 1, 0, 1, 1, -1, 1); break; // ffma sanity
	case 1: ok &= test_gemm(128, 128, 32, 0, 1, 0, 0, -1, 1); ok &= test_gemm(256, 256, 128, 0, 1, 1, 0, -1, 1); ok &= test_gemm(256, 256, 128, 0, 1, 0, 0, 0, 1); break;
	case 2: ok &= test_gemm(256, 256, 128, 0, 0, 0, 0, -1, 1); break;
	case 3: ok &= test_gemm(256, 256, 128, 1, 1, 0, 0, -1, 1); break;
	case 4: ok &= test_gemm(256, 256, 128, 1, 0, 0, 0, -1, 1); break;
	case 5: ok &= test_gemm(200, 136, 100, 0, 1, 1, 0, -1, 1); ok &= test_gemm(200, 136, 100, 0, 0, 1, 0, -1, 1); ok &= test_gemm(200, 136, 100, 1, 0, 1, 0, -1, 1); ok &= test_gemm(1024, 1000, 2048, 0, 1, 1, 0, -1, 1); ok &= test_gemm(64, 48, 4096, 0, 1, 0, 0, -1, 1); break;
	case 6: ok &= test_conv(small_s1, 0, 1, -1, 1); ok &= test_conv(small_s2, 1, 1, -1, 1); ok &= test_conv(small_s2, 2, 1, -1, 1); ok &= test_conv(stem, 0, 1, -1, 1); ok &= test_conv(stem, 2, 1, -1, 1); ok &= test_conv(dil2, 1, 1, -1, 1); break; // ffma conv sanity
	case 7: ok &= test_conv(small_s1, 0, 0, -1, 1); break;
	case 8: ok &= test_conv(small_s2, 0, 0, -1, 1); ok &= test_conv(odd, 0, 0, -1, 1); ok &= test_conv(odd_s2, 0, 0, -1, 1); ok &= test_conv(dil2, 0, 0, -1, 1); ok &= test_conv(k5, 0, 0, -1, 1); ok &= test_conv(pw, 0, 0, -1, 1); break;
	case 9: ok &= test_conv(small_s1, 1, 0, -1, 1); break;
	case 10: ok &= test_conv(small_s2, 1, 0, -1, 1); ok &= test_conv(odd, 1, 0, -1, 1); ok &= test_conv(odd_s2, 1, 0, -1, 1); ok &= test_conv(dil2, 1, 0, -1, 1); ok &= test_conv(pw, 1, 0, -1, 1); break;
	case 11: ok &= test_conv(small_s1, 2, 0, -1, 1); break;
	case 12: ok &= test_conv(small_s2, 2, 0, -1, 1); ok &= test_conv(odd, 2, 0, -1, 1); ok &= test_conv(odd_s2, 2, 0, -1, 1); ok &= test_conv(dil2, 2, 0, -1, 1); ok &= test_conv(pw, 2, 0, -1, 1); break;
	case 13: rounding_probe(); break;
	case 14: ok &= test_conv(r50a, 0, 0, 0, 1); ok &= test_conv(r50a, 1, 0, 0, 1); ok &= test_conv(r50a, 2, 0, 0, 1); break; // positive data: bias check
	case 15: time_gemm(1024, 1024, 1024, 0, 1); time_gemm(1024, 1024, 1024, 0, 0); time_gemm(4096, 4096, 4096, 0, 1); time_gemm(4096, 4096, 4096, 0, 0); time_gemm(8192, 8192, 8192, 0, 1); time_gemm(200704, 256, 64, 0, 1); time_gemm(200704, 64, 256, 0, 1); break;
	case 16: time_conv(cfg2, 0); time_conv(cfg2, 1); time_conv(cfg2, 2); { ConvCase c = { 256, 56, 56, 64, 64, 3, 3, 1, 1, 1 }; time_conv(c, 0); time_conv(c, 1); time_conv(c, 2); } { ConvCase c = { 256, 14, 14, 256, 256, 3, 3, 1, 1, 1 }; time_conv(c, 0); time_conv(c, 1); time_conv(c, 2); } { ConvCase c = { 256, 56, 56, 128, 128, 3, 3, 2, 1, 1 }; time_conv(c, 0); time_conv(c, 1); time_conv(c, 2); } break;
	case 17: time_gemm(200704, 256, 64, 0, 1); break; // the 1x1 64->256 convolution of ResNet layer1 at N=64 (write-dominated)
	case 18: { ConvCase c = { 64, 56, 56, 64, 64, 3, 3, 1, 1, 1 }; time_conv(c, 0); time_conv(c, 2); } break; // BASELINE configs[1]
	case 19: time_gemm(8192, 8192, 8192, 0, 1); break;
	case 20: { // filter gradients with few filters: the taps-along-M kernel (sm100_umma_wgrad.cuh)
		const ConvCase a = { 2, 12, 12, 32, 32, 3, 3, 1, 1, 1 }, b = { 3, 13, 9, 128, 48, 3, 3, 1, 1, 1 }, c = { 3, 13, 9, 64, 32, 3, 3, 2, 1, 1 }, d = { 5, 28, 28, 32, 64, 3, 3, 1, 1, 1 };
		ok &= test_conv(a, 2, 0, -1, 1); ok &= test_conv(b, 2, 0, -1, 1); ok &= test_conv(c, 2, 0, -1, 1); ok &= test_conv(d, 2, 0, -1, 1); ok &= test_conv(k5, 2, 0, -1, 1);
		ok &= test_conv(small_s1, 2, 0, -1, 1); ok &= test_conv(small_s2, 2, 0, -1, 1); ok &= test_conv(dil2, 2, 0, -1, 1);
	} break;
	case 21: { // ResNet-50 stem / layer-1 3x3 filter gradients at N = 256
		const ConvCase c2 = { 256, 112, 112, 32, 32, 3, 3, 1, 1, 1 }, c3 = { 256, 112, 112, 32, 64, 3, 3, 1, 1, 1 }, l1 = { 256, 56, 56, 64, 64, 3, 3, 1, 1, 1 };
		time_conv(c2, 2); time_conv(c3, 2); time_conv(l1, 2); time_conv(c2, 1); time_conv(c3, 1); time_conv(c2, 0); time_conv(c3, 0);
	} break;
	default: printf("unknown test id\n"); return 1;
	}
	printf("probe %d: %s\n", id, ok ? "ALL PASS" : "SOME FAIL");
	return ok ? 0 : 1;
}
