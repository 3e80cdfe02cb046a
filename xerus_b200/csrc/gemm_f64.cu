// FP64 GEMM for sm_100a on the double-precision tensor pipe (DMMA, mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).
//
// Replaces cblas_dgemm / dgemv / dger behind blasWrapper::matrix_matrix_product
// (reference: src/xerus/blasLapackWrapper.cpp:149-195; row-major, beta = 0 there, general beta here because the
// blocked QR / Cholesky trailing updates reuse this kernel).
//
// Tiling: CTA = 4 warps (2 x 2), warp tile (WMF*8) x (WNF*8) built from 8x8x4 DMMA fragments, BK = 16.
// Shared-memory tiles are stored k-major with a +4 double pad so that the fragment loads (lane -> (k = lane%4,
// row = lane/4)) are bank-conflict free; global->shared goes through registers with the next tile prefetched
// while the current one is multiplied (one __syncthreads per k-tile, two smem stages).
#include "xb_internal.cuh"

namespace xb {

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	             : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct GemmArgs {
	double* C; const double* A; const double* B;
	long long ldc, lda, ldb, strideC, strideA, strideB;
	int m, n, k;
	int transA, transB;
	double alpha, beta;
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 128;

template <int WMF, int WNF>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f64_kernel(const GemmArgs g) {
	constexpr int BM = 2 * WMF * 8, BN = 2 * WNF * 8, BK = GEMM_BK;
	constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
	constexpr int A_PER_THREAD = BM * BK / GEMM_THREADS, B_PER_THREAD = BN * BK / GEMM_THREADS;
	__shared__ double As[2][BK][LDA_S];
	__shared__ double Bs[2][BK][LDB_S];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wm = warp >> 1, wn = warp & 1;
	const int grp = lane >> 2, tig = lane & 3;
	const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
	const double* __restrict__ A = g.A + (long long)blockIdx.z * g.strideA;
	const double* __restrict__ B = g.B + (long long)blockIdx.z * g.strideB;
	double* __restrict__ C = g.C + (long long)blockIdx.z * g.strideC;

	double acc[WMF][WNF][2];
#pragma unroll
	for (int i = 0; i < WMF; ++i)
#pragma unroll
		for (int j = 0; j < WNF; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

	double ra[A_PER_THREAD], rb[B_PER_THREAD];
	const int nt = (g.k + BK - 1) / BK;

	auto load_tiles = [&](int t) {
		const int k0 = t * BK;
#pragma unroll
		for (int it = 0; it < A_PER_THREAD; ++it) {
			const int e = tid + it * GEMM_THREADS;
			int i, kk;
			if (g.transA) { kk = e / BM; i = e % BM; } else { i = e / BK; kk = e % BK; }
			const int gi = m0 + i, gk = k0 + kk;
			double v = 0.0;
			if (gi < g.m && gk < g.k) v = g.transA ? A[(long long)gk * g.lda + gi] : A[(long long)gi * g.lda + gk];
			ra[it] = v;
		}
#pragma unroll
		for (int it = 0; it < B_PER_THREAD; ++it) {
			const int e = tid + it * GEMM_THREADS;
			int j, kk;
			if (g.transB) { j = e / BK; kk = e % BK; } else { kk = e / BN; j = e % BN; }
			const int gj = n0 + j, gk = k0 + kk;
			double v = 0.0;
			if (gj < g.n && gk < g.k) v = g.transB ? B[(long long)gj * g.ldb + gk] : B[(long long)gk * g.ldb + gj];
			rb[it] = v;
		}
	};
	auto store_tiles = [&](int buf) {
#pragma unroll
		for (int it = 0; it < A_PER_THREAD; ++it) {
			const int e = tid + it * GEMM_THREADS;
			int i, kk;
			if (g.transA) { kk = e / BM; i = e % BM; } else { i = e / BK; kk = e % BK; }
			As[buf][kk][i] = ra[it];
		}
#pragma unroll
		for (int it = 0; it < B_PER_THREAD; ++it) {
			const int e = tid + it * GEMM_THREADS;
			int j, kk;
			if (g.transB) { j = e / BK; kk = e % BK; } else { kk = e / BN; j = e % BN; }
			Bs[buf][kk][j] = rb[it];
		}
	};

	if (nt > 0) { load_tiles(0); store_tiles(0); }
	__syncthreads();

	for (int t = 0; t < nt; ++t) {
		const int buf = t & 1;
		if (t + 1 < nt) load_tiles(t + 1);
#pragma unroll
		for (int k4 = 0; k4 < BK / 4; ++k4) {
			double af[WMF], bf[WNF];
#pragma unroll
			for (int i = 0; i < WMF; ++i) af[i] = As[buf][k4 * 4 + tig][wm * WMF * 8 + i * 8 + grp];
#pragma unroll
			for (int j = 0; j < WNF; ++j) bf[j] = Bs[buf][k4 * 4 + tig][wn * WNF * 8 + j * 8 + grp];
#pragma unroll
			for (int i = 0; i < WMF; ++i)
#pragma unroll
				for (int j = 0; j < WNF; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
		}
		if (t + 1 < nt) store_tiles(buf ^ 1);
		__syncthreads();
	}

#pragma unroll
	for (int i = 0; i < WMF; ++i) {
		const int row = m0 + wm * WMF * 8 + i * 8 + grp;
		if (row >= g.m) continue;
#pragma unroll
		for (int j = 0; j < WNF; ++j) {
			const int col = n0 + wn * WNF * 8 + j * 8 + tig * 2;
#pragma unroll
			for (int c = 0; c < 2; ++c) {
				if (col + c < g.n) {
					double* p = C + (long long)row * g.ldc + col + c;
					double v = g.alpha * acc[i][j][c];
					if (g.beta != 0.0) v += g.beta * (*p);
					*p = v;
				}
			}
		}
	}
}

void gemm_batched(double* C, size_t ldc, size_t strideC, size_t m, size_t n, double alpha, const double* A, size_t lda,
                  size_t strideA, bool transA, size_t k, const double* B, size_t ldb, size_t strideB, bool transB,
                  double beta, size_t batch) {
	if (m == 0 || n == 0 || batch == 0) return;
	ProfScope prof("gemm");
	XB_REQUIRE(m <= 0x7fffffffULL && n <= 0x7fffffffULL && k <= 0x7fffffffULL, "Dimension too large for GEMM");
	XB_REQUIRE(batch <= 65535, "batch too large");
	GemmArgs g;
	g.C = C; g.A = A; g.B = B;
	g.ldc = (long long)ldc; g.lda = (long long)lda; g.ldb = (long long)ldb;
	g.strideC = (long long)strideC; g.strideA = (long long)strideA; g.strideB = (long long)strideB;
	g.m = int(m); g.n = int(n); g.k = int(k);
	g.transA = transA; g.transB = transB; g.alpha = alpha; g.beta = beta;
	const size_t tiles64 = ((m + 63) / 64) * ((n + 63) / 64) * batch;
	const bool small = ctx().gemm_force_small || tiles64 < size_t(ctx().num_sms);
	if (small) {
		dim3 grid(unsigned((n + 31) / 32), unsigned((m + 31) / 32), unsigned(batch));
		XB_REQUIRE(grid.y <= 65535, "m too large for the small-tile GEMM");
		gemm_f64_kernel<2, 2><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
	} else {
		dim3 grid(unsigned((n + 63) / 64), unsigned((m + 63) / 64), unsigned(batch));
		XB_REQUIRE(grid.y <= 65535, "m too large for the GEMM grid");
		gemm_f64_kernel<4, 4><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
	}
	XB_LAUNCH_CHECK();
}

void gemm(double* C, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, bool transA, size_t k,
          const double* B, size_t ldb, bool transB, double beta) {
	gemm_batched(C, ldc, 0, m, n, alpha, A, lda, 0, transA, k, B, ldb, 0, transB, beta, 1);
}

} // namespace xb
