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
	// optional scattered output (beta = 0): row i goes to Cblk[i / rpb] + (i % rpb) * ldc.  The blocks may live on peer GPUs
	// (bond-split application: the epilogue of the last contraction writes each rank's row block straight into that rank's
	// receive buffer over NVLink, so the transfer runs while the remaining tiles are still being multiplied).
	double* Cblk[8];
	int rpb;
	// optional replicated output (beta = 0): every element is written to Cblk[0 .. rep) — the row-split bond application stores
	// its row block of the result into the result area of every rank (the all-gather happens in the epilogue, over NVLink)
	int rep;
	// tall outputs (TT unfoldings: m = prod n_i reaches millions of rows): the m tiles go on grid.x, whose limit is 2^31 - 1
	// (grid.y stops at 65535), the n tiles on grid.y
	int swap_xy;
};
__device__ __forceinline__ double* gemm_out_row(const GemmArgs& g, double* C, const int row) {
	return g.rpb ? g.Cblk[row / g.rpb] + (long long)(row % g.rpb) * g.ldc : C + (long long)row * g.ldc;
}

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 128;

template <int WMF, int WNF, bool REP = false>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f64_kernel(const GemmArgs g) {
	constexpr int BM = 2 * WMF * 8, BN = 2 * WNF * 8, BK = GEMM_BK;
	constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
	constexpr int A_PER_THREAD = BM * BK / GEMM_THREADS, B_PER_THREAD = BN * BK / GEMM_THREADS;
	__shared__ double As[2][BK][LDA_S];
	__shared__ double Bs[2][BK][LDB_S];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wm = warp >> 1, wn = warp & 1;
	const int grp = lane >> 2, tig = lane & 3;
	const int m0 = (g.swap_xy ? blockIdx.x : blockIdx.y) * BM, n0 = (g.swap_xy ? blockIdx.y : blockIdx.x) * BN;
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
					double v = g.alpha * acc[i][j][c];
					if (REP) {
						for (int dst = 0; dst < g.rep; ++dst) g.Cblk[dst][(long long)row * g.ldc + col + c] = v;
						continue;
					}
					double* p = gemm_out_row(g, C, row) + col + c;
					if (g.beta != 0.0) v += g.beta * (*p);
					*p = v;
				}
			}
		}
	}
}


// ---- large-tile kernel: 128 x 128 CTA tile, cp.async pipeline ---------------------------------------------------------------
// The 64 x 64 kernel above moves 16 bytes from L2 per 8 x 8 x 4 DMMA-tile of work: with the whole GPU multiplying, that is
// 4.6 TB/s of L2 -> SM traffic at the tensor pipe's peak, more than the L2 delivers, and it is why the bond-512 contractions of
// the DMRG local apply ran at 46 % of the measured DGEMM peak.  Here a CTA of 8 warps (2 x 4, warp tile 64 x 32 = 8 x 4 DMMA
// fragments, 128 accumulator registers per thread) owns a 128 x 128 tile: half the L2 traffic per flop, 16-byte cp.async.cg
// copies straight into shared memory (no register staging), three stages of BK = 16.  An operand whose k index is
// contiguous in global memory is stored [row][k] with a row stride of 20 doubles, one whose row/column index is contiguous
// is stored [k][row] with a stride of 132: in both layouts the DMMA fragment loads (lane -> (row = lane / 4, k = lane % 4))
// of a half-warp hit 16 distinct 8-byte banks.  The k loop adds the products in the same order as the kernel above (ascending
// k, one accumulator per output element), so both kernels and all four transposition cases give bit-identical results.
constexpr int BIG_BM = 128, BIG_BN = 128, BIG_BK = 16, BIG_STAGES = 3, BIG_THREADS = 256;
constexpr int BIG_LDK = BIG_BK + 4;       // [row][k] layout
constexpr int BIG_LDR = BIG_BM + 4;       // [k][row] layout
constexpr int BIG_TILE = (BIG_BM * BIG_LDK > BIG_BK * BIG_LDR) ? BIG_BM * BIG_LDK : BIG_BK * BIG_LDR;   // doubles per operand and stage

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src, const bool valid) {
	const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
	const int bytes = valid ? 16 : 0;      // src-size 0: the 16 destination bytes are zero-filled, nothing is read
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

// KMINOR: element (row r of the tile, k) sits at base[r * ld + k] in global memory (k contiguous) -> shared [row][k];
// otherwise at base[k * ld + r] (r contiguous) -> shared [k][row].  rows = valid tile rows, kk = valid k entries (both even
// or the chunk grid is aligned with them: the host checks that).
template <bool KMINOR>
__device__ __forceinline__ void big_load_tile(double* __restrict__ sm, const double* __restrict__ base, const long long ld,
                                              const int rows, const int kk, const int tid) {
#pragma unroll
	for (int it = 0; it < BIG_BM * BIG_BK / 2 / BIG_THREADS; ++it) {
		const int c = tid + it * BIG_THREADS;
		if (KMINOR) {
			const int r = c >> 3, k = (c & 7) * 2;
			const bool ok = r < rows && k < kk;
			cp_async16(sm + r * BIG_LDK + k, ok ? base + (long long)r * ld + k : base, ok);
		} else {
			const int k = c >> 6, r = (c & 63) * 2;
			const bool ok = r < rows && k < kk;
			cp_async16(sm + k * BIG_LDR + r, ok ? base + (long long)k * ld + r : base, ok);
		}
	}
}

// REP: replicated output (GemmArgs::rep destinations) — a separate instantiation so that the ordinary epilogue stays as it was
template <bool TA, bool TB, bool REP>
__global__ void __launch_bounds__(BIG_THREADS, 1) gemm_f64_big_kernel(const GemmArgs g, const int vec_store) {
	extern __shared__ __align__(16) double big_smem[];
	double* As = big_smem;                               // [STAGES][BIG_TILE]
	double* Bs = big_smem + BIG_STAGES * BIG_TILE;       // [STAGES][BIG_TILE]
	constexpr bool A_KMINOR = !TA;                       // A stored m x k row-major: k contiguous
	constexpr bool B_KMINOR = TB;                        // B stored n x k row-major when transposed

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wm = warp >> 2, wn = warp & 3;             // 2 x 4 warps, warp tile 64 x 32
	const int grp = lane >> 2, tig = lane & 3;
	const int m0 = (g.swap_xy ? blockIdx.x : blockIdx.y) * BIG_BM, n0 = (g.swap_xy ? blockIdx.y : blockIdx.x) * BIG_BN;
	const double* __restrict__ A = g.A + (long long)blockIdx.z * g.strideA;
	const double* __restrict__ B = g.B + (long long)blockIdx.z * g.strideB;
	double* __restrict__ C = g.C + (long long)blockIdx.z * g.strideC;
	const int mrows = min(BIG_BM, g.m - m0), ncols = min(BIG_BN, g.n - n0);
	const double* Abase = TA ? A + m0 : A + (long long)m0 * g.lda;
	const double* Bbase = TB ? B + (long long)n0 * g.ldb : B + n0;

	double acc[8][4][2];
#pragma unroll
	for (int i = 0; i < 8; ++i)
#pragma unroll
		for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

	const int nt = (g.k + BIG_BK - 1) / BIG_BK;
	auto load_stage = [&](const int t) {
		const int s = t % BIG_STAGES, k0 = t * BIG_BK, kk = min(BIG_BK, g.k - k0);
		big_load_tile<A_KMINOR>(As + s * BIG_TILE, TA ? Abase + (long long)k0 * g.lda : Abase + k0, g.lda, mrows, kk, tid);
		big_load_tile<B_KMINOR>(Bs + s * BIG_TILE, TB ? Bbase + k0 : Bbase + (long long)k0 * g.ldb, g.ldb, ncols, kk, tid);
	};
#pragma unroll
	for (int s = 0; s < BIG_STAGES - 1; ++s) {
		if (s < nt) load_stage(s);
		asm volatile("cp.async.commit_group;\n" ::: "memory");
	}
	for (int t = 0; t < nt; ++t) {
		asm volatile("cp.async.wait_group %0;\n" :: "n"(BIG_STAGES - 2) : "memory");
		__syncthreads();                                 // tile t has landed for everybody; stage (t - 1) % STAGES is free again
		if (t + BIG_STAGES - 1 < nt) load_stage(t + BIG_STAGES - 1);
		asm volatile("cp.async.commit_group;\n" ::: "memory");
		const double* as = As + (t % BIG_STAGES) * BIG_TILE;
		const double* bs = Bs + (t % BIG_STAGES) * BIG_TILE;
#pragma unroll
		for (int k4 = 0; k4 < BIG_BK / 4; ++k4) {
			double af[8], bf[4];
#pragma unroll
			for (int i = 0; i < 8; ++i)
				af[i] = A_KMINOR ? as[(wm * 64 + i * 8 + grp) * BIG_LDK + k4 * 4 + tig] : as[(k4 * 4 + tig) * BIG_LDR + wm * 64 + i * 8 + grp];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				bf[j] = B_KMINOR ? bs[(wn * 32 + j * 8 + grp) * BIG_LDK + k4 * 4 + tig] : bs[(k4 * 4 + tig) * BIG_LDR + wn * 32 + j * 8 + grp];
#pragma unroll
			for (int i = 0; i < 8; ++i)
#pragma unroll
				for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
		}
	}
	asm volatile("cp.async.wait_group 0;\n" ::: "memory");

#pragma unroll
	for (int i = 0; i < 8; ++i) {
		const int row = m0 + wm * 64 + i * 8 + grp;
		if (row >= g.m) continue;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int col = n0 + wn * 32 + j * 8 + tig * 2;
			if (REP) {
				const long long off = (long long)row * g.ldc + col;
				if (vec_store && col + 1 < g.n) {
					const double2 v = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
					for (int dst = 0; dst < g.rep; ++dst) *reinterpret_cast<double2*>(g.Cblk[dst] + off) = v;
				} else {
					for (int c = 0; c < 2; ++c) if (col + c < g.n) for (int dst = 0; dst < g.rep; ++dst) g.Cblk[dst][off + c] = g.alpha * acc[i][j][c];
				}
				continue;
			}
			double* p = gemm_out_row(g, C, row) + col;
			if (vec_store && col + 1 < g.n) {
				double2 v = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
				if (g.beta != 0.0) { const double2 o = *reinterpret_cast<const double2*>(p); v.x += g.beta * o.x; v.y += g.beta * o.y; }
				*reinterpret_cast<double2*>(p) = v;
			} else {
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					if (col + c < g.n) {
						double v = g.alpha * acc[i][j][c];
						if (g.beta != 0.0) v += g.beta * p[c];
						p[c] = v;
					}
				}
			}
		}
	}
}

template <bool TA, bool TB, bool REP>
static void launch_big_impl(const GemmArgs& g, const dim3 grid, const int vec_store) {
	constexpr size_t smem = size_t(2) * BIG_STAGES * BIG_TILE * sizeof(double);
	static bool attr = false;
	if (!attr) {
		XB_CUDA(cudaFuncSetAttribute(gemm_f64_big_kernel<TA, TB, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
		attr = true;
	}
	gemm_f64_big_kernel<TA, TB, REP><<<grid, BIG_THREADS, smem, ctx().stream>>>(g, vec_store);
}
template <bool TA, bool TB>
static void launch_big(const GemmArgs& g, const dim3 grid, const int vec_store) {
	if (g.rep) launch_big_impl<TA, TB, true>(g, grid, vec_store); else launch_big_impl<TA, TB, false>(g, grid, vec_store);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static thread_local const GemmScatter* tl_scatter = nullptr;

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
	g.rpb = 0; g.swap_xy = 0; g.rep = 0;
	for (int i = 0; i < 8; ++i) g.Cblk[i] = nullptr;
	auto make_grid = [&](size_t tile) {
		const size_t tn = (n + tile - 1) / tile, tm = (m + tile - 1) / tile;
		g.swap_xy = tm > 65535 ? 1 : 0;
		XB_REQUIRE((g.swap_xy ? tn : tm) <= 65535, "both dimensions of the GEMM output are too large for the launch grid");
		return g.swap_xy ? dim3(unsigned(tm), unsigned(tn), unsigned(batch)) : dim3(unsigned(tn), unsigned(tm), unsigned(batch));
	};
	bool blocks_aligned = true;
	if (tl_scatter && tl_scatter->replicate > 0) {
		XB_REQUIRE(beta == 0.0 && batch == 1 && tl_scatter->replicate <= 8, "replicated GEMM output: beta = 0, one problem, at most 8 destinations");
		g.rep = tl_scatter->replicate;
		for (int i = 0; i < 8; ++i) { g.Cblk[i] = tl_scatter->blk[i]; if (g.Cblk[i] && (reinterpret_cast<uintptr_t>(g.Cblk[i]) & 15)) blocks_aligned = false; }
		C = g.Cblk[0]; g.C = C;
	} else if (tl_scatter) {
		XB_REQUIRE(beta == 0.0 && batch == 1 && tl_scatter->rows_per_block > 0 && (m + tl_scatter->rows_per_block - 1) / tl_scatter->rows_per_block <= 8,
		           "scattered GEMM output: beta = 0, one problem, at most 8 row blocks");
		g.rpb = int(tl_scatter->rows_per_block);
		for (int i = 0; i < 8; ++i) { g.Cblk[i] = tl_scatter->blk[i]; if (g.Cblk[i] && (reinterpret_cast<uintptr_t>(g.Cblk[i]) & 15)) blocks_aligned = false; }
		C = g.Cblk[0]; g.C = C;
	}
	// large-tile path: at least ~0.8 waves of 128 x 128 tiles, and every 16-byte cp.async chunk aligned and either fully
	// inside or fully outside its operand (the choice never changes the bits of the result, see gemm_f64_big_kernel)
	const size_t tiles128 = ((m + 127) / 128) * ((n + 127) / 128) * batch;
	if (ctx().gemm_big && !ctx().gemm_force_small && tiles128 * 5 >= size_t(ctx().num_sms) * 4 && k >= 64 && k % 2 == 0 &&
	    lda % 2 == 0 && ldb % 2 == 0 && strideA % 2 == 0 && strideB % 2 == 0 && aligned16(A) && aligned16(B) &&
	    (!transA || m % 2 == 0) && (transB || n % 2 == 0)) {
		const dim3 grid = make_grid(128);
		const int vec_store = (ldc % 2 == 0 && strideC % 2 == 0 && aligned16(C) && blocks_aligned) ? 1 : 0;
		if (transA) { if (transB) launch_big<true, true>(g, grid, vec_store); else launch_big<true, false>(g, grid, vec_store); }
		else { if (transB) launch_big<false, true>(g, grid, vec_store); else launch_big<false, false>(g, grid, vec_store); }
		XB_LAUNCH_CHECK();
		return;
	}
	const size_t tiles64 = ((m + 63) / 64) * ((n + 63) / 64) * batch;
	const bool small = ctx().gemm_force_small || tiles64 < size_t(ctx().num_sms);
	if (small) {
		const dim3 grid = make_grid(32);
		if (g.rep) gemm_f64_kernel<2, 2, true><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
		else gemm_f64_kernel<2, 2><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
	} else {
		const dim3 grid = make_grid(64);
		if (g.rep) gemm_f64_kernel<4, 4, true><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
		else gemm_f64_kernel<4, 4><<<grid, GEMM_THREADS, 0, ctx().stream>>>(g);
	}
	XB_LAUNCH_CHECK();
}

void gemm_scatter(const GemmScatter& sc, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, bool transA, size_t k,
                  const double* B, size_t ldb, bool transB) {
	struct Reset { ~Reset() { tl_scatter = nullptr; } } reset;
	tl_scatter = &sc;
	gemm_batched(sc.blk[0], ldc, 0, m, n, alpha, A, lda, 0, transA, k, B, ldb, 0, transB, 0.0, 1);
}

void gemm(double* C, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, bool transA, size_t k,
          const double* B, size_t ldb, bool transB, double beta) {
	gemm_batched(C, ldc, 0, m, n, alpha, A, lda, 0, transA, k, B, ldb, 0, transB, beta, 1);
}

} // namespace xb
