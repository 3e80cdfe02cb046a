// Dense direct solvers behind blasWrapper::solve (reference: src/xerus/blasLapackWrapper.cpp:542-651):
// Cholesky (dpotrf2 + dpotrs, :593-610) for symmetric matrices with a definite diagonal, LU with partial pivoting
// (dgesv, :570) otherwise.  These serve the per-call drop-in layer at the small/medium sizes the reference's tests
// use; the ALS hot path never densifies its local operator (it is solved matrix-free, see als.cu), so these kernels
// are deliberately simple: one CTA, right-looking, operating in L2-resident global memory.
#include "xb_internal.cuh"

namespace xb {

// out[0] = max entry (signed, as the reference's is_symmetric :501-505), out[1] = max |A - A^T|,
// out[2] = 1 if the diagonal is all > eps or all < -eps (pos_neg_definite_diagonal :519-537) else 0, out[3] = A[0][0]
__global__ void __launch_bounds__(256) sym_probe_kernel(const double* __restrict__ A, const int n, double* __restrict__ part) {
	__shared__ double s_max[8], s_asym[8];
	__shared__ int s_bad[8];
	double mx = -HUGE_VAL, asym = 0.0;
	int bad = 0;
	const bool positive = A[0] > 0.0;
	const double eps = 2.220446049250313e-16;
	// CTA b owns rows b, b + gridDim.x, ...; a warp reads a row segment (coalesced) and the mirrored column segment
	for (int i = blockIdx.x; i < n; i += gridDim.x) {
		const double* row = A + (size_t)i * n;
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			const double v = row[j];
			mx = fmax(mx, v);
			if (j > i) asym = fmax(asym, fabs(v - A[(size_t)j * n + i]));
			if (i == j && i > 0) { if (positive ? (v < eps) : (v > -eps)) bad = 1; }
		}
	}
	for (int o = 16; o > 0; o >>= 1) {
		mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
		bad |= __shfl_xor_sync(0xffffffffu, bad, o);
	}
	if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = mx; s_asym[threadIdx.x >> 5] = asym; s_bad[threadIdx.x >> 5] = bad; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < int(blockDim.x >> 5); ++w) { mx = fmax(mx, s_max[w]); asym = fmax(asym, s_asym[w]); bad |= s_bad[w]; }
		part[blockIdx.x * 4 + 0] = mx; part[blockIdx.x * 4 + 1] = asym; part[blockIdx.x * 4 + 2] = bad ? 0.0 : 1.0;
	}
}
__global__ void sym_probe_final_kernel(const double* __restrict__ part, const int nparts, const double* __restrict__ A, double* __restrict__ out) {
	double mx = -HUGE_VAL, asym = 0.0, ok = 1.0;
	for (int p = threadIdx.x; p < nparts; p += 32) { mx = fmax(mx, part[p * 4]); asym = fmax(asym, part[p * 4 + 1]); ok = fmin(ok, part[p * 4 + 2]); }
	for (int o = 16; o > 0; o >>= 1) {
		mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
		ok = fmin(ok, __shfl_xor_sync(0xffffffffu, ok, o));
	}
	if (threadIdx.x == 0) { out[0] = mx; out[1] = asym; out[2] = ok; out[3] = A[0]; }
}

// In-place upper Cholesky A = U^T U (row-major, upper triangle), then solves U^T U X = B in place.
// status[0] = 0 on success, j+1 if the leading minor of order j+1 is not positive definite.
__global__ void __launch_bounds__(1024) cholesky_solve_kernel(double* __restrict__ A, double* __restrict__ B, const int n, const int nrhs,
                                                             const double sign, int* __restrict__ status) {
	__shared__ double s_d;
	__shared__ int s_fail;
	if (threadIdx.x == 0) s_fail = 0;
	__syncthreads();
	for (int j = 0; j < n; ++j) {
		if (threadIdx.x == 0) {
			const double a = sign * A[(size_t)j * n + j];
			if (!(a > 0.0)) s_fail = j + 1;
			s_d = sqrt(a);
		}
		__syncthreads();
		if (s_fail) { if (threadIdx.x == 0) status[0] = s_fail; return; }
		const double d = s_d;
		for (int c = j + threadIdx.x; c < n; c += blockDim.x) A[(size_t)j * n + c] = (c == j) ? d : sign * A[(size_t)j * n + c] / d;
		__syncthreads();
		// trailing update of the upper triangle: A[i][c] -= sign * U[j][i] * U[j][c]  (i > j, c >= i), in the signed matrix
		{
			const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
			for (int i = j + 1 + ty; i < n; i += nty) {
				const double uji = sign * A[(size_t)j * n + i];
				for (int c = i + tx; c < n; c += 32) A[(size_t)i * n + c] -= uji * A[(size_t)j * n + c];
			}
		}
		__syncthreads();
	}
	// forward substitution U^T Y = sign * B, then back substitution U X = Y ; threads own rhs columns / row sweeps
	for (int j = 0; j < n; ++j) {
		for (int r = threadIdx.x; r < nrhs; r += blockDim.x) B[(size_t)j * nrhs + r] = sign * B[(size_t)j * nrhs + r] / A[(size_t)j * n + j];
		__syncthreads();
		for (size_t e = threadIdx.x; e < (size_t)(n - j - 1) * nrhs; e += blockDim.x) {
			const int i = j + 1 + int(e / nrhs), r = int(e % nrhs);
			B[(size_t)i * nrhs + r] -= sign * A[(size_t)j * n + i] * B[(size_t)j * nrhs + r] * sign;
		}
		__syncthreads();
	}
	for (int j = n - 1; j >= 0; --j) {
		for (int r = threadIdx.x; r < nrhs; r += blockDim.x) B[(size_t)j * nrhs + r] /= A[(size_t)j * n + j];
		__syncthreads();
		for (size_t e = threadIdx.x; e < (size_t)j * nrhs; e += blockDim.x) {
			const int i = int(e / nrhs), r = int(e % nrhs);
			B[(size_t)i * nrhs + r] -= A[(size_t)i * n + j] * B[(size_t)j * nrhs + r];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) status[0] = 0;
}

// LU with partial (row) pivoting in place, then solves for B in place.  status[0] = j+1 if a zero pivot was met.
__global__ void __launch_bounds__(1024) lu_solve_kernel(double* __restrict__ A, double* __restrict__ B, const int n, const int nrhs,
                                                       int* __restrict__ status) {
	__shared__ double s_val[32];
	__shared__ int s_idx[32];
	__shared__ int s_piv;
	__shared__ int s_fail;
	if (threadIdx.x == 0) s_fail = 0;
	__syncthreads();
	for (int j = 0; j < n; ++j) {
		double best = -1.0; int bi = j;
		for (int i = j + threadIdx.x; i < n; i += blockDim.x) { const double v = fabs(A[(size_t)i * n + j]); if (v > best) { best = v; bi = i; } }
		for (int o = 16; o > 0; o >>= 1) {
			const double ov = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
		}
		if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
		__syncthreads();
		if (threadIdx.x == 0) {
			for (int w = 1; w < int(blockDim.x >> 5); ++w) if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
			s_piv = bi;
			if (!(best > 0.0)) s_fail = j + 1;
		}
		__syncthreads();
		if (s_fail) { if (threadIdx.x == 0) status[0] = s_fail; return; }
		const int p = s_piv;
		if (p != j) {
			for (int c = threadIdx.x; c < n; c += blockDim.x) { const double t = A[(size_t)j * n + c]; A[(size_t)j * n + c] = A[(size_t)p * n + c]; A[(size_t)p * n + c] = t; }
			for (int r = threadIdx.x; r < nrhs; r += blockDim.x) { const double t = B[(size_t)j * nrhs + r]; B[(size_t)j * nrhs + r] = B[(size_t)p * nrhs + r]; B[(size_t)p * nrhs + r] = t; }
		}
		__syncthreads();
		const double piv = A[(size_t)j * n + j];
		for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[(size_t)i * n + j] /= piv;
		__syncthreads();
		const int rem = n - j - 1;
		{
			const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
			for (int i = j + 1 + ty; i < n; i += nty) {
				const double lij = A[(size_t)i * n + j];
				for (int c = j + 1 + tx; c < n; c += 32) A[(size_t)i * n + c] -= lij * A[(size_t)j * n + c];
			}
		}
		for (size_t e = threadIdx.x; e < (size_t)rem * nrhs; e += blockDim.x) {
			const int i = j + 1 + int(e / nrhs), r = int(e % nrhs);
			B[(size_t)i * nrhs + r] -= A[(size_t)i * n + j] * B[(size_t)j * nrhs + r];
		}
		__syncthreads();
	}
	for (int j = n - 1; j >= 0; --j) {
		for (int r = threadIdx.x; r < nrhs; r += blockDim.x) B[(size_t)j * nrhs + r] /= A[(size_t)j * n + j];
		__syncthreads();
		for (size_t e = threadIdx.x; e < (size_t)j * nrhs; e += blockDim.x) {
			const int i = int(e / nrhs), r = int(e % nrhs);
			B[(size_t)i * nrhs + r] -= A[(size_t)i * n + j] * B[(size_t)j * nrhs + r];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) status[0] = 0;
}

// ---- blocked Cholesky --------------------------------------------------------------------------------------------------
constexpr int CH_NB = 64;

// In-place upper Cholesky of the nb x nb diagonal block at A (row stride lda).  status[0] = k0 + j + 1 on a non-positive pivot.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, const long long lda, const int nb, const int k0, int* __restrict__ status) {
	__shared__ double U[CH_NB][CH_NB + 1];
	__shared__ int s_fail;
	for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) { const int i = e / nb, c = e % nb; U[i][c] = (c >= i) ? A[(long long)i * lda + c] : 0.0; }
	if (threadIdx.x == 0) s_fail = 0;
	__syncthreads();
	for (int j = 0; j < nb; ++j) {
		const double a = U[j][j];
		if (!(a > 0.0)) { if (threadIdx.x == 0) s_fail = k0 + j + 1; }
		__syncthreads();
		if (s_fail) break;
		const double inv = 1.0 / sqrt(a);
		if (threadIdx.x >= j && threadIdx.x < nb) U[j][threadIdx.x] = (threadIdx.x == j) ? sqrt(a) : U[j][threadIdx.x] * inv;
		__syncthreads();
		// trailing update of the upper triangle: U[i][c] -= U[j][i] * U[j][c]   (i > j, c >= i)
		const int rem = nb - j - 1;
		for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
			const int i = j + 1 + e / rem, c = j + 1 + e % rem;
			if (c >= i) U[i][c] -= U[j][i] * U[j][c];
		}
		__syncthreads();
	}
	if (s_fail) { if (threadIdx.x == 0 && status[0] == 0) status[0] = s_fail; return; }
	for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) { const int i = e / nb, c = e % nb; if (c >= i) A[(long long)i * lda + c] = U[i][c]; }
}

// X = U^-T X (LOWER = true: forward substitution with U^T) or X = U^-1 X (back substitution) for the nb x nb upper block U
// and the nb x ncols matrix X (row stride ldx): one thread per column of X, U broadcast from shared memory.
template <bool LOWER>
__global__ void __launch_bounds__(128) chol_trsm_kernel(const double* __restrict__ Ublk, const long long ldu, double* __restrict__ X, const long long ldx,
                                                       const int nb, const int ncols) {
	__shared__ double U[CH_NB][CH_NB];
	__shared__ double invd[CH_NB];
	for (int e = threadIdx.x; e < CH_NB * CH_NB; e += blockDim.x) {
		const int i = e / CH_NB, c = e % CH_NB;
		U[i][c] = (i < nb && c < nb && c >= i) ? Ublk[(long long)i * ldu + c] : 0.0;
	}
	__syncthreads();
	if (threadIdx.x < CH_NB) invd[threadIdx.x] = (threadIdx.x < nb) ? 1.0 / U[threadIdx.x][threadIdx.x] : 0.0;
	__syncthreads();
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncols) return;
	double a[CH_NB];
#pragma unroll
	for (int i = 0; i < CH_NB; ++i) a[i] = (i < nb) ? X[(long long)i * ldx + c] : 0.0;
	if (LOWER) {
#pragma unroll
		for (int i = 0; i < CH_NB; ++i) {
			const double xi = a[i] * invd[i];
			a[i] = xi;
#pragma unroll
			for (int j = i + 1; j < CH_NB; ++j) a[j] -= U[i][j] * xi;
		}
	} else {
#pragma unroll
		for (int i = CH_NB - 1; i >= 0; --i) {
			const double xi = a[i] * invd[i];
			a[i] = xi;
#pragma unroll
			for (int j = 0; j < i; ++j) a[j] -= U[j][i] * xi;
		}
	}
#pragma unroll
	for (int i = 0; i < CH_NB; ++i) if (i < nb) X[(long long)i * ldx + c] = a[i];
}

static bool cholesky_solve_blocked(double* A, double* B, size_t n, size_t nrhs) {
	cudaStream_t st = ctx().stream;
	int* d_status = static_cast<int*>(dalloc_bytes(sizeof(int)));
	XB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), st));
	// factorisation A = U^T U, right looking
	for (size_t k0 = 0; k0 < n; k0 += CH_NB) {
		const size_t nb = std::min<size_t>(CH_NB, n - k0), n2 = n - k0 - nb;
		double* Akk = A + k0 * n + k0;
		chol_diag_kernel<<<1, 256, 0, st>>>(Akk, (long long)n, int(nb), int(k0), d_status);
		XB_LAUNCH_CHECK();
		if (n2 == 0) break;
		double* A12 = Akk + nb;
		chol_trsm_kernel<true><<<unsigned((n2 + 127) / 128), 128, 0, st>>>(Akk, (long long)n, A12, (long long)n, int(nb), int(n2));
		XB_LAUNCH_CHECK();
		gemm(Akk + nb * n + nb, n, n2, n2, -1.0, A12, n, true, nb, A12, n, false, 1.0);        // A22 -= U12^T U12
	}
	const int st_fact = [&] { int* h = reinterpret_cast<int*>(ctx().h_scratch);
		XB_CUDA(cudaMemcpyAsync(h, d_status, sizeof(int), cudaMemcpyDeviceToHost, st)); XB_CUDA(cudaStreamSynchronize(st)); return h[0]; }();
	dfree(d_status);
	if (st_fact != 0) return false;
	// U^T Y = B (forward), U X = Y (backward), block by block
	for (size_t k0 = 0; k0 < n; k0 += CH_NB) {
		const size_t nb = std::min<size_t>(CH_NB, n - k0), n2 = n - k0 - nb;
		const double* Akk = A + k0 * n + k0;
		double* B1 = B + k0 * nrhs;
		chol_trsm_kernel<true><<<unsigned((nrhs + 127) / 128), 128, 0, st>>>(Akk, (long long)n, B1, (long long)nrhs, int(nb), int(nrhs));
		XB_LAUNCH_CHECK();
		if (n2) gemm(B1 + nb * nrhs, nrhs, n2, nrhs, -1.0, Akk + nb, n, true, nb, B1, nrhs, false, 1.0);     // B2 -= U12^T Y1
	}
	for (size_t kb = (n + CH_NB - 1) / CH_NB; kb-- > 0;) {
		const size_t k0 = kb * CH_NB, nb = std::min<size_t>(CH_NB, n - k0), n2 = n - k0 - nb;
		const double* Akk = A + k0 * n + k0;
		double* B1 = B + k0 * nrhs;
		if (n2) gemm(B1, nrhs, nb, nrhs, -1.0, Akk + nb, n, false, n2, B1 + nb * nrhs, nrhs, false, 1.0);    // B1 -= U12 X2
		chol_trsm_kernel<false><<<unsigned((nrhs + 127) / 128), 128, 0, st>>>(Akk, (long long)n, B1, (long long)nrhs, int(nb), int(nrhs));
		XB_LAUNCH_CHECK();
	}
	return true;
}

static int read_status(int* d_status) {
	Context& c = ctx();
	int* h = reinterpret_cast<int*>(c.h_scratch);
	XB_CUDA(cudaMemcpyAsync(h, d_status, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	return h[0];
}

// ---- blocked LU with partial pivoting ---------------------------------------------------------------------------------
// Right-looking, 64-column panels: the panel is factored by one CTA working in L2-resident global memory (pivot search,
// row swap inside the panel, scaling, rank-1 update), the row swaps are then applied to the rest of the matrix and to B,
// U12 = L11^-1 A12 is one thread per column with L11 in shared memory, and the trailing update A22 -= L21 U12 is the DMMA
// GEMM.  B rides along as extra columns, so only the back substitution is left after the loop.
__global__ void __launch_bounds__(1024) lu_panel_kernel(double* __restrict__ A, const long long n, const int k0, const int nb, int* __restrict__ ipiv,
                                                       int* __restrict__ status) {
	__shared__ double s_val[32];
	__shared__ int s_idx[32];
	__shared__ double s_row[CH_NB];
	__shared__ int s_piv, s_fail;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_fail = 0;
	__syncthreads();
	for (int j = 0; j < nb; ++j) {
		const long long col = k0 + j;
		double best = -1.0; int bi = int(col);
		for (long long i = col + threadIdx.x; i < n; i += blockDim.x) { const double v = fabs(A[i * n + col]); if (v > best) { best = v; bi = int(i); } }
		for (int o = 16; o > 0; o >>= 1) {
			const double ov = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
		}
		if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
		__syncthreads();
		if (threadIdx.x == 0) {
			for (int w = 1; w < 32; ++w) if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
			s_piv = bi;
			ipiv[col] = bi;
			if (!(best > 0.0)) s_fail = int(col) + 1;
		}
		__syncthreads();
		if (s_fail) { if (threadIdx.x == 0 && status[0] == 0) status[0] = s_fail; return; }
		const long long p = s_piv;
		if (threadIdx.x < nb) {
			const long long c = k0 + threadIdx.x;
			const double top = A[col * n + c], bot = A[p * n + c];
			if (p != col) { A[col * n + c] = bot; A[p * n + c] = top; }
			s_row[threadIdx.x] = bot;                        // the pivot row inside the panel (after the swap)
		}
		__syncthreads();
		const double inv = 1.0 / s_row[j];
		// rows below the pivot: l = a / pivot, then a rank-1 update of the panel columns right of j
		for (long long i = col + 1 + warp; i < n; i += 32) {
			double* row = A + i * n + k0;
			const double l = row[j] * inv;
			__syncwarp();
			if (lane == 0) row[j] = l;
			for (int c = j + 1 + lane; c < nb; c += 32) row[c] -= l * s_row[c];
		}
		__syncthreads();
	}
}

__global__ void lu_iota_kernel(int* __restrict__ ipiv, const int n) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ipiv[i] = i;
}

// applies the nb row interchanges of a panel to `cols` columns starting at M (row stride ld): one thread per column
__global__ void lu_swap_kernel(double* __restrict__ M, const long long ld, const int cols, const int* __restrict__ ipiv, const int k0, const int nb) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= cols) return;
	for (int j = 0; j < nb; ++j) {
		const long long r = k0 + j, p = ipiv[r];
		if (p != r) { const double t = M[r * ld + c]; M[r * ld + c] = M[p * ld + c]; M[p * ld + c] = t; }
	}
}

// X = L^-1 X for the unit lower nb x nb block L (row stride ldl) and the nb x ncols matrix X: one thread per column
__global__ void __launch_bounds__(128) lu_trsm_kernel(const double* __restrict__ Lblk, const long long ldl, double* __restrict__ X, const long long ldx,
                                                     const int nb, const int ncols) {
	__shared__ double L[CH_NB][CH_NB + 1];
	for (int e = threadIdx.x; e < CH_NB * CH_NB; e += blockDim.x) {
		const int i = e / CH_NB, c = e % CH_NB;
		L[i][c] = (i < nb && c < i) ? Lblk[(long long)i * ldl + c] : 0.0;
	}
	__syncthreads();
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncols) return;
	double a[CH_NB];
#pragma unroll
	for (int i = 0; i < CH_NB; ++i) a[i] = (i < nb) ? X[(long long)i * ldx + c] : 0.0;
#pragma unroll
	for (int j = 0; j < CH_NB; ++j) {
#pragma unroll
		for (int i = j + 1; i < CH_NB; ++i) a[i] -= L[i][j] * a[j];
	}
#pragma unroll
	for (int i = 0; i < CH_NB; ++i) if (i < nb) X[(long long)i * ldx + c] = a[i];
}

static void lu_solve_blocked(double* A, double* B, size_t n, size_t nrhs) {
	cudaStream_t st = ctx().stream;
	int* d_status = static_cast<int*>(dalloc_bytes(sizeof(int)));
	int* ipiv = static_cast<int*>(dalloc_bytes(n * sizeof(int)));
	XB_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), st));
	// identity pivots: a panel that meets a zero pivot returns early, the kernels queued behind it must stay harmless
	lu_iota_kernel<<<unsigned(std::min<size_t>((n + 255) / 256, 1024)), 256, 0, st>>>(ipiv, int(n));
	XB_LAUNCH_CHECK();
	for (size_t k0 = 0; k0 < n; k0 += CH_NB) {
		const size_t nb = std::min<size_t>(CH_NB, n - k0), n2 = n - k0 - nb;
		lu_panel_kernel<<<1, 1024, 0, st>>>(A, (long long)n, int(k0), int(nb), ipiv, d_status);
		XB_LAUNCH_CHECK();
		if (k0) { lu_swap_kernel<<<unsigned((k0 + 127) / 128), 128, 0, st>>>(A, (long long)n, int(k0), ipiv, int(k0), int(nb)); XB_LAUNCH_CHECK(); }
		if (n2) { lu_swap_kernel<<<unsigned((n2 + 127) / 128), 128, 0, st>>>(A + k0 + nb, (long long)n, int(n2), ipiv, int(k0), int(nb)); XB_LAUNCH_CHECK(); }
		lu_swap_kernel<<<unsigned((nrhs + 127) / 128), 128, 0, st>>>(B, (long long)nrhs, int(nrhs), ipiv, int(k0), int(nb));
		XB_LAUNCH_CHECK();
		const double* L11 = A + k0 * n + k0;
		double* B1 = B + k0 * nrhs;
		lu_trsm_kernel<<<unsigned((nrhs + 127) / 128), 128, 0, st>>>(L11, (long long)n, B1, (long long)nrhs, int(nb), int(nrhs));
		XB_LAUNCH_CHECK();
		if (n2 == 0) break;
		double* A12 = A + k0 * n + k0 + nb;
		const double* L21 = A + (k0 + nb) * n + k0;
		lu_trsm_kernel<<<unsigned((n2 + 127) / 128), 128, 0, st>>>(L11, (long long)n, A12, (long long)n, int(nb), int(n2));
		XB_LAUNCH_CHECK();
		gemm(A + (k0 + nb) * n + k0 + nb, n, n2, n2, -1.0, L21, n, false, nb, A12, n, false, 1.0);       // A22 -= L21 U12
		gemm(B1 + nb * nrhs, nrhs, n2, nrhs, -1.0, L21, n, false, nb, B1, nrhs, false, 1.0);             // B2  -= L21 Y1
	}
	const int stat = read_status(d_status);
	dfree(d_status);
	dfree(ipiv);
	if (stat != 0) throw Error(XB_ERR_NUMERIC, "Unable to solve Ax = b (PLU solver): zero pivot in column " + std::to_string(stat - 1));
	// U X = Y, block by block from the bottom
	for (size_t kb = (n + CH_NB - 1) / CH_NB; kb-- > 0;) {
		const size_t k0 = kb * CH_NB, nb = std::min<size_t>(CH_NB, n - k0), n2 = n - k0 - nb;
		const double* Akk = A + k0 * n + k0;
		double* B1 = B + k0 * nrhs;
		if (n2) gemm(B1, nrhs, nb, nrhs, -1.0, Akk + nb, n, false, n2, B1 + nb * nrhs, nrhs, false, 1.0);    // B1 -= U12 X2
		chol_trsm_kernel<false><<<unsigned((nrhs + 127) / 128), 128, 0, st>>>(Akk, (long long)n, B1, (long long)nrhs, int(nb), int(nrhs));
		XB_LAUNCH_CHECK();
	}
}

// sign = +1 for a positive, -1 for a negative diagonal (the reference's dpotrf2 only succeeds for +1)
bool cholesky_solve(double* A, double* B, size_t n, size_t nrhs) {
	if (n > size_t(CH_NB)) return cholesky_solve_blocked(A, B, n, nrhs);
	int* d_status = static_cast<int*>(dalloc_bytes(sizeof(int)));
	cholesky_solve_kernel<<<1, 1024, 0, ctx().stream>>>(A, B, int(n), int(nrhs), 1.0, d_status);
	XB_LAUNCH_CHECK();
	const int st = read_status(d_status);
	dfree(d_status);
	return st == 0;
}

void lu_solve(double* A, double* B, size_t n, size_t nrhs) {
	if (n > size_t(CH_NB)) { lu_solve_blocked(A, B, n, nrhs); return; }
	int* d_status = static_cast<int*>(dalloc_bytes(sizeof(int)));
	lu_solve_kernel<<<1, 1024, 0, ctx().stream>>>(A, B, int(n), int(nrhs), d_status);
	XB_LAUNCH_CHECK();
	const int st = read_status(d_status);
	dfree(d_status);
	if (st != 0) throw Error(XB_ERR_NUMERIC, "Unable to solve Ax = b (PLU solver): zero pivot in column " + std::to_string(st - 1));
}

// returns {symmetric, definite diagonal}
void probe_symmetry(const double* A, size_t n, bool& symmetric, bool& definite_diag) {
	DBuf out(4);
	const unsigned parts = unsigned(std::min<size_t>(n, size_t(ctx().num_sms) * 4));
	DBuf part(size_t(parts) * 4);
	sym_probe_kernel<<<parts, 256, 0, ctx().stream>>>(A, int(n), part);
	XB_LAUNCH_CHECK();
	sym_probe_final_kernel<<<1, 32, 0, ctx().stream>>>(part, int(parts), A, out);
	XB_LAUNCH_CHECK();
	Context& c = ctx();
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, out.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	const double mx = std::max(0.0, c.h_scratch[0]);
	symmetric = (n == 1) || (c.h_scratch[1] < 4.0 * mx * 2.220446049250313e-16);   // blasLapackWrapper.cpp:509
	definite_diag = c.h_scratch[2] != 0.0;
}

} // namespace xb
