// Small-matrix factorisations: one CTA, one launch, everything in shared memory.
//
// The ends of every tensor train are rank ramps (2, 4, 8, ... in config 3; 4, 16 in configs 1 and 5) and config 1 is nothing
// but ramps: matrices of at most a few dozen columns.  On those the general path (cluster QR panels, multi-CTA cooperative
// Jacobi, pre-conditioning QR, Newton-Schulz polish: ~8 launches per QR, ~25 per SVD, 3-25 us each) is pure launch latency —
// a config-1 round() was 190 launches for 7 MFLOP.  Here a thin QR (Householder, explicit Q) and a one-sided Jacobi SVD of a
// matrix with min(m, n) <= 32 columns and max(m, n) <= 256 rows are one launch each:
//
//   qr_small   replaces dgeqrf + dorgqr (blasLapackWrapper.cpp:374-437) / their transposed use for LQ on these shapes
//   svd_small  replaces dgesdd (blasLapackWrapper.cpp:201-232); it leaves the factor in the layout Svd::extract() reads
//
// Both work on A * 2^-e (exact power of two; TT cores carry norms like 1e33 .. 1e150 and the kernels sum squares).
#include "xb_internal.cuh"

namespace xb {

constexpr int SM_WARPS = 16;
constexpr int SM_THREADS = SM_WARPS * 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// max |a| over the CTA -> power-of-two scale (scl = 2^-e with max * scl in [0.5, 1), inv = 2^e); zero / non-finite input: 1
__device__ void cta_pow2_scale(const double amax_local, double* red /* SM_WARPS + 2 */, double& scl, double& inv) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const double w = warp_max(amax_local);
	if (lane == 0) red[warp] = w;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int i = 0; i < SM_WARPS; ++i) s = fmax(s, red[i]);
		double a = 1.0, b = 1.0;
		if (s > 0.0 && s < HUGE_VAL) { int e; frexp(s, &e); a = ldexp(1.0, -e); b = ldexp(1.0, e); }
		red[SM_WARPS] = a; red[SM_WARPS + 1] = b;
	}
	__syncthreads();
	scl = red[SM_WARPS]; inv = red[SM_WARPS + 1];
}

// ---- thin QR --------------------------------------------------------------------------------------------------------------
// G(i, j) = A[i * ars + j * acs] (m x n), G = Q R with Q (m x k) -> Qo[i * qrs + c * qcs], R (k x n) -> Ro[c * rrs + j * rcs],
// k = min(m, n).  Strides make the transposed uses (LQ: A = L Q') free.  Shared memory: W[n][ldw] (columns of G, reflectors
// below the diagonal as in LAPACK), Qs[k][ldw] (columns of Q), tau[k].
// Schedule: one warp per column and look-ahead — while the other warps apply reflector j to their columns, warp 0 applies it to
// column j+1 and turns that column into reflector j+1 straight away, so a step costs one block barrier and the reflector
// arithmetic (reduction, rsqrt, reciprocal) is off the other warps' critical path.
constexpr int QS_WARPS = 32;
constexpr int QS_THREADS = QS_WARPS * 32;

__device__ __forceinline__ void qs_make_reflector(double* x, const int j, const int m, const int lane, double* tau) {
	// dlarfg on x[j..m): beta = -sign(alpha) |x|, v = x / (alpha - beta) below the diagonal, tau = (beta - alpha) / beta
	double s = 0.0;
	for (int i = j + 1 + lane; i < m; i += 32) s += x[i] * x[i];
	s = warp_sum(s);
	const double alpha = x[j];
	double t = 0.0, beta = alpha, f = 0.0;
	if (s > 0.0) {
		const double n2 = alpha * alpha + s, r = rsqrt(n2), nrm = n2 * r;     // |x| = n2 / sqrt(n2)
		beta = -copysign(nrm, alpha);
		t = 1.0 + fabs(alpha) * r;                                            // (beta - alpha) / beta
		f = copysign(__drcp_rn(fabs(alpha) + nrm), alpha);                    // 1 / (alpha - beta)
	}
	__syncwarp();
	for (int i = j + 1 + lane; i < m; i += 32) x[i] *= f;
	if (lane == 0) { x[j] = beta; tau[j] = t; }
	__syncwarp();
}
// a -= tau (v' a) v with v = [1; x[j+1..m)]
__device__ __forceinline__ void qs_apply(const double* x, const double t, double* a, const int j, const int m, const int lane) {
	double w = 0.0;
	for (int i = j + 1 + lane; i < m; i += 32) w += x[i] * a[i];
	const double aj = a[j];
	w = (warp_sum(w) + aj) * t;
	for (int i = j + 1 + lane; i < m; i += 32) a[i] -= w * x[i];
	if (lane == 0) a[j] = aj - w;
}

__global__ void __launch_bounds__(QS_THREADS, 1) qr_small_kernel(const double* __restrict__ A, const long long ars, const long long acs, const int m, const int n,
                                                              double* __restrict__ Qo, const long long qrs, const long long qcs,
                                                              double* __restrict__ Ro, const long long rrs, const long long rcs) {
	extern __shared__ __align__(16) double sm[];
	const int k = min(m, n), ldw = m + 1;
	double* W = sm;                        // n * ldw
	double* Qs = W + (size_t)n * ldw;      // k * ldw
	double* tau = Qs + (size_t)k * ldw;    // k
	double* red = tau + k;                 // QS_WARPS + 2
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	double amax = 0.0;
	if (acs == 1) {      // row-major source: neighbouring threads read neighbouring columns
		for (int e = threadIdx.x; e < m * n; e += QS_THREADS) { const int i = e / n, j = e - i * n; const double v = A[i * ars + j]; W[(size_t)j * ldw + i] = v; amax = fmax(amax, fabs(v)); }
	} else {
		for (int e = threadIdx.x; e < m * n; e += QS_THREADS) { const int j = e / m, i = e - j * m; const double v = A[i * ars + j * acs]; W[(size_t)j * ldw + i] = v; amax = fmax(amax, fabs(v)); }
	}
	// power-of-two scale from the block maximum
	{
		const double w = warp_max(amax);
		if (lane == 0) red[warp] = w;
		__syncthreads();
		if (threadIdx.x == 0) {
			double s = 0.0;
			for (int i = 0; i < QS_WARPS; ++i) s = fmax(s, red[i]);
			double a = 1.0, b = 1.0;
			if (s > 0.0 && s < HUGE_VAL) { int e; frexp(s, &e); a = ldexp(1.0, -e); b = ldexp(1.0, e); }
			red[QS_WARPS] = a; red[QS_WARPS + 1] = b;
		}
		__syncthreads();
	}
	const double scl = red[QS_WARPS], inv = red[QS_WARPS + 1];
	for (int e = threadIdx.x; e < m * n; e += QS_THREADS) { const int j = e / m, i = e - j * m; W[(size_t)j * ldw + i] *= scl; }
	for (int e = threadIdx.x; e < m * k; e += QS_THREADS) { const int c = e / m, i = e - c * m; Qs[(size_t)c * ldw + i] = (i == c) ? 1.0 : 0.0; }
	__syncthreads();
	if (warp == 0) qs_make_reflector(W, 0, m, lane, tau);
	__syncthreads();

	for (int j = 0; j < k; ++j) {
		const double* x = W + (size_t)j * ldw;
		const double t = tau[j];
		if (warp == 0) {
			if (j + 1 < n) {
				double* a = W + (size_t)(j + 1) * ldw;
				if (t != 0.0) qs_apply(x, t, a, j, m, lane);
				__syncwarp();
				if (j + 1 < k) qs_make_reflector(a, j + 1, m, lane, tau);
			}
		} else if (t != 0.0) {
			for (int c = j + 1 + warp; c < n; c += QS_WARPS - 1) qs_apply(x, t, W + (size_t)c * ldw, j, m, lane);
		}
		__syncthreads();
	}
	// R: upper trapezoid, unscaled
	for (int e = threadIdx.x; e < k * n; e += QS_THREADS) {
		const int c = e / n, j = e - c * n;
		Ro[c * rrs + j * rcs] = (j >= c) ? W[(size_t)j * ldw + c] * inv : 0.0;
	}
	// Q = H_0 ... H_{k-1} [I; 0]: reflectors in reverse order; H_j only touches rows >= j, and columns < j of the
	// partially formed Q are still unit vectors e_c with c < j: untouched
	for (int j = k - 1; j >= 0; --j) {
		const double t = tau[j];
		const double* x = W + (size_t)j * ldw;
		if (t != 0.0) for (int c = j + warp; c < k; c += QS_WARPS) qs_apply(x, t, Qs + (size_t)c * ldw, j, m, lane);
		__syncthreads();
	}
	if (qcs == 1) {
		for (int e = threadIdx.x; e < m * k; e += QS_THREADS) { const int i = e / k, c = e - i * k; Qo[i * qrs + c] = Qs[(size_t)c * ldw + i]; }
	} else {
		for (int e = threadIdx.x; e < m * k; e += QS_THREADS) { const int c = e / m, i = e - c * m; Qo[i * qrs + c * qcs] = Qs[(size_t)c * ldw + i]; }
	}
}

static size_t qr_small_smem(size_t m, size_t n) {
	const size_t k = std::min(m, n);
	return ((n + k) * (m + 1) + k + QS_WARPS + 2) * sizeof(double);
}

bool qr_small_fits(size_t m, size_t n) {
	if (!ctx().small_kernels) return false;
	if (std::min(m, n) > 32 || std::max(m, n) > 512) return false;
	return qr_small_smem(m, n) <= 200 * 1024;
}

void qr_small(double* Q, long long qrs, long long qcs, double* R, long long rrs, long long rcs, const double* A, long long ars, long long acs, size_t m, size_t n) {
	static bool attr_set = false;
	if (!attr_set) {
		XB_CUDA(cudaFuncSetAttribute(qr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		attr_set = true;
	}
	qr_small_kernel<<<1, QS_THREADS, qr_small_smem(m, n), ctx().stream>>>(A, ars, acs, int(m), int(n), Q, qrs, qcs, R, rrs, rcs);
	XB_LAUNCH_CHECK();
}

// ---- one-sided Jacobi SVD ---------------------------------------------------------------------------------------------------
// Working matrix G (mw x nw, nw <= mw): G(i, j) = A[i * rs + j * cs].  Columns are rotated pairwise (Hestenes) in a round-robin
// tournament, one warp per pair, squared norms cached and refreshed every sweep; V accumulates the rotations.  A sweep in which
// no pair had |cos| above `last_cos` is the last one (quadratic convergence), as in the multi-CTA kernels (svd_f64.cu).
// Output, in the layout Svd::extract() reads: GT row j = [ x_j (mw) ... | v_j (nw) ... ] with leading dimension ld and the V part
// at voff; Ssorted (descending, unscaled) and perm (rank -> row); scale2 = {2^-e, 2^e}; info = {0, sweeps, not converged}.
__global__ void __launch_bounds__(SM_THREADS) svd_small_kernel(const double* __restrict__ A, const long long rs, const long long cs, const int mw, const int nw,
                                                               double* __restrict__ GT, const int ld, const int voff,
                                                               double* __restrict__ Ssorted, int* __restrict__ perm, double* __restrict__ scale2,
                                                               unsigned int* __restrict__ info, const double tol, const double last_cos, const int max_sweeps,
                                                               const int polish) {
	extern __shared__ __align__(16) double sm[];
	const int np = (nw + 1) & ~1;                      // even number of columns (a zero column pads odd ones)
	const int ldx = mw + 1, ldv = nw + 1;
	double* X = sm;                                    // np * ldx
	double* V = X + (size_t)np * ldx;                  // np * ldv
	double* nrm = V + (size_t)np * ldv;                // np : squared norms
	double* red = nrm + np;                            // SM_WARPS + 2
	double* T = red + SM_WARPS + 2;                    // nw * ldv : Newton-Schulz factor, then the polished V
	__shared__ double sweep_max[SM_WARPS];
	__shared__ int stop_flag;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	double amax = 0.0;
	if (cs == 1) {
		for (int e = threadIdx.x; e < mw * nw; e += SM_THREADS) { const int i = e / nw, j = e - i * nw; const double v = A[i * rs + j]; X[(size_t)j * ldx + i] = v; amax = fmax(amax, fabs(v)); }
	} else {
		for (int e = threadIdx.x; e < mw * nw; e += SM_THREADS) { const int j = e / mw, i = e - j * mw; const double v = A[i * rs + j * cs]; X[(size_t)j * ldx + i] = v; amax = fmax(amax, fabs(v)); }
	}
	if (np > nw) for (int i = threadIdx.x; i < mw; i += SM_THREADS) X[(size_t)nw * ldx + i] = 0.0;
	for (int e = threadIdx.x; e < np * nw; e += SM_THREADS) { const int j = e / nw, c = e - j * nw; V[(size_t)j * ldv + c] = (j == c) ? 1.0 : 0.0; }
	double scl, inv;
	cta_pow2_scale(amax, red, scl, inv);
	for (int e = threadIdx.x; e < mw * nw; e += SM_THREADS) { const int j = e / mw, i = e - j * mw; X[(size_t)j * ldx + i] *= scl; }
	__syncthreads();

	const int npairs = np / 2;
	const double tol2 = tol * tol;
	int sweeps = 0;
	bool converged = (nw == 1);
	while (!converged && sweeps < max_sweeps) {
		for (int j = warp; j < np; j += SM_WARPS) {      // refresh the cached squared norms
			const double* x = X + (size_t)j * ldx;
			double s = 0.0;
			for (int i = lane; i < mw; i += 32) s += x[i] * x[i];
			s = warp_sum(s);
			if (lane == 0) nrm[j] = s;
		}
		if (lane == 0) sweep_max[warp] = 0.0;
		__syncthreads();
		double my_max = 0.0;                       // largest cos^2 met in this sweep
		for (int round = 0; round < np - 1; ++round) {
			for (int pi = warp; pi < npairs; pi += SM_WARPS) {
				// circle method: slot 0 is fixed, slots 1 .. np-1 rotate; pair pi = (slot pi, slot np-1-pi)
				int p = 0, q = np - 2 - pi + round;
				if (pi != 0) { p = pi - 1 + round; if (p >= np - 1) p -= np - 1; p += 1; }
				if (q >= np - 1) q -= np - 1;
				q += 1;
				if (p > q) { const int t = p; p = q; q = t; }
				double* xp = X + (size_t)p * ldx; double* xq = X + (size_t)q * ldx;
				const double a = nrm[p], b = nrm[q];
				if (a == 0.0 || b == 0.0) continue;
				double c = 0.0;
				for (int i = lane; i < mw; i += 32) c += xp[i] * xq[i];
				c = warp_sum(c);
				const double c2 = c * c, ab = a * b;
				if (c2 > my_max * ab) my_max = c2 / ab;       // rare after the first sweeps: the maximum only ever grows
				if (!(c2 > tol2 * ab)) continue;
				// rotation that annihilates c in [[a, c], [c, b]], two rsqrt and no division: with d = b - a, h = sqrt(d^2 + 4 c^2):
				// cos^2 = (1 + |d| / h) / 2, sin cos = sign(d) c / h
				const double d = b - a;
				const double r1 = rsqrt(d * d + 4.0 * c2);
				const double cs2 = 0.5 + 0.5 * fabs(d) * r1;
				const double r2 = rsqrt(cs2);
				const double csr = cs2 * r2;
				const double snr = copysign(c * r1 * r2, d * c);
				const double tc = copysign(c2 * r1 * r2 * r2, d);      // tan * c
				double* vp = V + (size_t)p * ldv; double* vq = V + (size_t)q * ldv;
				for (int i = lane; i < mw; i += 32) { const double u = xp[i], w = xq[i]; xp[i] = csr * u - snr * w; xq[i] = snr * u + csr * w; }
				for (int i = lane; i < nw; i += 32) { const double u = vp[i], w = vq[i]; vp[i] = csr * u - snr * w; vq[i] = snr * u + csr * w; }
				if (lane == 0) { nrm[p] = fmax(0.0, a - tc); nrm[q] = b + tc; }
			}
			__syncthreads();
		}
		if (lane == 0) sweep_max[warp] = my_max;
		__syncthreads();
		if (threadIdx.x == 0) {
			double mx = 0.0;
			for (int i = 0; i < SM_WARPS; ++i) mx = fmax(mx, sweep_max[i]);
			stop_flag = (mx <= last_cos * last_cos) ? 1 : 0;
		}
		__syncthreads();
		++sweeps;
		converged = stop_flag != 0;
		__syncthreads();
	}
	if (polish && nw > 1) {
		// Polish (as Svd::factor does after the multi-CTA kernels): hundreds of plane rotations leave V orthogonal to ~eps * sqrt(#rotations)
		// and X = G V with the same drift.  One Newton-Schulz step V <- V (1.5 I - 0.5 V'V) makes V orthogonal to eps^2-level, and the
		// left part is recomputed from the untouched input, X = G0 V: backward error back at the eps * sqrt(n) of LAPACK.
		for (int e = threadIdx.x; e < nw * nw; e += SM_THREADS) {
			const int j = e / nw, l = e - j * nw;
			const double* vj = V + (size_t)j * ldv; const double* vl = V + (size_t)l * ldv;
			double g = 0.0;
			for (int c = 0; c < nw; ++c) g += vj[c] * vl[c];
			T[(size_t)j * ldv + l] = (j == l ? 1.5 : 0.0) - 0.5 * g;
		}
		__syncthreads();
		// Vnew[j][c] = sum_l T[l][j] V[l][c]  (T symmetric); staged through X's first rows?  no: X is still needed -> use registers
		double vnew[2];
		int cnt = 0;
		for (int e = threadIdx.x; e < nw * nw; e += SM_THREADS, ++cnt) {
			const int j = e / nw, c = e - j * nw;
			double a = 0.0;
			for (int l = 0; l < nw; ++l) a += T[(size_t)l * ldv + j] * V[(size_t)l * ldv + c];
			vnew[cnt] = a;                          // nw <= 32: at most two entries per thread
		}
		__syncthreads();
		cnt = 0;
		for (int e = threadIdx.x; e < nw * nw; e += SM_THREADS, ++cnt) { const int j = e / nw, c = e - j * nw; V[(size_t)j * ldv + c] = vnew[cnt]; }
		__syncthreads();
		// X[j][i] = sum_c G0(i, c) V[j][c], G0 = scl * A
		for (int e = threadIdx.x; e < nw * mw; e += SM_THREADS) {
			const int j = e / mw, i = e - j * mw;
			const double* vj = V + (size_t)j * ldv;
			double a = 0.0;
			for (int c = 0; c < nw; ++c) a += A[i * rs + c * cs] * vj[c];
			X[(size_t)j * ldx + i] = a * scl;
		}
		__syncthreads();
	}
	// singular values = column norms (recomputed), ranked in descending order
	for (int j = warp; j < nw; j += SM_WARPS) {
		const double* x = X + (size_t)j * ldx;
		double s = 0.0;
		for (int i = lane; i < mw; i += 32) s += x[i] * x[i];
		s = warp_sum(s);
		if (lane == 0) nrm[j] = sqrt(s);
	}
	__syncthreads();
	for (int j = threadIdx.x; j < nw; j += SM_THREADS) {
		const double v = nrm[j];
		int rank = 0;
		for (int i = 0; i < nw; ++i) { const double u = nrm[i]; rank += (u > v || (u == v && i < j)) ? 1 : 0; }
		Ssorted[rank] = v * inv;
		perm[rank] = j;
	}
	for (int e = threadIdx.x; e < nw * mw; e += SM_THREADS) { const int j = e / mw, i = e - j * mw; GT[(size_t)j * ld + i] = X[(size_t)j * ldx + i]; }
	for (int e = threadIdx.x; e < nw * nw; e += SM_THREADS) { const int j = e / nw, c = e - j * nw; GT[(size_t)j * ld + voff + c] = V[(size_t)j * ldv + c]; }
	if (threadIdx.x == 0) { scale2[0] = scl; scale2[1] = inv; info[0] = 0u; info[1] = unsigned(sweeps); info[2] = converged ? 0u : 1u; }
}

static size_t svd_small_smem(size_t mw, size_t nw) {
	const size_t np = (nw + 1) & ~size_t(1);
	return (np * (mw + 1) + np * (nw + 1) + np + SM_WARPS + 2 + nw * (nw + 1)) * sizeof(double);
}

bool svd_small_fits(size_t mw, size_t nw) {
	if (!ctx().small_kernels) return false;
	if (nw > 32 || mw > 512) return false;
	return svd_small_smem(mw, nw) <= 200 * 1024;
}

void svd_small(const double* A, long long rs, long long cs, size_t mw, size_t nw, double* GT, size_t ld, size_t voff, double* Ssorted, int* perm,
               double* scale2, unsigned int* info, double tol, double last_cos, int max_sweeps, int polish) {
	static bool attr_set = false;
	if (!attr_set) {
		XB_CUDA(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
		attr_set = true;
	}
	svd_small_kernel<<<1, SM_THREADS, svd_small_smem(mw, nw), ctx().stream>>>(A, rs, cs, int(mw), int(nw), GT, int(ld), int(voff), Ssorted, perm, scale2, info,
	                                                                         tol, last_cos, max_sweeps, polish);
	XB_LAUNCH_CHECK();
}


// ---- Cholesky factor and its inverse -------------------------------------------------------------------------------------------
// G = sum of `nparts` n x n row-major slices (the split-K partial sums of a Gram matrix A^T A; symmetric, only the upper triangle
// is used) = R^T R:  R -> Ro (upper triangular) and W^T = R^-T -> Wo (lower triangular: that is how it sits in the registers, so
// both stores are coalesced and the GEMM that applies W reads it transposed), explicit zeros in the other triangle, n <= 128.
// The building block of the Cholesky-QR2 path for tall matrices (qr_f64.cu: cholqr2), where the Householder route is a chain of
// ~10 dependent panel / block-reflector launches per 32 columns and this is one launch between two GEMMs.
//
// The matrix lives in registers: 512 threads as a 16 x 32 grid, thread (ty, tx) owns the elements (ty + 16 a, tx + 32 b) —
// cyclic in both directions, so the work stays balanced while the active part of the matrix moves.  One Gauss step per column
// and one block barrier per step: the pivot row is published through a double-buffered row in shared memory, and the warp that
// owns the *next* pivot row updates and publishes it (with the reciprocal of its pivot) before it touches its other rows, so
// that latency hides behind everybody else's updates.  The same row operations are applied to the identity in the strict lower
// triangle of the array (G = L D L^T: the upper triangle ends as D L^T, the lower as L^-1 without its unit diagonal), which gives
// W^T = D^-1/2 L^-1 with no triangular solve afterwards; R = D^-1/2 (D L^T).
// near_identity (the second pass of Cholesky-QR2, G = I + E): ||E||_F <= 1e-8 takes R = I + U, W = I - U with U = triu(E) minus
// half its diagonal, exact to O(||E||^2) <= 1e-16 — no factorisation at all.
// Status word: *flag = flag_value when G is not *safely* positive definite — a diagonal entry outside [1e-200, 1e200] (the caller
// works unscaled) or NaN, a pivot <= 0, min pivot < ratio * max pivot, or (near_identity) ||G - I||_F > 1/2, which is what bounds
// the orthogonality of the second pass.  With clear != 0 a clean run writes 0 (the ordinary path reads the word back; a speculated
// run shares one sticky word with the other checks of its graph).
constexpr int CH_THREADS = 512;

// Steps j = 16 Q .. 16 Q + 15 of the elimination, for the thread grid of chol_inv_kernel (ty = warp = row residue mod 16,
// tx = lane = column residue mod 32; register row a holds matrix row ty + 16 a).  Inside such a block everything that shapes the
// code is known at compile time: the block of 32 columns that holds column j (Q / 2), the register row that may hold the pivot
// row or rows above it (Q; rows a > Q are below the pivot for every warp), and the register row of the next pivot (Q, or Q + 1
// after the last step) — no jump tables in the step, only warp-uniform predicates.
// Step j, row i > j: v(i, k) -= f_i r_k with f_i = U(j, i) / p_j and r the published pivot row, on the columns k < j (the L^-1
// part), k = j (the publisher stores r_j = 1, and v(i, j) is zeroed first: L^-1(i, j) = -f_i) and k >= i (the Schur complement).
// The columns j < k < i in between are not live yet — they pick up garbage that the step j' = k overwrites — so a row is CB
// unconditional DFMAs.  The warp that owns the next pivot row does that row first and publishes it; its lane 0 reads the new
// pivot back from shared memory, adds the reciprocal and puts the 1 in its place, while the other warps are busy with their rows.
template <int CB, int Q>
__device__ __forceinline__ bool chol_inv_block(double (&v)[2 * CB][CB], double (*rowbuf)[32 * CB], double* pv, double* pinv, double* piv,
                                               const int n, const int tx, const int ty, unsigned int* flag, const unsigned int flag_value) {
	constexpr int RA = 2 * CB, JB = (Q >> 1) < CB ? (Q >> 1) : 0, QN = (Q + 1 < RA) ? Q + 1 : Q;
	const int jend = min(n, 16 * Q + 16);
#define CH_ROW(A)                                                                                   \
	{                                                                                               \
		const double fa_ = rowbuf[buf][ty + 16 * (A)] * inv;                                        \
		if (tx == jl) v[A][JB] = 0.0;                                                               \
		_Pragma("unroll") for (int b = 0; b < CB; ++b) v[A][b] = fma(-fa_, rk[b], v[A][b]);         \
	}
#define CH_PUB(A)                                                                                   \
	{                                                                                               \
		_Pragma("unroll") for (int b = 0; b < CB; ++b) rowbuf[buf ^ 1][tx + 32 * b] = v[A][b];      \
	}
	for (int j = 16 * Q; j < jend; ++j) {
		const int buf = j & 1, t = j & 15, jl = j & 31, jn = j + 1;
		const double p = pv[buf];                        // the same word for every thread: the exit below is uniform
		const double inv = pinv[buf];
		double rk[CB];
#pragma unroll
		for (int b = 0; b < CB; ++b) rk[b] = rowbuf[buf][tx + 32 * b];
		if (!(p > 0.0)) { if (threadIdx.x == 0) *flag = flag_value; return false; }
		if (threadIdx.x == 0) piv[j] = p;
		const bool own_next = jn < n && ((t + 1) & 15) == ty;
		if (own_next) {
			if (t < 15) { CH_ROW(Q) CH_PUB(Q) } else { CH_ROW(QN) CH_PUB(QN) }
			__syncwarp();
			if (tx == 0) { const double pc = rowbuf[buf ^ 1][jn]; pv[buf ^ 1] = pc; pinv[buf ^ 1] = __drcp_rn(pc); rowbuf[buf ^ 1][jn] = 1.0; }
		}
		if (ty > t && !(own_next && t < 15)) CH_ROW(Q)
#pragma unroll
		for (int a = Q + 1; a < RA; ++a) {
			if (!(own_next && t == 15 && a == QN)) CH_ROW(a)
		}
		__syncthreads();
	}
#undef CH_ROW
#undef CH_PUB
	return true;
}

template <int CB>
__global__ void __launch_bounds__(CH_THREADS, 1) chol_inv_kernel(const double* __restrict__ G, const int nparts, const long long part_stride, const int n,
                                                                 double* __restrict__ Ro, double* __restrict__ Wo, const int near_identity, const double ratio,
                                                                 unsigned int* __restrict__ flag, const unsigned int flag_value, const int clear) {
	constexpr int RA = 2 * CB;
	__shared__ double rowbuf[2][32 * CB];
	__shared__ double pv[2], pinv[2];
	__shared__ double piv[32 * CB];
	__shared__ double red[CH_THREADS / 32];
	const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
	double v[RA][CB];
#pragma unroll
	for (int a = 0; a < RA; ++a)
#pragma unroll
		for (int b = 0; b < CB; ++b) v[a][b] = 0.0;
	// slices outermost: the RA * CB loads of one slice are independent and in flight together
	if (n == 32 * CB) {
		// full tile (n = 32, 64, 96, 128 — the bond dimensions of config 5): no bounds, one base pointer and immediate offsets
		const double* base = G + (size_t)ty * n + tx;
		for (int s = 0; s < nparts; ++s, base += part_stride) {
#pragma unroll
			for (int a = 0; a < RA; ++a)
#pragma unroll
				for (int b = 0; b < CB; ++b) v[a][b] += __ldg(base + (size_t)a * 16 * (32 * CB) + 32 * b);
		}
	} else
	for (int s = 0; s < nparts; ++s) {
		const double* src = G + (long long)s * part_stride;
#pragma unroll
		for (int a = 0; a < RA; ++a) {
			const int i = ty + 16 * a;
#pragma unroll
			for (int b = 0; b < CB; ++b) {
				const int k = tx + 32 * b;
				// clamped address, masked value: no branch per element, so the loads of a slice issue back to back
				const double g = __ldg(src + (size_t)min(i, n - 1) * n + min(k, n - 1));
				v[a][b] += (i < n && k < n) ? g : 0.0;
			}
		}
	}
	double dv = 0.0;
	int bad = 0;
#pragma unroll
	for (int a = 0; a < RA; ++a) {
		const int i = ty + 16 * a;
#pragma unroll
		for (int b = 0; b < CB; ++b) {
			const int k = tx + 32 * b;
			if (i < n && k < n) {
				const double g = v[a][b];
				const double t = g - (i == k ? 1.0 : 0.0);
				dv += t * t;
				if (i == k && !(g >= 1e-200 && g <= 1e200)) bad = 1;
			}
		}
	}
	if (tid < 2 * 32 * CB) (&rowbuf[0][0])[tid] = 0.0;
	dv = warp_sum(dv);
	if (tx == 0) red[ty] = dv;
	bad = __syncthreads_or(bad);
	double dev2 = 0.0;
#pragma unroll
	for (int w = 0; w < CH_THREADS / 32; ++w) dev2 += red[w];
	if (near_identity && !(dev2 <= 0.25)) bad = 1;
	if (bad) { if (tid == 0) *flag = flag_value; return; }
	if (near_identity && dev2 <= 1e-16) {
#pragma unroll
		for (int a = 0; a < RA; ++a) {
			const int i = ty + 16 * a;
#pragma unroll
			for (int b = 0; b < CB; ++b) {
				const int k = tx + 32 * b;
				if (i >= n || k >= n) continue;
				const double g = v[a][b];
				const size_t e = (size_t)i * n + k;
				if (k > i) { Ro[e] = g; Wo[e] = 0.0; }
				else if (k == i) { Ro[e] = 0.5 * (1.0 + g); Wo[e] = 0.5 * (3.0 - g); }
				else { Ro[e] = 0.0; Wo[e] = -g; }                  // G(i, k) = G(k, i)
			}
		}
		if (tid == 0 && clear) *flag = 0u;
		return;
	}
	if (ty == 0) {
#pragma unroll
		for (int b = 0; b < CB; ++b) rowbuf[0][tx + 32 * b] = v[0][b];
		__syncwarp();
		if (tx == 0) { pv[0] = v[0][0]; pinv[0] = __drcp_rn(v[0][0]); rowbuf[0][0] = 1.0; }
	}
	__syncthreads();
	{
		bool ok = chol_inv_block<CB, 0>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
		if (ok && n > 16) ok = chol_inv_block<CB, 1>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
		if constexpr (CB >= 2) {
			if (ok && n > 32) ok = chol_inv_block<CB, 2>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
			if (ok && n > 48) ok = chol_inv_block<CB, 3>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
		}
		if constexpr (CB >= 3) {
			if (ok && n > 64) ok = chol_inv_block<CB, 4>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
			if (ok && n > 80) ok = chol_inv_block<CB, 5>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
		}
		if constexpr (CB >= 4) {
			if (ok && n > 96) ok = chol_inv_block<CB, 6>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
			if (ok && n > 112) ok = chol_inv_block<CB, 7>(v, rowbuf, pv, pinv, piv, n, tx, ty, flag, flag_value);
		}
		if (!ok) return;
	}
	bool flagged = false;
	if (tid == 0) {
		double pmin = HUGE_VAL, pmax = 0.0;
		for (int i = 0; i < n; ++i) { pmin = fmin(pmin, piv[i]); pmax = fmax(pmax, piv[i]); }
		flagged = !(pmin >= ratio * pmax);
	}
#pragma unroll
	for (int a = 0; a < RA; ++a) {
		const int i = ty + 16 * a;
		if (i >= n) continue;
		const double rs = rsqrt(piv[i]);
#pragma unroll
		for (int b = 0; b < CB; ++b) {
			const int k = tx + 32 * b;
			if (k >= n) continue;
			const double x = v[a][b] * rs;
			const size_t e = (size_t)i * n + k;
			if (k > i) { Ro[e] = x; Wo[e] = 0.0; }
			else if (k == i) { Ro[e] = x; Wo[e] = rs; }
			else { Ro[e] = 0.0; Wo[e] = x; }
		}
	}
	if (tid == 0) { if (flagged) *flag = flag_value; else if (clear) *flag = 0u; }
}

bool chol_inv_fits(size_t n) { return n >= 1 && n <= 128; }

void chol_inv(const double* G, size_t nparts, size_t n, double* R, double* W, bool near_identity, double ratio, unsigned int* flag, unsigned int flag_value, bool clear) {
	const long long ps = (long long)(n * n);
	const int ni = near_identity ? 1 : 0, cl = clear ? 1 : 0;
	cudaStream_t st = ctx().stream;
	if (n <= 32) chol_inv_kernel<1><<<1, CH_THREADS, 0, st>>>(G, int(nparts), ps, int(n), R, W, ni, ratio, flag, flag_value, cl);
	else if (n <= 64) chol_inv_kernel<2><<<1, CH_THREADS, 0, st>>>(G, int(nparts), ps, int(n), R, W, ni, ratio, flag, flag_value, cl);
	else if (n <= 96) chol_inv_kernel<3><<<1, CH_THREADS, 0, st>>>(G, int(nparts), ps, int(n), R, W, ni, ratio, flag, flag_value, cl);
	else chol_inv_kernel<4><<<1, CH_THREADS, 0, st>>>(G, int(nparts), ps, int(n), R, W, ni, ratio, flag, flag_value, cl);
	XB_LAUNCH_CHECK();
}

// out[e] = sum over `parts` slices of P[q * len + e]: one block per element group, fixed summation tree (deterministic)
__global__ void __launch_bounds__(256) sum_parts_kernel(double* __restrict__ out, const double* __restrict__ P, const int parts, const size_t len) {
	// 8 elements per block, 32 lanes each: lane l adds the slices l, l + 32, ...; a shuffle tree finishes
	const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
	const size_t e = (size_t)blockIdx.x * 8 + sub;
	if (e >= len) return;
	double s = 0.0;
	for (int q = lane; q < parts; q += 32) s += P[(size_t)q * len + e];
	s = warp_sum(s);
	if (lane == 0) out[e] = s;
}
void sum_parts(double* out, const double* P, size_t parts, size_t len) {
	sum_parts_kernel<<<unsigned((len + 7) / 8), 256, 0, ctx().stream>>>(out, P, int(parts), len);
	XB_LAUNCH_CHECK();
}

} // namespace xb
