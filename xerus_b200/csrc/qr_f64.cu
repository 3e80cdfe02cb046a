// Blocked Householder QR for sm_100a (FP64), plus LQ / RQ / rank-revealing QC / CQ on top of it.
//
// Replaces LAPACKE_dgeqrf + dorgqr (blasWrapper::qr, reference: src/xerus/blasLapackWrapper.cpp:388-438),
// dgerqf + dorgrq (:455-498) and dgeqp3 + dorgqr (qc/cq, :243-371).
//
// Structure (compact WY, panel width 32):
//   qr_panel_kernel   one CTA, panel resident in shared memory (lane = panel column, warps stride the rows).
//                     Per column ONE fused pass computes g_c = x^T a_c for all remaining columns c (c = j gives
//                     ||x||^2), from which beta, tau and w = v^T A follow algebraically — no separate norm pass —
//                     and one pass applies the rank-1 update.  T (compact WY) is built at the end from V^T V.
//   qr_vtc_kernel     partial W_p = V^T C over row chunks (grid: column blocks x row chunks, deterministic)
//   qr_update_kernel  sums the partials, W2 = op(T) W, C -= V W2
// The trailing update and the formation of the explicit thin Q use the same two kernels (op(T) = T^T resp. T).
#include "xb_internal.cuh"
#include <cooperative_groups.h>

namespace xb {

constexpr int QR_NB = 32;          // panel width == warp width
constexpr int QR_PANEL_WARPS = 32; // 1024 threads
constexpr int QR_CHUNK = 64;       // rows per CTA in the trailing kernels

template <bool SMEM>
__global__ void __launch_bounds__(QR_PANEL_WARPS * 32) qr_panel_kernel(double* __restrict__ W, const long long ldw, const int mp,
                                                                      const int nbe, double* __restrict__ Vbuf, double* __restrict__ Tout) {
	extern __shared__ double sm[];
	double* red = sm;                         // [32][33]
	double* GV = red + 32 * 33;               // [32][33]
	double* Ts = GV + 32 * 33;                // [32][33]
	double* s_wv = Ts + 32 * 33;              // [32]
	double* s_tau = s_wv + 32;                // [32]
	double* s_misc = s_tau + 32;              // [4]
	double* Ps = s_misc + 4;                  // [mp][32] when SMEM

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	constexpr int NW = QR_PANEL_WARPS;      // the launch always uses QR_PANEL_WARPS warps
	const bool active = lane < nbe;
	auto pe = [&](int i, int c) -> double& { return SMEM ? Ps[i * QR_NB + c] : W[(long long)i * ldw + c]; };

	if (SMEM) {
		for (int i = warp; i < mp; i += NW) Ps[i * QR_NB + lane] = active ? W[(long long)i * ldw + lane] : 0.0;
	}
	for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) { Ts[i] = 0.0; GV[i] = 0.0; }
	if (threadIdx.x < 32) s_tau[threadIdx.x] = 0.0;
	__syncthreads();

	const int kmax = min(nbe, mp);
	for (int j = 0; j < kmax; ++j) {
		// pass 1: g_c = sum_{i >= j} P[i][j] * P[i][c] for c > j; lane j accumulates the TAIL sum_{i > j} P[i][j]^2 only, so
		// that ||x||^2 - alpha^2 is never formed by subtraction (a nearly triangular input would lose its sub-diagonal mass)
		double g = 0.0;
		if (active && lane >= j) {
			for (int i = j + warp; i < mp; i += NW) { if (lane != j || i != j) g += pe(i, j) * pe(i, lane); }
		}
		red[warp * 33 + lane] = g;
		__syncthreads();
		if (warp == 0) {
			double G0 = 0.0, G1 = 0.0, G2 = 0.0, G3 = 0.0;
#pragma unroll
			for (int w = 0; w < NW; w += 4) {
				G0 += red[w * 33 + lane]; G1 += red[(w + 1) * 33 + lane]; G2 += red[(w + 2) * 33 + lane]; G3 += red[(w + 3) * 33 + lane];
			}
			const double G = (G0 + G1) + (G2 + G3);
			const double tail = __shfl_sync(0xffffffffu, G, j);
			const double alpha = pe(j, j);
			const double sigma = tail + alpha * alpha;
			double beta = alpha, tau = 0.0, scl = 0.0;
			if (tail > 0.0 && j + 1 < mp) {
				// beta = -sign(alpha) ||x||, tau = (beta - alpha)/beta = 1 + |alpha|/||x||, 1/(alpha - beta) = sign(alpha)/(|alpha| + ||x||)
				const double rs = rsqrt(sigma), nrmx = sigma * rs;
				beta = -copysign(nrmx, alpha);
				tau = 1.0 + fabs(alpha) * rs;
				scl = copysign(1.0, alpha) / (fabs(alpha) + nrmx);
			}
			double wv = 0.0;
			if (active && lane > j && tau != 0.0) wv = tau * (G - beta * pe(j, lane)) * scl;
			s_wv[lane] = wv;
			if (lane == 0) { s_tau[j] = tau; s_misc[0] = beta; s_misc[1] = scl; }
		}
		__syncthreads();
		const double tau = s_tau[j], beta = s_misc[0], scl = s_misc[1];
		// pass 2: A[:, c] -= tau * w_c * v for c > j ; store v below the diagonal of column j, beta on it
		for (int i = j + warp; i < mp; i += NW) {
			const double vi = (i == j) ? 1.0 : pe(i, j) * scl;
			__syncwarp();
			if (active) {
				if (lane > j) { if (tau != 0.0) pe(i, lane) -= s_wv[lane] * vi; }
				else if (lane == j) pe(i, j) = (i == j) ? beta : ((tau != 0.0) ? vi : 0.0);
			}
		}
		__syncthreads();
	}

	// explicit V (unit lower trapezoidal, zero padded to 32 columns)
	auto vval = [&](int i, int c) -> double {
		if (c >= kmax || i < c) return 0.0;
		return (i == c) ? 1.0 : pe(i, c);
	};
	for (int i = warp; i < mp; i += NW) Vbuf[(long long)i * QR_NB + lane] = vval(i, lane);
	// GV = V^T V : warp w owns row w
	if (warp < kmax) {
		// rows below the top 32 x 32 block hold plain V entries (no unit-diagonal / zero masking needed)
		double acc0 = 0.0, acc1 = 0.0;
		const int top = min(mp, 32);
		for (int i = warp; i < top; ++i) acc0 += vval(i, warp) * vval(i, lane);
		if (lane < kmax) {
			int i = 32;
			for (; i + 1 < mp; i += 2) { acc0 += pe(i, warp) * pe(i, lane); acc1 += pe(i + 1, warp) * pe(i + 1, lane); }
			if (i < mp) acc0 += pe(i, warp) * pe(i, lane);
		}
		GV[warp * 33 + lane] = acc0 + acc1;
	}
	__syncthreads();
	// T(0:j, j) = -tau_j * T(0:j, 0:j) * GV(0:j, j) ; T(j, j) = tau_j
	if (warp == 0) {
		for (int j = 0; j < kmax; ++j) {
			// T[lane][c] is zero for c < lane and for c >= j (not written yet): the full-length sum needs no per-lane bounds
			double t0 = 0.0, t1 = 0.0;
#pragma unroll
			for (int c = 0; c < 32; c += 2) { t0 += Ts[lane * 33 + c] * GV[c * 33 + j]; t1 += Ts[lane * 33 + c + 1] * GV[(c + 1) * 33 + j]; }
			__syncwarp();
			if (lane < j) Ts[lane * 33 + j] = -s_tau[j] * (t0 + t1);
			else if (lane == j) Ts[j * 33 + j] = s_tau[j];
			__syncwarp();
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) Tout[i] = Ts[(i >> 5) * 33 + (i & 31)];
	if (SMEM) {
		for (int i = warp; i < mp; i += NW) if (active) W[(long long)i * ldw + lane] = Ps[i * QR_NB + lane];
	}
}


// ---- cluster panel kernel ---------------------------------------------------------------------------------------------
// The one-CTA panel kernel above is bound by the FP64 issue rate of a single SM and by its shared-memory read-modify-write
// chain.  This variant spreads the panel rows over a thread-block cluster of QRC_CS CTAs (one SM each) and keeps every
// row strip in REGISTERS (lane = panel column, RPT rows per thread); a column entry of another lane comes from a shuffle.
// Per column there is ONE fused pass — apply reflector j, accumulate g_c = x_{j+1}^T a_c for the next column on the fly —
// and one cluster-wide reduction: CTA partials are written into every CTA's shared memory with st.async (DSMEM stores that
// complete a transaction count on the receiving CTA's mbarrier — no cluster barrier, no fence in the column loop), every
// warp waits on its own CTA's mbarrier and derives beta / tau / w redundantly.  Lanes c < j are idle in the Householder step, so they accumulate
// v_c^T x_j in the same pass, which gives column j of V^T V (needed for the compact-WY T) without a pass over V.
// ---- DSMEM signalling: remote 8-byte stores that complete a transaction count on the remote CTA's mbarrier ---------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
__device__ __forceinline__ void st_async_f64(unsigned raddr, double v, unsigned rmbar) {
	asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
	             :: "r"(raddr), "l"(__double_as_longlong(v)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
	unsigned done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
	} while (!done);
}

constexpr int QRC_CS = 8;        // CTAs per cluster (portable maximum)
constexpr int QRC_WARPS = 8;     // warps per CTA

template <int RPT>
__global__ void __cluster_dims__(QRC_CS, 1, 1) __launch_bounds__(QRC_WARPS * 32)
qr_panel_cluster_kernel(double* __restrict__ W, const long long ldw, const int mp, const int nbe, double* __restrict__ Vbuf,
                        double* __restrict__ Tout, long long* __restrict__ dbg) {
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ double red[QRC_WARPS][32];
	__shared__ double slots[2][QRC_CS][32];     // [parity][source CTA][lane] partial sums of the current column step
	__shared__ double rowbuf[2][32];            // [parity][lane] row j of the panel as it stands before step j
	__shared__ double GVs[32][33], Ts[32][33];  // used by CTA 0 only
	__shared__ double s_tau[32];
	__shared__ __align__(8) unsigned long long mbar[2];   // [parity] all partial sums + row j of a column step have arrived
	constexpr unsigned TX_BYTES = (QRC_CS * 32 + 32) * sizeof(double);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int rank = int(cluster.block_rank());
	// row of slot k of this warp: k * 64 + gw (cyclic over the 64 warps of the cluster).  Only slot 0 can hold one of the
	// top 32 rows, so slots 1.. are always strictly below the current and the next column's diagonal: no masking there.
	constexpr int NWC = QRC_CS * QRC_WARPS;
	const int gw = rank * QRC_WARPS + warp;
	const bool active = lane < nbe;
	const int kmax = min(nbe, mp);

	double P[RPT];
#pragma unroll
	for (int k = 0; k < RPT; ++k) { const int i = k * NWC + gw; P[k] = (active && i < mp) ? W[(long long)i * ldw + lane] : 0.0; }
	if (rank == 0) {
		for (int e = threadIdx.x; e < 32 * 33; e += blockDim.x) { (&GVs[0][0])[e] = 0.0; (&Ts[0][0])[e] = 0.0; }
		if (threadIdx.x < 32) s_tau[threadIdx.x] = 0.0;
	}
	// g for column 0 (lane 0 accumulates the tail below the diagonal only); the owner of row 0 publishes it
	double g = 0.0;
#pragma unroll
	for (int k = 0; k < RPT; ++k) {
		const double pn = __shfl_sync(0xffffffffu, P[k], 0);
		if (!(lane == 0 && k == 0 && gw == 0)) g += pn * P[k];
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&mbar[0]), 1); mbar_init(smem_u32(&mbar[1]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		mbar_arrive_expect_tx(smem_u32(&mbar[0]), TX_BYTES);         // column steps 0 and 1
		mbar_arrive_expect_tx(smem_u32(&mbar[1]), TX_BYTES);
	}
	cluster.sync();                              // every CTA is resident and its barriers are initialised before the first remote store
	if (gw == 0) {
		const unsigned a = smem_u32(&rowbuf[0][lane]), mb = smem_u32(&mbar[0]);
#pragma unroll
		for (int c = 0; c < QRC_CS; ++c) st_async_f64(mapa_u32(a, c), P[0], mapa_u32(mb, c));
	}

	// optional phase timing (XB_QR_TIMING=1): cycles of [reduce + publish | cluster barrier | parameters | update]
	const bool timing = dbg != nullptr && rank == 0 && threadIdx.x == 0;
	long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0;
	for (int j = 0; j < kmax; ++j) {
		const int par = j & 1;
		if (timing) t0 = clock64();
		red[warp][lane] = g;
		__syncthreads();
		if (warp == 0) {
			const double s = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
			const unsigned a = smem_u32(&slots[par][rank][lane]), mb = smem_u32(&mbar[par]);
#pragma unroll
			for (int c = 0; c < QRC_CS; ++c) st_async_f64(mapa_u32(a, c), s, mapa_u32(mb, c));
		}
		if (timing) { const long long t1 = clock64(); tk[0] += t1 - t0; t0 = t1; }
		mbar_wait(smem_u32(&mbar[par]), (j >> 1) & 1);
		// re-arm this parity for column step j + 2 (its stores may already be under way: the transaction count may run negative)
		if (threadIdx.x == 0 && j + 2 < kmax) mbar_arrive_expect_tx(smem_u32(&mbar[par]), TX_BYTES);
		if (timing) { const long long t1 = clock64(); tk[1] += t1 - t0; t0 = t1; }
		const double G = ((slots[par][0][lane] + slots[par][1][lane]) + (slots[par][2][lane] + slots[par][3][lane])) +
		                 ((slots[par][4][lane] + slots[par][5][lane]) + (slots[par][6][lane] + slots[par][7][lane]));
		const double rowj = rowbuf[par][lane];
		const double tail = __shfl_sync(0xffffffffu, G, j);
		const double alpha = __shfl_sync(0xffffffffu, rowj, j);
		const double sigma = tail + alpha * alpha;
		double beta = alpha, tau = 0.0, scl = 0.0;
		if (tail > 0.0 && j + 1 < mp) {
			// beta = -sign(alpha) ||x||, tau = 1 + |alpha|/||x||, 1/(alpha - beta) = sign(alpha)/(|alpha| + ||x||)
			const double rs = rsqrt(sigma), nrmx = sigma * rs;
			beta = -copysign(nrmx, alpha);
			tau = 1.0 + fabs(alpha) * rs;
			scl = copysign(__drcp_rn(fabs(alpha) + nrmx), alpha);
		}
		// w_c = tau * v^T a_c for the columns right of j; zero elsewhere, and zero altogether when tau == 0 (then scl == 0 too)
		const double wv = (lane > j) ? (tau * scl) * (G - beta * rowj) : 0.0;
		if (rank == 0 && warp == 0) {
			// v_c^T v_j = v_c[j] + scl * (v_c^T x - v_c[j] * alpha)   for c < j   (rowj holds v_c[j] in lane c)
			if (lane < j) GVs[lane][j] = (tau != 0.0) ? rowj + scl * (G - rowj * alpha) : 0.0;
			if (lane == 0) s_tau[j] = tau;
		}
		if (timing) { const long long t1 = clock64(); tk[2] += t1 - t0 + (long long)(wv != wv); t0 = t1; }
		// apply H_j to the strip, store v_j / beta in column j, accumulate the sums of step j + 1.  The shuffles of all rows
		// are issued back to back (a shuffle inside the per-row dependency chain would serialise the rows).
		const int jn = j + 1;
		double pj[RPT];
#pragma unroll
		for (int k = 0; k < RPT; ++k) pj[k] = __shfl_sync(0xffffffffu, P[k], j);
		// slot 0: row gw may be above (done), on, or below the diagonal — warp-uniform branches
		if (gw == j) P[0] = (lane > j) ? P[0] - wv : ((lane == j) ? beta : P[0]);
		else if (gw > j) { const double vi = pj[0] * scl; P[0] = (lane == j) ? vi : P[0] - wv * vi; }
#pragma unroll
		for (int k = 1; k < RPT; ++k) { const double vi = pj[k] * scl; P[k] = (lane == j) ? vi : P[k] - wv * vi; }
		if (timing) { const long long t1 = clock64(); tk[4] += t1 - t0 + (long long)(P[0] != P[0]) + (long long)(P[RPT - 1] != P[RPT - 1]); t0 = t1; }
		double pn[RPT];
#pragma unroll
		for (int k = 0; k < RPT; ++k) pn[k] = __shfl_sync(0xffffffffu, P[k], jn & 31);
		double ga[4] = {0.0, 0.0, 0.0, 0.0};
		if (gw > jn || (gw == jn && lane != jn)) ga[0] = pn[0] * P[0];      // row jn: lane jn keeps the tail below the diagonal only
#pragma unroll
		for (int k = 1; k < RPT; ++k) ga[k & 3] += pn[k] * P[k];
		g = (ga[0] + ga[1]) + (ga[2] + ga[3]);
		if (timing) { const long long t1 = clock64(); tk[5] += t1 - t0 + (long long)(g != g); t0 = t1; }
		if (gw == jn && jn < kmax) {
			const unsigned a = smem_u32(&rowbuf[par ^ 1][lane]), mb = smem_u32(&mbar[par ^ 1]);
#pragma unroll
			for (int c = 0; c < QRC_CS; ++c) st_async_f64(mapa_u32(a, c), P[0], mapa_u32(mb, c));
		}
		if (timing) { const long long t1 = clock64(); tk[3] += t1 - t0; t0 = t1; }
	}
	if (timing) { for (int q = 0; q < 8; ++q) dbg[q] = tk[q]; }
	cluster.sync();                              // no CTA leaves while stores to its shared memory could still be in flight

	// results: panel in place (R on and above the diagonal, reflectors below), explicit V, compact-WY T
#pragma unroll
	for (int k = 0; k < RPT; ++k) {
		const int i = k * NWC + gw;
		if (i < mp) {
			if (active) W[(long long)i * ldw + lane] = P[k];
			Vbuf[(long long)i * QR_NB + lane] = (lane >= kmax || i < lane) ? 0.0 : ((i == lane) ? 1.0 : P[k]);
		}
	}
	if (rank == 0 && warp == 0) {
		__syncwarp();
		// T(0:j, j) = -tau_j * T(0:j, 0:j) * GV(0:j, j) ; T(j, j) = tau_j      (GV strictly upper, zero elsewhere)
		for (int j = 0; j < kmax; ++j) {
			double t0 = 0.0, t1 = 0.0;
#pragma unroll
			for (int c = 0; c < 32; c += 2) { t0 += Ts[lane][c] * GVs[c][j]; t1 += Ts[lane][c + 1] * GVs[c + 1][j]; }
			__syncwarp();
			if (lane < j) Ts[lane][j] = -s_tau[j] * (t0 + t1);
			else if (lane == j) Ts[j][j] = s_tau[j];
			__syncwarp();
		}
		for (int e = lane; e < 32 * 32; e += 32) Tout[e] = Ts[e >> 5][e & 31];
	}
}

static bool launch_panel_cluster(double* Wpanel, long long ldw, size_t mp, size_t nbe, double* Vp, double* Tp) {
	if (!ctx().qr_cluster || mp < size_t(ctx().qr_cluster_min_rows) || mp > size_t(QRC_CS * QRC_WARPS * 32)) return false;
	const size_t rpt = (mp + QRC_CS * QRC_WARPS - 1) / (QRC_CS * QRC_WARPS);
	cudaStream_t st = ctx().stream;
	static const bool timing = getenv("XB_QR_TIMING") != nullptr;
	long long* dbg = nullptr;
	if (timing) dbg = static_cast<long long*>(dalloc_bytes(8 * sizeof(long long)));
	struct Report { long long* d; size_t mp; ~Report() { if (!d) return; long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[qr panel] mp=%zu cycles: reduce+publish %lld barrier %lld params %lld apply %lld next-sums %lld row-publish %lld\n", mp, h[0], h[1], h[2], h[4], h[5], h[3]); dfree(d); } } report{dbg, mp};
	if (rpt <= 2) qr_panel_cluster_kernel<2><<<QRC_CS, QRC_WARPS * 32, 0, st>>>(Wpanel, ldw, int(mp), int(nbe), Vp, Tp, dbg);
	else if (rpt <= 4) qr_panel_cluster_kernel<4><<<QRC_CS, QRC_WARPS * 32, 0, st>>>(Wpanel, ldw, int(mp), int(nbe), Vp, Tp, dbg);
	else if (rpt <= 8) qr_panel_cluster_kernel<8><<<QRC_CS, QRC_WARPS * 32, 0, st>>>(Wpanel, ldw, int(mp), int(nbe), Vp, Tp, dbg);
	else if (rpt <= 16) qr_panel_cluster_kernel<16><<<QRC_CS, QRC_WARPS * 32, 0, st>>>(Wpanel, ldw, int(mp), int(nbe), Vp, Tp, dbg);
	else qr_panel_cluster_kernel<32><<<QRC_CS, QRC_WARPS * 32, 0, st>>>(Wpanel, ldw, int(mp), int(nbe), Vp, Tp, dbg);
	return true;
}

// Wp[chunk][kk][c] = sum_{rows of chunk} V[row][kk] * C[row][c]     block: 32 x 32 threads (warp = kk, lane = c)
__global__ void __launch_bounds__(1024) qr_vtc_kernel(const double* __restrict__ Vbuf, const double* __restrict__ C, const long long ldc,
                                                     const int mp, const int nc, double* __restrict__ Wp, const int ncpad) {
	__shared__ double Vs[QR_CHUNK][32];
	__shared__ double Cs[QR_CHUNK][33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int r0 = blockIdx.y * QR_CHUNK;
	const int c = blockIdx.x * 32 + lane;
	for (int ii = warp; ii < QR_CHUNK; ii += 32) {
		const int row = r0 + ii;
		const bool ok = row < mp;
		Vs[ii][lane] = ok ? Vbuf[(long long)row * QR_NB + lane] : 0.0;
		Cs[ii][lane] = (ok && c < nc) ? C[(long long)row * ldc + c] : 0.0;
	}
	__syncthreads();
	double acc = 0.0;
#pragma unroll 8
	for (int ii = 0; ii < QR_CHUNK; ++ii) acc += Vs[ii][warp] * Cs[ii][lane];
	Wp[((long long)blockIdx.y * 32 + warp) * ncpad + c] = acc;
}

// C[rows of chunk][col block] -= V * (op(T) * sum_p Wp)
__global__ void __launch_bounds__(1024) qr_update_kernel(double* __restrict__ C, const long long ldc, const int mp, const int nc,
                                                        const double* __restrict__ Vbuf, const double* __restrict__ T, const int transT,
                                                        const double* __restrict__ Wp, const int nchunks, const int ncpad) {
	__shared__ double Tsm[32][33];
	__shared__ double Ws[32][33];
	__shared__ double W2[32][33];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int c = blockIdx.x * 32 + lane;
	Tsm[warp][lane] = T[warp * 32 + lane];
	double s = 0.0;
	for (int p = 0; p < nchunks; ++p) s += Wp[((long long)p * 32 + warp) * ncpad + c];
	Ws[warp][lane] = s;
	__syncthreads();
	double acc = 0.0;
#pragma unroll 8
	for (int l = 0; l < 32; ++l) acc += (transT ? Tsm[l][warp] : Tsm[warp][l]) * Ws[l][lane];
	W2[warp][lane] = acc;
	__syncthreads();
	const int r0 = blockIdx.y * QR_CHUNK, r1 = min(mp, r0 + QR_CHUNK);
	if (c < nc) {
		for (int row = r0 + warp; row < r1; row += 32) {
			const double* v = Vbuf + (long long)row * QR_NB;
			double d = 0.0;
#pragma unroll 8
			for (int kk = 0; kk < 32; ++kk) d += v[kk] * W2[kk][lane];
			C[(long long)row * ldc + c] -= d;
		}
	}
}

__global__ void qr_extract_r_kernel(double* __restrict__ R, const double* __restrict__ W, const size_t k, const size_t n, const double* __restrict__ unscale) {
	const size_t total = k * n;
	const double f = *unscale;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / n, c = i % n;
		R[i] = (c >= r) ? W[i] * f : 0.0;
	}
}

// min / max of |diag| of a (rows x cols, ld) matrix -> out[0] = min, out[1] = max      (single warp)
__global__ void diag_minmax_kernel(const double* __restrict__ A, const size_t k, const size_t ld, double* __restrict__ out) {
	double mn = HUGE_VAL, mx = 0.0;
	for (size_t i = threadIdx.x; i < k; i += 32) { const double v = fabs(A[i * ld + i]); mn = fmin(mn, v); mx = fmax(mx, v); }
	for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
	if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
}


// ---- cluster block-reflector kernel -----------------------------------------------------------------------------------
// C -= V * (op(T) * (V^T C)) for one 32-column block of C per cluster: the QRC_CS CTAs of a cluster split the rows, each
// forms its partial V^T C (32 x 32), the partials meet in an L2-resident scratch across ONE cluster barrier, every CTA
// sums them in a fixed order (deterministic), applies op(T) and updates its rows — which are still in shared memory.
// One launch instead of two, no second pass over C from global memory.
constexpr int QRA_TILE = 64;
constexpr int QRA_THREADS = 256;
constexpr size_t QRA_SMEM = (size_t(QRA_TILE) * 32 + size_t(QRA_TILE) * 33 + 3 * 32 * 33) * sizeof(double);

__global__ void __cluster_dims__(QRC_CS, 1, 1) __launch_bounds__(QRA_THREADS)
qr_apply_cluster_kernel(double* __restrict__ C, const long long ldc, const int mp, const int nc, const double* __restrict__ Vbuf,
                        const double* __restrict__ T, const int transT, double* Wp) {
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	extern __shared__ __align__(16) double qra_sm[];
	double (*Vs)[32] = reinterpret_cast<double (*)[32]>(qra_sm);
	double (*Cs)[33] = reinterpret_cast<double (*)[33]>(qra_sm + QRA_TILE * 32);
	double (*Ws)[33] = reinterpret_cast<double (*)[33]>(qra_sm + QRA_TILE * 32 + QRA_TILE * 33);
	double (*W2s)[33] = Ws + 32;
	double (*Tsm)[33] = Ws + 64;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int rank = int(cluster.block_rank()), cb = blockIdx.x / QRC_CS;
	const int c = cb * 32 + lane;
	const bool cok = c < nc;
	const int rpc = (mp + QRC_CS - 1) / QRC_CS;
	const int r_lo = min(mp, rank * rpc), r_hi = min(mp, r_lo + rpc);
	const int ntiles = (r_hi - r_lo + QRA_TILE - 1) / QRA_TILE;
	auto load_tile = [&](int t) {
		for (int ii = warp; ii < QRA_TILE; ii += QRA_THREADS / 32) {
			const int row = r_lo + t * QRA_TILE + ii;
			const bool ok = row < r_hi;
			Vs[ii][lane] = ok ? Vbuf[(long long)row * QR_NB + lane] : 0.0;
			Cs[ii][lane] = (ok && cok) ? C[(long long)row * ldc + c] : 0.0;
		}
	};
	for (int e = threadIdx.x; e < 32 * 32; e += QRA_THREADS) Tsm[e >> 5][e & 31] = T[e];
	// phase 1: partial W = V^T C over this CTA's rows; thread (warp, lane) owns W[4 warp .. 4 warp + 3][lane]
	double acc[4] = {0.0, 0.0, 0.0, 0.0};
	for (int t = 0; t < ntiles; ++t) {
		if (t > 0) __syncthreads();
		load_tile(t);
		__syncthreads();
#pragma unroll 8
		for (int ii = 0; ii < QRA_TILE; ++ii) {
			const double cv = Cs[ii][lane];
			const double2 va = *reinterpret_cast<const double2*>(&Vs[ii][4 * warp]);
			const double2 vb = *reinterpret_cast<const double2*>(&Vs[ii][4 * warp + 2]);
			acc[0] += va.x * cv; acc[1] += va.y * cv; acc[2] += vb.x * cv; acc[3] += vb.y * cv;
		}
	}
	double* mine = Wp + ((size_t)blockIdx.x * 32 + 4 * warp) * 32 + lane;
#pragma unroll
	for (int q = 0; q < 4; ++q) mine[q * 32] = acc[q];
	__threadfence();
	cluster.sync();
	// phase 2: W = sum of the partials (fixed order), W2 = op(T) W
	{
		const double* base = Wp + ((size_t)cb * QRC_CS * 32 + 4 * warp) * 32 + lane;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			double s = 0.0;
#pragma unroll
			for (int r = 0; r < QRC_CS; ++r) s += __ldcg(base + (size_t)r * 1024 + q * 32);
			Ws[4 * warp + q][lane] = s;
		}
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const int kk = 4 * warp + q;
		double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
		for (int l = 0; l < 32; l += 2) {
			a0 += (transT ? Tsm[l][kk] : Tsm[kk][l]) * Ws[l][lane];
			a1 += (transT ? Tsm[l + 1][kk] : Tsm[kk][l + 1]) * Ws[l + 1][lane];
		}
		W2s[kk][lane] = a0 + a1;
	}
	__syncthreads();
	double w2[32];
#pragma unroll
	for (int kk = 0; kk < 32; ++kk) w2[kk] = W2s[kk][lane];
	// phase 3: C -= V W2 on this CTA's rows
	for (int t = 0; t < ntiles; ++t) {
		if (ntiles > 1) { __syncthreads(); load_tile(t); __syncthreads(); }
		for (int ii = warp; ii < QRA_TILE; ii += QRA_THREADS / 32) {
			const int row = r_lo + t * QRA_TILE + ii;
			if (row < r_hi && cok) {
				double d0 = 0.0, d1 = 0.0;
#pragma unroll
				for (int kk = 0; kk < 32; kk += 2) {
					const double2 v = *reinterpret_cast<const double2*>(&Vs[ii][kk]);
					d0 += v.x * w2[kk]; d1 += v.y * w2[kk + 1];
				}
				C[(long long)row * ldc + c] = Cs[ii][lane] - (d0 + d1);
			}
		}
	}
}

static void apply_block_reflector(double* C, size_t ldc, size_t mp, size_t nc, const double* Vbuf, const double* T, bool transT, double* Wp) {
	if (nc == 0 || mp == 0) return;
	if (ctx().qr_cluster && mp >= 64) {
		static bool attr_set = false;
		if (!attr_set) {
			XB_CUDA(cudaFuncSetAttribute(qr_apply_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(QRA_SMEM)));
			attr_set = true;
		}
		const unsigned ncb = unsigned((nc + 31) / 32);
		qr_apply_cluster_kernel<<<ncb * QRC_CS, QRA_THREADS, QRA_SMEM, ctx().stream>>>(C, (long long)ldc, int(mp), int(nc), Vbuf, T, transT ? 1 : 0, Wp);
		XB_LAUNCH_CHECK();
		return;
	}
	const unsigned ncb = unsigned((nc + 31) / 32), nch = unsigned((mp + QR_CHUNK - 1) / QR_CHUNK);
	const int ncpad = int(ncb * 32);
	dim3 grid(ncb, nch);
	qr_vtc_kernel<<<grid, 1024, 0, ctx().stream>>>(Vbuf, C, (long long)ldc, int(mp), int(nc), Wp, ncpad);
	XB_LAUNCH_CHECK();
	qr_update_kernel<<<grid, 1024, 0, ctx().stream>>>(C, (long long)ldc, int(mp), int(nc), Vbuf, T, transT ? 1 : 0, Wp, int(nch), ncpad);
	XB_LAUNCH_CHECK();
}

// ---- Cholesky-QR2 -----------------------------------------------------------------------------------------------------------
// Tall matrices of up to 128 columns — the first unfoldings of a TT-SVD, cores with large mode dimensions, anything with
// thousands of rows — are where the Householder path is weakest: ~10 dependent launches per 32 columns, twice (factor, then
// form Q), and beyond 2048 rows the panel no longer fits the cluster kernel (49 ms at 20000 x 100).  For these,
//     G = A^T A,  R1 = chol(G),  Q1 = A R1^-1,     G2 = Q1^T Q1,  R2 = chol(G2),  Q = Q1 R2^-1,  R = R2 R1
// is 7..9 launches of GEMM-rich work whose cost hardly depends on m (0.39 ms at 20000 x 100; table in DESIGN.md 3.2).  The
// Gram matrices are split-K batched GEMMs over row slabs; the factorisation, its inverse and the sum of up to 4 slab results
// are one single-CTA launch (small_f64.cu: chol_inv_kernel).  The second pass restores orthogonality to O(u) provided
// cond(A)^2 u << 1 (Yamamoto et al., "Roundoff error analysis of the CholeskyQR2 algorithm", ETNA 44 (2015)), so the path
// polices itself on the device: the first Cholesky must see a pivot ratio >= 1e-10 (cond(A) <~ 1e5), the second must start from
// ||Q1^T Q1 - I||_F <= 1/2.  Anything else — ill-conditioned or rank-deficient input, entries so large or small that the Gram
// matrix leaves the range of a double — is *declined* and the Householder path below runs instead: the decision is read back
// on the ordinary path (the callers synchronise for the rank decision right after anyway) and speculated under a round plan
// (reason 4), which replays the accept / decline decisions its recording run made (Context::chol_tape).  After two declines in
// a row the next 16 candidates of the same call go straight to Householder.  R has a positive diagonal here and a mixed-sign
// one from Householder; both are QRs of A (dgeqrf's signs are not part of the contract the reference's callers rely on:
// blasLapackWrapper.cpp:374-437 only promises A = Q R with orthonormal Q).
// split-K over row slabs: S slabs of `slab` rows plus `rest` rows that go on top of the first partial sum
struct RowSplit { size_t S, slab, rest; };
static RowSplit split_rows(size_t m, size_t out_elems) {
	size_t S = std::min<size_t>(std::max<size_t>(m / 128, 1), std::min<size_t>(1024, std::max<size_t>((size_t(1) << 22) / out_elems, 64)));
	for (size_t t = S; t > S / 2 && t >= 2; --t) if (m % t == 0) { S = t; break; }
	return RowSplit{S, m / S, m - S * (m / S)};
}
// P[0 .. parts) (na x nb each) <- partial sums of X^T Y over the m rows (X: m x na, ldx; Y: m x nb, ldy) as one batched GEMM (a
// single product would be a handful of CTAs walking all m rows).  Up to 4 partial sums are left for the consumer to add up while
// it loads (chol_inv_kernel), more — or any number with sum_all — go through one summation launch.  Returns the parts left.
static size_t xty_split(double* P, const double* X, size_t ldx, size_t na, const double* Y, size_t ldy, size_t nb, const RowSplit& rs, bool sum_all) {
	ProfScope ps("chol_gram");
	gemm_batched(P, nb, na * nb, na, nb, 1.0, X, ldx, rs.slab * ldx, true, rs.slab, Y, ldy, rs.slab * ldy, false, 0.0, rs.S);
	if (rs.rest) gemm(P, nb, na, nb, 1.0, X + rs.S * rs.slab * ldx, ldx, true, rs.rest, Y + rs.S * rs.slab * ldy, ldy, false, 1.0);
	if (rs.S == 1 || (rs.S <= 4 && !sum_all)) return rs.S;
	sum_parts(P, P, rs.S, na * nb);
	return 1;
}
// enqueues Q (m x n, ldq), R (n x n) of A (m x n, lda), n <= 128; the two status words f1, f2 receive 4 on a decline
static void cholqr2_enqueue(double* Q, size_t ldq, double* R, const double* A, size_t lda, size_t m, size_t n, unsigned int* f1, unsigned int* f2, bool clear) {
	const RowSplit rs = split_rows(m, n * n);
	DBuf P(rs.S * n * n), R1(n * n), R2(n * n), W(n * n), Q1(m * n);
	size_t parts = xty_split(P, A, lda, n, A, lda, n, rs, false);
	{ ProfScope ps("chol_fact"); chol_inv(P, parts, n, R1, W, false, 1e-10, f1, 4u, clear); }
	{ ProfScope ps("chol_apply"); gemm(Q1, n, m, n, 1.0, A, lda, false, n, W, n, true, 0.0); }
	parts = xty_split(P, Q1, n, n, Q1, n, n, rs, false);
	{ ProfScope ps("chol_fact"); chol_inv(P, parts, n, R2, W, true, 0.25, f2, 4u, clear); }
	{ ProfScope ps("chol_apply"); gemm(Q, ldq, m, n, 1.0, Q1, n, false, n, W, n, true, 0.0); }
	{ ProfScope ps("chol_rr"); gemm(R, n, n, n, 1.0, R2, n, false, n, R1, n, false, 0.0); }
}

static bool cholqr2(double* Q, double* R, const double* A, const size_t m, const size_t n) {
	Context& c = ctx();
	if (!c.qr_chol || c.chol_off || n > m) return false;
	// 129..256 columns beyond the reach of the cluster panel kernel (> 2048 rows: 15 ms at 4096 x 150): two column halves, block
	// Gram-Schmidt between them with the projection applied twice, Cholesky-QR2 on each half
	const bool wide = n > 128 && n <= 256 && m > 2048;
	if (!wide) {
		if (!chol_inv_fits(n)) return false;
		// measured crossover against the cluster panel kernels (profiles/r2_qr_cholqr2.txt), per call on the ordinary path: ~250 us
		// whatever m at 128 columns against 265 / 353 / 538 us at 512 / 768 / 2048 rows; 175 us against 136 / 185 / 261 us at 64
		// columns; beyond 2048 rows (where the panels leave the cluster kernel) 0.15 ms against 1.8 ms at 4096 x 16 and 0.39 ms
		// against 49 ms at 20000 x 100.  Up to 32 columns are a single panel (90 us at 1024 x 32) until that limit.  Inside a round
		// plan, where launches cost nothing and what counts is SM time, 96+ columns win from 256 rows on (config 5: 1310 -> 1535
		// items/s), so that class starts there on both paths (240 -> 225 us at 256 x 128 on the ordinary one) and plan and
		// ordinary path keep taking the same decisions.
		const size_t min_rows = size_t(std::max(c.qr_chol_min_rows, 0));
		const size_t thr = n >= 96 ? std::min<size_t>(min_rows, 256) : (n <= 32 && min_rows > 0) ? std::max<size_t>(min_rows, 2049) : min_rows;
		if (m < thr) return false;
	}
	if (c.chol_tape_mode == 2) {                                       // capture of a plan: do what the recording run did
		const bool take = c.chol_tape_pos < c.chol_tape.size() && c.chol_tape[c.chol_tape_pos] != 0;
		c.chol_tape_pos += 1;
		if (!take) return false;
	} else if (c.chol_skip > 0) {                                      // back-off after repeated declines (ill-conditioned workload)
		c.chol_skip -= 1;
		if (c.chol_tape_mode == 1) c.chol_tape.push_back(0);
		return false;
	}
	ProfScope prof("qr_chol");
	const bool spec = c.speculate;
	DBuf fl(2);
	unsigned int* fw = reinterpret_cast<unsigned int*>(fl.p);
	auto word = [&](int i) { return spec ? c.spec_flag : fw + i; };
	const int nwords = wide ? 4 : 2;
	if (!wide) {
		cholqr2_enqueue(Q, n, R, A, n, m, n, word(0), word(1), !spec);
	} else {
		const size_t n1 = (n + 1) / 2, n2 = n - n1;
		const RowSplit rs = split_rows(m, n1 * n2);
		DBuf R11(n1 * n1), R22(n2 * n2), R12(n1 * n2), Rb(n1 * n2), T(m * n2), P(rs.S * n1 * n2);
		cholqr2_enqueue(Q, n, R11, A, n, m, n1, word(0), word(1), !spec);                  // Q1 = Q(:, :n1)
		copy2d(T, n2, A + n1, n, m, n2);
		for (int pass = 0; pass < 2; ++pass) {                                              // T -= Q1 (Q1^T T), twice
			double* Rx = pass ? Rb.p : R12.p;
			xty_split(P, Q, n, n1, T, n2, n2, rs, true);
			copy(Rx, P, n1 * n2);
			gemm(T, n2, m, n2, -1.0, Q, n, false, n1, Rx, n2, false, 1.0);
		}
		axpy(R12, 1.0, Rb, n1 * n2);
		cholqr2_enqueue(Q + n1, n, R22, T, n2, m, n2, word(2), word(3), !spec);             // Q2 = Q(:, n1:)
		fill(R, 0.0, n * n);
		copy2d(R, n, R11, n1, n1, n1);
		copy2d(R + n1, n, R12, n2, n1, n2);
		copy2d(R + n1 * n + n1, n, R22, n2, n2, n2);
	}
	if (spec) return true;
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, fl.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	unsigned int w[4];
	std::memcpy(w, c.h_scratch, sizeof w);
	bool ok = true;
	for (int i = 0; i < nwords; ++i) ok = ok && w[i] == 0u;
	if (c.chol_tape_mode == 1) c.chol_tape.push_back(ok ? 1 : 0);
	if (ok) { c.chol_declines = 0; return true; }
	ProfScope declined("qr_chol_declined");
	if (++c.chol_declines >= 2) c.chol_skip = 16;
	return false;
}

void qr(double* Q, double* R, const double* A, size_t m, size_t n, bool defer_q) {
	XB_REQUIRE(m > 0 && n > 0, "Dimension m and n must be larger than zero");    // blasLapackWrapper.cpp:392-393
	XB_REQUIRE(m <= 0x7fffffffULL && n <= 0x7fffffffULL, "Dimension to large for QR");
	ProfScope prof("qr");
	const size_t k = std::min(m, n);
	if (qr_small_fits(m, n)) {           // rank ramps: factor + explicit Q in one single-CTA launch (small_f64.cu)
		qr_small(Q, (long long)k, 1, R, (long long)n, 1, A, (long long)n, 1, m, n);
		return;
	}
	if (cholqr2(Q, R, A, m, n)) return;
	const size_t npanels = (k + QR_NB - 1) / QR_NB;
	DBuf W(m * n), Vall(npanels * m * QR_NB), Tall(npanels * QR_NB * QR_NB);
	const size_t nch_max = (m + QR_CHUNK - 1) / QR_CHUNK;
	const size_t ncpad_max = ((std::max(n, k) + 31) / 32) * 32;
	DBuf Wp(std::max<size_t>(nch_max, QRC_CS) * 32 * ncpad_max);     // partial V^T C: [row chunk or cluster rank][32][columns]
	// work on A * 2^-e (exact): the reflector norms are sums of squares and TT cores carry norms like 1e33 .. 1e150
	DBuf sc(2);
	amax_scale_dev(sc, A, m * n);
	scale_by_dev(W, A, m * n, sc);

	static bool attr_set = false;
	const size_t fixed_smem = (3 * 32 * 33 + 32 + 32 + 4) * sizeof(double);
	const size_t smem_cap = std::min<size_t>(ctx().max_smem_optin, 227 * 1024);
	if (!attr_set) {
		XB_CUDA(cudaFuncSetAttribute(qr_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
		attr_set = true;
	}
	for (size_t p = 0; p < npanels; ++p) {
		const size_t j0 = p * QR_NB, nbe = std::min<size_t>(QR_NB, k - j0), mp = m - j0;
		double* Wpanel = W.p + j0 * n + j0;
		double* Vp = Vall.p + p * m * QR_NB;
		double* Tp = Tall.p + p * QR_NB * QR_NB;
		const size_t need = fixed_smem + mp * QR_NB * sizeof(double);
		if (launch_panel_cluster(Wpanel, (long long)n, mp, nbe, Vp, Tp)) {
		} else if (need <= smem_cap) {
			qr_panel_kernel<true><<<1, QR_PANEL_WARPS * 32, need, ctx().stream>>>(Wpanel, (long long)n, int(mp), int(nbe), Vp, Tp);
		} else {
			qr_panel_kernel<false><<<1, QR_PANEL_WARPS * 32, fixed_smem, ctx().stream>>>(Wpanel, (long long)n, int(mp), int(nbe), Vp, Tp);
		}
		XB_LAUNCH_CHECK();
		const size_t nc = n - (j0 + nbe);
		apply_block_reflector(Wpanel + nbe, n, mp, nc, Vp, Tp, true, Wp);
	}
	// R = upper trapezoid of the first k rows
	{
		const size_t total = k * n;
		const unsigned blocks = unsigned(std::min<size_t>((total + 255) / 256, size_t(ctx().num_sms) * 8));
		qr_extract_r_kernel<<<blocks, 256, 0, ctx().stream>>>(R, W, k, n, sc.p + 1);
		XB_LAUNCH_CHECK();
	}
	// Q = H_1 ... H_k [I; 0] : apply the block reflectors in reverse order to the identity.  In a sweep nothing downstream of
	// R needs Q (it only becomes the new core), so with defer_q it is formed on the side stream, next to the main stream's
	// push of R and the next factorization; the reflector storage is released on the stream that read it last.
	const bool defer = defer_q && ctx().qr_defer != 0 && npanels > 1;
	if (defer) aux_fork();
	{
		AuxScope side(defer);
		set_identity(Q, m, k, k);
		for (size_t p = npanels; p-- > 0;) {
			const size_t j0 = p * QR_NB, mp = m - j0;
			apply_block_reflector(Q + j0 * k + j0, k, mp, k - j0, Vall.p + p * m * QR_NB, Tall.p + p * QR_NB * QR_NB, false, Wp);
		}
		Vall.reset(); Tall.reset(); Wp.reset();
	}
}

void lq(double* L, double* Q, const double* A, size_t m, size_t n) {
	const size_t k = std::min(m, n);
	if (qr_small_fits(n, m)) {           // QR of A^T through strides: Q = Qt^T (k x n), L = Rt^T (m x k); no transposed copies
		ProfScope prof("qr");
		qr_small(Q, 1, (long long)n, L, 1, (long long)k, A, 1, (long long)n, n, m);
		return;
	}
	DBuf At(m * n), Qt(n * k), Rt(k * m);
	transpose(At, A, m, n);            // n x m
	qr(Qt, Rt, At, n, m);              // A^T = Qt * Rt
	transpose(Q, Qt, n, k);            // k x n
	transpose(L, Rt, k, m);            // m x k
}

void rq(double* R, double* Q, const double* A, size_t m, size_t n) {
	// RQ in LAPACK's convention through the QR of the doubly reversed transpose: with Ar(j,i) = A(m-1-i, n-1-j) = Qr Rr,
	// R(i,l) = Rr(k-1-l, m-1-i) is upper trapezoidal (bottom-right aligned) and Q(l,j) = Qr(n-1-j, k-1-l).
	const size_t k = std::min(m, n);
	DBuf Ar(m * n), Qr(n * k), Rr(k * m);
	transpose_reverse(Ar, A, m, n);
	qr(Qr, Rr, Ar, n, m);
	transpose_reverse(Q, Qr, n, k);
	transpose_reverse(R, Rr, k, m);
}

// ---- rank revealing variants ------------------------------------------------------------------------------------
// speculative variant: raises the flag instead of reporting to the host
__global__ void diag_spec_kernel(const double* __restrict__ A, const size_t k, const size_t ld, unsigned int* __restrict__ flag) {
	double mn = HUGE_VAL, mx = 0.0;
	for (size_t i = threadIdx.x; i < k; i += 32) { const double v = fabs(A[i * ld + i]); mn = fmin(mn, v); mx = fmax(mx, v); }
	for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
	if (threadIdx.x == 0 && !(mx > 0.0 && mn >= 16.0 * 2.220446049250313e-16 * mx)) *flag = 1u;
}

static bool diag_is_full_rank(const double* M, size_t k, size_t ld) {
	if (ctx().speculate) {
		diag_spec_kernel<<<1, 32, 0, ctx().stream>>>(M, k, ld, ctx().spec_flag);
		XB_LAUNCH_CHECK();
		return true;
	}
	DBuf mm(2);
	diag_minmax_kernel<<<1, 32, 0, ctx().stream>>>(M, k, ld, mm);
	XB_LAUNCH_CHECK();
	Context& c = ctx();
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, mm.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	const double mn = c.h_scratch[0], mx = c.h_scratch[1];
	return mx > 0.0 && mn >= 16.0 * 2.220446049250313e-16 * mx;
}

static size_t numerical_rank(const std::vector<double>& S) {
	size_t rank = 1;
	for (size_t j = 1; j < S.size(); ++j) if (S[j] >= 16.0 * 2.220446049250313e-16 * S[0] && S[j] > 0.0) rank = j + 1; else break;
	return rank;
}

size_t qc(double* Q, double* C, const double* A, size_t m, size_t n, bool defer_q) {
	const size_t k = std::min(m, n);
	qr(Q, C, A, m, n, defer_q);
	if (diag_is_full_rank(C, k, n)) return k;
	aux_join();
	// (near) rank deficient: reveal the rank through the SVD of the triangular factor, R = U S Vt:
	//   A = (Q U_r) (S_r Vt_r)
	Svd s;
	s.factor(C, k, n);
	const size_t rank = numerical_rank(s.S);
	DBuf U(k * rank), Cn(rank * n), Qn(m * rank);
	s.extract(U, Cn, rank, false, true, nullptr);
	gemm(Qn, rank, m, rank, 1.0, Q, k, false, k, U, rank, false, 0.0);
	copy(Q, Qn, m * rank);
	copy(C, Cn, rank * n);
	return rank;
}

size_t cq(double* C, double* Q, const double* A, size_t m, size_t n) {
	const size_t k = std::min(m, n);
	lq(C, Q, A, m, n);
	if (diag_is_full_rank(C, k, k)) return k;
	// L = U S Vt  ->  A = (U_r S_r) (Vt_r Q)
	Svd s;
	s.factor(C, m, k);
	const size_t rank = numerical_rank(s.S);
	DBuf Cn(m * rank), Vt(rank * k), Qn(rank * n);
	s.extract(Cn, Vt, rank, true, false, nullptr);
	gemm(Qn, n, rank, n, 1.0, Vt, k, false, k, Q, n, false, 0.0);
	copy(C, Cn, m * rank);
	copy(Q, Qn, rank * n);
	return rank;
}

} // namespace xb
