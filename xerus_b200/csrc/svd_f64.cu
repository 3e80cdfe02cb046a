// One-sided (Hestenes) block-Jacobi SVD for sm_100a, FP64 — replaces LAPACKE_dgesdd behind blasWrapper::svd
// (reference: src/xerus/blasLapackWrapper.cpp:201-232) and carries the truncation of calculate_svd
// (src/xerus/tensor.cpp:1464-1489: rank cap, eps rule, crop of U / Vt, optional Sigma fold-in of round_edge,
// src/xerus/tensorNetwork.cpp:769) in its epilogue, so cropped factors are never materialised at full size.
//
// Layout: the working matrix is kept TRANSPOSED in HBM, GT[j] = [ x_j (mdot entries) ; v_j (nw entries) ], one
// contiguous row per column of [X; V], so that a block of columns is a contiguous slab.  A CTA owns one pair of
// column blocks, keeps the slab in shared memory (2*bw rows of mt doubles, up to ~200 KB), and runs the inner
// rotations with one warp per column pair: three fused dot products (warp-shuffle reduction), one rotation of the
// stacked vector.  Every pair is met exactly once per sweep: the first round of the outer tournament runs the full
// 2*bw-player inner tournament (intra + cross pairs), later rounds only the bw cross rounds.
// Tall inputs are first reduced by QR (SVD of the triangular factor), wide inputs are handled as the transpose.
#include "xb_internal.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace xb {

constexpr double DBL_EPS = 2.220446049250313e-16;

__global__ void svd_init_kernel(double* __restrict__ GT, const int ld, const int npad, const int mdot, const int voff, const int nw,
                                const double* __restrict__ src, const long long rs, const long long cs, const double* __restrict__ scale) {
	const double sc = *scale;
	const size_t total = (size_t)npad * ld;
	for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const int j = int(e / ld), i = int(e % ld);
		double v = 0.0;
		if (j < nw) {
			if (i < mdot) v = src[(long long)i * rs + (long long)j * cs] * sc;
			else if (i - voff == j) v = 1.0;
		}
		GT[e] = v;
	}
}

// blockIdx.x = pair index inside outer round `round`; nblk even.  mode_full: all pairs among the 2*bw columns,
// otherwise only the bw*bw cross pairs.  With loop != 0 (single block pair) the kernel sweeps until no rotation
// happened in a sweep, and reports the number of sweeps in info[1].
__global__ void jacobi_block_kernel(double* __restrict__ GT, const int ldg, const int mt, const int mdot, const int bw,
                                    const int nblk, const int round, const int mode_full, const double tol,
                                    unsigned int* __restrict__ info, const int loop, const int max_sweeps) {
	extern __shared__ double S[];
	__shared__ unsigned int s_rot;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	int pb, qb;
	{
		const int pi = blockIdx.x, N1 = nblk - 1;
		if (nblk == 2) { pb = 0; qb = 1; }
		else if (pi == 0) { pb = N1; qb = round % N1; }
		else { pb = (round + pi) % N1; qb = (round - pi + N1) % N1; }
	}
	// load the two column blocks (rows of GT)
	for (int r = warp; r < 2 * bw; r += nwarps) {
		const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
		const double* src = GT + (size_t)grow * ldg;
		double* dst = S + (size_t)r * mt;
		for (int i = lane; i < mt; i += 32) dst[i] = src[i];
	}
	if (threadIdx.x == 0) s_rot = 0;
	__syncthreads();

	const int N = 2 * bw;
	const int inner_rounds = mode_full ? (N - 1) : bw;
	unsigned int total_rot = 0;
	int sweeps = 0;
	for (;;) {
		for (int rr = 0; rr < inner_rounds; ++rr) {
			for (int pi = warp; pi < bw; pi += nwarps) {
				int a, b;
				if (mode_full) {
					if (pi == 0) { a = N - 1; b = rr; }
					else { a = (rr + pi) % (N - 1); b = (rr - pi + N - 1) % (N - 1); }
				} else {
					a = pi; b = bw + (pi + rr) % bw;
				}
				double* x = S + (size_t)a * mt;
				double* y = S + (size_t)b * mt;
				double aa = 0.0, bb = 0.0, ab = 0.0;
				for (int i = lane; i < mdot; i += 32) {
					const double xi = x[i], yi = y[i];
					aa += xi * xi; bb += yi * yi; ab += xi * yi;
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) {
					aa += __shfl_xor_sync(0xffffffffu, aa, o);
					bb += __shfl_xor_sync(0xffffffffu, bb, o);
					ab += __shfl_xor_sync(0xffffffffu, ab, o);
				}
				if (fabs(ab) > tol * sqrt(aa) * sqrt(bb)) {
					const double zeta = (bb - aa) / (2.0 * ab);
					const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
					const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
					for (int i = lane; i < mt; i += 32) {
						const double xi = x[i], yi = y[i];
						x[i] = c * xi - s * yi;
						y[i] = s * xi + c * yi;
					}
					if (lane == 0) atomicAdd(&s_rot, 1u);
				}
			}
			__syncthreads();
		}
		++sweeps;
		const unsigned int rot = s_rot;
		__syncthreads();
		if (threadIdx.x == 0) s_rot = 0;
		total_rot += rot;
		if (!loop || rot == 0 || sweeps >= max_sweeps) {
			if (threadIdx.x == 0 && loop) { info[1] = (unsigned)sweeps; info[2] = rot; }
			break;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0 && total_rot) atomicAdd(&info[0], total_rot);
	// write back
	for (int r = warp; r < 2 * bw; r += nwarps) {
		const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
		double* dst = GT + (size_t)grow * ldg;
		const double* src = S + (size_t)r * mt;
		for (int i = lane; i < mt; i += 32) dst[i] = src[i];
	}
}

// Persistent variant: ONE cooperative launch runs every round of every sweep (grid = nblk/2 CTAs, all co-resident),
// separated by grid-wide barriers; convergence is decided on the device.
//  * a warp owns one column pair per inner round and keeps both stacked columns in REGISTERS (EPL doubles per lane
//    each): one batch of shared-memory loads, the cross product from the leading part, the rotation, one batch of
//    stores — every element is read and written once per pair visit and all loads are in flight together;
//  * only the cross product is reduced (warp shuffles): the squared column norms are cached in shared memory and
//    updated with the rotation (alpha' = alpha - t*gamma, beta' = beta + t*gamma), refreshed whenever a block is loaded;
//  * the rotation needs two rsqrt and one division: t = sign(d) 2g / (|d| + sqrt(d^2 + 4 g^2)), c = rsqrt(1 + t^2);
//  * a sweep in which no pair had |cos| > 1e-9 is the last one (cyclic Jacobi converges quadratically), so no
//    verification sweep is spent.
// Rows of GT are [ x (mdot, zero padded to a multiple of 32) | v (nw, zero padded to a multiple of 32) ].
__device__ __forceinline__ double xb_rsqrt(double x) { return rsqrt(x); }

// T = double (the kernel is generic in the element type; an FP32 pre-sweep variant was measured in round 1, gave nothing and is gone).
// EH >= max(voff, ld - voff) / 32 : elements per lane of one column part (register tile of a warp = one column pair).
template <typename T, int EH, int MAXT>
__global__ void __launch_bounds__(MAXT) jacobi_persistent_kernel(T* __restrict__ GT, const int ld, const int epl_x, const int epl_v,
                                                                const int bw, const int nblk, const T tol2, const T big2,
                                                                unsigned int* __restrict__ counters, unsigned int* __restrict__ info,
                                                                const int max_sweeps, unsigned int* ready, const int recursive) {
	extern __shared__ double S_raw[];
	T* S = reinterpret_cast<T*>(S_raw);
	__shared__ unsigned int s_rot, s_big;
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int N = 2 * bw;
	const int voff = 32 * epl_x;
	T* nrm = S + (size_t)N * ld;                // [N] cached squared norms of the resident columns
	// schedule of the full 2*bw-player tournament, precomputed once: keeps the integer modulo out of the per-round chain
	unsigned short* sched = reinterpret_cast<unsigned short*>(nrm + N + (N & 1));      // [(N-1) * bw] : a | b << 8
	for (int e = threadIdx.x; e < (N - 1) * bw; e += blockDim.x) {
		const int rr = e / bw, pi = e % bw;
		int a, b;
		if (pi == 0) { a = N - 1; b = rr; }
		else { a = (rr + pi) % (N - 1); b = (rr - pi + N - 1) % (N - 1); }
		sched[e] = (unsigned short)(a | (b << 8));
	}
	const int nrounds = (nblk == 2) ? 1 : nblk - 1;
	if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
	__syncthreads();
	int sweeps = 0, ground = 0;                  // ground: global round counter (version of the block flags)
	unsigned int last_rot = 1, last_big = 1;
	unsigned int my_rot = 0, my_big = 0;         // per-warp counters (lane 0), flushed once per visit
	long long tk_load = 0, tk_inner = 0, tk_store = 0, tk_sync = 0, tk0 = 0;   // phase cycle counters (thread 0 of block 0)
	const bool timing = (info[0] == 0xC10C) && blockIdx.x == 0 && threadIdx.x == 0;
	for (; sweeps < max_sweeps; ) {
		for (int round = 0; round < nrounds; ++round, ++ground) {
			int pb, qb;
			bool loadp = true, storep = true, full = (round == 0);
			if (recursive) {
				// Recursive bipartite tournament (nblk a power of two): phase with groups of g blocks pairs the lower half of a
				// group (stationary in this CTA for the whole phase) with the upper half, which rotates through the group's
				// CTAs; then both halves are split again.  nblk/2 + nblk/4 + ... + 1 = nblk - 1 rounds, every block pair met
				// once; the last phase (g = 2) also runs the pairs inside the two blocks.  Per round only ONE block moves, and
				// a CTA depends only on the neighbour that held that block: point-to-point flags instead of a grid barrier.
				int g = nblk, t = round;
				while (t >= (g >> 1)) { t -= (g >> 1); g >>= 1; }
				const int h = g >> 1, G = blockIdx.x / h, j = blockIdx.x % h;
				pb = G * g + j; qb = G * g + h + ((j + t) & (h - 1));
				loadp = (t == 0); storep = (t == h - 1); full = (g == 2);
				if (threadIdx.x == 0) {
					// wait until the global copies reflect round ground - 1 (bounded spin: never hang the GPU)
					unsigned int spins = 0;
					volatile unsigned int* rd = ready;
					while ((rd[qb] < (unsigned)ground || (loadp && rd[pb] < (unsigned)ground)) && spins < (1u << 27)) ++spins;
					if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
					__threadfence();
				}
				__syncthreads();
			} else {
				const int pi = blockIdx.x, N1 = nblk - 1;
				if (nblk == 2) { pb = 0; qb = 1; }
				else if (pi == 0) { pb = N1; qb = round % N1; }
				else { pb = (round + pi) % N1; qb = (round - pi + N1) % N1; }
			}
			if (timing) tk0 = clock64();
			for (int r = (loadp ? 0 : bw) + warp; r < N; r += nwarps) {
				const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
				const T* src = GT + (size_t)grow * ld + lane;
				T* dst = S + (size_t)r * ld + lane;
				T ss = T(0);
#pragma unroll
				for (int k0 = 0; k0 < EH; k0 += 4) {
					T v[4], w[4];
#pragma unroll
					for (int k = 0; k < 4; ++k) { v[k] = (k0 + k < epl_x) ? __ldcg(src + 32 * (k0 + k)) : T(0); w[k] = (k0 + k < epl_v) ? __ldcg(src + voff + 32 * (k0 + k)) : T(0); }
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						if (k0 + k < epl_x) { dst[32 * (k0 + k)] = v[k]; ss += v[k] * v[k]; }
						if (k0 + k < epl_v) dst[voff + 32 * (k0 + k)] = w[k];
					}
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
				if (lane == 0) nrm[r] = ss;
			}
			__syncthreads();
			if (timing) { const long long t1 = clock64(); tk_load += t1 - tk0; tk0 = t1; }
			const int inner_rounds = full ? (N - 1) : bw;
			for (int rr = 0; rr < inner_rounds; ++rr) {
				for (int pi = warp; pi < bw; pi += nwarps) {
					int a, b;
					if (full) { const unsigned int ab_ = sched[rr * bw + pi]; a = ab_ & 255; b = ab_ >> 8; }
					else { a = pi; b = bw + ((pi + rr) & (bw - 1)); }      // bw is a power of two
					T* x = S + (size_t)a * ld + lane;
					T* y = S + (size_t)b * ld + lane;
					T xr[EH], yr[EH];
#pragma unroll
					for (int k = 0; k < EH; ++k) { xr[k] = (k < epl_x) ? x[32 * k] : T(0); yr[k] = (k < epl_x) ? y[32 * k] : T(0); }
					T gacc[4] = {T(0), T(0), T(0), T(0)};                  // independent chains: DFMA latency is ~20 cycles here
#pragma unroll
					for (int k = 0; k < EH; ++k) gacc[k & 3] += xr[k] * yr[k];      // padding entries are zero
					T g = (gacc[0] + gacc[1]) + (gacc[2] + gacc[3]);
					const T aa = nrm[a], bb = nrm[b];
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
					const T gg = g * g, ab = aa * bb;
					if (gg > tol2 * ab) {
						// the accumulated-rotation part is only touched by pairs that rotate: fetch it while the parameters are computed
						T xv[EH], yv[EH];
#pragma unroll
						for (int k = 0; k < EH; ++k) { xv[k] = (k < epl_v) ? x[voff + 32 * k] : T(0); yv[k] = (k < epl_v) ? y[voff + 32 * k] : T(0); }
						// c^2 = (1 + |d|/h)/2, s = sign(d) 2g / (2 h c), t = s/c  with h = sqrt(d^2 + 4 g^2): two rsqrt, no division
						const T d = bb - aa;
						const T rh = xb_rsqrt(d * d + T(4) * gg);
						const T c2 = T(0.5) + T(0.5) * fabs(d) * rh;
						const T rc = xb_rsqrt(c2);
						const T c = c2 * rc;
						const T s = (d >= T(0) ? g : -g) * rh * rc;
#pragma unroll
						for (int k = 0; k < EH; ++k) {
							if (k < epl_x) { x[32 * k] = c * xr[k] - s * yr[k]; y[32 * k] = s * xr[k] + c * yr[k]; }
							if (k < epl_v) { x[voff + 32 * k] = c * xv[k] - s * yv[k]; y[voff + 32 * k] = s * xv[k] + c * yv[k]; }
						}
						// cached norms: alpha' = alpha - t*gamma, beta' = beta + t*gamma.  When a rotation moves most of a
						// column's mass the update cancels; then both norms are recomputed from the rotated registers.
						const T t = s * rc;
						T na = aa - t * g, nb = bb + t * g;
						if (na < T(0.25) * aa || nb < T(0.25) * bb) {
							T sa = T(0), sb = T(0);
#pragma unroll
							for (int k = 0; k < EH; ++k) {
								const T xn = c * xr[k] - s * yr[k], yn = s * xr[k] + c * yr[k];
								sa += xn * xn; sb += yn * yn;
							}
#pragma unroll
							for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
							na = sa; nb = sb;
						}
						if (lane == 0) {
							nrm[a] = na; nrm[b] = nb;
							my_rot += 1;
							if (gg > big2 * ab) my_big += 1;
						}
					}
				}
				__syncthreads();
			}
			if (timing) { const long long t1 = clock64(); tk_inner += t1 - tk0; tk0 = t1; }
			if (lane == 0 && my_rot) { atomicAdd(&s_rot, my_rot); atomicAdd(&s_big, my_big); my_rot = 0; my_big = 0; }
			for (int r = (storep ? 0 : bw) + warp; r < N; r += nwarps) {
				const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
				T* dst = GT + (size_t)grow * ld + lane;
				const T* src = S + (size_t)r * ld + lane;
#pragma unroll
				for (int k0 = 0; k0 < EH; k0 += 4) {
					T v[4], w[4];
#pragma unroll
					for (int k = 0; k < 4; ++k) { v[k] = (k0 + k < epl_x) ? src[32 * (k0 + k)] : T(0); w[k] = (k0 + k < epl_v) ? src[voff + 32 * (k0 + k)] : T(0); }
#pragma unroll
					for (int k = 0; k < 4; ++k) { if (k0 + k < epl_x) dst[32 * (k0 + k)] = v[k]; if (k0 + k < epl_v) dst[voff + 32 * (k0 + k)] = w[k]; }
				}
			}
			if (timing) { const long long t1 = clock64(); tk_store += t1 - tk0; tk0 = t1; }
			if (recursive) {
				// the barrier orders every thread's row stores before thread 0's fence; the fence is cumulative, so one
				// thread publishing is enough (a fence in all 256 threads was 8 % of the kernel's stall samples)
				__syncthreads();
				if (threadIdx.x == 0) {
					__threadfence();
					volatile unsigned int* rd = ready;
					if (storep) rd[pb] = (unsigned)ground + 1u;
					rd[qb] = (unsigned)ground + 1u;
				}
			} else if (nblk > 2) { __threadfence(); grid.sync(); }
			else __syncthreads();
			if (timing) { const long long t1 = clock64(); tk_sync += t1 - tk0; tk0 = t1; }
		}
		++sweeps;
		unsigned int rot, big;
		if (nblk > 2) {
			if (threadIdx.x == 0) { atomicAdd(&counters[2 * (sweeps - 1)], s_rot); atomicAdd(&counters[2 * (sweeps - 1) + 1], s_big); s_rot = 0; s_big = 0; }
			__threadfence();
			grid.sync();
			rot = *((volatile unsigned int*)&counters[2 * (sweeps - 1)]);
			big = *((volatile unsigned int*)&counters[2 * (sweeps - 1) + 1]);
		} else {
			rot = s_rot; big = s_big;
			__syncthreads();
			if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
			__syncthreads();
		}
		last_rot = rot; last_big = big;
		if (big == 0) break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { info[1] = (unsigned)sweeps; info[2] = last_big; info[3] = last_rot; }
	if (timing) { info[4] = (unsigned)(tk_load >> 10); info[5] = (unsigned)(tk_inner >> 10); info[6] = (unsigned)(tk_store >> 10); info[7] = (unsigned)(tk_sync >> 10); }
}

// D(8x8) += A(8x4) B(4x8) on the FP64 tensor pipe.  Lane (g = lane >> 2, t = lane & 3) holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_884(double& d0, double& d1, const double a, const double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// V_new(r, :) = sum_r' Js(r, r') V_old(r', :) for the N = 8 MT = 4 KT resident rows (accumulated-rotation part).  Every warp
// owns 8-element tiles of the row length: it loads its tile of all N rows, multiplies on the tensor pipe, writes the tile back.
template <int MT, int KT>
__device__ __forceinline__ void apply_rotations_to_v(double* __restrict__ Sv, const int lds, const double* __restrict__ Js, const int ldj,
                                                    const int ntiles, const int warp, const int nwarps, const int lane) {
	const int g = lane >> 2, t = lane & 3;
	double afr[MT][KT];
#pragma unroll
	for (int mt = 0; mt < MT; ++mt)
#pragma unroll
		for (int kt = 0; kt < KT; ++kt) afr[mt][kt] = Js[(8 * mt + g) * ldj + 4 * kt + t];
	for (int tile = warp; tile < ntiles; tile += nwarps) {
		double* base = Sv + 8 * tile;
		double bfr[KT];
#pragma unroll
		for (int kt = 0; kt < KT; ++kt) bfr[kt] = base[(size_t)(4 * kt + t) * lds + g];
		double acc[MT][2];
#pragma unroll
		for (int mt = 0; mt < MT; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
#pragma unroll
		for (int kt = 0; kt < KT; ++kt)
#pragma unroll
			for (int mt = 0; mt < MT; ++mt) dmma_884(acc[mt][0], acc[mt][1], afr[mt][kt], bfr[kt]);
		__syncwarp();
#pragma unroll
		for (int mt = 0; mt < MT; ++mt) *reinterpret_cast<double2*>(base + (size_t)(8 * mt + g) * lds + 2 * t) = make_double2(acc[mt][0], acc[mt][1]);
	}
}

// As apply_rotations_to_v, but TB tiles per step so that consecutive DMMAs are independent (a chain of dependent DMMAs
// costs their full latency each): the warp's tile count must be a multiple of TB.
template <int MT, int KT, int TB>
__device__ __forceinline__ void apply_rotations_tiles(double* __restrict__ Sv, const int lds, const double* __restrict__ Js, const int ldj,
                                                     const int ntiles, const int warp, const int nwarps, const int lane) {
	const int g = lane >> 2, t = lane & 3;
	double afr[MT][KT];
#pragma unroll
	for (int mt = 0; mt < MT; ++mt)
#pragma unroll
		for (int kt = 0; kt < KT; ++kt) afr[mt][kt] = Js[(8 * mt + g) * ldj + 4 * kt + t];
	for (int tile0 = warp; tile0 < ntiles; tile0 += nwarps * TB) {
		double bfr[TB][KT];
#pragma unroll
		for (int tb = 0; tb < TB; ++tb)
#pragma unroll
			for (int kt = 0; kt < KT; ++kt) bfr[tb][kt] = Sv[(size_t)(4 * kt + t) * lds + 8 * (tile0 + tb * nwarps) + g];
		double acc[TB][MT][2];
#pragma unroll
		for (int tb = 0; tb < TB; ++tb)
#pragma unroll
			for (int mt = 0; mt < MT; ++mt) { acc[tb][mt][0] = 0.0; acc[tb][mt][1] = 0.0; }
#pragma unroll
		for (int kt = 0; kt < KT; ++kt)
#pragma unroll
			for (int tb = 0; tb < TB; ++tb)
#pragma unroll
				for (int mt = 0; mt < MT; ++mt) dmma_884(acc[tb][mt][0], acc[tb][mt][1], afr[mt][kt], bfr[tb][kt]);
		__syncwarp();
#pragma unroll
		for (int tb = 0; tb < TB; ++tb)
#pragma unroll
			for (int mt = 0; mt < MT; ++mt)
				*reinterpret_cast<double2*>(Sv + (size_t)(8 * mt + g) * lds + 8 * (tile0 + tb * nwarps) + 2 * t) = make_double2(acc[tb][mt][0], acc[tb][mt][1]);
	}
}

// ---- specialised persistent kernel (FP64, both row parts 64 * EP2 doubles long) ------------------------------------------
// Same algorithm and schedule as jacobi_persistent_kernel, rebuilt around what bounds an inner round on sm_100a (measured,
// profiles/micro/fp64_micro.cu): the shared-memory pipe (128 B/clk: moving both 4 KB rows of 8 pairs in and out is 1024
// cycles), instruction issue, and a ~650-cycle dependent chain (load, dot, 5-stage butterfly, two rsqrt, rotate, barrier).
//  * row length is a template parameter, a lane owns PAIRS of consecutive elements: 128-bit accesses, no bounds predicates;
//  * in a cross visit (block p against block q) warp w keeps column w of p in REGISTERS for all rounds; only the q column
//    travels through shared memory;
//  * JACC (8-column blocks): the accumulated-rotation part of the rows is not rotated pair by pair.  The rotations of a visit
//    are multiplied up in a 16 x 16 matrix (rows of 16 entries) and applied once per visit as Js * V with DMMA: the same
//    arithmetic, off the latency chain and off the shared-memory pipe.
template <int EP2, int MAXT, bool JACC>
__global__ void __launch_bounds__(MAXT) jacobi_fast_kernel(double* __restrict__ GT, const int bw, const int nblk, const double tol2, const double big2,
                                                          unsigned int* __restrict__ counters, unsigned int* __restrict__ info,
                                                          const int max_sweeps, unsigned int* ready, const int recursive) {
	constexpr int LD = 128 * EP2;               // doubles per row in global memory: [x : 64 EP2 | v : 64 EP2]
	constexpr int LDS = LD + 4;                 // shared-memory row stride: 4 rows x 8 elements of a DMMA operand on distinct banks
	constexpr int V2 = 32 * EP2;                // offset of the v part in double2 units
	extern __shared__ double S_fast[];
	double* S = S_fast;
	__shared__ unsigned int s_rot, s_big;
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int N = 2 * bw;
	const int ldj = N + 4;
	double* nrm = S + (size_t)N * LDS;
	unsigned short* sched = reinterpret_cast<unsigned short*>(nrm + N + (N & 1));
	double* Js = reinterpret_cast<double*>(reinterpret_cast<char*>(sched) + ((size_t(N - 1) * bw * sizeof(unsigned short) + 15) / 16) * 16);
	for (int e = threadIdx.x; e < (N - 1) * bw; e += blockDim.x) {
		const int rr = e / bw, pi = e % bw;
		int a, b;
		if (pi == 0) { a = N - 1; b = rr; }
		else { a = (rr + pi) % (N - 1); b = (rr - pi + N - 1) % (N - 1); }
		sched[e] = (unsigned short)(a | (b << 8));
	}
	const int nrounds = (nblk == 2) ? 1 : nblk - 1;
	if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
	__syncthreads();
	int sweeps = 0, ground = 0;
	unsigned int last_rot = 1, last_big = 1;
	unsigned int my_rot = 0, my_big = 0;
	long long tk_load = 0, tk_inner = 0, tk_store = 0, tk_sync = 0, tk0 = 0;
	const bool timing = (info[0] == 0xC10C) && blockIdx.x == 0 && threadIdx.x == 0;

	// rotation parameters of a pair with squared norms aa, bb and cross product g (gg = g^2):
	//   c^2 = (1 + |d|/h)/2, s = sign(d) 2g / (2 h c) with d = bb - aa, h = sqrt(d^2 + 4 g^2): two rsqrt, no division
	auto rotation = [](const double aa, const double bb, const double g, const double gg, double& c, double& s, double& t) {
		const double d = bb - aa;
		const double rh = rsqrt(d * d + 4.0 * gg);
		const double c2 = 0.5 + 0.5 * fabs(d) * rh;
		const double rc = rsqrt(c2);
		c = c2 * rc;
		s = (d >= 0.0 ? g : -g) * rh * rc;
		t = s * rc;
	};

	for (; sweeps < max_sweeps; ) {
		for (int round = 0; round < nrounds; ++round, ++ground) {
			int pb, qb;
			bool loadp = true, storep = true, full = (round == 0);
			if (recursive) {
				// recursive bipartite tournament with point-to-point block flags: see jacobi_persistent_kernel
				int g = nblk, t = round;
				while (t >= (g >> 1)) { t -= (g >> 1); g >>= 1; }
				const int h = g >> 1, G = blockIdx.x / h, j = blockIdx.x % h;
				pb = G * g + j; qb = G * g + h + ((j + t) & (h - 1));
				loadp = (t == 0); storep = (t == h - 1); full = (g == 2);
				if (threadIdx.x == 0) {
					unsigned int spins = 0;
					volatile unsigned int* rd = ready;
					while ((rd[qb] < (unsigned)ground || (loadp && rd[pb] < (unsigned)ground)) && spins < (1u << 27)) ++spins;
					if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
					__threadfence();
				}
				__syncthreads();
			} else {
				const int pi = blockIdx.x, N1 = nblk - 1;
				if (nblk == 2) { pb = 0; qb = 1; }
				else if (pi == 0) { pb = N1; qb = round % N1; }
				else { pb = (round + pi) % N1; qb = (round - pi + N1) % N1; }
			}
			if (timing) tk0 = clock64();
			for (int r = (loadp ? 0 : bw) + warp; r < N; r += nwarps) {
				const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
				const double2* src = reinterpret_cast<const double2*>(GT + (size_t)grow * LD) + lane;
				double2* dst = reinterpret_cast<double2*>(S + (size_t)r * LDS) + lane;
				double2 v[EP2], w[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) { v[k] = __ldcg(src + 32 * k); w[k] = __ldcg(src + V2 + 32 * k); }
				double s0 = 0.0, s1 = 0.0;
#pragma unroll
				for (int k = 0; k < EP2; ++k) { dst[32 * k] = v[k]; dst[V2 + 32 * k] = w[k]; s0 += v[k].x * v[k].x; s1 += v[k].y * v[k].y; }
				double ss = s0 + s1;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
				if (lane == 0) nrm[r] = ss;
			}
			if (JACC) { for (int e = threadIdx.x; e < N * N; e += blockDim.x) Js[(e / N) * ldj + (e % N)] = (e / N == e % N) ? 1.0 : 0.0; }
			unsigned int visit_rot = 0;
			__syncthreads();
			if (timing) { const long long t1 = clock64(); tk_load += t1 - tk0; tk0 = t1; }

			if (full) {
				// all pairs among the 2 bw resident columns: both columns of a pair change owner every round
				for (int rr = 0; rr < N - 1; ++rr) {
					if (warp < bw) {
						const unsigned int ab_ = sched[rr * bw + warp];
						const int a = ab_ & 255, b = ab_ >> 8;
						double2* x = reinterpret_cast<double2*>(S + (size_t)a * LDS) + lane;
						double2* y = reinterpret_cast<double2*>(S + (size_t)b * LDS) + lane;
						double2 xr[EP2], yr[EP2];
#pragma unroll
						for (int k = 0; k < EP2; ++k) { xr[k] = x[32 * k]; yr[k] = y[32 * k]; }
						double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
						for (int k = 0; k < EP2; ++k) {
							if (k & 1) { g2 += xr[k].x * yr[k].x; g3 += xr[k].y * yr[k].y; }
							else { g0 += xr[k].x * yr[k].x; g1 += xr[k].y * yr[k].y; }
						}
						double g = (g0 + g1) + (g2 + g3);
						const double aa = nrm[a], bb = nrm[b];
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
						const double gg = g * g, ab = aa * bb;
						if (gg > tol2 * ab) {
							double2 xv[EP2], yv[EP2];
							double ja = 0.0, jb = 0.0;
							if (JACC) { if (lane < N) { ja = Js[a * ldj + lane]; jb = Js[b * ldj + lane]; } }
							else {
#pragma unroll
								for (int k = 0; k < EP2; ++k) { xv[k] = x[V2 + 32 * k]; yv[k] = y[V2 + 32 * k]; }
							}
							double c, s, t;
							rotation(aa, bb, g, gg, c, s, t);
#pragma unroll
							for (int k = 0; k < EP2; ++k) {
								x[32 * k] = make_double2(c * xr[k].x - s * yr[k].x, c * xr[k].y - s * yr[k].y);
								y[32 * k] = make_double2(s * xr[k].x + c * yr[k].x, s * xr[k].y + c * yr[k].y);
								if (!JACC) {
									x[V2 + 32 * k] = make_double2(c * xv[k].x - s * yv[k].x, c * xv[k].y - s * yv[k].y);
									y[V2 + 32 * k] = make_double2(s * xv[k].x + c * yv[k].x, s * xv[k].y + c * yv[k].y);
								}
							}
							if (JACC && lane < N) { Js[a * ldj + lane] = c * ja - s * jb; Js[b * ldj + lane] = s * ja + c * jb; }
							// cached norms: alpha' = alpha - t*gamma, beta' = beta + t*gamma; recomputed when the update cancels
							double na = aa - t * g, nb = bb + t * g;
							if (na < 0.25 * aa || nb < 0.25 * bb) {
								double sa = 0.0, sb = 0.0;
#pragma unroll
								for (int k = 0; k < EP2; ++k) {
									const double xn0 = c * xr[k].x - s * yr[k].x, yn0 = s * xr[k].x + c * yr[k].x;
									const double xn1 = c * xr[k].y - s * yr[k].y, yn1 = s * xr[k].y + c * yr[k].y;
									sa += xn0 * xn0 + xn1 * xn1; sb += yn0 * yn0 + yn1 * yn1;
								}
#pragma unroll
								for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
								na = sa; nb = sb;
							}
							visit_rot = 1;
							if (lane == 0) {
								nrm[a] = na; nrm[b] = nb;
								my_rot += 1;
								if (gg > big2 * ab) my_big += 1;
							}
						}
					}
					__syncthreads();
				}
			} else {
				// block p against block q: warp w keeps column w of p (x part, its norm, its row of Js; the v part too without JACC)
				// in registers for all bw rounds and meets column (w + rr) mod bw of q in round rr
				const bool has = warp < bw;
				const int a = warp;
				double2* x = reinterpret_cast<double2*>(S + (size_t)a * LDS) + lane;
				double2 xr[EP2], xv[EP2];
				double aa = 0.0, ja = 0.0;
				if (has) {
#pragma unroll
					for (int k = 0; k < EP2; ++k) { xr[k] = x[32 * k]; if (!JACC) xv[k] = x[V2 + 32 * k]; }
					aa = nrm[a];
					if (JACC && lane < N) ja = Js[a * ldj + lane];
				}
				for (int rr = 0; rr < bw; ++rr) {
					if (has) {
						const int b = bw + ((warp + rr) & (bw - 1));
						double2* y = reinterpret_cast<double2*>(S + (size_t)b * LDS) + lane;
						double2 yr[EP2];
#pragma unroll
						for (int k = 0; k < EP2; ++k) yr[k] = y[32 * k];
						double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
						for (int k = 0; k < EP2; ++k) {
							if (k & 1) { g2 += xr[k].x * yr[k].x; g3 += xr[k].y * yr[k].y; }
							else { g0 += xr[k].x * yr[k].x; g1 += xr[k].y * yr[k].y; }
						}
						double g = (g0 + g1) + (g2 + g3);
						const double bb = nrm[b];
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
						const double gg = g * g, ab = aa * bb;
						if (gg > tol2 * ab) {
							double2 yv[EP2];
							double jb = 0.0;
							if (JACC) { if (lane < N) jb = Js[b * ldj + lane]; }
							else {
#pragma unroll
								for (int k = 0; k < EP2; ++k) yv[k] = y[V2 + 32 * k];
							}
							double c, s, t;
							rotation(aa, bb, g, gg, c, s, t);
							double sa = 0.0, sb = 0.0;
#pragma unroll
							for (int k = 0; k < EP2; ++k) {
								const double2 xn = make_double2(c * xr[k].x - s * yr[k].x, c * xr[k].y - s * yr[k].y);
								const double2 yn = make_double2(s * xr[k].x + c * yr[k].x, s * xr[k].y + c * yr[k].y);
								xr[k] = xn; y[32 * k] = yn;
								if (!JACC) {
									const double2 vn = make_double2(c * xv[k].x - s * yv[k].x, c * xv[k].y - s * yv[k].y);
									y[V2 + 32 * k] = make_double2(s * xv[k].x + c * yv[k].x, s * xv[k].y + c * yv[k].y);
									xv[k] = vn;
								}
							}
							if (JACC && lane < N) { Js[b * ldj + lane] = s * ja + c * jb; ja = c * ja - s * jb; }
							double na = aa - t * g, nb = bb + t * g;
							if (na < 0.25 * aa || nb < 0.25 * bb) {
								// the cached-norm update cancelled: both norms from the rotated columns (y is re-read: this path is rare)
#pragma unroll
								for (int k = 0; k < EP2; ++k) {
									const double2 yn = y[32 * k];
									sa += xr[k].x * xr[k].x + xr[k].y * xr[k].y; sb += yn.x * yn.x + yn.y * yn.y;
								}
#pragma unroll
								for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
								na = sa; nb = sb;
							}
							aa = na;
							visit_rot = 1;
							if (lane == 0) {
								nrm[b] = nb;
								my_rot += 1;
								if (gg > big2 * ab) my_big += 1;
							}
						}
					}
					__syncthreads();
				}
				if (has) {
#pragma unroll
					for (int k = 0; k < EP2; ++k) { x[32 * k] = xr[k]; if (!JACC) x[V2 + 32 * k] = xv[k]; }
					if (lane == 0) nrm[a] = aa;
					if (JACC && lane < N) Js[a * ldj + lane] = ja;
				}
				__syncthreads();
			}
			if (JACC) {
				if (__syncthreads_or(int(visit_rot))) {
					if (N == 16) apply_rotations_to_v<2, 4>(S + 64 * EP2, LDS, Js, ldj, 8 * EP2, warp, nwarps, lane);
					else apply_rotations_to_v<1, 2>(S + 64 * EP2, LDS, Js, ldj, 8 * EP2, warp, nwarps, lane);
					__syncthreads();
				}
			}
			if (timing) { const long long t1 = clock64(); tk_inner += t1 - tk0; tk0 = t1; }
			if (lane == 0 && my_rot) { atomicAdd(&s_rot, my_rot); atomicAdd(&s_big, my_big); my_rot = 0; my_big = 0; }
			for (int r = (storep ? 0 : bw) + warp; r < N; r += nwarps) {
				const int grow = (r < bw ? pb * bw + r : qb * bw + (r - bw));
				double2* dst = reinterpret_cast<double2*>(GT + (size_t)grow * LD) + lane;
				const double2* src = reinterpret_cast<const double2*>(S + (size_t)r * LDS) + lane;
				double2 v[EP2], w[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) { v[k] = src[32 * k]; w[k] = src[V2 + 32 * k]; }
#pragma unroll
				for (int k = 0; k < EP2; ++k) { dst[32 * k] = v[k]; dst[V2 + 32 * k] = w[k]; }
			}
			if (timing) { const long long t1 = clock64(); tk_store += t1 - tk0; tk0 = t1; }
			if (recursive) {
				// the barrier orders every thread's row stores before thread 0's fence; the fence is cumulative, so one
				// thread publishing is enough (a fence in all 256 threads was 8 % of the kernel's stall samples)
				__syncthreads();
				if (threadIdx.x == 0) {
					__threadfence();
					volatile unsigned int* rd = ready;
					if (storep) rd[pb] = (unsigned)ground + 1u;
					rd[qb] = (unsigned)ground + 1u;
				}
			} else if (nblk > 2) { __threadfence(); grid.sync(); }
			else __syncthreads();
			if (timing) { const long long t1 = clock64(); tk_sync += t1 - tk0; tk0 = t1; }
		}
		++sweeps;
		unsigned int rot, big;
		if (nblk > 2) {
			if (threadIdx.x == 0) { atomicAdd(&counters[2 * (sweeps - 1)], s_rot); atomicAdd(&counters[2 * (sweeps - 1) + 1], s_big); s_rot = 0; s_big = 0; }
			__threadfence();
			grid.sync();
			rot = *((volatile unsigned int*)&counters[2 * (sweeps - 1)]);
			big = *((volatile unsigned int*)&counters[2 * (sweeps - 1) + 1]);
		} else {
			rot = s_rot; big = s_big;
			__syncthreads();
			if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
			__syncthreads();
		}
		last_rot = rot; last_big = big;
		if (big == 0) break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { info[1] = (unsigned)sweeps; info[2] = last_big; info[3] = last_rot; }
	if (timing) { info[4] = (unsigned)(tk_load >> 10); info[5] = (unsigned)(tk_inner >> 10); info[6] = (unsigned)(tk_store >> 10); info[7] = (unsigned)(tk_sync >> 10); }
}

// ---- split kernel: X workers and V workers -------------------------------------------------------------------------------
// With the rotations of a visit accumulated in Js, the accumulated-rotation halves of the rows are pure bookkeeping: nothing
// in the Jacobi iteration reads them.  Here they leave the critical path altogether.  The grid is doubled: CTA w < nblk/2 is
// the X worker of jacobi_fast_kernel reduced to the x halves (load, rounds, store, hand-over); CTA nblk/2 + w is its V worker,
// which follows the same tournament one step behind: it waits for the product Js_k its X worker logs after visit k (a ring of
// JS_DEPTH slots in global memory, with back-pressure), applies it to the v halves with DMMA and hands the travelling block's
// v half to the next V worker through its own ready flags.  Measured upper bound of the gain (V work deleted): -16 %.
constexpr int JS_DEPTH = 4;
// DSMEM helpers of the hand-over (same instructions as the QR panel kernel uses for its reductions)
__device__ __forceinline__ unsigned jsm_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned jmapa(unsigned addr, unsigned rank) {
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
__device__ __forceinline__ void jst_async_v2(unsigned raddr, const double2 v, unsigned rmbar) {
	asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
	             :: "r"(raddr), "l"(__double_as_longlong(v.x)), "l"(__double_as_longlong(v.y)), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void jst_async_f64(unsigned raddr, const double v, unsigned rmbar) {
	asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
	             :: "r"(raddr), "l"(__double_as_longlong(v)), "r"(rmbar) : "memory");
}
// bounded wait (never hang the GPU): returns false on time-out
__device__ __forceinline__ bool jmbar_wait(unsigned mbar, unsigned parity) {
	unsigned done = 0, spins = 0;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
	} while (!done && ++spins < (1u << 22));
	return done != 0;
}
// DS: the X workers form one thread-block cluster (the V workers a second one) and, inside a phase of the tournament, hand the
// travelling block's x half to the next worker by pushing it straight into that CTA's shared memory (st.async + mbarrier
// transaction count; a variant with one DSMEM bulk copy per row measured 2 % slower) instead of storing it to global memory, fencing, raising a flag and having the neighbour poll and load
// it: one SM-to-SM transfer replaces three L2 round trips on the critical path.  The travelling block is double buffered
// (buffer = step parity); a worker pushes into its neighbour's other buffer only after that neighbour has published (a
// progress counter in global memory, read off the critical path) that it is done with the step that used it.  Phase changes
// (both blocks re-dealt) and the V workers keep the global-memory path.
template <int EP2, bool DS>
__global__ void __launch_bounds__(256) jacobi_split_kernel(double* __restrict__ GT, const int nblk, const double tol2, const double big2,
                                                          unsigned int* __restrict__ counters, unsigned int* __restrict__ info,
                                                          const int max_sweeps, unsigned int* flags, double* jlog) {
	constexpr int BW = 8, N = 16;
	constexpr int LD = 128 * EP2;               // doubles per row in global memory: [x : 64 EP2 | v : 64 EP2]
	constexpr int HL = 64 * EP2;                // one half
	constexpr int HLS = HL + 4;                 // shared-memory row stride
	constexpr int ldj = N + 4;
	constexpr int SLOT = N * N + 8;
	constexpr int TB = (EP2 % 4 == 0) ? 4 : ((EP2 % 2 == 0) ? 2 : 1);
	extern __shared__ double S_split[];
	double* S = S_split;                        // [N][HLS] (+ [BW][HLS]: second buffer of the travelling block with DS)
	double* nrm = S + (size_t)(DS ? N + BW : N) * HLS;   // [N] (+ [BW] with DS)
	unsigned short* sched = reinterpret_cast<unsigned short*>(nrm + (DS ? N + BW : N));              // [(N-1)][BW]
	double* Js = reinterpret_cast<double*>(sched + ((N - 1) * BW + 8));              // [N][ldj]   ((N-1)*BW + 8 = 128 shorts)
	__shared__ unsigned int s_rot, s_big;
	__shared__ __align__(8) unsigned long long full_bar[2];
	__shared__ unsigned int prog[16];           // DS: steps completed by every X worker of the cluster (each worker writes its entry in all CTAs)
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nw = nblk >> 1;
	const bool vrole = int(blockIdx.x) >= nw;
	const int wid = vrole ? int(blockIdx.x) - nw : int(blockIdx.x);
	unsigned int* xready = flags;
	unsigned int* vready = flags + nblk;
	unsigned int* jready = flags + 2 * nblk;
	unsigned int* jdone = jready + nw;
	unsigned int* rdy = vrole ? vready : xready;
	const size_t hoff = vrole ? HL : 0;
	constexpr unsigned PUSH_BYTES = BW * HL * 8 + BW * 8;
	unsigned int use0 = 0, use1 = 0;            // completed DSMEM receptions per buffer (mbarrier phase parity)
	if (DS) {
		if (threadIdx.x < 16) prog[threadIdx.x] = 0;
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(jsm_u32(&full_bar[0])), "r"(1u) : "memory");
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(jsm_u32(&full_bar[1])), "r"(1u) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		cooperative_groups::this_cluster().sync();  // every CTA of the cluster is resident and its barriers are initialised
	}
	for (int e = threadIdx.x; e < (N - 1) * BW; e += blockDim.x) {
		const int rr = e / BW, pi = e % BW;
		int a, b;
		if (pi == 0) { a = N - 1; b = rr; }
		else { a = (rr + pi) % (N - 1); b = (rr - pi + N - 1) % (N - 1); }
		sched[e] = (unsigned short)(a | (b << 8));
	}
	const int nrounds = nblk - 1;
	if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
	__syncthreads();
	int sweeps = 0, ground = 0;
	unsigned int last_rot = 1, last_big = 1;
	unsigned int my_rot = 0, my_big = 0;
	long long tk_load = 0, tk_inner = 0, tk_store = 0, tk_sync = 0, tk0 = 0;
	const bool timing = (info[0] == 0xC10C) && blockIdx.x == 0 && threadIdx.x == 0;

	auto rotation = [](const double aa, const double bb, const double g, const double gg, double& c, double& s, double& t) {
		const double d = bb - aa;
		const double rh = rsqrt(d * d + 4.0 * gg);
		const double c2 = 0.5 + 0.5 * fabs(d) * rh;
		const double rc = rsqrt(c2);
		c = c2 * rc;
		s = (d >= 0.0 ? g : -g) * rh * rc;
		t = s * rc;
	};

	for (; sweeps < max_sweeps; ) {
		for (int round = 0; round < nrounds; ++round, ++ground) {
			// recursive bipartite tournament with point-to-point block flags: see jacobi_persistent_kernel
			int g = nblk, t = round;
			while (t >= (g >> 1)) { t -= (g >> 1); g >>= 1; }
			const int h = g >> 1, G = wid / h, j = wid % h;
			const int pb = G * g + j, qb = G * g + h + ((j + t) & (h - 1));
			const bool loadp = (t == 0), storep = (t == h - 1), full = (g == 2);
			double* slot = jlog + ((size_t)wid * JS_DEPTH + (ground % JS_DEPTH)) * SLOT;
			// travelling block: rows BW..N-1 of S, or (DS, X workers, odd steps) the second buffer behind S
			const bool dsx = DS && !vrole;
			const int buf = dsx ? (ground & 1) : 0;
			double* Qc = S + (size_t)(buf ? N : BW) * HLS;
			double* nq = nrm + (buf ? N : BW);
			const bool by_push = dsx && !loadp;       // this step's travelling block arrives through DSMEM
			const bool push = dsx && !storep;         // ... and leaves through DSMEM, into the next worker's other buffer
			const int recv = G * h + ((j + h - 1) & (h - 1));
			if (by_push) {
				const unsigned mb = jsm_u32(&full_bar[buf]);
				if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(PUSH_BYTES) : "memory");
				if (!jmbar_wait(mb, (buf ? use1 : use0) & 1u)) { if (threadIdx.x == 0) atomicOr(&counters[2 * max_sweeps], 0xDEADu); }
				if (buf) ++use1; else ++use0;
			}
			if (threadIdx.x == 0 && by_push) {
				// only the ring slot of the rotation-product log has to be free
				unsigned int spins = 0;
				volatile unsigned int* jd = jdone;
				while (jd[wid] + JS_DEPTH < (unsigned)ground + 1u && ++spins < (1u << 27)) {}
				if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
			}
			if (threadIdx.x == 0 && !by_push) {
				unsigned int spins = 0;
				volatile unsigned int* rd = rdy;
				volatile unsigned int* jr = jready;
				volatile unsigned int* jd = jdone;
				for (;;) {
					bool ok = rd[qb] >= (unsigned)ground && (!loadp || rd[pb] >= (unsigned)ground);
					if (vrole) ok = ok && jr[wid] >= (unsigned)ground + 1u;                      // the product of this visit is logged
					else ok = ok && jd[wid] + JS_DEPTH >= (unsigned)ground + 1u;                  // the ring slot is free again
					if (ok || ++spins >= (1u << 27)) break;
				}
				if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
				__threadfence();
			}
			__syncthreads();
			if (timing) tk0 = clock64();
			for (int r = (by_push ? N : (loadp ? 0 : BW)) + warp; r < N; r += 8) {
				const int grow = (r < BW ? pb * BW + r : qb * BW + (r - BW));
				const double2* src = reinterpret_cast<const double2*>(GT + (size_t)grow * LD + hoff) + lane;
				double2* dst = reinterpret_cast<double2*>(r < BW ? S + (size_t)r * HLS : Qc + (size_t)(r - BW) * HLS) + lane;
				double2 v[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) v[k] = __ldcg(src + 32 * k);
				double s0 = 0.0, s1 = 0.0;
#pragma unroll
				for (int k = 0; k < EP2; ++k) { dst[32 * k] = v[k]; s0 += v[k].x * v[k].x; s1 += v[k].y * v[k].y; }
				if (!vrole) {
					double ss = s0 + s1;
#pragma unroll
					for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
					if (lane == 0) { if (r < BW) nrm[r] = ss; else nq[r - BW] = ss; }
				}
			}
			if (vrole) {
				// V worker: fetch the logged product, apply it to the v halves
				Js[(threadIdx.x >> 4) * ldj + (threadIdx.x & 15)] = __ldcg(slot + threadIdx.x);
				const bool rotated = __ldcg(slot + N * N) != 0.0;
				__syncthreads();
				if (timing) { const long long t1 = clock64(); tk_load += t1 - tk0; tk0 = t1; }
				if (rotated) {
					apply_rotations_tiles<2, 4, TB>(S, HLS, Js, ldj, HL / 8, warp, 8, lane);
					__syncthreads();
				}
			} else {
				Js[(threadIdx.x >> 4) * ldj + (threadIdx.x & 15)] = ((threadIdx.x >> 4) == (threadIdx.x & 15)) ? 1.0 : 0.0;
				unsigned int visit_rot = 0;
				__syncthreads();
				if (timing) { const long long t1 = clock64(); tk_load += t1 - tk0; tk0 = t1; }
				if (full) {
					// all pairs among the 16 resident columns: both columns of a pair change owner every round
					for (int rr = 0; rr < N - 1; ++rr) {
						const unsigned int ab_ = sched[rr * BW + warp];
						const int a = ab_ & 255, b = ab_ >> 8;
						double2* x = reinterpret_cast<double2*>(a < BW ? S + (size_t)a * HLS : Qc + (size_t)(a - BW) * HLS) + lane;
						double2* y = reinterpret_cast<double2*>(b < BW ? S + (size_t)b * HLS : Qc + (size_t)(b - BW) * HLS) + lane;
						double* na_p = a < BW ? nrm + a : nq + (a - BW);
						double* nb_p = b < BW ? nrm + b : nq + (b - BW);
						double2 xr[EP2], yr[EP2];
#pragma unroll
						for (int k = 0; k < EP2; ++k) { xr[k] = x[32 * k]; yr[k] = y[32 * k]; }
						double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
						for (int k = 0; k < EP2; ++k) {
							if (k & 1) { g2 += xr[k].x * yr[k].x; g3 += xr[k].y * yr[k].y; }
							else { g0 += xr[k].x * yr[k].x; g1 += xr[k].y * yr[k].y; }
						}
						double gs = (g0 + g1) + (g2 + g3);
						const double aa = *na_p, bb = *nb_p;
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
						const double gg = gs * gs, ab = aa * bb;
						if (gg > tol2 * ab) {
							double ja = 0.0, jb = 0.0;
							if (lane < N) { ja = Js[a * ldj + lane]; jb = Js[b * ldj + lane]; }
							double c, s, tt;
							rotation(aa, bb, gs, gg, c, s, tt);
#pragma unroll
							for (int k = 0; k < EP2; ++k) {
								x[32 * k] = make_double2(c * xr[k].x - s * yr[k].x, c * xr[k].y - s * yr[k].y);
								y[32 * k] = make_double2(s * xr[k].x + c * yr[k].x, s * xr[k].y + c * yr[k].y);
							}
							if (lane < N) { Js[a * ldj + lane] = c * ja - s * jb; Js[b * ldj + lane] = s * ja + c * jb; }
							double na = aa - tt * gs, nb = bb + tt * gs;
							if (na < 0.25 * aa || nb < 0.25 * bb) {
								double sa = 0.0, sb = 0.0;
#pragma unroll
								for (int k = 0; k < EP2; ++k) {
									const double xn0 = c * xr[k].x - s * yr[k].x, yn0 = s * xr[k].x + c * yr[k].x;
									const double xn1 = c * xr[k].y - s * yr[k].y, yn1 = s * xr[k].y + c * yr[k].y;
									sa += xn0 * xn0 + xn1 * xn1; sb += yn0 * yn0 + yn1 * yn1;
								}
#pragma unroll
								for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
								na = sa; nb = sb;
							}
							visit_rot = 1;
							if (lane == 0) {
								*na_p = na; *nb_p = nb;
								my_rot += 1;
								if (gg > big2 * ab) my_big += 1;
							}
						}
						__syncthreads();
					}
				} else {
					// block p against block q: warp w keeps column w of p in registers and meets column (w + rr) mod 8 of q in round rr
					const int a = warp;
					double2* x = reinterpret_cast<double2*>(S + (size_t)a * HLS) + lane;
					double2 xr[EP2];
#pragma unroll
					for (int k = 0; k < EP2; ++k) xr[k] = x[32 * k];
					double aa = nrm[a], ja = (lane < N) ? Js[a * ldj + lane] : 0.0;
					for (int rr = 0; rr < BW; ++rr) {
						const int b = BW + ((warp + rr) & (BW - 1));
						double2* y = reinterpret_cast<double2*>(Qc + (size_t)(b - BW) * HLS) + lane;
						double2 yr[EP2];
#pragma unroll
						for (int k = 0; k < EP2; ++k) yr[k] = y[32 * k];
						double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
						for (int k = 0; k < EP2; ++k) {
							if (k & 1) { g2 += xr[k].x * yr[k].x; g3 += xr[k].y * yr[k].y; }
							else { g0 += xr[k].x * yr[k].x; g1 += xr[k].y * yr[k].y; }
						}
						double gs = (g0 + g1) + (g2 + g3);
						const double bb = nq[b - BW];
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
						const double gg = gs * gs, ab = aa * bb;
						if (gg > tol2 * ab) {
							const double jb = (lane < N) ? Js[b * ldj + lane] : 0.0;
							// bookkeeping first (every lane counts, lane 0's counters are the ones flushed): it then overlaps the latency of
							// the rotation parameters instead of trailing the stores in front of the barrier
							my_rot += 1;
							my_big += (gg > big2 * ab) ? 1u : 0u;
							visit_rot = 1;
							double c, s, tt;
							rotation(aa, bb, gs, gg, c, s, tt);
#pragma unroll
							for (int k = 0; k < EP2; ++k) {
								const double2 xn = make_double2(c * xr[k].x - s * yr[k].x, c * xr[k].y - s * yr[k].y);
								yr[k] = make_double2(s * xr[k].x + c * yr[k].x, s * xr[k].y + c * yr[k].y);
								y[32 * k] = yr[k];
								xr[k] = xn;
							}
							if (lane < N) { Js[b * ldj + lane] = s * ja + c * jb; ja = c * ja - s * jb; }
							double na = aa - tt * gs, nb = bb + tt * gs;
							if (na < 0.25 * aa || nb < 0.25 * bb) {
								double sa = 0.0, sb = 0.0;
#pragma unroll
								for (int k = 0; k < EP2; ++k) {
									sa += xr[k].x * xr[k].x + xr[k].y * xr[k].y; sb += yr[k].x * yr[k].x + yr[k].y * yr[k].y;
								}
#pragma unroll
								for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
								na = sa; nb = sb;
							}
							aa = na;
							if (lane == 0) nq[b - BW] = nb;
						}
						__syncthreads();
					}
#pragma unroll
					for (int k = 0; k < EP2; ++k) x[32 * k] = xr[k];
					if (lane == 0) nrm[a] = aa;
					if (lane < N) Js[a * ldj + lane] = ja;
				}
				// log the product of this visit for the V worker
				const int any = __syncthreads_or(int(visit_rot));
				slot[threadIdx.x] = Js[(threadIdx.x >> 4) * ldj + (threadIdx.x & 15)];
				if (threadIdx.x == 0) slot[N * N] = any ? 1.0 : 0.0;
				__syncthreads();
				if (threadIdx.x == 0) { __threadfence(); *((volatile unsigned int*)&jready[wid]) = (unsigned)ground + 1u; }
			}
			if (timing) { const long long t1 = clock64(); tk_inner += t1 - tk0; tk0 = t1; }
			if (lane == 0 && my_rot) { atomicAdd(&s_rot, my_rot); atomicAdd(&s_big, my_big); my_rot = 0; my_big = 0; }
			if (push) {
				// every thread pushes 16-byte pieces of the travelling block (and lane 0 of each warp a cached norm) into the
				// receiver's other buffer; completion is counted in bytes on the receiver's mbarrier.  The receiver meets the block
				// in step ground + 1 in its buffer (ground + 1) & 1, which it used in step ground - 1: prog[] says it has left it.
				if (threadIdx.x == 0) {
					unsigned int spins = 0;
					volatile unsigned int* xp = prog;
					while (xp[recv] < (unsigned)ground && ++spins < (1u << 27)) {}
					if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
				}
				__syncthreads();
				const int nb_ = buf ^ 1;
				const unsigned rmb = jmapa(jsm_u32(&full_bar[nb_]), (unsigned)recv);
				const double2* src = reinterpret_cast<const double2*>(Qc + (size_t)warp * HLS) + lane;
				const unsigned rdst = jmapa(jsm_u32(S + (size_t)((nb_ ? N : BW) + warp) * HLS) + 16u * lane, (unsigned)recv);
				double2 v[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) v[k] = src[32 * k];
#pragma unroll
				for (int k = 0; k < EP2; ++k) jst_async_v2(rdst + 512u * k, v[k], rmb);
				if (lane == 0) jst_async_f64(jmapa(jsm_u32(nrm + (nb_ ? N : BW) + warp), (unsigned)recv), nq[warp], rmb);
			} else {
				for (int r = (storep ? 0 : BW) + warp; r < N; r += 8) {
					const int grow = (r < BW ? pb * BW + r : qb * BW + (r - BW));
					double2* dst = reinterpret_cast<double2*>(GT + (size_t)grow * LD + hoff) + lane;
					const double2* src = reinterpret_cast<const double2*>(r < BW ? S + (size_t)r * HLS : Qc + (size_t)(r - BW) * HLS) + lane;
					double2 v[EP2];
#pragma unroll
					for (int k = 0; k < EP2; ++k) v[k] = src[32 * k];
#pragma unroll
					for (int k = 0; k < EP2; ++k) dst[32 * k] = v[k];
				}
			}
			if (timing) { const long long t1 = clock64(); tk_store += t1 - tk0; tk0 = t1; }
			__syncthreads();
			if (threadIdx.x == 0) {
				if (!push) {
					__threadfence();
					volatile unsigned int* rd = rdy;
					if (storep) rd[pb] = (unsigned)ground + 1u;
					rd[qb] = (unsigned)ground + 1u;
				}
				if (vrole) *((volatile unsigned int*)&jdone[wid]) = (unsigned)ground + 1u;
			}
			// this step's buffer has been read out (the pushes took their data from registers): publish the step count in every
			// X worker's shared memory (plain DSMEM stores)
			if (dsx && threadIdx.x < nw)
				asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(jmapa(jsm_u32(&prog[wid]), (unsigned)threadIdx.x)), "r"((unsigned)ground + 1u) : "memory");
			if (timing) { const long long t1 = clock64(); tk_sync += t1 - tk0; tk0 = t1; }
		}
		++sweeps;
		if (threadIdx.x == 0 && !vrole) { atomicAdd(&counters[2 * (sweeps - 1)], s_rot); atomicAdd(&counters[2 * (sweeps - 1) + 1], s_big); s_rot = 0; s_big = 0; }
		__threadfence();
		grid.sync();
		const unsigned int rot = *((volatile unsigned int*)&counters[2 * (sweeps - 1)]);
		const unsigned int big = *((volatile unsigned int*)&counters[2 * (sweeps - 1) + 1]);
		last_rot = rot; last_big = big;
		if (big == 0) break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { info[1] = (unsigned)sweeps; info[2] = last_big; info[3] = last_rot; }
	if (timing) { info[4] = (unsigned)(tk_load >> 10); info[5] = (unsigned)(tk_inner >> 10); info[6] = (unsigned)(tk_store >> 10); info[7] = (unsigned)(tk_sync >> 10); }
}

// ---- Gram-space block Jacobi kernel (FP64, 8-column blocks, both row parts 64 * EP2 doubles long) -------------------------
// jacobi_fast_kernel pays a ~1 100-cycle dependent chain (load, dot, 5-stage butterfly, two rsqrt, rotate, barrier) for every
// one of the 8 rounds of a block visit, on full-length vectors.  Here a visit (block p against block q, 16 columns) is
//   1. G = X^T X of the 16 resident columns (x part), once, on the tensor pipe (partials per warp, fixed-order sum);
//   2. the same 8 rounds of 8 disjoint rotations, but carried out on the 16 x 16 Gram matrix: the cross product and the
//      norms of a pair are matrix entries (no dot, no butterfly), G <- R^T G R and Js <- R^T Js are one entry per thread;
//   3. rows <- Js * rows for the x AND the accumulated-rotation part, once, with DMMA.
// In exact arithmetic this is the same sequence of rotations as the column-space kernel.  In floating point the entries of
// G used by the later rounds of a visit carry errors of eps * |G|_max instead of eps * |a||b|; every visit starts from a
// fresh Gram matrix of the actual columns, so the convergence criterion is evaluated on true cosines for the pairs that
// matter and errors never accumulate across visits.  Sweep counts and final accuracy are unchanged (numpy emulation of
// both variants on Gaussian, graded 1e-8 / 1e-14 and rank-deficient inputs; GPU tests), absolute accuracy eps * sigma_max as
// for LAPACK's dgesdd, which is the reference.
template <int EP2>
__global__ void __launch_bounds__(256) jacobi_gram_kernel(double* __restrict__ GT, const int nblk, const double tol2, const double big2,
                                                         unsigned int* __restrict__ counters, unsigned int* __restrict__ info,
                                                         const int max_sweeps, unsigned int* ready, const int recursive) {
	constexpr int BW = 8, N = 16;
	constexpr int LD = 128 * EP2;               // doubles per row in global memory: [x : 64 EP2 | v : 64 EP2]
	constexpr int LDS = LD + 4;                 // shared-memory row stride (DMMA operand loads of 4 rows x 8 elements: distinct banks)
	constexpr int V2 = 32 * EP2;                // offset of the v part in double2 units
	constexpr int LG = 17;
	extern __shared__ double S_gram[];
	double* S = S_gram;                         // [N][LDS]
	double* Gp = S + (size_t)N * LDS;           // [8][256] per-warp partial Gram matrices
	double* Gm = Gp + 8 * 256;                  // [2][N][LG]  Gram matrix, double buffered
	double* Jm = Gm + 2 * N * LG;               // [2][N][LG]  product of the rotations of the visit (row operations)
	unsigned short* sched = reinterpret_cast<unsigned short*>(Jm + 2 * N * LG);   // [(N-1)][8] : a | b << 8   (all-pairs visits)
	unsigned char* role = reinterpret_cast<unsigned char*>(sched + (N - 1) * BW + 8); // [(N-1)][16] : pair | is_b << 7
	__shared__ double tabc[N], tabs[N];
	__shared__ int tab_any_s;
	int* tab_any = &tab_any_s;
	__shared__ unsigned int s_rot, s_big;
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int e = threadIdx.x; e < (N - 1) * BW; e += blockDim.x) {
		const int rr = e / BW, pi = e % BW;
		int a, b;
		if (pi == 0) { a = N - 1; b = rr; }
		else { a = (rr + pi) % (N - 1); b = (rr - pi + N - 1) % (N - 1); }
		sched[e] = (unsigned short)(a | (b << 8));
		role[rr * N + a] = (unsigned char)pi;
		role[rr * N + b] = (unsigned char)(pi | 128);
	}
	const int nrounds = (nblk == 2) ? 1 : nblk - 1;
	if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
	__syncthreads();
	int sweeps = 0, ground = 0;
	unsigned int last_rot = 1, last_big = 1;
	unsigned int my_rot = 0, my_big = 0;
	long long tk_load = 0, tk_inner = 0, tk_store = 0, tk_sync = 0, tk0 = 0, tk_gram = 0, tk_solve = 0, tk_apply = 0, tq = 0;
	const bool timing = (info[0] == 0xC10C) && blockIdx.x == 0 && threadIdx.x == 0;
	const int fg = lane >> 2, ft = lane & 3;    // DMMA fragment coordinates

	for (; sweeps < max_sweeps; ) {
		for (int round = 0; round < nrounds; ++round, ++ground) {
			int pb, qb;
			bool loadp = true, storep = true, full = (round == 0);
			if (recursive) {
				// recursive bipartite tournament with point-to-point block flags: see jacobi_persistent_kernel
				int g = nblk, t = round;
				while (t >= (g >> 1)) { t -= (g >> 1); g >>= 1; }
				const int h = g >> 1, G = blockIdx.x / h, j = blockIdx.x % h;
				pb = G * g + j; qb = G * g + h + ((j + t) & (h - 1));
				loadp = (t == 0); storep = (t == h - 1); full = (g == 2);
				if (threadIdx.x == 0) {
					unsigned int spins = 0;
					volatile unsigned int* rd = ready;
					while ((rd[qb] < (unsigned)ground || (loadp && rd[pb] < (unsigned)ground)) && spins < (1u << 27)) ++spins;
					if (spins >= (1u << 27)) atomicOr(&counters[2 * max_sweeps], 0xDEADu);
					__threadfence();
				}
				__syncthreads();
			} else {
				const int pi = blockIdx.x, N1 = nblk - 1;
				if (nblk == 2) { pb = 0; qb = 1; }
				else if (pi == 0) { pb = N1; qb = round % N1; }
				else { pb = (round + pi) % N1; qb = (round - pi + N1) % N1; }
			}
			if (timing) tk0 = clock64();
			for (int r = (loadp ? 0 : BW) + warp; r < N; r += 8) {
				const int grow = (r < BW ? pb * BW + r : qb * BW + (r - BW));
				const double2* src = reinterpret_cast<const double2*>(GT + (size_t)grow * LD) + lane;
				double2* dst = reinterpret_cast<double2*>(S + (size_t)r * LDS) + lane;
				double2 v[EP2], w[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) { v[k] = __ldcg(src + 32 * k); w[k] = __ldcg(src + V2 + 32 * k); }
#pragma unroll
				for (int k = 0; k < EP2; ++k) { dst[32 * k] = v[k]; dst[V2 + 32 * k] = w[k]; }
			}
			__syncthreads();
			if (timing) { const long long t1 = clock64(); tk_load += t1 - tk0; tk0 = t1; }

			if (timing) tq = clock64();
			// 1. Gram matrix of the x parts: warp w takes the k-steps w, w + 8, ... (two interleaved accumulator sets: a DMMA
			//    depends on the previous one of its chain); tile (1,0) is the mirror of (0,1)
			{
				double a00[2] = {0.0, 0.0}, a01[2] = {0.0, 0.0}, a11[2] = {0.0, 0.0};
				double b00[2] = {0.0, 0.0}, b01[2] = {0.0, 0.0}, b11[2] = {0.0, 0.0};
				const double* s0 = S + (size_t)fg * LDS + ft;
				const double* s1 = S + (size_t)(8 + fg) * LDS + ft;
#pragma unroll
				for (int i = 0; i < EP2; ++i) {
					const int ks = warp + 16 * i;
					const double f0 = s0[4 * ks], f1 = s1[4 * ks], h0 = s0[4 * (ks + 8)], h1 = s1[4 * (ks + 8)];
					dmma_884(a00[0], a00[1], f0, f0);
					dmma_884(b00[0], b00[1], h0, h0);
					dmma_884(a01[0], a01[1], f0, f1);
					dmma_884(b01[0], b01[1], h0, h1);
					dmma_884(a11[0], a11[1], f1, f1);
					dmma_884(b11[0], b11[1], h1, h1);
				}
				double* gp = Gp + warp * 256;
				*reinterpret_cast<double2*>(gp + fg * 16 + 2 * ft) = make_double2(a00[0] + b00[0], a00[1] + b00[1]);
				*reinterpret_cast<double2*>(gp + fg * 16 + 8 + 2 * ft) = make_double2(a01[0] + b01[0], a01[1] + b01[1]);
				*reinterpret_cast<double2*>(gp + (8 + fg) * 16 + 8 + 2 * ft) = make_double2(a11[0] + b11[0], a11[1] + b11[1]);
			}
			__syncthreads();
			{
				const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
				const int e = (r >= 8 && c < 8) ? c * 16 + r : r * 16 + c;
				double sum = 0.0;
#pragma unroll
				for (int w = 0; w < 8; ++w) sum += Gp[w * 256 + e];
				Gm[r * LG + c] = sum;
				Jm[r * LG + c] = (r == c) ? 1.0 : 0.0;
			}
			__syncthreads();

			if (timing) { const long long t1 = clock64(); tk_gram += t1 - tq; tq = t1; }
			// 2. the rounds of the visit on G.  Lanes 0..7 of warp 0 derive the 8 rotations of a round and publish, per column
			//    index r, the row operation  row_r' = tabc[r] row_r + tabs[r] row_partner(r)  (row_a' = c row_a - s row_b,
			//    row_b' = s row_a + c row_b); every thread then updates one entry of G (R^T G R) and of Js (R^T Js).  The entries a
			//    thread needs depend on the schedule only, so they are fetched while warp 0 works on the parameters.
			int cur = 0;
			bool any_rot = false;
			const int nr = full ? (N - 1) : BW;
			const int r0 = threadIdx.x >> 4, c0 = threadIdx.x & 15;
			for (int rr = 0; rr < nr; ++rr) {
				const double* Gc = Gm + cur * N * LG;
				const double* Jc = Jm + cur * N * LG;
				int rp, cp;                                  // partners of this thread's row and column index in this round
				if (full) {
					const unsigned int rl = role[rr * N + r0], cl = role[rr * N + c0];
					const unsigned int pr_ = sched[rr * BW + (rl & 127)], pc_ = sched[rr * BW + (cl & 127)];
					rp = (rl >> 7) ? (pr_ & 255) : (pr_ >> 8);
					cp = (cl >> 7) ? (pc_ & 255) : (pc_ >> 8);
				} else {
					rp = (r0 < BW) ? BW + ((r0 + rr) & 7) : ((r0 - BW - rr) & 7);
					cp = (c0 < BW) ? BW + ((c0 + rr) & 7) : ((c0 - BW - rr) & 7);
				}
				const double g00 = Gc[r0 * LG + c0], g01 = Gc[r0 * LG + cp], g10 = Gc[rp * LG + c0], g11 = Gc[rp * LG + cp];
				const double j0 = Jc[r0 * LG + c0], j1 = Jc[rp * LG + c0];
				if (warp == 0) {
					int a, b;
					if (full) { const unsigned int ab_ = sched[rr * BW + (lane & 7)]; a = ab_ & 255; b = ab_ >> 8; }
					else { a = lane & 7; b = BW + (((lane & 7) + rr) & 7); }
					const double aa = Gc[a * LG + a], bb = Gc[b * LG + b], g = Gc[a * LG + b];
					const double gg = g * g, ab = aa * bb;
					const bool rot = gg > tol2 * ab;
					const bool any = __any_sync(0xffffffffu, rot);
					if (any && lane < 8) {
						double c = 1.0, s = 0.0;
						if (rot) {
							// c^2 = (1 + |d|/h)/2, s = sign(d) 2g / (2 h c) with d = bb - aa, h = sqrt(d^2 + 4 g^2): two rsqrt, no division
							const double d = bb - aa;
							const double rh = rsqrt(d * d + 4.0 * gg);
							const double c2 = 0.5 + 0.5 * fabs(d) * rh;
							const double rc = rsqrt(c2);
							c = c2 * rc;
							s = (d >= 0.0 ? g : -g) * rh * rc;
							my_rot += 1;
							if (gg > big2 * ab) my_big += 1;
						}
						tabc[a] = c; tabs[a] = -s; tabc[b] = c; tabs[b] = s;
					}
					if (lane == 0) *tab_any = any ? 1 : 0;
				}
				__syncthreads();
				if (*tab_any == 0) { __syncthreads(); continue; }     // (second barrier: warp 0 may not overwrite the flag before all have read it)
				any_rot = true;
				const double rc_ = tabc[r0], rs_ = tabs[r0], tc = tabc[c0], ts = tabs[c0];
				double* Gn = Gm + (cur ^ 1) * N * LG;
				double* Jn = Jm + (cur ^ 1) * N * LG;
				Gn[r0 * LG + c0] = rc_ * (tc * g00 + ts * g01) + rs_ * (tc * g10 + ts * g11);
				Jn[r0 * LG + c0] = rc_ * j0 + rs_ * j1;
				cur ^= 1;
				__syncthreads();
			}
			if (timing) { const long long t1 = clock64(); tk_solve += t1 - tq; tq = t1; }
			// 3. rows <- Js * rows (both parts)
			if (any_rot) {
				apply_rotations_tiles<2, 4, (EP2 % 2 == 0) ? 4 : 2>(S, LDS, Jm + cur * N * LG, LG, LD / 8, warp, 8, lane);
				__syncthreads();
			}
			if (timing) { const long long t1 = clock64(); tk_apply += t1 - tq; tk_inner += t1 - tk0; tk0 = t1; }
			if (warp == 0 && my_rot) { atomicAdd(&s_rot, my_rot); atomicAdd(&s_big, my_big); my_rot = 0; my_big = 0; }
			for (int r = (storep ? 0 : BW) + warp; r < N; r += 8) {
				const int grow = (r < BW ? pb * BW + r : qb * BW + (r - BW));
				double2* dst = reinterpret_cast<double2*>(GT + (size_t)grow * LD) + lane;
				const double2* src = reinterpret_cast<const double2*>(S + (size_t)r * LDS) + lane;
				double2 v[EP2], w[EP2];
#pragma unroll
				for (int k = 0; k < EP2; ++k) { v[k] = src[32 * k]; w[k] = src[V2 + 32 * k]; }
#pragma unroll
				for (int k = 0; k < EP2; ++k) { dst[32 * k] = v[k]; dst[V2 + 32 * k] = w[k]; }
			}
			if (timing) { const long long t1 = clock64(); tk_store += t1 - tk0; tk0 = t1; }
			if (recursive) {
				__syncthreads();
				if (threadIdx.x == 0) {
					__threadfence();
					volatile unsigned int* rd = ready;
					if (storep) rd[pb] = (unsigned)ground + 1u;
					rd[qb] = (unsigned)ground + 1u;
				}
			} else if (nblk > 2) { __threadfence(); grid.sync(); }
			else __syncthreads();
			if (timing) { const long long t1 = clock64(); tk_sync += t1 - tk0; tk0 = t1; }
		}
		++sweeps;
		unsigned int rot, big;
		if (nblk > 2) {
			if (threadIdx.x == 0) { atomicAdd(&counters[2 * (sweeps - 1)], s_rot); atomicAdd(&counters[2 * (sweeps - 1) + 1], s_big); s_rot = 0; s_big = 0; }
			__threadfence();
			grid.sync();
			rot = *((volatile unsigned int*)&counters[2 * (sweeps - 1)]);
			big = *((volatile unsigned int*)&counters[2 * (sweeps - 1) + 1]);
		} else {
			rot = s_rot; big = s_big;
			__syncthreads();
			if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
			__syncthreads();
		}
		last_rot = rot; last_big = big;
		if (big == 0) break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { info[1] = (unsigned)sweeps; info[2] = last_big; info[3] = last_rot; }
	if (timing) {
		info[4] = (unsigned)(tk_load >> 10); info[5] = (unsigned)(tk_inner >> 10); info[6] = (unsigned)(tk_store >> 10); info[7] = (unsigned)(tk_sync >> 10);
		counters[2 * max_sweeps + 1] = (unsigned)(tk_gram >> 10); counters[2 * max_sweeps + 2] = (unsigned)(tk_solve >> 10); counters[2 * max_sweeps + 3] = (unsigned)(tk_apply >> 10);
	}
}

// singular values = column norms; rank them (descending, ties by index) -> Ssorted, perm.   single CTA
__global__ void svd_sort_kernel(const double* __restrict__ GT, const int ldg, const int mdot, const int nw,
                                double* __restrict__ Ssorted, int* __restrict__ perm, const double* __restrict__ unscale) {
	extern __shared__ double nrm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	for (int j = warp; j < nw; j += nwarps) {
		const double* x = GT + (size_t)j * ldg;
		double s = 0.0;
		for (int i = lane; i < mdot; i += 32) s += x[i] * x[i];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (lane == 0) nrm[j] = sqrt(s);
	}
	__syncthreads();
	for (int j = threadIdx.x; j < nw; j += blockDim.x) {
		const double v = nrm[j];
		int rank = 0;
		for (int i = 0; i < nw; ++i) { const double u = nrm[i]; rank += (u > v || (u == v && i < j)) ? 1 : 0; }
		Ssorted[rank] = v * (*unscale);
		perm[rank] = j;
	}
}

// blockIdx.x = output index r (< k).  X part -> outX[i*sxi + r*sxr] (i < mdot), V part -> outV[c*svc + r*svr] (c < nw).
// The X part is normalised by sigma_r unless scale_x (then it keeps Sigma); the V part is multiplied by sigma_r if scale_v.
__global__ void svd_extract_kernel(const double* __restrict__ GT, const int ldg, const int mdot, const int voff, const int nw,
                                   const double* __restrict__ Ssorted, const int* __restrict__ perm,
                                   double* __restrict__ outX, const long long sxi, const long long sxr, const int scale_x,
                                   double* __restrict__ outV, const long long svc, const long long svr, const int scale_v,
                                   double* __restrict__ dS, const double soft, const double* __restrict__ scale2,
                                   const int* __restrict__ mapX, const int* __restrict__ mapV) {
	const int r = blockIdx.x;
	const double sigma = Ssorted[r];                          // unscaled singular value; the X part of GT holds x * scale2[0]
	const double sigma_eff = fmax(0.0, sigma - soft);        // soft thresholding (tensorNetwork.cpp:766)
	const double* g = GT + (size_t)perm[r] * ldg;
	const double inv = sigma > 0.0 ? 1.0 / (sigma * scale2[0]) : 0.0;        // 1 / (scaled sigma)
	const double fx = scale_x ? (soft == 0.0 ? scale2[1] : sigma_eff * inv) : inv;
	const double fv = scale_v ? sigma_eff : 1.0;
	// mapX / mapV (optional): the part that holds right vectors of the column-sorted working matrix goes back to the original order
	for (int i = threadIdx.x; i < mdot; i += blockDim.x) outX[(long long)(mapX ? mapX[i] : i) * sxi + (long long)r * sxr] = g[i] * fx;
	for (int c = threadIdx.x; c < nw; c += blockDim.x) outV[(long long)(mapV ? mapV[c] : c) * svc + (long long)r * svr] = g[voff + c] * fv;
	if (dS && threadIdx.x == 0) dS[r] = sigma_eff;
}

// Column ordering of the QR pre-conditioning (Drmac & Veselic use a pivoted QR; sorting the columns by norm first captures
// most of its effect): the working matrix Mw(i, j) = A[i * rs + j * cs] (m x n) gets its columns ranked by descending
// Euclidean norm (ties by index), perm[rank] = column.  Single CTA; squares are taken of the power-of-two scaled entries.
__global__ void __launch_bounds__(1024) svd_colrank_kernel(const double* __restrict__ A, const int m, const int n, const long long rs,
                                                            const long long cs, const double* __restrict__ scale, int* __restrict__ perm) {
	extern __shared__ double cn[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const double sc = scale[0];
	if (cs == 1) {
		// row-major working matrix: a thread sums every fourth row of one column (neighbouring threads read neighbouring
		// columns: coalesced), the four partial sums of a column are added in a fixed order
		double* part = cn + n;                               // [4][n]
		for (int e = threadIdx.x; e < 4 * n; e += blockDim.x) {
			const int p4 = e / n, j = e % n;
			double a0 = 0.0, a1 = 0.0;
			int i = p4;
			for (; i + 4 < m; i += 8) {
				const double v0 = A[(long long)i * rs + j] * sc, v1 = A[(long long)(i + 4) * rs + j] * sc;
				a0 += v0 * v0; a1 += v1 * v1;
			}
			if (i < m) { const double v0 = A[(long long)i * rs + j] * sc; a0 += v0 * v0; }
			part[e] = a0 + a1;
		}
		__syncthreads();
		for (int j = threadIdx.x; j < n; j += blockDim.x) cn[j] = (part[j] + part[n + j]) + (part[2 * n + j] + part[3 * n + j]);
	} else {
		for (int j = warp; j < n; j += nwarps) {
			double acc = 0.0;
			for (int i = lane; i < m; i += 32) { const double v = A[(long long)i * rs + (long long)j * cs] * sc; acc += v * v; }
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
			if (lane == 0) cn[j] = acc;
		}
	}
	__syncthreads();
	for (int j = threadIdx.x; j < n; j += blockDim.x) {
		const double v = cn[j];
		int rank = 0;
		for (int i = 0; i < n; ++i) { const double u = cn[i]; rank += (u > v || (u == v && i < j)) ? 1 : 0; }
		perm[rank] = j;
	}
}
// out (m x n, packed) = Mw[:, perm]
__global__ void svd_gather_cols_kernel(double* __restrict__ out, const double* __restrict__ A, const int m, const int n, const long long rs,
                                       const long long cs, const int* __restrict__ perm) {
	const size_t total = (size_t)m * n;
	for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const int i = int(e / n), j = int(e % n);
		out[e] = A[(long long)i * rs + (long long)perm[j] * cs];
	}
}

__global__ void add_diag_kernel(double* __restrict__ M, const size_t n, const double v) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) M[i * n + i] += v;
}

// ---- launch plan of the persistent Jacobi kernel --------------------------------------------------------------------
struct JacobiPlan {
	int EH = 4, bw = 1, threads = 64;
	size_t nblk = 2, npad = 2, smem = 0;
	bool persistent = false;
	int ep2 = 0;             // > 0: jacobi_fast_kernel<ep2> applies (both row parts 64 * ep2 doubles)
};

static JacobiPlan plan_jacobi(size_t ld, size_t nw, size_t voff, size_t elem, size_t smem_cap) {
	Context& c = ctx();
	JacobiPlan p;
	const int eh = int(std::max(voff, ld - voff) / 32);
	p.EH = eh <= 4 ? 4 : (eh <= 8 ? 8 : 16);
	// The kernel is latency / FP64-issue bound: an inner round costs a dependent chain (reduce -> 2 rsqrt -> rotate ->
	// barrier) whose length barely depends on the vector length, and the FP64 pipe is shared by the warps of an SM.
	// More, smaller CTAs only pay off from 256 columns on (measured: scratch/svd_time.py).
	int max_bw = 8;      // 16-column blocks (4 warps per scheduler) were slower at every size (profiles/bw_svd.py)
	if (c.svd_max_bw > 0) max_bw = std::min(max_bw, c.svd_max_bw);
	const int maxt = p.EH == 16 ? 256 : 512;
	max_bw = std::min(max_bw, maxt / 32);
	int bw = max_bw;
	// rows [2b][ld + 4] | norms [2b] | schedule (2b - 1) b u16 | rotation product [2b][2b + 4]
	auto need = [&](int b) { return size_t(2 * b) * (ld + 4 + 2) * elem + size_t(4 * b) * b + size_t(2 * b) * (2 * b + 4) * elem + 96; };
	while (bw > 1 && (need(bw) > smem_cap || size_t(bw) >= nw)) bw >>= 1;
	p.bw = bw;
	p.nblk = (nw + bw - 1) / bw;
	if (p.nblk < 2) p.nblk = 2;
	if (p.nblk & 1) ++p.nblk;
	p.npad = p.nblk * bw;
	p.threads = std::max(64, 32 * bw);
	p.smem = need(bw);
	p.persistent = c.svd_persistent && eh <= 16 && (p.nblk / 2) <= size_t(c.num_sms) && p.smem <= smem_cap;
	return p;
}

template <typename T, int EH, int MAXT>
static void launch_persistent(T* gt, int ld, int epl_x, int epl_v, const JacobiPlan& p, T tol2, T big2, unsigned int* d_cnt,
                              unsigned int* d_info, int max_sweeps, size_t smem_cap) {
	static bool attr = false;
	if (!attr) {
		XB_CUDA(cudaFuncSetAttribute(jacobi_persistent_kernel<T, EH, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
		attr = true;
	}
	XB_REQUIRE(p.threads <= MAXT, "internal: Jacobi launch exceeds its launch bound");
	int bw = p.bw, nblk = int(p.nblk);
	unsigned int* ready = d_cnt + 2 * max_sweeps + 12;
	int recursive = (ctx().svd_recursive && nblk > 2 && (nblk & (nblk - 1)) == 0) ? 1 : 0;
	void* args[] = {&gt, &ld, &epl_x, &epl_v, &bw, &nblk, &tol2, &big2, &d_cnt, &d_info, &max_sweeps, &ready, &recursive};
	XB_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_persistent_kernel<T, EH, MAXT>, dim3(unsigned(p.nblk / 2)), dim3(p.threads), args, p.smem, ctx().stream));
	ctx().launches++;
}

template <int EP2, int MAXT, bool JACC>
static void launch_fast(double* gt, const JacobiPlan& p, double tol2, double big2, unsigned int* d_cnt, unsigned int* d_info, int max_sweeps,
                        size_t smem_cap) {
	static bool attr = false;
	if (!attr) {
		XB_CUDA(cudaFuncSetAttribute(jacobi_fast_kernel<EP2, MAXT, JACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
		attr = true;
	}
	XB_REQUIRE(p.threads <= MAXT, "internal: Jacobi launch exceeds its launch bound");
	int bw = p.bw, nblk = int(p.nblk);
	unsigned int* ready = d_cnt + 2 * max_sweeps + 12;
	int recursive = (ctx().svd_recursive && nblk > 2 && (nblk & (nblk - 1)) == 0) ? 1 : 0;
	void* args[] = {&gt, &bw, &nblk, &tol2, &big2, &d_cnt, &d_info, &max_sweeps, &ready, &recursive};
	XB_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_fast_kernel<EP2, MAXT, JACC>, dim3(unsigned(p.nblk / 2)), dim3(p.threads), args, p.smem, ctx().stream));
	ctx().launches++;
}

template <int EP2>
static void launch_gram(double* gt, const JacobiPlan& p, double tol2, double big2, unsigned int* d_cnt, unsigned int* d_info, int max_sweeps,
                        size_t smem_cap) {
	static bool attr = false;
	if (!attr) {
		XB_CUDA(cudaFuncSetAttribute(jacobi_gram_kernel<EP2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
		attr = true;
	}
	int nblk = int(p.nblk);
	unsigned int* ready = d_cnt + 2 * max_sweeps + 12;
	int recursive = (ctx().svd_recursive && nblk > 2 && (nblk & (nblk - 1)) == 0) ? 1 : 0;
	const size_t smem = (size_t(16) * (128 * EP2 + 4) + 8 * 256 + 4 * 16 * 17) * sizeof(double) + 15 * 8 * 2 + 16 + 15 * 16 + 64;
	XB_REQUIRE(smem <= smem_cap, "internal: Gram Jacobi kernel exceeds shared memory");
	void* args[] = {&gt, &nblk, &tol2, &big2, &d_cnt, &d_info, &max_sweeps, &ready, &recursive};
	XB_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_gram_kernel<EP2>, dim3(unsigned(p.nblk / 2)), dim3(256), args, smem, ctx().stream));
	ctx().launches++;
}

template <int EP2, bool DS>
static bool launch_split_impl(double* gt, const JacobiPlan& p, double tol2, double big2, unsigned int* d_cnt, unsigned int* d_info, int max_sweeps,
                              size_t smem_cap) {
	int nblk = int(p.nblk);
	const int nw = nblk / 2;
	const size_t rows = DS ? 24 : 16;
	const size_t smem = (rows * (64 * EP2 + 4) + rows + 16 * 20) * sizeof(double) + 256 + 64;
	if (smem > smem_cap) { XB_REQUIRE(DS, "internal: split Jacobi kernel exceeds shared memory"); return false; }
	static bool attr = false;
	if (!attr) {
		XB_CUDA(cudaFuncSetAttribute(jacobi_split_kernel<EP2, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
		if (DS) XB_CUDA(cudaFuncSetAttribute(jacobi_split_kernel<EP2, DS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
		attr = true;
	}
	unsigned int* flags = d_cnt + 2 * max_sweeps + 12;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(unsigned(nblk)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = ctx().stream;
	cudaLaunchAttribute at[2];
	at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
	unsigned nat = 1;
	if (DS) {
		at[1].id = cudaLaunchAttributeClusterDimension;
		at[1].val.clusterDim.x = unsigned(nw); at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
		nat = 2;
	}
	cfg.attrs = at; cfg.numAttrs = nat;
	if (DS) {
		// both clusters (X workers, V workers) must be resident at the same time
		int nclusters = 0;
		if (cudaOccupancyMaxActiveClusters(&nclusters, jacobi_split_kernel<EP2, DS>, &cfg) != cudaSuccess || nclusters < 2) { cudaGetLastError(); return false; }
	}
	DBuf jlog(size_t(nw) * JS_DEPTH * (16 * 16 + 8));
	double* jl = jlog.p;
	const cudaError_t le = cudaLaunchKernelEx(&cfg, jacobi_split_kernel<EP2, DS>, gt, nblk, tol2, big2, d_cnt, d_info, max_sweeps, flags, jl);
	if (le != cudaSuccess && DS) { cudaGetLastError(); return false; }        // the caller falls back to the global-memory hand-over
	XB_CUDA(le);
	ctx().launches++;
	return true;
}
template <int EP2>
static void launch_split(double* gt, const JacobiPlan& p, double tol2, double big2, unsigned int* d_cnt, unsigned int* d_info, int max_sweeps,
                         size_t smem_cap) {
	const int nw = int(p.nblk) / 2;
	// Nsight Compute cannot launch the cluster + cooperative variant (it aborts the process with LaunchFailed): under the
	// profiler (its injection variables are in the environment) the hand-over goes through global memory
	static const bool profiler_attached = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr;
	if (ctx().svd_dsmem && !profiler_attached && nw >= 2 && nw <= 16 &&
	    launch_split_impl<EP2, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap)) return;
	launch_split_impl<EP2, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap);
}

static void launch_fast_any(double* gt, const JacobiPlan& p, double tol2, double big2, unsigned int* d_cnt, unsigned int* d_info, int max_sweeps,
                            size_t smem_cap) {
	const bool jacc = ctx().svd_jacc && (p.bw == 8 || p.bw == 4);
	const bool pow2 = p.nblk > 2 && (p.nblk & (p.nblk - 1)) == 0;
	if (ctx().svd_split && jacc && p.bw == 8 && pow2 && ctx().svd_recursive && p.nblk <= size_t(ctx().num_sms)) {
		switch (p.ep2) {
			case 1: launch_split<1>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 2: launch_split<2>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 3: launch_split<3>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 4: launch_split<4>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 6: launch_split<6>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 8: launch_split<8>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			default: break;
		}
	}
	if (ctx().svd_gram && p.bw == 8) {
		switch (p.ep2) {
			case 1: launch_gram<1>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 2: launch_gram<2>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 3: launch_gram<3>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 4: launch_gram<4>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 6: launch_gram<6>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			case 8: launch_gram<8>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); return;
			default: break;
		}
	}
	switch (p.ep2) {
		case 1: if (jacc) launch_fast<1, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<1, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		case 2: if (jacc) launch_fast<2, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<2, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		case 3: if (jacc) launch_fast<3, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<3, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		case 4: if (jacc) launch_fast<4, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<4, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		case 6: if (jacc) launch_fast<6, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<6, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		case 8: if (jacc) launch_fast<8, 256, true>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); else launch_fast<8, 256, false>(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap); break;
		default: throw Error(XB_ERR_UNSUPPORTED, "internal: no specialised Jacobi kernel for this row length");
	}
}

// Runs sweeps until convergence (or max_sweeps) in one cooperative launch; returns {converged, sweeps}.
struct JacobiPending {          // convergence record of a launch whose read-back was left to the caller's next synchronisation
	bool active = false; std::string tag; size_t ld = 0, nblk = 0; int bw = 0;
};
// evaluates the record once the stream has been synchronised (it sits in the last 16 doubles of the pinned scratch)
static bool jacobi_finish(const JacobiPending& pend, int& sweeps_out) {
	Context& c = ctx();
	unsigned int* h_info = reinterpret_cast<unsigned int*>(c.h_scratch + c.h_scratch_len - 16);
	const bool dead = h_info[0] == 0xDEADu;
	h_info += 4;
	sweeps_out = int(h_info[1]);
	const bool converged = h_info[2] == 0;
	if (dead) throw Error(XB_ERR_CUDA, "Jacobi SVD: a block hand-over flag was never raised (internal scheduling error)");
	if (getenv("XB_JACOBI_TIMING") != nullptr) {
		if (h_info[-3]) fprintf(stderr, "[jacobi %s] inner = gram %u + rounds on G %u + apply %u kcycles\n", pend.tag.c_str(), h_info[-3], h_info[-2], h_info[-1]);
		fprintf(stderr, "[jacobi %s] ld=%zu bw=%d ctas=%zu sweeps=%d kcycles: load %u inner %u store %u sync %u\n", pend.tag.c_str(), pend.ld, pend.bw,
		        pend.nblk / 2, sweeps_out, h_info[4], h_info[5], h_info[6], h_info[7]);
	}
	return converged;
}
// speculative execution: the convergence record is checked on the device
__global__ void jacobi_spec_kernel(const unsigned int* __restrict__ rec, unsigned int* __restrict__ flag) {
	if (threadIdx.x == 0 && (rec[0] == 0xDEADu || rec[4 + 2] != 0u)) *flag = 2u;
}
__global__ void small_spec_kernel(const unsigned int* __restrict__ info, unsigned int* __restrict__ flag) {
	if (threadIdx.x == 0 && info[2] != 0u) *flag = 2u;
}
// raises the flag if the eps rule (tensor.cpp:1468-1473) would cut inside the first k singular values
__global__ void svd_spec_rank_kernel(const double* __restrict__ S, const int k, const double eps, unsigned int* __restrict__ flag) {
	const double s0 = S[0];
	bool bad = !(s0 == s0);
	for (int j = 1 + threadIdx.x; j < k; j += blockDim.x) if (S[j] <= eps * s0) bad = true;
	if (bad) *flag = 3u;
}
// pending != nullptr: the kernel and the read-back of its convergence record are enqueued, nothing is waited for; the caller calls
// jacobi_finish() after its next synchronisation of the stream (one host round trip per SVD instead of two)
template <typename T>
static bool run_persistent(T* gt, size_t ld, size_t voff, const JacobiPlan& p, double tol, double big, int max_sweeps, int& sweeps_out,
                           size_t smem_cap, const char* tag, JacobiPending* pending = nullptr) {
	Context& c = ctx();
	const size_t n_u32 = 2 * max_sweeps + 12 + 4 * p.nblk;  // sweep counters | info | ready flags (x blocks, v blocks, product log)
	unsigned int* d_cnt = static_cast<unsigned int*>(dalloc_bytes(n_u32 * sizeof(unsigned int)));
	unsigned int* d_info = d_cnt + 2 * max_sweeps + 4;
	XB_CUDA(cudaMemsetAsync(d_cnt, 0, n_u32 * sizeof(unsigned int), c.stream));
	const bool timing = getenv("XB_JACOBI_TIMING") != nullptr;
	if (timing) { const unsigned int flag = 0xC10C; XB_CUDA(cudaMemcpyAsync(d_info, &flag, 4, cudaMemcpyHostToDevice, c.stream)); }
	const int epl_x = int(voff / 32), epl_v = int((ld - voff) / 32);
	const T tol2 = T(tol * tol), big2 = T(big * big);
	if (p.ep2 > 0) launch_fast_any(gt, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap);
	else if (p.EH == 4) launch_persistent<T, 4, 512>(gt, int(ld), epl_x, epl_v, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap);
	else if (p.EH == 8) launch_persistent<T, 8, 512>(gt, int(ld), epl_x, epl_v, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap);
	else launch_persistent<T, 16, 256>(gt, int(ld), epl_x, epl_v, p, tol2, big2, d_cnt, d_info, max_sweeps, smem_cap);
	if (c.speculate) {
		jacobi_spec_kernel<<<1, 32, 0, c.stream>>>(d_info - 4, c.spec_flag);
		XB_LAUNCH_CHECK();
		dfree(d_cnt);
		sweeps_out = 0;
		return true;
	}
	unsigned int* h_info = reinterpret_cast<unsigned int*>(c.h_scratch + c.h_scratch_len - 16);
	XB_CUDA(cudaMemcpyAsync(h_info, d_info - 4, 12 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
	dfree(d_cnt);
	JacobiPending local;
	JacobiPending& pend = pending ? *pending : local;
	pend.active = true; pend.tag = tag; pend.ld = ld; pend.nblk = p.nblk; pend.bw = p.bw;
	if (pending) return true;
	XB_CUDA(cudaStreamSynchronize(c.stream));
	return jacobi_finish(pend, sweeps_out);
}

void Svd::factor(const double* A, size_t m_, size_t n_) {
	XB_REQUIRE(m_ > 0 && n_ > 0, "SVD of an empty matrix");
	ProfScope prof_total("svd");
	m = m_; n = n_;
	swapped = m < n;
	mw = std::max(m, n); nw = std::min(m, n);
	kmax = nw;
	Context& c = ctx();
	// square inputs take the QR step too (from 64 columns on): it is what makes the flipped orientation below possible
	reduced = (mw > 32) && (mw > nw || (c.svd_flip != 0 && c.svd_square_qr != 0 && nw >= 64));
	const size_t smem_cap = std::min<size_t>(c.max_smem_optin, 227 * 1024) - 1024;

	if (svd_small_fits(mw, nw)) {
		// rank ramps: scaling, Jacobi sweeps and the ranking of the singular values in one single-CTA launch (small_f64.cu);
		// the factor is left in the layout extract() reads
		reduced = false; flipped = false; permuted = false; q_deferred = false;
		mdot = mw;
		voff = (mw + 3) / 4 * 4; ld = voff + (nw + 3) / 4 * 4; mt = ld; npad = nw;
		GT.resize(nw * ld); Ssorted.resize(nw); perm.resize((nw + 1) / 2 + 1); scale.resize(2);
		unsigned int* d_info = static_cast<unsigned int*>(dalloc_bytes(4 * sizeof(unsigned int)));
		const long long srs = swapped ? 1 : (long long)n, scs = swapped ? (long long)n : 1;     // Gw(i, j) = A[i * srs + j * scs]
		{
			ProfScope prof_jacobi("svd_jacobi");
			svd_small(A, srs, scs, mw, nw, GT, ld, voff, Ssorted, reinterpret_cast<int*>(perm.p), scale, d_info,
			          std::sqrt(double(mdot)) * DBL_EPS, c.svd_last_sweep_cos, c.svd_max_sweeps, (c.svd_polish && polish > 0) ? 1 : 0);
		}
		S.resize(nw);
		if (c.speculate) {
			small_spec_kernel<<<1, 32, 0, c.stream>>>(d_info, c.spec_flag);
			XB_LAUNCH_CHECK();
			dfree(d_info);
			S.clear();
			return;
		}
		unsigned int* h_info = reinterpret_cast<unsigned int*>(c.h_scratch + c.h_scratch_len - 16);
		XB_CUDA(cudaMemcpyAsync(h_info, d_info, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
		XB_CUDA(cudaMemcpyAsync(c.h_scratch, Ssorted.p, nw * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		dfree(d_info);
		XB_CUDA(cudaStreamSynchronize(c.stream));
		std::copy(c.h_scratch, c.h_scratch + nw, S.begin());
		sweeps = int(h_info[1]);
		if (h_info[2] != 0) throw Error(XB_ERR_NUMERIC, "Jacobi SVD did not converge within svd_max_sweeps sweeps");
		return;
	}

	const double* src; long long rs, cs;
	DBuf At, Rr;
	if (reduced) {
		Qred.resize(mw * nw); Rr.resize(nw * nw);
		// Qred is not needed before extract(): it is formed on the side stream while the Jacobi kernel runs
		const long long ars = swapped ? 1 : (long long)n, acs = swapped ? (long long)n : 1;     // Mw(i, j) = A[i * ars + j * acs]
		if (c.svd_colsort && nw >= 32 && nw <= 1200) {
			// columns of the working matrix in order of descending norm: the poor man's pivoted QR.  The TT sweeps hand over
			// matrices like [Q1 W, Q2 W] with W = U Sigma, and products of many random cores at the first edges: column norms
			// spread over up to nine decades in no particular order, on which the unpivoted factor needs 2 - 4 x the sweeps.
			colperm.resize((nw + 1) / 2 + 1);
			int* perm_ = reinterpret_cast<int*>(colperm.p);
			DBuf sc0(2);
			amax_scale_dev(sc0, A, m * n);
			svd_colrank_kernel<<<1, 1024, 5 * nw * sizeof(double), c.stream>>>(A, int(mw), int(nw), ars, acs, sc0.p, perm_);
			XB_LAUNCH_CHECK();
			At.resize(m * n);
			svd_gather_cols_kernel<<<unsigned(std::min<size_t>((m * n + 255) / 256, size_t(c.num_sms) * 8)), 256, 0, c.stream>>>(At, A, int(mw), int(nw), ars, acs, perm_);
			XB_LAUNCH_CHECK();
			permuted = true;
			qr(Qred, Rr, At, mw, nw, true);
		} else {
			permuted = false;
			if (swapped) { At.resize(m * n); transpose(At, A, m, n); qr(Qred, Rr, At, mw, nw, true); }
			else qr(Qred, Rr, A, mw, nw, true);
		}
		q_deferred = true;
		// Jacobi runs on the columns of R^T (the rows of R), not of R: after a QR step the rows of the triangular factor are
		// far closer to the left singular directions than its columns are to the right ones (Drmac & Veselic's
		// pre-conditioning), so graded inputs need half the sweeps.  The two vector parts swap roles in extract().
		flipped = c.svd_flip != 0;
		src = Rr; mdot = nw;
		if (flipped) { rs = 1; cs = (long long)nw; } else { rs = (long long)nw; cs = 1; }
	} else {
		flipped = false; permuted = false;
		src = A; mdot = mw;
		if (swapped) { rs = 1; cs = (long long)n; } else { rs = (long long)n; cs = 1; }
	}
	// row layout: [ x (mdot) padded to 32 | v (nw) padded to 32 ]; up to 512 elements per part both parts are padded to
	// the same multiple of 64 that the specialised kernel is instantiated for
	voff = (mdot + 31) / 32 * 32;
	ld = voff + (nw + 31) / 32 * 32;
	int ep2 = 0;
	if (c.svd_fast && std::max(mdot, nw) <= 512) {
		ep2 = int((std::max(mdot, nw) + 63) / 64);
		if (ep2 == 5) ep2 = 6;
		if (ep2 == 7) ep2 = 8;
		voff = size_t(64) * ep2; ld = 2 * voff;
	}
	mt = ld;
	XB_REQUIRE(2 * (ld + 2) * sizeof(double) + 128 <= smem_cap, "SVD: matrix too large for the shared-memory Jacobi kernel (min(m,n) <= ~7000)");
	JacobiPlan plan = plan_jacobi(ld, nw, voff, sizeof(double), smem_cap);
	if (plan.persistent) plan.ep2 = ep2;
	npad = plan.npad;
	GT.resize(npad * ld);
	// exact power-of-two scaling of the Jacobi input: norms and cross products are sums of squares
	scale.resize(2);
	amax_scale_dev(scale, src, reduced ? nw * nw : m * n);
	const double tol = std::sqrt(double(mdot)) * DBL_EPS;
	const unsigned init_blocks = unsigned(std::min<size_t>((npad * ld + 255) / 256, size_t(c.num_sms) * 8));
	// G0 V = X recomputed from the untouched input (transposed storage: XT = VT G0^T)
	auto recompute_left = [&]() {
		double* VT = GT.p + voff;
		if (reduced) gemm(GT.p, ld, nw, mdot, 1.0, VT, ld, false, nw, Rr, nw, !flipped, 0.0);
		else if (!swapped) gemm(GT.p, ld, nw, mdot, 1.0, VT, ld, false, nw, A, n, true, 0.0);
		else gemm(GT.p, ld, nw, mdot, 1.0, VT, ld, false, nw, A, n, false, 0.0);
		scale_block_by_dev(GT.p, ld, nw, mdot, scale.p);
	};
	// one Newton-Schulz step on V (quadratic): VT <- (1.5 I - 0.5 VT VT^T) VT
	auto newton_schulz = [&]() {
		double* VT = GT.p + voff;
		DBuf M(nw * nw), T2(nw * nw);
		gemm(M, nw, nw, nw, -0.5, VT, ld, false, nw, VT, ld, true, 0.0);
		add_diag_kernel<<<unsigned((nw + 255) / 256), 256, 0, c.stream>>>(M, nw, 1.5);
		XB_LAUNCH_CHECK();
		gemm(T2, nw, nw, nw, 1.0, M, nw, false, nw, VT, ld, false, 0.0);
		copy2d(VT, ld, T2, nw, nw, nw);
	};

	sweeps = 0;
	bool converged = false;
	JacobiPending pending;
	ProfScope* prof_jacobi = new ProfScope("svd_jacobi");
	svd_init_kernel<<<init_blocks, 256, 0, c.stream>>>(GT, int(ld), int(npad), int(mdot), int(voff), int(nw), src, rs, cs, scale.p);
	XB_LAUNCH_CHECK();
	if (plan.persistent) {
		// without a clean-up run (sweep layer) the convergence record is read back together with the singular values below
		const bool defer = !(c.svd_polish && polish > 1);
		converged = run_persistent<double>(GT.p, ld, voff, plan, tol, c.svd_last_sweep_cos, c.svd_max_sweeps, sweeps, smem_cap, "f64",
		                                   defer ? &pending : nullptr);
	} else {
		// fallback for shapes the cooperative kernel cannot hold: one launch per tournament round
		if (c.speculate) { delete prof_jacobi; throw SpecUnsupported("launch-per-round Jacobi needs the host between sweeps"); }
		const int bw = plan.bw;
		const size_t nblk = plan.nblk;
		const size_t smem = size_t(2 * bw) * ld * sizeof(double);
		static bool attr_set = false;
		if (!attr_set) {
			XB_CUDA(cudaFuncSetAttribute(jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cap)));
			attr_set = true;
		}
		unsigned int* d_info = static_cast<unsigned int*>(dalloc_bytes(8 * sizeof(unsigned int)));
		unsigned int* h_info = reinterpret_cast<unsigned int*>(c.h_scratch);
		const int threads = std::max(64, std::min(1024, 32 * bw));
		for (int sw = 0; sw < c.svd_max_sweeps && !converged; ++sw) {
			XB_CUDA(cudaMemsetAsync(d_info, 0, 8 * sizeof(unsigned int), c.stream));
			if (nblk == 2) {
				jacobi_block_kernel<<<1, threads, smem, c.stream>>>(GT, int(ld), int(ld), int(voff), bw, 2, 0, 1, tol, d_info, 0, 1);
				XB_LAUNCH_CHECK();
			} else {
				for (size_t round = 0; round + 1 < nblk; ++round) {
					jacobi_block_kernel<<<unsigned(nblk / 2), threads, smem, c.stream>>>(GT, int(ld), int(ld), int(voff), bw, int(nblk),
					                                                                     int(round), round == 0 ? 1 : 0, tol, d_info, 0, 1);
					XB_LAUNCH_CHECK();
				}
			}
			XB_CUDA(cudaMemcpyAsync(h_info, d_info, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
			XB_CUDA(cudaStreamSynchronize(c.stream));
			++sweeps;
			converged = (h_info[0] == 0);
		}
		dfree(d_info);
	}
	if (!converged) { delete prof_jacobi; throw Error(XB_ERR_NUMERIC, "Jacobi SVD did not converge within svd_max_sweeps sweeps"); }

	if (c.svd_polish && polish > 0) {
		// Polish: thousands of plane rotations leave V orthogonal only to ~eps*sqrt(#rotations) and X = G V with the same
		// drift.  One Newton-Schulz step re-orthogonalises V, the left part is recomputed from the untouched input, and one
		// clean-up sweep of tiny rotations restores |cos| <= tol between the left vectors: backward error back at the
		// eps*sqrt(n) level of LAPACK.
		newton_schulz();
		recompute_left();
		if (polish > 1 && plan.persistent && nw > 1) { int extra = 0; run_persistent<double>(GT.p, ld, voff, plan, tol, 1e-7, 2, extra, smem_cap, "clean"); sweeps += extra; }
	}
	delete prof_jacobi;

	Ssorted.resize(nw);
	perm.resize((nw + 1) / 2 + 1);   // nw ints
	{
		const size_t sm = nw * sizeof(double);
		XB_REQUIRE(sm <= 48 * 1024, "SVD: too many columns for the sort kernel");
		svd_sort_kernel<<<1, 1024, sm, c.stream>>>(GT, int(ld), int(mdot), int(nw), Ssorted, reinterpret_cast<int*>(perm.p), scale.p + 1);
		XB_LAUNCH_CHECK();
	}
	S.resize(nw);
	if (c.speculate) { S.clear(); return; }                    // rank_for() checks the values on the device
	if (nw + 16 <= c.h_scratch_len) {
		XB_CUDA(cudaMemcpyAsync(c.h_scratch, Ssorted.p, nw * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		XB_CUDA(cudaStreamSynchronize(c.stream));
		std::copy(c.h_scratch, c.h_scratch + nw, S.begin());
	} else {
		XB_CUDA(cudaMemcpyAsync(S.data(), Ssorted.p, nw * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
		XB_CUDA(cudaStreamSynchronize(c.stream));
	}
	if (pending.active && !jacobi_finish(pending, sweeps)) throw Error(XB_ERR_NUMERIC, "Jacobi SVD did not converge within svd_max_sweeps sweeps");
}

size_t Svd::rank_for(size_t max_rank, double eps) {
	Context& c = ctx();
	if (!c.speculate) return truncation_rank(S, max_rank, eps);
	const size_t k = max_rank ? std::min(kmax, max_rank) : kmax;
	svd_spec_rank_kernel<<<1, 256, 0, c.stream>>>(Ssorted, int(k), eps, c.spec_flag);
	XB_LAUNCH_CHECK();
	return k;
}

void Svd::extract(double* U, double* Vt, size_t k, bool scale_u, bool scale_vt, double* dS) {
	XB_REQUIRE(k >= 1 && k <= kmax, "SVD extract: rank out of range");
	Context& c = ctx();
	const int* p = reinterpret_cast<const int*>(perm.p);
	// The X part holds the left vectors of the working matrix G, the V part its right vectors;
	// G = A (not swapped) or A^T (swapped); if reduced, G = Qred * (working matrix).
	double* outX; long long sxi, sxr; int scale_x;
	double* outV; long long svc, svr; int scale_v;
	DBuf Xk;
	if (reduced) { Xk.resize(nw * k); outX = Xk; sxi = (long long)k; sxr = 1; }
	if (!swapped) {
		if (!reduced) { outX = U; sxi = (long long)k; sxr = 1; }
		scale_x = scale_u;
		outV = Vt; svc = 1; svr = (long long)n; scale_v = scale_vt;
	} else {
		if (!reduced) { outX = Vt; sxi = 1; sxr = (long long)n; }
		scale_x = scale_vt;
		outV = U; svc = (long long)k; svr = 1; scale_v = scale_u;
	}
	if (flipped) {
		// working matrix was R^T: its left vectors (X part) are the right vectors of R and vice versa
		std::swap(outX, outV); std::swap(sxi, svc); std::swap(sxr, svr); std::swap(scale_x, scale_v);
	}
	// with sorted columns the right vectors of the working matrix carry the permuted index: the X part if flipped, else the V part
	const int* cp = permuted ? reinterpret_cast<const int*>(colperm.p) : nullptr;
	svd_extract_kernel<<<unsigned(k), 256, 0, c.stream>>>(GT, int(ld), int(mdot), int(voff), int(nw), Ssorted, p, outX, sxi, sxr, scale_x,
	                                                       outV, svc, svr, scale_v, dS, soft_threshold, scale.p, flipped ? cp : nullptr, flipped ? nullptr : cp);
	XB_LAUNCH_CHECK();
	if (reduced) {
		if (q_deferred) { aux_join(); q_deferred = false; }
		if (!swapped) gemm(U, k, m, k, 1.0, Qred, nw, false, nw, Xk, k, false, 0.0);          // U = Qred * Xk
		else gemm(Vt, n, k, n, 1.0, Xk, k, true, nw, Qred, nw, true, 0.0);                    // Vt = Xk^T * Qred^T
	}
}

} // namespace xb
