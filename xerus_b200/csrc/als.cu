// ALS / DMRG sweeps on device-resident tensor trains (reference: src/xerus/algorithms/als.cpp).
//
// The sweep driver follows ALSVariant::solve (als.cpp:483-553) step by step: prepare_x_for_als (:105-182),
// prepare_stacks (:217-253), energy functional (:255-320), move_to_next_index (:340-380), check_for_end_of_sweep
// (:426-475).  The environments are the reference's small dense tensors (SPD: (r, r_A, r); general: (r, r_A, r_A, r);
// rhs: (r_b, r) / (r_b, r_A, r)), kept in HBM for the whole run.
//
// What differs on purpose: the reference densifies the local operator ((r n^s r)^2 doubles, als.cpp:44 — 5 GB at
// BASELINE config 2, 8.8 TB at config 4) and calls a direct solver.  Here the un-contracted network
// {left env, A_p ..., right env} (exactly what construct_local_operator returns, als.cpp:383-401) is applied
// matrix-free as a chain of GEMMs, and the local system is solved by conjugate gradients on the device (the local
// operator is SPD for assumeSPD, and the normal-equation operator P^T A^T A P otherwise), warm-started from the current
// core.  All CG scalars live in device memory; the host only reads the residual norm every few iterations.
// The two-site driver applies the obvious fix for the reference's sweep-turn defect (als.cpp:371,:376 push the slice
// of site currIndex where currIndex + sites - 1 is needed).
#include "tt_internal.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

using namespace xb;

namespace {

constexpr double EPSILON = 8 * 2.220446049250313e-16;   // include/xerus/basic.h:50

// ---- a small dense device tensor + pairwise contraction (the role TensorNetwork::contract(a,b) plays in the reference,
// src/xerus/tensorNetwork.cpp:1037-1229: pick transposition flags, reshuffle only if unavoidable, one GEMM)
struct DT {
	std::vector<size_t> dims;
	const double* p = nullptr;
	DBuf own;
	size_t size() const { size_t s = 1; for (size_t d : dims) s *= d; return s; }
	double* data() { return own.p; }
};

DT dt_view(const double* p, std::vector<size_t> dims) { DT t; t.dims = std::move(dims); t.p = p; return t; }
DT dt_alloc(std::vector<size_t> dims) { DT t; t.dims = std::move(dims); t.own.resize(t.size()); t.p = t.own.p; return t; }
DT dt_copy(const DT& s) { DT t = dt_alloc(s.dims); copy(t.data(), s.p, s.size()); return t; }
DT dt_ones(std::vector<size_t> dims) { DT t = dt_alloc(std::move(dims)); fill(t.data(), 1.0, t.size()); return t; }

// new mode j = old mode order[j]
DT dt_permute(const DT& s, const std::vector<int>& order) {
	const size_t n = s.dims.size();
	std::vector<size_t> shuffle(n), nd(n);
	for (size_t j = 0; j < n; ++j) { shuffle[order[j]] = j; nd[j] = s.dims[order[j]]; }
	DT t = dt_alloc(nd);
	permute(t.data(), s.p, s.dims.data(), shuffle.data(), n);
	return t;
}

bool is_range(const std::vector<int>& v, int start) {
	for (size_t i = 0; i < v.size(); ++i) if (v[i] != start + int(i)) return false;
	return true;
}

// result modes: free modes of X (in order) followed by free modes of Y (in order); cx[i] is contracted with cy[i]
DT dt_contract(const DT& X, const std::vector<int>& cx, const DT& Y, const std::vector<int>& cy) {
	const int nx = int(X.dims.size()), ny = int(Y.dims.size()), k = int(cx.size());
	XB_REQUIRE(cx.size() == cy.size(), "contract: mode lists differ in length");
	std::vector<int> fx, fy;
	for (int i = 0; i < nx; ++i) if (std::find(cx.begin(), cx.end(), i) == cx.end()) fx.push_back(i);
	for (int i = 0; i < ny; ++i) if (std::find(cy.begin(), cy.end(), i) == cy.end()) fy.push_back(i);
	size_t M = 1, N = 1, K = 1;
	std::vector<size_t> od;
	for (int i : fx) { M *= X.dims[i]; od.push_back(X.dims[i]); }
	for (int i : fy) { N *= Y.dims[i]; od.push_back(Y.dims[i]); }
	for (int i = 0; i < k; ++i) {
		XB_REQUIRE(X.dims[cx[i]] == Y.dims[cy[i]], "Index dimensions do not coincide");   // tensorNetwork.cpp:625
		K *= X.dims[cx[i]];
	}
	DT Xp, Yp;
	const double* xa = X.p; const double* yb = Y.p;
	bool tA = false, tB = false;
	if (is_range(cx, nx - k)) tA = false;
	else if (is_range(cx, 0)) tA = true;
	else { std::vector<int> o = fx; o.insert(o.end(), cx.begin(), cx.end()); Xp = dt_permute(X, o); xa = Xp.p; }
	if (is_range(cy, 0)) tB = false;
	else if (is_range(cy, ny - k)) tB = true;
	else { std::vector<int> o = cy; o.insert(o.end(), fy.begin(), fy.end()); Yp = dt_permute(Y, o); yb = Yp.p; }
	DT R = dt_alloc(od);
	gemm(R.data(), N, M, N, 1.0, xa, tA ? M : K, tA, K, yb, tB ? K : N, tB, 0.0);
	return R;
}

double dt_dot(const DT& a, const DT& b) {
	XB_REQUIRE(a.size() == b.size(), "dot: sizes differ");
	return dot(a.p, b.p, a.size());
}

// ---- CG vector kernels: scalars stay on the device -------------------------------------------------------------------
// sc[0] = rr (current r.r), sc[1] = p.q, sc[2] = rr_new
__global__ void cg_step1_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                const double* __restrict__ q, const double* __restrict__ sc, const size_t n) {
	const double pq = sc[1];
	const double alpha = (pq != 0.0) ? sc[0] / pq : 0.0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		x[i] += alpha * p[i];
		r[i] -= alpha * q[i];
	}
}
__global__ void cg_step2_kernel(double* __restrict__ p, const double* __restrict__ r, double* __restrict__ sc, const size_t n) {
	const double rr = sc[0], rrn = sc[2];
	const double beta = (rr != 0.0) ? rrn / rr : 0.0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = r[i] + beta * p[i];
}
__global__ void cg_shift_kernel(double* __restrict__ sc) { if (threadIdx.x == 0 && blockIdx.x == 0) sc[0] = sc[2]; }

unsigned vec_grid(size_t n) { return unsigned(std::min<size_t>(std::max<size_t>(1, (n + 255) / 256), size_t(ctx().num_sms) * 4)); }

// One CG iteration's vector work in a single CTA (local problems up to CG_FUSED_MAX unknowns: at BASELINE config 2 a vector
// is 200 KB and the five separate launches cost more than the arithmetic): with q = A p,
//   alpha = rr / p.q ; x += alpha p ; r -= alpha q ; rr' = r.r ; p = r + (rr'/rr) p ; sc[0] = rr', sc[1] = p.q
// Reductions run in a fixed order (deterministic).
constexpr size_t CG_FUSED_MAX = 262144;
constexpr int CG_CS = 8;           // CTAs per cluster
// sum over the whole cluster: warp shuffle, CTA partial to every CTA's slot (DSMEM), one cluster barrier, fixed-order sum
__device__ __forceinline__ double cluster_sum(double v, double (*slots)[CG_CS], double* sh, const int phase,
                                              cooperative_groups::cluster_group& cluster) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) sh[phase * 32 + (threadIdx.x >> 5)] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		double t = sh[phase * 32 + threadIdx.x];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (threadIdx.x < CG_CS) *cluster.map_shared_rank(&slots[phase][cluster.block_rank()], threadIdx.x) = t;
	}
	cluster.sync();
	double s = 0.0;
#pragma unroll
	for (int c = 0; c < CG_CS; ++c) s += slots[phase][c];
	return s;
}
__global__ void __cluster_dims__(CG_CS, 1, 1) __launch_bounds__(1024) cg_fused_kernel(double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                                                                                     const double* __restrict__ q, double* __restrict__ sc, const int n) {
	cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
	__shared__ double sh[64];
	__shared__ double slots[2][CG_CS];
	const double rr = sc[0];
	const int stride = CG_CS * 1024, first = blockIdx.x * 1024 + threadIdx.x;
	double acc = 0.0;
	for (int i = first; i < n; i += stride) acc += p[i] * q[i];
	const double pq = cluster_sum(acc, slots, sh, 0, cluster);
	const double alpha = (pq != 0.0) ? rr / pq : 0.0;
	acc = 0.0;
	for (int i = first; i < n; i += stride) {
		x[i] += alpha * p[i];
		const double rn = r[i] - alpha * q[i];
		r[i] = rn;
		acc += rn * rn;
	}
	const double rrn = cluster_sum(acc, slots, sh, 1, cluster);
	const double beta = (rr != 0.0) ? rrn / rr : 0.0;
	for (int i = first; i < n; i += stride) p[i] = r[i] + beta * p[i];
	// sc[0] is read by every CTA at entry: nobody may overwrite it before all have passed the second barrier (they have)
	if (blockIdx.x == 0 && threadIdx.x == 0) { sc[0] = rrn; sc[1] = pq; sc[2] = rrn; }
}

// One-site SPD local operator prepared once per site: y(l,m,r) = sum L(l,a,l') A(a,m,n,b) R(r,b,r') v(l',n,r').
// With the operator core pre-shuffled to (m,b | a,n) the application is three GEMMs into fixed workspaces and no
// reshuffle:  t1(l,a | n,r') = L v ;  u(l | m,b | r') = A2 t1(l)  (batched over l) ;  y(l,m | r) = u R^T.
struct SpdSiteApply {
	size_t l = 0, a = 0, m = 0, n = 0, b = 0, r = 0;
	const double* L = nullptr; const double* R = nullptr;
	DBuf A2, t1, u;
	void prepare(const DT& Lenv, const DT& Acore, const DT& Renv) {
		l = Lenv.dims[0]; a = Lenv.dims[1]; r = Renv.dims[0]; b = Renv.dims[1];
		m = Acore.dims[1]; n = Acore.dims[2];
		XB_REQUIRE(Lenv.dims[2] == l && Renv.dims[2] == r && Acore.dims[0] == a && Acore.dims[3] == b, "internal: local operator shapes");
		L = Lenv.p; R = Renv.p;
		DT s = dt_permute(Acore, {1, 3, 0, 2});          // (m, b, a, n)
		A2 = std::move(s.own);
		t1.resize(l * a * n * r); u.resize(l * m * b * r);
	}
	void apply(const double* v, double* y) {
		gemm(t1, n * r, l * a, n * r, 1.0, L, l, false, l, v, n * r, false, 0.0);
		gemm_batched(u, r, m * b * r, m * b, r, 1.0, A2, a * n, 0, false, a * n, t1, r, a * n * r, false, 0.0, l);
		gemm(y, r, l * m, r, 1.0, u, b * r, false, b * r, R, b * r, true, 0.0);
	}
};

// ---- persistent CG for the one-site SPD local problem -----------------------------------------------------------------------
// At BASELINE config 2 an operator application is 12 MFLOP: as three GEMM launches plus the fused vector kernel a CG iteration
// costs ~35 us of launch and drain latency for ~1 us of arithmetic.  Here a whole CG run is ONE cooperative launch.  CTA g owns
// the slices li = g, g + G, ... of the left bond: for a slice the three factors chain without leaving the SM,
//   t1(a,n | r') = L(li, a, :) p          (reads all of p from L2 — the only all-to-all step of an iteration)
//   u(m,b | r')  = A2 t1                  (shared memory)
//   y(li, m, r)  = u(m | b,r') R(r | b,r')^T   (R stays transposed in shared memory for the whole run)
// and the CTA also owns the matching segments of x, r, p.  Per iteration there are two reductions (p.q and r.r: per-CTA
// partials in global memory, summed by every CTA in the same fixed order, so all CTAs take identical decisions) and two
// grid barriers (three in the textbook arrangement, SpdCgArgs::merged = 0).  The stopping logic of the host loop (target, stagnation at the rounding floor, re-anchoring) runs on the
// device; the host reads one status record per launch.
struct SpdCgArgs {
	const double* L; const double* A2; const double* R;
	const double* rhs;               // right-hand side of the local problem
	double* x; double* p; double* partial; double* sc; unsigned int* info; unsigned int* barrier;
	int l, a, n, m, b, rr;           // L (l, a, l), A2 (m b | a n), R (rr, b, rr); vectors (l, n, rr) with m == n
	int nsl;                         // slices per CTA (1 whenever l <= number of co-resident CTAs)
	int max_it;
	int merged;                      // 1: two grid barriers per iteration (q = A r + beta q, see the loop)
	double target, bnorm2;
};
constexpr int CGP_THREADS = 256;
constexpr int CGP_AMAX = 4;          // operator bond dimension handled by the register accumulators of step 1
constexpr int CGP_ROWS = 25;         // rows of p in flight per thread in step 1

__device__ __forceinline__ double cgp_block_sum(double v, double* red) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double t = 0.0;
#pragma unroll
	for (int w = 0; w < CGP_THREADS / 32; ++w) t += red[w];
	return t;
}
// sum of the G per-CTA partials, same order in every CTA (bit-identical results everywhere)
__device__ __forceinline__ double cgp_grid_sum(const double* partial, const int G, double* red) {
	double t = 0.0;
	if (threadIdx.x < 32) {
		for (int i = threadIdx.x; i < G; i += 32) t += __ldcg(partial + i);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (threadIdx.x == 0) red[16] = t;
	}
	__syncthreads();
	const double s = red[16];
	__syncthreads();
	return s;
}
// Grid barrier on a monotone counter (the launch is cooperative, so all CTAs are resident).  The CTA barrier orders every
// thread's global stores before thread 0's fence, which is cumulative; readers use ld.global.cg afterwards.  Bounded spin.
__device__ __forceinline__ void cgp_grid_barrier(unsigned int* counter, unsigned int& expected, const int G, unsigned int* info) {
	__syncthreads();
	if (threadIdx.x == 0) {
		expected += (unsigned)G;
		__threadfence();
		atomicAdd(counter, 1u);
		unsigned int spins = 0;
		while (*((volatile unsigned int*)counter) < expected && ++spins < (1u << 26)) {}
		if (spins >= (1u << 26)) info[2] = 0xDEADu;
		__threadfence();
	}
	__syncthreads();
}

__device__ __forceinline__ unsigned cgp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cgp_mapa(unsigned addr, unsigned rank) {
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
// CSZ > 1: the CTAs form clusters of CSZ.  Step 1 of an application (t1 = L(li) src, which needs ALL of src in every CTA and is
// bound by L2 latency: 40 % of an iteration) is shared inside a cluster: a CTA reads only its 1/CSZ of the rows of src, forms
// the partial t1 of all CSZ slices of its cluster from them, and the partials are exchanged through DSMEM — one bulk copy
// (cp.async.bulk shared::cta -> shared::cluster, 8 KB at config 2) per destination, completion counted on the destination's
// mbarrier — and summed in member order.  L2 traffic and load latency per CTA drop by CSZ, the arithmetic per CTA is unchanged.
template <int CSZ>
__global__ void __launch_bounds__(CGP_THREADS) spd_cg_kernel(const SpdCgArgs g) {
	extern __shared__ double cgp_smem[];
	const int l = g.l, a = g.a, n = g.n, m = g.m, b = g.b, R = g.rr, nsl = g.nsl;
	const int Cin = n * R, Cout = m * R, KA = a * n, QA = m * b, KR = b * R;
	const int Rp = (R + 1) & ~1;                   // row strides padded to even: 16-byte shared-memory accesses
	double* Rt = cgp_smem;                         // [KR][Rp]   Rt[(bb, r')][ro] = R[ro][bb][r']
	double* A2s = Rt + (size_t)KR * Rp;            // [QA][KA]
	double* Ls = A2s + ((QA * KA + 1) & ~1);       // [l][CGP_AMAX]  weights of the current slice, transposed, zero padded
	double* t1s = Ls + CGP_AMAX * l;               // [KA][Rp]
	double* us = t1s + KA * Rp;                    // [m][KR]  (KR even or odd: scalar reads of u)
	double* xs = us + ((m * KR + 1) & ~1);         // [nsl][Cout] owned segments of x, r, p, q
	double* rs = xs + (size_t)nsl * Cout;
	double* ps = rs + (size_t)nsl * Cout;
	double* qs = ps + (size_t)nsl * Cout;
	double* bs = qs + (size_t)nsl * Cout;          //            ... and of the right-hand side
	double* hs = bs + (size_t)nsl * Cout;          //            ... and of q = A p carried by recurrence (merged variant)
	double* red = hs + (size_t)nsl * Cout;         // [32]
	// cluster variant: weights of all slices of the cluster over this CTA's rows, staged partials, received partials
	const int crows = (l + CSZ - 1) / CSZ + 1;
	double* Lc = red + 32;                         // [CSZ][crows][2]
	double* stage = Lc + ((CSZ * crows * 2 + 1) & ~1);   // [CSZ][KA * Rp]
	double* recvp = stage + (size_t)CSZ * KA * Rp; // [CSZ][KA * Rp]
	__shared__ __align__(8) unsigned long long xbar;
	unsigned int applies = 0;
	if (CSZ > 1) {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(cgp_smem_u32(&xbar)), "r"(1u) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		cooperative_groups::this_cluster().sync();
	}
	const int tid = threadIdx.x, G = gridDim.x;
	for (int e = tid; e < KR * R; e += CGP_THREADS) { const int ro = e / KR, k = e % KR; Rt[k * Rp + ro] = g.R[e]; }
	if (Rp != R) for (int k = tid; k < KR; k += CGP_THREADS) Rt[k * Rp + R] = 0.0;
	for (int e = tid; e < QA * KA; e += CGP_THREADS) A2s[e] = g.A2[e];
	for (int s = 0; s < nsl; ++s) {
		const int li = blockIdx.x + s * G;
		if (li < l) for (int e = tid; e < Cout; e += CGP_THREADS) {
			const size_t o = (size_t)li * Cout + e;
			xs[s * Cout + e] = g.x[o]; bs[s * Cout + e] = g.rhs[o];
		}
	}
	__syncthreads();

	double rr = 0.0, rr0 = -1.0, best = 0.0;
	int since_best = 0, it = 0;
	unsigned int reason = 0;                       // 1 target, 2 stagnation, 3 re-anchor, 4 NaN, 0 iteration budget
	unsigned int bar_expected = 0;
	const bool timing = g.info[3] == 0xC10C && blockIdx.x == 0 && tid == 0;
	long long tk[6] = {0, 0, 0, 0, 0, 0}, t0 = 0;  // step 1 | steps 2-3 | barrier 1 | update | barrier 2 | barrier 3
	const bool vec_ok = (Cin & 1) == 0;
	// qs <- (A src) on the owned slices; dot_local += <wsm, qs> with wsm the shared-memory copy of src's owned segments.
	// src is read through L2 (every CTA needs all of it, and other CTAs wrote it).
	auto apply_owned = [&](const double* __restrict__ src, const double* __restrict__ wsm, double& dot_local) {
		if (timing) t0 = clock64();
		for (int s = 0; s < nsl; ++s) {
			const int li = blockIdx.x + s * G;
			if (li >= l) break;                      // uniform per CTA
			if (CSZ > 1) {
				// ---- step 1 shared inside the cluster (host guarantees nsl == 1, a <= 2, Cin even, G % CSZ == 0)
				const int cr = blockIdx.x % CSZ, cbase = blockIdx.x - cr;
				const int lr0 = (cr * l) / CSZ, lr1 = ((cr + 1) * l) / CSZ, nr = lr1 - lr0;
				const int TS = KA * Rp;                                   // doubles of one partial t1
				if (tid < CSZ) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // last application's copies have read `stage`
				for (int e = tid; e < CSZ * nr * 2; e += CGP_THREADS) {
					const int c = e / (nr * 2), i = (e / 2) % nr, aa = e & 1;
					Lc[(c * crows + i) * 2 + aa] = (aa < a) ? g.L[((size_t)(cbase + c) * a + aa) * l + lr0 + i] : 0.0;
				}
				__syncthreads();
				for (int j2 = tid; j2 < (Cin >> 1); j2 += CGP_THREADS) {
					double acc[CSZ][2][2];
#pragma unroll
					for (int c = 0; c < CSZ; ++c) { acc[c][0][0] = 0.0; acc[c][0][1] = 0.0; acc[c][1][0] = 0.0; acc[c][1][1] = 0.0; }
					const double2* pj = reinterpret_cast<const double2*>(src) + j2;
					for (int base = 0; base < nr; base += CGP_ROWS) {
						double2 pv[CGP_ROWS];
#pragma unroll
						for (int i = 0; i < CGP_ROWS; ++i) pv[i] = (base + i < nr) ? __ldcg(pj + (size_t)(lr0 + base + i) * (Cin >> 1)) : make_double2(0.0, 0.0);
#pragma unroll
						for (int i = 0; i < CGP_ROWS; ++i) {
							if (base + i < nr) {
#pragma unroll
								for (int c = 0; c < CSZ; ++c) {
									const double2 w = *reinterpret_cast<const double2*>(Lc + (c * crows + base + i) * 2);
									acc[c][0][0] += w.x * pv[i].x; acc[c][0][1] += w.x * pv[i].y;
									acc[c][1][0] += w.y * pv[i].x; acc[c][1][1] += w.y * pv[i].y;
								}
							}
						}
					}
					const int j = 2 * j2, n0 = j / R, r0 = j - n0 * R;
					const int n1 = (r0 + 1 < R) ? n0 : n0 + 1, r1 = (r0 + 1 < R) ? r0 + 1 : 0;
#pragma unroll
					for (int c = 0; c < CSZ; ++c)
#pragma unroll
						for (int aa = 0; aa < 2; ++aa) if (aa < a) {
							stage[(size_t)c * TS + (aa * n + n0) * Rp + r0] = acc[c][aa][0];
							stage[(size_t)c * TS + (aa * n + n1) * Rp + r1] = acc[c][aa][1];
						}
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				__syncthreads();
				const unsigned mb = cgp_smem_u32(&xbar);
				if (tid < CSZ && tid != cr) {
					asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					             :: "r"(cgp_mapa(cgp_smem_u32(recvp + (size_t)cr * TS), (unsigned)tid)), "r"(cgp_smem_u32(stage + (size_t)tid * TS)),
					                "r"((unsigned)(TS * 8)), "r"(cgp_mapa(mb, (unsigned)tid)) : "memory");
					asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				}
				if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"((unsigned)((CSZ - 1) * TS * 8)) : "memory");
				{
					unsigned done = 0, spins = 0;
					const unsigned parity = applies & 1u;
					do {
						asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
						             : "=r"(done) : "r"(mb), "r"(parity) : "memory");
					} while (!done && ++spins < (1u << 24));
					if (!done && tid == 0) g.info[2] = 0xDEADu;
				}
				++applies;
				// t1 = sum of the partials in member order (the own one from `stage`)
				for (int e = tid; e < TS; e += CGP_THREADS) {
					double t = 0.0;
#pragma unroll
					for (int c = 0; c < CSZ; ++c) t += (c == cr) ? stage[(size_t)cr * TS + e] : recvp[(size_t)c * TS + e];
					t1s[e] = t;
				}
			} else {
			// weights of the slice, transposed and zero padded to CGP_AMAX per row: Ls[lp][aa] = L(li, aa, lp)
			for (int e = tid; e < CGP_AMAX * l; e += CGP_THREADS) { const int lp = e / CGP_AMAX, aa = e % CGP_AMAX; Ls[e] = (aa < a) ? g.L[((size_t)li * a + aa) * l + lp] : 0.0; }
			__syncthreads();
			// step 1: t1(aa, j) = sum_l' L(li, aa, l') p(l', j).  Every CTA needs all of p (ld.global.cg: other CTAs wrote it):
			// CGP_ROWS 16-byte loads in flight per thread, the weights of a row come as one broadcast 16-byte shared-memory load.
			if (vec_ok) {
				for (int j2 = tid; j2 < (Cin >> 1); j2 += CGP_THREADS) {
					double acc[CGP_AMAX][2];
#pragma unroll
					for (int aa = 0; aa < CGP_AMAX; ++aa) { acc[aa][0] = 0.0; acc[aa][1] = 0.0; }
					const double2* pj = reinterpret_cast<const double2*>(src) + j2;
					for (int base = 0; base < l; base += CGP_ROWS) {
						double2 pv[CGP_ROWS];
#pragma unroll
						for (int i = 0; i < CGP_ROWS; ++i) pv[i] = (base + i < l) ? __ldcg(pj + (size_t)(base + i) * (Cin >> 1)) : make_double2(0.0, 0.0);
#pragma unroll
						for (int i = 0; i < CGP_ROWS; ++i) {
							if (base + i < l) {
								const double2 w01 = *reinterpret_cast<const double2*>(Ls + (base + i) * CGP_AMAX);
								acc[0][0] += w01.x * pv[i].x; acc[0][1] += w01.x * pv[i].y;
								acc[1][0] += w01.y * pv[i].x; acc[1][1] += w01.y * pv[i].y;
								if (a > 2) {
									const double2 w23 = *reinterpret_cast<const double2*>(Ls + (base + i) * CGP_AMAX + 2);
									acc[2][0] += w23.x * pv[i].x; acc[2][1] += w23.x * pv[i].y;
									acc[3][0] += w23.y * pv[i].x; acc[3][1] += w23.y * pv[i].y;
								}
							}
						}
					}
					// t1s rows are indexed (aa, nidx) with row stride Rp: column j = nidx * R + r'
					const int j = 2 * j2, n0 = j / R, r0 = j - n0 * R;
					const int n1 = (r0 + 1 < R) ? n0 : n0 + 1, r1 = (r0 + 1 < R) ? r0 + 1 : 0;
#pragma unroll
					for (int aa = 0; aa < CGP_AMAX; ++aa) if (aa < a) { t1s[(aa * n + n0) * Rp + r0] = acc[aa][0]; t1s[(aa * n + n1) * Rp + r1] = acc[aa][1]; }
				}
			} else {
				for (int j = tid; j < Cin; j += CGP_THREADS) {
					double acc[CGP_AMAX] = {0.0, 0.0, 0.0, 0.0};
					const double* pj = src + j;
#pragma unroll 8
					for (int lp = 0; lp < l; ++lp) {
						const double pv = __ldcg(pj + (size_t)lp * Cin);
#pragma unroll
						for (int aa = 0; aa < CGP_AMAX; ++aa) acc[aa] += Ls[lp * CGP_AMAX + aa] * pv;
					}
#pragma unroll
					for (int aa = 0; aa < CGP_AMAX; ++aa) if (aa < a) t1s[(aa * n + j / R) * Rp + j % R] = acc[aa];
				}
			}
			}
			__syncthreads();
			if (timing) { const long long t1 = clock64(); tk[0] += t1 - t0; t0 = t1; }
			// step 2: u(q, r') = sum_k A2(q, k) t1(k, r'): 2 x 2 outputs per thread
			for (int e = tid; e < ((QA + 1) / 2) * (Rp / 2); e += CGP_THREADS) {
				const int q0 = (e / (Rp / 2)) * 2, rp = (e % (Rp / 2)) * 2;
				const int q1 = (q0 + 1 < QA) ? q0 + 1 : q0;
				double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll 4
				for (int k = 0; k < KA; ++k) {
					const double2 tv = *reinterpret_cast<const double2*>(t1s + k * Rp + rp);
					const double w0 = A2s[q0 * KA + k], w1 = A2s[q1 * KA + k];
					a00 += w0 * tv.x; a01 += w0 * tv.y; a10 += w1 * tv.x; a11 += w1 * tv.y;
				}
				// us is indexed [mm][(bb, r')] = flat q * R + r'
				us[q0 * R + rp] = a00; if (rp + 1 < R) us[q0 * R + rp + 1] = a01;
				if (q1 != q0) { us[q1 * R + rp] = a10; if (rp + 1 < R) us[q1 * R + rp + 1] = a11; }
			}
			__syncthreads();
			// step 3: y(mm, ro) = sum_kk u(mm, kk) Rt(kk, ro): 2 x 2 outputs per thread; p.q on the fly
			for (int e = tid; e < ((m + 1) / 2) * (Rp / 2); e += CGP_THREADS) {
				const int m0 = (e / (Rp / 2)) * 2, ro = (e % (Rp / 2)) * 2;
				const int m1 = (m0 + 1 < m) ? m0 + 1 : m0;
				const double* u0 = us + (size_t)m0 * KR;
				const double* u1 = us + (size_t)m1 * KR;
				double y00 = 0.0, y01 = 0.0, y10 = 0.0, y11 = 0.0;
#pragma unroll 4
				for (int kk = 0; kk < KR; ++kk) {
					const double2 rv = *reinterpret_cast<const double2*>(Rt + kk * Rp + ro);
					const double ua = u0[kk], ub = u1[kk];
					y00 += ua * rv.x; y01 += ua * rv.y; y10 += ub * rv.x; y11 += ub * rv.y;
				}
				const int o0 = s * Cout + m0 * R + ro;
				qs[o0] = y00; dot_local += wsm[o0] * y00;
				if (ro + 1 < R) { qs[o0 + 1] = y01; dot_local += wsm[o0 + 1] * y01; }
				if (m1 != m0) {
					qs[o0 + R] = y10; dot_local += wsm[o0 + R] * y10;
					if (ro + 1 < R) { qs[o0 + R + 1] = y11; dot_local += wsm[o0 + R + 1] * y11; }
				}
			}
			__syncthreads();
		}
	};

	// The host loop of local_solve_cg, on the device: true residual of the current x, CG from it until the recurrence residual
	// meets the target / stagnates / has dropped 20 decades, re-anchor at the true residual, at most six times.
	for (int restart = 0; restart < 6; ++restart) {
		if (restart > 0) {
			for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
				const int li = blockIdx.x + (e / Cout) * G;
				if (li < l) __stcg(g.x + (size_t)li * Cout + (e % Cout), xs[e]);
			}
			cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
		}
		double dummy = 0.0, rl = 0.0;
		apply_owned(g.x, xs, dummy);
		for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
			if (blockIdx.x + (e / Cout) * G < l) { const double rn = bs[e] - qs[e]; rs[e] = rn; rl += rn * rn; }
		}
		double* rpart = g.partial + (size_t)(4 + (restart & 1)) * G;
		const double rl_cta = cgp_block_sum(rl, red);
		if (tid == 0) __stcg(rpart + blockIdx.x, rl_cta);
		cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
		rr = cgp_grid_sum(rpart, G, red);
		if (rr0 < 0.0) rr0 = rr;
		if (!(rr == rr)) { reason = 4; break; }
		if (rr <= g.target) { reason = 1; break; }
		if (it >= g.max_it) { reason = 0; break; }
		if (restart == 0 && rr > 1e4 * g.bnorm2) {       // useless warm start: begin from zero instead
			for (int e = tid; e < nsl * Cout; e += CGP_THREADS) { xs[e] = 0.0; rs[e] = bs[e]; }
			rr = g.bnorm2;
		}
		for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
			const int li = blockIdx.x + (e / Cout) * G;
			if (li < l) { ps[e] = rs[e]; __stcg(g.p + (size_t)li * Cout + (e % Cout), rs[e]); }
		}
		cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
		best = rr; since_best = 0; reason = 0;
		const double rr_start = rr;
		double beta_m = 0.0;                          // merged variant: p_{-1} = q_{-1} = 0
		if (g.merged) for (int e = tid; e < nsl * Cout; e += CGP_THREADS) { ps[e] = 0.0; hs[e] = 0.0; }
	for (; it < g.max_it; ++it) {
		double* part = g.partial + (size_t)(it & 1) * 2 * G;
		if (timing) t0 = clock64();
		if (g.merged) {
			// Two barriers per iteration instead of three: what the other CTAs need is r, not p — A p = A r + beta A p_old, so
			// the operator is applied to the published residual and q, p follow by recurrence on the owned segments.
			// Same iterates in exact arithmetic; the recurrence for q adds to the gap between recurrence and true residual,
			// which the re-anchoring below absorbs.
			double dummy = 0.0, pq_local = 0.0;
			apply_owned(g.p, rs, dummy);                                  // qs = A r   (g.p holds r)
			for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
				if (blockIdx.x + (e / Cout) * G < l) {
					const double h = qs[e] + beta_m * hs[e], pn = rs[e] + beta_m * ps[e];
					hs[e] = h; ps[e] = pn;
					pq_local += pn * h;
				}
			}
			const double pq_cta = cgp_block_sum(pq_local, red);
			if (tid == 0) __stcg(part + blockIdx.x, pq_cta);
			if (timing) { const long long t1 = clock64(); tk[1] += t1 - t0; t0 = t1; }
			cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
			if (timing) { const long long t1 = clock64(); tk[2] += t1 - t0; t0 = t1; }
			const double pq = cgp_grid_sum(part, G, red);
			const double alpha = (pq != 0.0) ? rr / pq : 0.0;
			double rr_local = 0.0;
			for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
				const int li = blockIdx.x + (e / Cout) * G;
				if (li < l) {
					xs[e] += alpha * ps[e];
					const double rn = rs[e] - alpha * hs[e];
					rs[e] = rn;
					rr_local += rn * rn;
					__stcg(g.p + (size_t)li * Cout + (e % Cout), rn);
				}
			}
			const double rr_cta = cgp_block_sum(rr_local, red);
			if (tid == 0) __stcg(part + G + blockIdx.x, rr_cta);
			if (timing) { const long long t1 = clock64(); tk[3] += t1 - t0; t0 = t1; }
			cgp_grid_barrier(g.barrier, bar_expected, G, g.info);         // r.r is complete and the new r is visible to every CTA
			if (timing) { const long long t1 = clock64(); tk[4] += t1 - t0; t0 = t1; }
			const double rrn = cgp_grid_sum(part + G, G, red);
			beta_m = (rr != 0.0) ? rrn / rr : 0.0;
			rr = rrn;
			if (!(rr == rr)) { reason = 4; ++it; break; }
			if (rr <= g.target) { reason = 1; ++it; break; }
			if (rr < 0.5 * best) { best = rr; since_best = 0; } else if (++since_best >= 48) { reason = 2; ++it; break; }
			if (rr < 1e-20 * rr_start) { reason = 3; ++it; break; }
			continue;
		}
		// ---- q = A p on the owned slices, and the partial of p.q
		double pq_local = 0.0;
		apply_owned(g.p, ps, pq_local);
		const double pq_cta = cgp_block_sum(pq_local, red);
		if (tid == 0) __stcg(part + blockIdx.x, pq_cta);
		if (timing) { const long long t1 = clock64(); tk[1] += t1 - t0; t0 = t1; }
		cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
		if (timing) { const long long t1 = clock64(); tk[2] += t1 - t0; t0 = t1; }
		const double pq = cgp_grid_sum(part, G, red);
		const double alpha = (pq != 0.0) ? rr / pq : 0.0;
		// ---- x += alpha p ; r -= alpha q ; partial of r.r   (owned segments, shared memory)
		double rr_local = 0.0;
		for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
			if (blockIdx.x + (e / Cout) * G < l) {
				xs[e] += alpha * ps[e];
				const double rn = rs[e] - alpha * qs[e];
				rs[e] = rn;
				rr_local += rn * rn;
			}
		}
		const double rr_cta = cgp_block_sum(rr_local, red);
		if (tid == 0) __stcg(part + G + blockIdx.x, rr_cta);
		if (timing) { const long long t1 = clock64(); tk[3] += t1 - t0; t0 = t1; }
		cgp_grid_barrier(g.barrier, bar_expected, G, g.info);
		if (timing) { const long long t1 = clock64(); tk[4] += t1 - t0; t0 = t1; }
		const double rrn = cgp_grid_sum(part + G, G, red);
		const double beta = (rr != 0.0) ? rrn / rr : 0.0;
		for (int e = tid; e < nsl * Cout; e += CGP_THREADS) {
			const int li = blockIdx.x + (e / Cout) * G;
			if (li < l) { const double pn = rs[e] + beta * ps[e]; ps[e] = pn; __stcg(g.p + (size_t)li * Cout + (e % Cout), pn); }
		}
		rr = rrn;
		// stopping rules of the host loop (local_solve_cg), evaluated identically by every CTA
		if (!(rr == rr)) { reason = 4; ++it; break; }
		if (rr <= g.target) { reason = 1; ++it; break; }
		if (rr < 0.5 * best) { best = rr; since_best = 0; } else if (++since_best >= 48) { reason = 2; ++it; break; }
		if (rr < 1e-20 * rr_start) { reason = 3; ++it; break; }
		if (timing) { const long long t1 = clock64(); tk[3] += t1 - t0; t0 = t1; }
		cgp_grid_barrier(g.barrier, bar_expected, G, g.info);     // the new p is visible to every CTA
		if (timing) { const long long t1 = clock64(); tk[5] += t1 - t0; t0 = t1; }
	}
		if (reason == 4) break;
		// (leaving the CG loop through a stopping rule skips its last barrier; the next residual starts with one after publishing x)
	}
	__syncthreads();
	for (int s = 0; s < nsl; ++s) {
		const int li = blockIdx.x + s * G;
		if (li < l) for (int e = tid; e < Cout; e += CGP_THREADS) {
			const size_t o = (size_t)li * Cout + e;
			g.x[o] = xs[s * Cout + e];
		}
	}
	if (blockIdx.x == 0 && tid == 0) { g.sc[0] = rr; g.sc[1] = rr0; g.info[0] = (unsigned)it; g.info[1] = reason; }
	if (timing) { for (int i = 0; i < 6; ++i) g.info[4 + i] = (unsigned)(tk[i] >> 4); }
}

// launches spd_cg_kernel if the shapes fit; returns false (nothing done) otherwise
// (x: warm start in, solution out; b: right-hand side; p: work vector; rr0_out: squared norm of the first true residual)
bool spd_cg_persistent(const SpdSiteApply& sa, double* x, const double* b, double* p, double* sc, size_t max_it, double target,
                       double bnorm2, size_t& iterations, unsigned& reason, double& rr_out, double& rr0_out) {
	Context& c = ctx();
	if (!c.als_persistent_cg || sa.m != sa.n || sa.a > size_t(CGP_AMAX) || sa.l > 4096 || sa.r > 4096) return false;
	const size_t KR = sa.b * sa.r, KA = sa.a * sa.n, QA = sa.m * sa.b, Rp = (sa.r + 1) & ~size_t(1), Cout = sa.m * sa.r;
	const size_t cap = std::min<size_t>(c.max_smem_optin, 227 * 1024) - 1024;
	auto smem_for = [&](size_t nsl, size_t csz) {
		const size_t crows = (sa.l + csz - 1) / csz + 1;
		const size_t cluster_part = csz > 1 ? ((csz * crows * 2 + 1) & ~size_t(1)) + 2 * csz * KA * Rp : 2;
		return (KR * Rp + ((QA * KA + 1) & ~size_t(1)) + size_t(CGP_AMAX) * sa.l + KA * Rp + ((sa.m * KR + 1) & ~size_t(1)) + 6 * nsl * Cout + 32 +
		        cluster_part) * sizeof(double);
	};
	struct Variant { const void* fn; int csz; };
	const Variant variants[4] = {{(const void*)spd_cg_kernel<8>, 8}, {(const void*)spd_cg_kernel<5>, 5}, {(const void*)spd_cg_kernel<4>, 4}, {(const void*)spd_cg_kernel<2>, 2}};
	static const bool profiler_attached = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr;
	DBuf partial;
	unsigned int* info = static_cast<unsigned int*>(dalloc_bytes(16 * sizeof(unsigned int)));
	const bool timing = getenv("XB_CG_TIMING") != nullptr;
	{ const unsigned int init[12] = {0, 0, 0, timing ? 0xC10Cu : 0u, 0, 0, 0, 0, 0, 0, 0, 0}; XB_CUDA(cudaMemcpyAsync(info, init, sizeof(init), cudaMemcpyHostToDevice, c.stream)); }
	SpdCgArgs g;
	g.L = sa.L; g.A2 = sa.A2; g.R = sa.R; g.x = x; g.rhs = b; g.p = p; g.sc = sc; g.info = info; g.barrier = info + 11;
	g.l = int(sa.l); g.a = int(sa.a); g.n = int(sa.n); g.m = int(sa.m); g.b = int(sa.b); g.rr = int(sa.r);
	g.max_it = int(std::min<size_t>(max_it, 1u << 30)); g.target = target; g.bnorm2 = bnorm2; g.merged = c.als_cg_merged ? 1 : 0;
	int G = 0;
	bool launched = false;
	// cluster variant: one slice per CTA, operator bond <= 2, even row length, 16-byte multiples for the bulk copies, and the
	// whole grid co-resident as clusters of csz (largest cluster size that divides the number of slices)
	if (c.als_cg_cluster && !profiler_attached && sa.a <= 2 && (sa.n * sa.r) % 2 == 0 && (KA * Rp) % 2 == 0 && sa.l <= size_t(c.num_sms)) {
		for (const Variant& v : variants) {
			if (sa.l % size_t(v.csz) != 0) continue;
			const size_t smem = smem_for(1, size_t(v.csz));
			if (smem > cap) continue;
			if (cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) { cudaGetLastError(); continue; }
			G = int(sa.l);
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(unsigned(G)); cfg.blockDim = dim3(CGP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = c.stream;
			cudaLaunchAttribute at[2];
			at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
			at[1].id = cudaLaunchAttributeClusterDimension;
			at[1].val.clusterDim.x = unsigned(v.csz); at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
			cfg.attrs = at; cfg.numAttrs = 2;
			int nclusters = 0;
			if (cudaOccupancyMaxActiveClusters(&nclusters, v.fn, &cfg) != cudaSuccess || nclusters * v.csz < G) { cudaGetLastError(); continue; }
			partial.resize(size_t(6) * G);
			g.partial = partial; g.nsl = 1;
			void* args[] = {&g};
			ProfScope prof("als_cg_kernel");
			if (cudaLaunchKernelExC(&cfg, v.fn, args) != cudaSuccess) { cudaGetLastError(); continue; }
			c.launches++;
			launched = true;
			break;
		}
	}
	if (!launched) {
		// one CTA per slice of the left bond when that many are co-resident, otherwise several slices per CTA
		size_t nsl = 1, smem = 0;
		for (;; ++nsl) {
			smem = smem_for(nsl, 1);
			if (smem > cap) { dfree(info); return false; }
			static size_t attr_smem = 0;
			if (smem > attr_smem) {
				XB_CUDA(cudaFuncSetAttribute(spd_cg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
				attr_smem = smem;
			}
			int per_sm = 0;
			XB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spd_cg_kernel<1>, CGP_THREADS, smem));
			if (per_sm < 1) { dfree(info); return false; }
			G = int((sa.l + nsl - 1) / nsl);
			if (size_t(G) <= size_t(c.num_sms) * per_sm) break;
		}
		partial.resize(size_t(6) * G);
		g.partial = partial; g.nsl = int(nsl);
		void* args[] = {&g};
		ProfScope prof("als_cg_kernel");
		XB_CUDA(cudaLaunchCooperativeKernel((void*)spd_cg_kernel<1>, dim3(unsigned(G)), dim3(CGP_THREADS), args, smem, c.stream));
		c.launches++;
	}
	unsigned int* h = reinterpret_cast<unsigned int*>(c.h_scratch + 8);
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, sc, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaMemcpyAsync(h, info, 10 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	dfree(info);
	rr_out = c.h_scratch[0]; rr0_out = c.h_scratch[1]; iterations = h[0]; reason = h[1];
	if (h[2] == 0xDEADu) throw Error(XB_ERR_CUDA, "persistent CG: a grid barrier timed out (internal scheduling error)");
	if (timing && h[0]) fprintf(stderr, "[cg] G=%d its=%u cycles/it: step1 %u steps2-3 %u barrier1 %u update %u barrier2 %u barrier3 %u\n", G, h[0],
	                            h[4] * 16 / h[0], h[5] * 16 / h[0], h[6] * 16 / h[0], h[7] * 16 / h[0], h[8] * 16 / h[0], h[9] * 16 / h[0]);
	return true;
}

// y = {L, A_1..A_s, R} applied to v, SPD environments: L(l, a, l'), A_p(a, m, n, b), R(r, b, r'), v(l', n_1..n_s, r' [, col])
// -> y(l, m_1..m_s, r [, col]).  One GEMM per factor, one reshuffle per operator core (als.cpp:383-401 un-contracted).
DT spd_env_apply(const DT& L, const std::vector<DT>& Acores, const DT& R, const DT& v, const GemmScatter* scatter = nullptr) {
	const int s = int(Acores.size());
	const bool batched = int(v.dims.size()) == s + 3;            // trailing column mode: (l, n.., r, col)
	DT t = dt_contract(L, {2}, v, {0});                           // (l, a, n_1..n_s, r'[, col])
	for (int p = 0; p < s; ++p) {
		// t modes: (l, m_1..m_p, a, n_{p+1}..n_s, r')  ->  contract (a, n_{p+1}) with A(a, m, n, b)
		{
			// the operator core as a (m b) x (a n) matrix applied to the middle of t in one pass over HBM (no reshuffle of t)
			const size_t P = t.dims[1 + p] * t.dims[2 + p], Q = Acores[p].dims[1] * Acores[p].dims[3];
			if (P <= 64 && Q <= 32) {
				XB_REQUIRE(Acores[p].dims[0] == t.dims[1 + p] && Acores[p].dims[2] == t.dims[2 + p], "Index dimensions do not coincide");
				size_t outer = 1, inner = 1;
				std::vector<size_t> nd;
				for (int i = 0; i <= p; ++i) { outer *= t.dims[i]; nd.push_back(t.dims[i]); }
				nd.push_back(Acores[p].dims[1]); nd.push_back(Acores[p].dims[3]);
				for (size_t i = 3 + p; i < t.dims.size(); ++i) { inner *= t.dims[i]; nd.push_back(t.dims[i]); }
				DT W = dt_permute(Acores[p], {1, 3, 0, 2});                   // (m, b | a, n)
				DT u = dt_alloc(nd);
				mid_apply(u.data(), t.p, W.p, outer, P, Q, inner);
				t = std::move(u);
				continue;
			}
		}
		DT u = dt_contract(t, {1 + p, 2 + p}, Acores[p], {0, 2});   // (l, m_1..m_p, n_{p+2}..n_s, r', m_{p+1}, b)
		const int nu = int(u.dims.size());
		std::vector<int> o;                                       // -> (l, m_1..m_{p+1}, b, n_{p+2}..n_s, r')
		for (int i = 0; i <= p; ++i) o.push_back(i);
		o.push_back(nu - 2); o.push_back(nu - 1);
		for (int i = p + 1; i < nu - 2; ++i) o.push_back(i);
		t = dt_permute(u, o);
	}
	if (scatter) {
		// last contraction with its row blocks written to the given destinations (peer receive buffers): y(rows | r) = t(rows | b r') R(r | b r')^T
		XB_REQUIRE(!batched, "scattered output: single vector only");
		const size_t K = R.dims[1] * R.dims[2], N = R.dims[0], M = t.size() / K;
		XB_REQUIRE(t.dims[1 + s] == R.dims[1] && t.dims[2 + s] == R.dims[2], "Index dimensions do not coincide");
		gemm_scatter(*scatter, N, M, N, 1.0, t.p, K, false, K, R.p, K, true);
		return DT();
	}
	DT y = dt_contract(t, {1 + s, 2 + s}, R, {1, 2});             // (l, m_1..m_s, [col,] r)
	if (!batched) return y;
	std::vector<int> o;                                           // (l, m.., col, r) -> (l, m.., r, col)
	for (int i = 0; i <= s; ++i) o.push_back(i);
	o.push_back(s + 2); o.push_back(s + 1);
	return dt_permute(y, o);
}

// ---- the algorithm ---------------------------------------------------------------------------------------------------
struct Als {
	const xb_tt* A; xb_tt* x; const xb_tt* b;
	xb_als_options opt;
	size_t d, sites;
	bool spd_like;          // SPD environments (assumeSPD or no operator)
	std::vector<DT> opL, opR, rhL, rhR;
	size_t first = 0, last = 0, cur = 0;
	bool increasing = true;
	std::vector<size_t> target_rank;
	double normB = 0.0;
	size_t cg_iterations = 0, direct_solves = 0;
	bool used_cg = false;
	SpdSiteApply* site_apply = nullptr;     // set while a one-site SPD local problem is being solved by CG
	unsigned long long graph_nodes = 0;     // kernels inside one captured chunk of CG iterations

	DT xcore(size_t i) const { return dt_view(x->core[i], {x->rank[i], x->dim_m[i], x->rank[i + 1]}); }
	DT bcore(size_t i) const { return dt_view(b->core[i], {b->rank[i], b->dim_m[i], b->rank[i + 1]}); }
	DT acore(size_t i) const { return dt_view(A->core[i], {A->rank[i], A->dim_m[i], A->dim_n[i], A->rank[i + 1]}); }

	// localOperatorSlice / localRhsSlice contracted with an environment (als.cpp:184-215, :238-251)
	DT op_left(const DT& env, size_t i) const {
		if (opt.assume_spd) {   // env(r1,r2,r3) x(r1,n1,c1) A(r2,n1,n2,c2) x(r3,n2,c3) -> (c1,c2,c3)
			DT t = dt_contract(env, {0}, xcore(i), {0});                 // (r2,r3,n1,c1)
			t = dt_contract(t, {0, 2}, acore(i), {0, 1});                // (r3,c1,n2,c2)
			t = dt_contract(t, {0, 2}, xcore(i), {0, 1});                // (c1,c2,c3)
			return t;
		}
		// env(r1,r2,r3,r4) x(r1,n1,c1) A(r2,n2,n1,c2) A(r3,n2,n3,c3) x(r4,n3,c4) -> (c1,c2,c3,c4)
		DT t = dt_contract(env, {0}, xcore(i), {0});                     // (r2,r3,r4,n1,c1)
		t = dt_contract(t, {0, 3}, acore(i), {0, 2});                    // (r3,r4,c1,n2,c2)
		t = dt_contract(t, {0, 3}, acore(i), {0, 1});                    // (r4,c1,c2,n3,c3)
		t = dt_contract(t, {0, 3}, xcore(i), {0, 1});                    // (c1,c2,c3,c4)
		return t;
	}
	DT op_right(const DT& env, size_t i) const {
		if (opt.assume_spd) {   // x(r1,n1,c1) A(r2,n1,n2,c2) x(r3,n2,c3) env(c1,c2,c3) -> (r1,r2,r3)
			DT t = dt_contract(xcore(i), {2}, env, {0});                 // (r1,n1,c2,c3)
			t = dt_contract(acore(i), {1, 3}, t, {1, 2});                // (r2,n2,r1,c3)
			t = dt_contract(t, {1, 3}, xcore(i), {1, 2});                // (r2,r1,r3)
			return dt_permute(t, {1, 0, 2});
		}
		DT t = dt_contract(xcore(i), {2}, env, {0});                     // (r1,n1,c2,c3,c4)
		t = dt_contract(acore(i), {2, 3}, t, {1, 2});                    // (r2,n2,r1,c3,c4)
		t = dt_contract(acore(i), {1, 3}, t, {1, 3});                    // (r3,n3,r2,r1,c4)
		t = dt_contract(t, {1, 4}, xcore(i), {1, 2});                    // (r3,r2,r1,r4)
		return dt_permute(t, {2, 1, 0, 3});
	}
	DT rhs_left(const DT& env, size_t i) const {
		if (spd_like) {        // env(r1,r2) b(r1,n1,c1) x(r2,n1,c2) -> (c1,c2)
			DT t = dt_contract(env, {0}, bcore(i), {0});                 // (r2,n1,c1)
			return dt_contract(t, {0, 1}, xcore(i), {0, 1});            // (c1,c2)
		}
		// env(r1,r2,r3) b(r1,n1,c1) A(r2,n1,n2,c2) x(r3,n2,c3) -> (c1,c2,c3)
		DT t = dt_contract(env, {0}, bcore(i), {0});                     // (r2,r3,n1,c1)
		t = dt_contract(t, {0, 2}, acore(i), {0, 1});                    // (r3,c1,n2,c2)
		return dt_contract(t, {0, 2}, xcore(i), {0, 1});                // (c1,c2,c3)
	}
	DT rhs_right(const DT& env, size_t i) const {
		if (spd_like) {        // b(r1,n1,c1) x(r2,n1,c2) env(c1,c2) -> (r1,r2)
			DT t = dt_contract(bcore(i), {2}, env, {0});                 // (r1,n1,c2)
			return dt_contract(t, {1, 2}, xcore(i), {1, 2});            // (r1,r2)
		}
		DT t = dt_contract(bcore(i), {2}, env, {0});                     // (r1,n1,c2,c3)
		t = dt_contract(acore(i), {1, 3}, t, {1, 2});                    // (r2,n2,r1,c3)
		t = dt_contract(t, {1, 3}, xcore(i), {1, 2});                    // (r2,r1,r3)
		return dt_permute(t, {1, 0, 2});
	}

	// ---- prepare_x_for_als (als.cpp:105-182) ---------------------------------------------------------------------------
	void prepare_x(bool canon_end, size_t core_end) {
		first = 0;
		size_t dim_prod = 1;
		while (first + 1 < d) {
			const size_t n_loc = x->dim_m[first], new_prod = dim_prod * n_loc;
			if (x->rank[first + 1] < new_prod) break;
			// component(first+1) = component(first) (as matrix) * component(first+1); component(first) = identity
			const size_t rows = x->rank[first] * n_loc, r = x->rank[first + 1], cols = x->dim_m[first + 1] * x->rank[first + 2];
			DBuf nxt(rows * cols);
			gemm(nxt, cols, rows, cols, 1.0, x->core[first], r, false, r, x->core[first + 1], cols, false, 0.0);
			x->core[first + 1] = std::move(nxt);
			x->core[first].resize(dim_prod * n_loc * new_prod);
			set_identity(x->core[first], new_prod, new_prod, new_prod);
			x->rank[first + 1] = new_prod;
			first += 1; dim_prod = new_prod;
		}
		last = d;
		dim_prod = 1;
		while (last > first + sites) {
			const size_t n_loc = x->dim_m[last - 1], new_prod = dim_prod * n_loc;
			if (x->rank[last - 1] < new_prod) break;
			const size_t rows = x->rank[last - 2] * x->dim_m[last - 2], r = x->rank[last - 1], cols = n_loc * x->rank[last];
			DBuf prv(rows * cols);
			gemm(prv, cols, rows, cols, 1.0, x->core[last - 2], r, false, r, x->core[last - 1], cols, false, 0.0);
			x->core[last - 2] = std::move(prv);
			x->core[last - 1].resize(new_prod * n_loc * dim_prod);
			set_identity(x->core[last - 1], new_prod, new_prod, new_prod);
			x->rank[last - 1] = new_prod;
			last -= 1; dim_prod = new_prod;
		}
		if (first > 0 || last < d) x->canonicalized = false;      // set_component on non-core indices (ttNetwork.cpp:491)
		if (canon_end && core_end < first) {
			x->canonicalized = true; x->core_position = first;     // assume_core_position (:171-172)
		} else {
			if (canon_end && core_end >= last) { x->canonicalized = true; x->core_position = last - 1; }
			move_core(x, first, true);                             // :178
		}
	}

	// ---- energy functional (als.cpp:255-320) ---------------------------------------------------------------------------
	double energy() const {
		if (A && opt.assume_spd) {                                   // |0.5 <x,Ax> - <x,b>| from the stacks (:265-278)
			DT xAx = dt_view(opL.back().p, opL.back().dims), bx = dt_view(rhL.back().p, rhL.back().dims);
			for (size_t i = 0; i < sites; ++i) { xAx = op_left(xAx, cur + i); bx = rhs_left(bx, cur + i); }
			return std::fabs(0.5 * dt_dot(xAx, opR.back()) - dt_dot(bx, rhR.back()));
		}
		if (A) {                                                     // residual from the stacks (:282-296)
			DT xAtAx = dt_view(opL.back().p, opL.back().dims), bAx = dt_view(rhL.back().p, rhL.back().dims);
			for (size_t i = 0; i < sites; ++i) { xAtAx = op_left(xAtAx, cur + i); bAx = rhs_left(bAx, cur + i); }
			const double v = dt_dot(xAtAx, opR.back()) - 2.0 * dt_dot(bAx, rhR.back());
			return std::sqrt(std::max(0.0, v + normB * normB)) / normB;
		}
		DT bx = dt_view(rhL.back().p, rhL.back().dims);                // (:305-316)
		for (size_t i = 0; i < sites; ++i) bx = rhs_left(bx, cur + i);
		DT xc = xcore(cur);
		return 0.5 * dt_dot(xc, xc) - dt_dot(bx, rhR.back());
	}

	// ---- local problem -------------------------------------------------------------------------------------------------
	// y = ATilde * v with ATilde = {opL.back(), A_cur .. A_cur+sites-1, opR.back()} un-contracted (als.cpp:383-401);
	// v, y have modes (l, n_1..n_s, r).
	DT local_apply(const DT& v) const {
		const DT& L = opL.back(); const DT& R = opR.back();
		const int s = int(sites);
		const bool batched = int(v.dims.size()) == s + 3;            // trailing column mode: (l, n.., r, col)
		auto finish = [&](DT y) -> DT {                              // (l, m.., col, r) -> (l, m.., r, col)
			if (!batched) return y;
			std::vector<int> o;
			for (int i = 0; i <= s; ++i) o.push_back(i);
			o.push_back(s + 2); o.push_back(s + 1);
			return dt_permute(y, o);
		};
		if (opt.assume_spd) {
			std::vector<DT> ac;
			for (int p = 0; p < s; ++p) ac.push_back(acore(cur + p));
			return spd_env_apply(L, ac, R, v);
		}
		// general: L(l, a1, a2, l'), per site A(a2, k, n, b2) then A(a1, k, m, b1); R(r, b1, b2, r')
		DT t = dt_contract(L, {3}, v, {0});                               // (l, a1, a2, n_1..n_s, r')
		for (int p = 0; p < s; ++p) {
			// t: (l, m_1..m_p, a1, a2, n_{p+1}.., r')
			DT u = dt_contract(t, {2 + p, 3 + p}, acore(cur + p), {0, 2});       // (l, m.., a1, n_{p+2}.., r', k, b2)
			int nu = int(u.dims.size());
			DT w = dt_contract(u, {1 + p, nu - 2}, acore(cur + p), {0, 1});     // (l, m.., n_{p+2}.., r', b2, m_{p+1}, b1)
			nu = int(w.dims.size());
			std::vector<int> o;                                       // -> (l, m_1..m_{p+1}, b1, b2, n_{p+2}.., r')
			for (int i = 0; i <= p; ++i) o.push_back(i);
			o.push_back(nu - 2); o.push_back(nu - 1); o.push_back(nu - 3);
			for (int i = p + 1; i < nu - 3; ++i) o.push_back(i);
			t = dt_permute(w, o);
		}
		return finish(dt_contract(t, {1 + s, 2 + s, 3 + s}, R, {1, 2, 3}));   // (l, m_1..m_s, r)
	}

	DT local_rhs() const {                                                // construct_local_RHS (als.cpp:404-423)
		const DT& L = rhL.back(); const DT& R = rhR.back();
		if (spd_like) {
			DT t = dt_contract(L, {0}, bcore(cur), {0});                   // (l, n_1, rb')
			for (size_t p = 1; p < sites; ++p) t = dt_contract(t, {int(t.dims.size()) - 1}, bcore(cur + p), {0});
			return dt_contract(t, {int(t.dims.size()) - 1}, R, {0});      // (l, n.., r)
		}
		// L(rb, a, l), b(rb, k, rb'), A(a, k, m, a'), R(rb', a', r)
		DT t = dt_permute(L, {2, 0, 1});                                   // (l, rb, a)
		for (size_t p = 0; p < sites; ++p) {
			const int nt = int(t.dims.size());
			DT u = dt_contract(t, {nt - 2}, bcore(cur + p), {0});          // (l, m.., a, k, rb')
			const int nu = int(u.dims.size());
			t = dt_contract(u, {nu - 3, nu - 2}, acore(cur + p), {0, 1}); // (l, m.., rb', m_p, a')
			const int nw = int(t.dims.size());
			std::vector<int> o;
			for (int i = 0; i < nw - 3; ++i) o.push_back(i);
			o.push_back(nw - 2); o.push_back(nw - 3); o.push_back(nw - 1);   // (l, m.., m_p, rb', a')
			t = dt_permute(t, o);
		}
		const int nt = int(t.dims.size());
		return dt_contract(t, {nt - 2, nt - 1}, R, {0, 1});               // (l, m.., r)
	}

	// Reference semantics for small local problems: densify the local operator (apply it to the identity, so it is by
	// construction the same operator the matrix-free path uses) and solve directly with the reference's dispatch
	// (symmetric + definite diagonal -> Cholesky, else LU; blasLapackWrapper.cpp:542-651).
	DT local_solve_direct(const DT& rhs) {
		const size_t n = rhs.size();
		std::vector<size_t> vd = rhs.dims; vd.push_back(n);
		DT I = dt_alloc(vd);
		set_identity(I.data(), n, n, n);
		DT Aloc = local_apply(I);                                     // n x n, row-major
		DT sol = dt_copy(rhs);
		bool symmetric = false, definite = false;
		probe_symmetry(Aloc.p, n, symmetric, definite);
		bool done = false;
		if (symmetric && definite) {
			DT Ac = dt_copy(Aloc);
			if (cholesky_solve(Ac.data(), sol.data(), n, 1)) done = true;
			else copy(sol.data(), rhs.p, n);
		}
		if (!done) lu_solve(Aloc.data(), sol.data(), n, 1);
		direct_solves += 1;
		return sol;
	}

	// conjugate gradients on the matrix-free local operator, warm start v0.  The recurrence residual drifts from the true
	// one by ~eps * ||r_0|| (a random start has ||r_0|| ~ 1e11 ||b||), so after the loop the true residual is recomputed and
	// CG is restarted from the current iterate until it meets the tolerance (residual replacement).
	DT local_solve_cg(const DT& rhs, const DT& v0) {
		const size_t n = rhs.size();
		const double tol = opt.local_tolerance > 0 ? opt.local_tolerance : 1e-15;
		const size_t max_it = opt.local_max_iterations ? opt.local_max_iterations : std::min<size_t>(std::max<size_t>(4 * n, 64), 4000);
		DT xv = dt_copy(v0);
		xv.dims = rhs.dims;
		DBuf sc(4);
		Context& c = ctx();
		SpdSiteApply fast_apply;
		DT qbuf;
		site_apply = nullptr;
		if (sites == 1 && opt.assume_spd && A && v0.dims.size() == 3) {
			fast_apply.prepare(opL.back(), acore(cur), opR.back());
			site_apply = &fast_apply;
			qbuf = dt_alloc(rhs.dims);
		}
		struct Reset { SpdSiteApply*& p; ~Reset() { p = nullptr; } } reset{site_apply};
		const bool fused_update = n <= CG_FUSED_MAX;
		const double bnorm2 = dt_dot(rhs, rhs);
		if (bnorm2 == 0.0) { fill(xv.data(), 0.0, n); return xv; }
		const double target = tol * tol * bnorm2;
		const unsigned grid = vec_grid(n);
		size_t it = 0;
		double rr = 0.0, rr0 = -1.0;
		if (site_apply) {
			// one cooperative launch for the whole solve: true residual, CG, re-anchoring at the true residual (spd_cg_kernel)
			DT pw = dt_alloc(rhs.dims);
			size_t done = 0; unsigned reason = 0;
			if (spd_cg_persistent(*site_apply, xv.data(), rhs.p, pw.data(), sc.p, max_it, target, bnorm2, done, reason, rr, rr0)) {
				if (reason == 4 || !(rr == rr)) throw Error(XB_ERR_NUMERIC, "local CG produced NaN (operator not positive definite?)");
				cg_iterations += done;
				if (getenv("XB_DEBUG_ALS")) fprintf(stderr, "[als] site %zu n=%zu cg its=%zu reason=%u rel res=%.3e (start %.3e) [one launch]\n", cur, n, done, reason, std::sqrt(rr / bnorm2), std::sqrt(rr0 / bnorm2));
				return xv;
			}
		}
		for (int restart = 0; restart < 6 && it < max_it; ++restart) {
			DT r = dt_copy(rhs);
			if (site_apply) { site_apply->apply(xv.p, qbuf.data()); axpy(r.data(), -1.0, qbuf.p, n); }
			else { DT Ax = local_apply(xv); axpy(r.data(), -1.0, Ax.p, n); }      // true residual
			dot_dev(sc.p + 0, r.p, r.p, n);
			rr = read_scalar(sc.p + 0);
			if (rr0 < 0.0) rr0 = rr;
			if (rr <= target) break;
			if (restart == 0 && rr > 1e4 * bnorm2) {                         // useless warm start: begin from zero instead
				fill(xv.data(), 0.0, n);
				copy(r.data(), rhs.p, n);
				dot_dev(sc.p + 0, r.p, r.p, n);
				rr = bnorm2;
			}
			DT p = dt_copy(r);
			double best = rr; size_t since_best = 0;
			const double rr_start = rr;
			// The iteration is a fixed launch sequence on fixed buffers (scalars stay on the device): chunks of 8 iterations
			// are captured once per (site, restart) into a CUDA graph and replayed — the host enqueue of 32 small kernels
			// costs more than their execution.
			struct GraphHolder { cudaGraph_t g = nullptr; cudaGraphExec_t e = nullptr;
				~GraphHolder() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); } } graph;
			const bool use_graph = site_apply && fused_update && c.als_graph && !c.profile;
			while (rr > target && it < max_it) {
				const size_t chunk = std::min<size_t>(8, max_it - it);
				if (use_graph && chunk == 8) {
					if (!graph.e) {
						const unsigned long long l0 = c.launches;
						XB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
						for (size_t j = 0; j < chunk; ++j) {
							site_apply->apply(p.p, qbuf.data());
							cg_fused_kernel<<<CG_CS, 1024, 0, c.stream>>>(xv.data(), r.data(), p.data(), qbuf.p, sc.p, int(n));
							c.launches++;
						}
						XB_CUDA(cudaStreamEndCapture(c.stream, &graph.g));
						XB_CUDA(cudaGraphInstantiate(&graph.e, graph.g, 0));
						graph_nodes = c.launches - l0;
						c.launches = l0;
					}
					XB_CUDA(cudaGraphLaunch(graph.e, c.stream));
					c.launches += graph_nodes;
				} else
				for (size_t j = 0; j < chunk; ++j) {
					DT q;
					const double* qp;
					if (site_apply) { site_apply->apply(p.p, qbuf.data()); qp = qbuf.p; }
					else { q = local_apply(p); qp = q.p; }
					if (fused_update) {
						cg_fused_kernel<<<CG_CS, 1024, 0, c.stream>>>(xv.data(), r.data(), p.data(), qp, sc.p, int(n));
						XB_LAUNCH_CHECK();
						continue;
					}
					dot_dev(sc.p + 1, p.p, qp, n);
					cg_step1_kernel<<<grid, 256, 0, c.stream>>>(xv.data(), r.data(), p.p, qp, sc.p, n);
					XB_LAUNCH_CHECK();
					dot_dev(sc.p + 2, r.p, r.p, n);
					cg_step2_kernel<<<grid, 256, 0, c.stream>>>(p.data(), r.p, sc.p, n);
					XB_LAUNCH_CHECK();
					cg_shift_kernel<<<1, 32, 0, c.stream>>>(sc.p);
					XB_LAUNCH_CHECK();
				}
				it += chunk;
				rr = read_scalar(sc.p + 0);
				if (!(rr == rr)) throw Error(XB_ERR_NUMERIC, "local CG produced NaN (operator not positive definite?)");
				if (rr < 0.5 * best) { best = rr; since_best = 0; } else { since_best += chunk; if (since_best >= 48) break; }   // stagnation at the rounding floor
				if (rr < 1e-20 * rr_start) break;                            // beyond what the recurrence can resolve: re-anchor
			}
		}
		cg_iterations += it;
		if (getenv("XB_DEBUG_ALS")) fprintf(stderr, "[als] site %zu n=%zu cg its=%zu rel res=%.3e (start %.3e)\n", cur, n, it, std::sqrt(rr / bnorm2), std::sqrt(rr0 / bnorm2));
		return xv;
	}

	void local_step() {
		if (!A) {   // x_core = rhs env contraction (als.cpp:541-545); sites == 1 only
			XB_REQUIRE(sites == 1, "approximation dmrg not implemented yet");     // als.cpp:543
			DT sol = local_rhs();
			x->core[cur] = std::move(sol.own);
			return;
		}
		DT rhs = local_rhs();
		// warm start: the current component(s) contracted to one tensor (l, n_1..n_s, r)
		DT v0 = dt_view(x->core[cur], {x->rank[cur], x->dim_m[cur], x->rank[cur + 1]});
		if (opt.local_solver == 1) {
			// ALSVariant::ASD_solver (als.cpp:73-103): x += alpha * grad with grad = b~ - A~ x and the step the reference takes:
			// SPD: <g,g>/<g,A~g>; otherwise g <- A~^T g first and alpha = ||g|| / ||A~ g|| (norms, not their squares, :97)
			XB_REQUIRE(sites == 1, "ASD only defined for single site alternation at the moment");   // :78
			const size_t n = rhs.size();
			DT g = dt_copy(rhs);
			{ DT Ax = local_apply(v0); axpy(g.data(), -1.0, Ax.p, n); }
			double alpha;
			if (opt.assume_spd) {
				DT Ag = local_apply(g);
				alpha = dt_dot(g, g) / dt_dot(g, Ag);
			} else {
				g = local_apply(g);                                      // the projected A^T A is symmetric
				DT Ag = local_apply(g);
				alpha = std::sqrt(dt_dot(g, g)) / std::sqrt(dt_dot(Ag, Ag));
			}
			if (alpha == alpha && std::isfinite(alpha)) axpy(x->core[cur].p, alpha, g.p, n);   // zero gradient: nothing to do
			return;
		}
		DT v0own;
		if (sites == 2) {
			v0own = dt_contract(v0, {2}, xcore(cur + 1), {0});
			XB_REQUIRE(v0own.size() == rhs.size(), "internal: two-site warm start has the wrong size");
			v0 = dt_view(v0own.p, v0own.dims);
		}
		const bool direct = rhs.size() <= size_t(ctx().als_direct_max);
		DT sol = direct ? local_solve_direct(rhs) : local_solve_cg(rhs, v0);
		if (sites == 1) {
			x->core[cur] = std::move(sol.own);
			return;
		}
		// two sites: split by SVD with the ranks x had on entry (als.cpp:51-70), Sigma pushed in sweep direction
		const size_t l = x->rank[cur], n1 = x->dim_m[cur], n2 = x->dim_m[cur + 1], r = x->rank[cur + 2];
		Svd svd;
		svd.polish = ctx().tt_svd_polish;
		svd.factor(sol.p, l * n1, n2 * r);
		// the reference truncates the split with eps = EPSILON (als.cpp:55,:65); below the accuracy of an iterative
		// local solve singular values are noise, so the CG path cuts at its own tolerance instead
		const double tol_cg = opt.local_tolerance > 0 ? opt.local_tolerance : 1e-15;
		const size_t k = truncation_rank(svd.S, target_rank[cur], direct ? EPSILON : std::max(EPSILON, 10.0 * tol_cg));
		DBuf U(l * n1 * k), Vt(k * n2 * r);
		svd.extract(U, Vt, k, !increasing, increasing, nullptr);
		x->core[cur] = std::move(U);
		x->core[cur + 1] = std::move(Vt);
		x->rank[cur + 1] = k;
		x->canonicalized = false;                                      // set_component on a non-core index (ttNetwork.cpp:491)
	}

	double run() {
		d = x->d; sites = opt.sites;
		XB_REQUIRE(sites == 1 || sites == 2, "only sites = 1 (ALS) and sites = 2 (DMRG) are implemented");
		XB_REQUIRE(d >= sites, "TT too short for this number of sites");
		spd_like = opt.assume_spd || !A;
		target_rank.assign(x->rank.begin() + 1, x->rank.end() - 1);   // als.cpp:324
		normB = tt_frob_norm(b);
		const bool canon_end = x->canonicalized; const size_t core_end = x->core_position;
		prepare_x(canon_end, core_end);
		// prepare_stacks (als.cpp:217-253)
		const std::vector<size_t> one_op = spd_like ? std::vector<size_t>{1, 1, 1} : std::vector<size_t>{1, 1, 1, 1};
		const std::vector<size_t> one_rhs = spd_like ? std::vector<size_t>{1, 1} : std::vector<size_t>{1, 1, 1};
		opL.push_back(dt_ones(one_op)); opR.push_back(dt_ones(one_op));
		rhL.push_back(dt_ones(one_rhs)); rhR.push_back(dt_ones(one_rhs));
		for (size_t i = d - 1; i > first + sites - 1; --i) {
			if (A) opR.push_back(op_right(opR.back(), i));
			rhR.push_back(rhs_right(rhR.back(), i));
		}
		for (size_t i = 0; i < first; ++i) {
			if (A) opL.push_back(op_left(opL.back(), i));
			rhL.push_back(rhs_left(rhL.back(), i));
		}
		cur = first;
		increasing = true;
		double last_e2 = 1e102, last_e = 1e101, e = energy();
		size_t half_sweeps = 0;
		for (;;) {
			{ ProfScope prof("als_local_step"); local_step(); }
			// check_for_end_of_sweep (als.cpp:426-475)
			if ((!increasing && cur == first) || (increasing && cur == last - sites)) {
				half_sweeps += 1;
				last_e2 = last_e; last_e = e; { ProfScope prof("als_energy"); e = energy(); }
				if (half_sweeps == opt.num_half_sweeps || std::fabs(last_e - e) < opt.convergence_epsilon ||
				    std::fabs(last_e2 - e) < opt.convergence_epsilon || last - first <= sites) {
					if (canon_end && opt.preserve_core_position) move_core(x, core_end, true);
					return e;
				}
				increasing = !increasing;
			}
			// move_to_next_index (als.cpp:340-380)
			ProfScope prof_move("als_move_to_next");
			if (increasing) {
				if (sites == 1) move_core(x, cur + 1, true);
				if (A) { opR.pop_back(); opL.push_back(op_left(opL.back(), cur)); }
				rhR.pop_back(); rhL.push_back(rhs_left(rhL.back(), cur));
				cur += 1;
			} else {
				if (sites == 1) move_core(x, cur - 1, true);
				const size_t pos = cur + sites - 1;                    // the reference pushes `cur` here (defect, see header)
				if (A) { opL.pop_back(); opR.push_back(op_right(opR.back(), pos)); }
				rhL.pop_back(); rhR.push_back(rhs_right(rhR.back(), pos));
				cur -= 1;
			}
		}
	}
};

} // namespace

extern "C" {

xb_status xb_als_default_options(xb_als_options* opt, uint32_t sites, int assume_spd) {
	return guard([&] {
		XB_REQUIRE(opt, "null");
		XB_REQUIRE(sites > 0, "sites must be positive");          // als.h:141
		opt->sites = sites;
		opt->assume_spd = assume_spd;
		opt->num_half_sweeps = 0;
		opt->convergence_epsilon = 1e-6;                           // als.h:137
		opt->preserve_core_position = 1;                           // als.h:138
		opt->local_tolerance = 0.0;
		opt->local_max_iterations = 0;
		opt->local_solver = 0;
	});
}

// Bond-split environment application (BASELINE config 4): the contribution of the right-bond slab r' in [begin, end)
// to y = {L, A_1..A_s, R} v.  The chain is linear in v, so slabs of r' give partial results of full size whose sum over
// the slabs (one NCCL all-reduce across the GPUs that own them) is the full application.  Device pointers.
xb_status xb_env_apply(double* y, const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims,
                       size_t sites, const double* R, size_t r, size_t a_right, const double* v, size_t slab_begin, size_t slab_end) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(y && L && A_cores && A_dims && R && v, "null");
		XB_REQUIRE(sites >= 1 && sites <= 4, "1 to 4 sites");
		XB_REQUIRE(slab_begin < slab_end && slab_end <= r, "illegal bond slab");
		std::vector<DT> ac;
		std::vector<size_t> vd = {l};
		for (size_t p = 0; p < sites; ++p) {
			const size_t* d = A_dims + 4 * p;
			XB_REQUIRE(d[0] == (p == 0 ? a_left : A_dims[4 * (p - 1) + 3]), "operator bond dimensions do not coincide");
			ac.push_back(dt_view(A_cores[p], {d[0], d[1], d[2], d[3]}));
			vd.push_back(d[2]);
		}
		XB_REQUIRE(A_dims[4 * (sites - 1) + 3] == a_right, "operator bond dimensions do not coincide");
		const size_t slab = slab_end - slab_begin;
		size_t rows_v = l;
		for (size_t p = 0; p < sites; ++p) rows_v *= A_dims[4 * p + 2];
		DT Lv = dt_view(L, {l, a_left, l});
		DT res;
		if (slab == r) {
			vd.push_back(r);
			res = spd_env_apply(Lv, ac, dt_view(R, {r, a_right, r}), dt_view(v, vd));
		} else {
			// pack the slab of the last mode of v and R (strided -> contiguous), then the same chain
			vd.push_back(slab);
			DT vs = dt_alloc(vd);
			copy2d(vs.data(), slab, v + slab_begin, r, rows_v, slab);
			DT Rs = dt_alloc({r, a_right, slab});
			copy2d(Rs.data(), slab, R + slab_begin, r, r * a_right, slab);
			res = spd_env_apply(Lv, ac, Rs, vs);
		}
		copy(y, res.p, res.size());
	});
}

// ---- bond-split application fused with its reduction over peer memory (BASELINE config 4, N GPUs of one box) -----------------
// Every rank owns a symmetric buffer that all ranks map (CUDA IPC):  [ flags 256 B | receive slots: world x (M/world x N) | y: M x N ].
//  1. the rank contracts its slab of the right bond; the epilogue of the last GEMM writes row block p of the partial result straight
//     into rank p's receive slot [rank] (remote stores over NVLink while the remaining tiles are multiplied), then a one-warp kernel
//     fences at system scope and bumps flag 0 of every peer;
//  2. after a one-thread wait for world - 1 bumps the reduce kernel sums the slots of its own row block in rank order (deterministic, identical on
//     all ranks) and writes the sum into the y area of EVERY rank; the last CTA to finish fences and bumps flag 1 everywhere;
//  3. a one-warp kernel waits for world bumps of flag 1: the y area of this rank is complete.
// Flags count monotonically (epoch = number of the call, same on all ranks), so nothing is ever reset; all waits are bounded.
constexpr size_t PEER_FLAG_BYTES = 256;
__global__ void peer_signal_kernel(unsigned int* const* flags, const int rank, const int world, const int which) {
	__threadfence_system();
	if (int(threadIdx.x) < world && int(threadIdx.x) != rank) atomicAdd_system(flags[threadIdx.x] + which, 1u);
}
struct PeerPtrs { unsigned int* flags[8]; double* y[8]; };
__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerPtrs pp, const double* __restrict__ recv, const size_t block_elems, const int rank,
                                                           const int world, const unsigned int epoch) {
	unsigned int* myflags = pp.flags[rank];
	// a wait on this buffer has timed out (a peer lagged or died): the receive slots may be incomplete — sum nothing, publish
	// nothing, bump nothing; the peers' waits then time out as well and xb_peer_buffer_check() reports it on every rank
	if (*((volatile unsigned int*)(myflags + 3)) == 0xDEADu) return;
	const size_t nv = block_elems / 2;                   // block_elems is even (checked on the host)
	const double2* rv = reinterpret_cast<const double2*>(recv);
	for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nv; e += (size_t)gridDim.x * blockDim.x) {
		double2 acc = __ldcg(rv + e);
		for (int s = 1; s < world; ++s) { const double2 t = __ldcg(rv + (size_t)s * nv + e); acc.x += t.x; acc.y += t.y; }
		for (int p = 0; p < world; ++p) reinterpret_cast<double2*>(pp.y[p] + (size_t)rank * block_elems)[e] = acc;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned int done = atomicInc(myflags + 2, gridDim.x - 1);
		if (done == gridDim.x - 1) {
			__threadfence_system();
			for (int p = 0; p < world; ++p) atomicAdd_system(pp.flags[p] + 1, 1u);
		}
	}
}
// one thread of one CTA polls (the waiting must not occupy the SMs the peers' kernels may need when "ranks" share a device)
__global__ void peer_wait_kernel(unsigned int* myflags, const int which, const unsigned int expected, const unsigned int max_spins) {
	if (threadIdx.x == 0) {
		unsigned int spins = 0;
		while (*((volatile unsigned int*)(myflags + which)) < expected && ++spins < max_spins) {}
		if (spins >= max_spins) myflags[3] = 0xDEADu;
		__threadfence_system();
	}
}

xb_status xb_peer_buffer_create(size_t bytes, void** dptr, unsigned char* handle64) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(dptr && handle64 && bytes > 0, "null");
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
		XB_CUDA(cudaMalloc(dptr, bytes));
		XB_CUDA(cudaMemset(*dptr, 0, bytes));
		cudaIpcMemHandle_t h;
		XB_CUDA(cudaIpcGetMemHandle(&h, *dptr));
		std::memcpy(handle64, &h, 64);
	});
}
xb_status xb_peer_buffer_open(const unsigned char* handle64, void** dptr) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(dptr && handle64, "null");
		cudaIpcMemHandle_t h;
		std::memcpy(&h, handle64, 64);
		XB_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
	});
}
// The waits of xb_env_apply_fused are bounded (2^28 polls); one that gives up poisons word 3 of the rank's flag block, the reduce
// kernel then publishes nothing.  This is the host-side check: synchronises the calling worker's stream and reports the poison.
xb_status xb_peer_buffer_check(void* dptr) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(dptr, "null");
		Context& c = ctx();
		unsigned int* h = reinterpret_cast<unsigned int*>(c.h_scratch);
		XB_CUDA(cudaMemcpyAsync(h, static_cast<unsigned int*>(dptr) + 3, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
		XB_CUDA(cudaStreamSynchronize(c.stream));
		if (*h == 0xDEADu) throw Error(XB_ERR_CUDA, "bond-split exchange: a wait for the peers timed out (a rank lagged by more than the bounded wait or died); the result is not valid");
	});
}
xb_status xb_peer_buffer_close(void* dptr) { return guard([&] { if (dptr) XB_CUDA(cudaIpcCloseMemHandle(dptr)); }); }
xb_status xb_peer_buffer_destroy(void* dptr) { return guard([&] { if (dptr) XB_CUDA(cudaFree(dptr)); }); }
xb_status xb_peer_buffer_bytes(size_t rows, size_t cols, int world, size_t* bytes) {
	return guard([&] {
		XB_REQUIRE(bytes && world >= 1 && world <= 8, "1 to 8 ranks");
		XB_REQUIRE(rows % size_t(world) == 0 && (rows / world * cols) % 2 == 0, "the rows of the result must split evenly over the ranks");
		*bytes = PEER_FLAG_BYTES + 2 * rows * cols * sizeof(double);
	});
}

xb_status xb_env_apply_fused(const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims, size_t sites,
                             const double* R, size_t r, size_t a_right, const double* v, size_t slab_begin, size_t slab_end,
                             int rank, int world, void* const* sym, unsigned int epoch, double** y_out) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(L && A_cores && A_dims && R && v && sym && y_out, "null");
		XB_REQUIRE(sites >= 1 && sites <= 4, "1 to 4 sites");
		XB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && epoch >= 1, "illegal rank / world / epoch");
		XB_REQUIRE(slab_begin < slab_end && slab_end <= r, "illegal bond slab");
		std::vector<DT> ac;
		std::vector<size_t> vd = {l};
		size_t rows = l;
		for (size_t p = 0; p < sites; ++p) {
			const size_t* d = A_dims + 4 * p;
			XB_REQUIRE(d[0] == (p == 0 ? a_left : A_dims[4 * (p - 1) + 3]), "operator bond dimensions do not coincide");
			ac.push_back(dt_view(A_cores[p], {d[0], d[1], d[2], d[3]}));
			vd.push_back(d[2]);
			rows *= d[1];
		}
		XB_REQUIRE(A_dims[4 * (sites - 1) + 3] == a_right, "operator bond dimensions do not coincide");
		XB_REQUIRE(rows % size_t(world) == 0 && (rows / world * r) % 2 == 0, "the rows of the result must split evenly over the ranks");
		const size_t rpb = rows / world, block_elems = rpb * r;
		size_t rows_v = l;
		for (size_t p = 0; p < sites; ++p) rows_v *= A_dims[4 * p + 2];
		auto flags_of = [&](int p) { return reinterpret_cast<unsigned int*>(sym[p]); };
		auto recv_of = [&](int p) { return reinterpret_cast<double*>(static_cast<char*>(sym[p]) + PEER_FLAG_BYTES); };
		auto y_of = [&](int p) { return recv_of(p) + rows * r; };
		GemmScatter sc;
		for (int i = 0; i < 8; ++i) sc.blk[i] = (i < world) ? recv_of(i) + size_t(rank) * block_elems : nullptr;
		sc.rows_per_block = rpb;
		const size_t slab = slab_end - slab_begin;
		DT Lv = dt_view(L, {l, a_left, l});
		if (slab == r) {
			vd.push_back(r);
			spd_env_apply(Lv, ac, dt_view(R, {r, a_right, r}), dt_view(v, vd), &sc);
		} else {
			vd.push_back(slab);
			DT vs = dt_alloc(vd);
			copy2d(vs.data(), slab, v + slab_begin, r, rows_v, slab);
			DT Rs = dt_alloc({r, a_right, slab});
			copy2d(Rs.data(), slab, R + slab_begin, r, r * a_right, slab);
			spd_env_apply(Lv, ac, Rs, vs, &sc);
		}
		Context& c = ctx();
		// device-side tables of the peers' flag words and result areas
		PeerPtrs pp;
		for (int p = 0; p < 8; ++p) { pp.flags[p] = (p < world) ? flags_of(p) : nullptr; pp.y[p] = (p < world) ? y_of(p) : nullptr; }
		DBuf table(8);
		XB_CUDA(cudaMemcpyAsync(table.p, pp.flags, 8 * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
		peer_signal_kernel<<<1, 32, 0, c.stream>>>(reinterpret_cast<unsigned int* const*>(table.p), rank, world, 0);
		XB_LAUNCH_CHECK();
		const unsigned grid = unsigned(std::min<size_t>((block_elems / 2 + 255) / 256, size_t(c.num_sms) * 4));
		const unsigned int max_spins = unsigned(std::max(1024.0, c.peer_wait_spins));
		peer_wait_kernel<<<1, 32, 0, c.stream>>>(flags_of(rank), 0, epoch * unsigned(world - 1), max_spins);      // every peer's partial block has landed
		XB_LAUNCH_CHECK();
		peer_reduce_kernel<<<grid, 256, 0, c.stream>>>(pp, recv_of(rank), block_elems, rank, world, epoch);
		XB_LAUNCH_CHECK();
		peer_wait_kernel<<<1, 32, 0, c.stream>>>(flags_of(rank), 1, epoch * unsigned(world), max_spins);
		XB_LAUNCH_CHECK();
		*y_out = y_of(rank);
	});
}

// ---- the same application split along the LEFT bond index (rows of the result) -------------------------------------------------
// y(l, m.., r) = sum L(l, a, l') ... : rank g takes the rows l in [l_begin, l_end) of L and with them 1/world of ALL three stages
// (L.v, operator cores, .R) — nothing of the chain is replicated, and the row blocks of y are disjoint, so there is nothing to
// reduce: the exchange is an all-gather.  Fused variant: the epilogue of the last GEMM stores the rank's row block into the result
// area of EVERY rank (remote stores over NVLink while the remaining tiles are still being multiplied), one warp fences at system
// scope and bumps flag 4 on every rank, one thread waits for world bumps.  Half the NVLink traffic of the bond split of the
// contracted index (no partial sums travel), no reduce kernel, one wait instead of two.  v, A, R are needed in full on every rank.
__global__ void peer_signal_all_kernel(unsigned int* const* flags, const int world, const int which) {
	__threadfence_system();
	if (int(threadIdx.x) < world) atomicAdd_system(flags[threadIdx.x] + which, 1u);
}

static void env_apply_rows_impl(double* y_rows, const GemmScatter* sc, const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims,
                                size_t sites, const double* R, size_t r, size_t a_right, const double* v, size_t l_begin, size_t l_end) {
	XB_REQUIRE(sites >= 1 && sites <= 4, "1 to 4 sites");
	XB_REQUIRE(l_begin < l_end && l_end <= l, "illegal slab of the left bond");
	std::vector<DT> ac;
	std::vector<size_t> vd = {l};
	for (size_t p = 0; p < sites; ++p) {
		const size_t* d = A_dims + 4 * p;
		XB_REQUIRE(d[0] == (p == 0 ? a_left : A_dims[4 * (p - 1) + 3]), "operator bond dimensions do not coincide");
		ac.push_back(dt_view(A_cores[p], {d[0], d[1], d[2], d[3]}));
		vd.push_back(d[2]);
	}
	XB_REQUIRE(A_dims[4 * (sites - 1) + 3] == a_right, "operator bond dimensions do not coincide");
	vd.push_back(r);
	DT Ls = dt_view(L + l_begin * a_left * l, {l_end - l_begin, a_left, l});       // rows of L are contiguous: a view, no copy
	if (sc) { spd_env_apply(Ls, ac, dt_view(R, {r, a_right, r}), dt_view(v, vd), sc); return; }
	DT res = spd_env_apply(Ls, ac, dt_view(R, {r, a_right, r}), dt_view(v, vd));
	copy(y_rows, res.p, res.size());
}

xb_status xb_env_apply_rows(double* y_rows, const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims,
                            size_t sites, const double* R, size_t r, size_t a_right, const double* v, size_t l_begin, size_t l_end) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(y_rows && L && A_cores && A_dims && R && v, "null");
		env_apply_rows_impl(y_rows, nullptr, L, l, a_left, A_cores, A_dims, sites, R, r, a_right, v, l_begin, l_end);
	});
}

xb_status xb_env_apply_rows_fused(const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims, size_t sites,
                                  const double* R, size_t r, size_t a_right, const double* v, size_t l_begin, size_t l_end,
                                  int rank, int world, void* const* sym, unsigned int epoch, double** y_out) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(L && A_cores && A_dims && R && v && sym && y_out, "null");
		XB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && epoch >= 1, "illegal rank / world / epoch");
		size_t rows = l, row_len = 1;                        // result: (l, m_1..m_s, r)
		for (size_t p = 0; p < sites; ++p) row_len *= A_dims[4 * p + 1];
		rows *= row_len;
		auto flags_of = [&](int p) { return reinterpret_cast<unsigned int*>(sym[p]); };
		auto y_of = [&](int p) { return reinterpret_cast<double*>(static_cast<char*>(sym[p]) + PEER_FLAG_BYTES) + rows * r; };   // same layout as xb_env_apply_fused
		GemmScatter sc;
		sc.rows_per_block = 0; sc.replicate = world;
		for (int i = 0; i < 8; ++i) sc.blk[i] = (i < world) ? y_of(i) + l_begin * row_len * r : nullptr;
		env_apply_rows_impl(nullptr, &sc, L, l, a_left, A_cores, A_dims, sites, R, r, a_right, v, l_begin, l_end);
		Context& c = ctx();
		unsigned int* table_h[8];
		for (int p = 0; p < 8; ++p) table_h[p] = (p < world) ? flags_of(p) : nullptr;
		DBuf table(8);
		XB_CUDA(cudaMemcpyAsync(table.p, table_h, 8 * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
		peer_signal_all_kernel<<<1, 32, 0, c.stream>>>(reinterpret_cast<unsigned int* const*>(table.p), world, 4);
		XB_LAUNCH_CHECK();
		const unsigned int max_spins = unsigned(std::max(1024.0, c.peer_wait_spins));
		peer_wait_kernel<<<1, 32, 0, c.stream>>>(flags_of(rank), 4, epoch * unsigned(world), max_spins);      // every rank's row block has landed here
		XB_LAUNCH_CHECK();
		*y_out = y_of(rank);
	});
}

xb_status xb_als_solve(const xb_tt* A, xb_tt* x, const xb_tt* b, const xb_als_options* opt, double* energy, size_t* local_iterations) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(x && b && opt && energy, "null");
		require_correct_format(x); require_correct_format(b);
		XB_REQUIRE(!x->is_operator && !b->is_operator, "x and b must be TTTensors");
		XB_REQUIRE(x->d == b->d && x->dim_m == b->dim_m, "x.dimensions != b.dimensions");        // als.cpp:489
		if (A) {
			require_correct_format(A);
			XB_REQUIRE(A->is_operator && A->d == x->d, "A must be a TTOperator of the same order");   // als.cpp:493
			for (size_t i = 0; i < x->d; ++i)
				XB_REQUIRE(A->dim_m[i] == x->dim_m[i] && A->dim_n[i] == x->dim_m[i], "operator and tensor dimensions differ");   // :495-496
		}
		size_t rmax = 0, nmax = 0;
		for (size_t i = 0; i < x->d; ++i) { rmax = std::max(rmax, x->rank[i]); nmax = std::max(nmax, x->dim_m[i]); }
		PerfScope pa("TT sweep", opt->sites == 2 ? "DMRG" : (opt->local_solver == 1 ? "ASD" : "ALS"),
		             "d=" + pa_str(x->d) + " n=" + pa_str(nmax) + " r=" + pa_str(rmax) + " half-sweeps=" + pa_str(opt->num_half_sweeps));
		Als als;
		als.A = A; als.x = x; als.b = b; als.opt = *opt;
		*energy = als.run();
		if (local_iterations) *local_iterations = als.cg_iterations;
	});
}

} // extern "C"
