// Internal declarations shared by the xb200 translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/xb200.h"

namespace xb {

struct Error : std::runtime_error {
	xb_status code;
	Error(xb_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define XB_STR2(x) #x
#define XB_STR(x) XB_STR2(x)
#define XB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
	throw ::xb::Error(XB_ERR_CUDA, std::string(#call " failed at " __FILE__ ":" XB_STR(__LINE__) ": ") + cudaGetErrorString(e__)); } while (0)
#define XB_REQUIRE(cond, msg) do { if (!(cond)) \
	throw ::xb::Error(XB_ERR_INVALID, std::string("REQUIRE(" #cond ") failed: ") + (msg)); } while (0)
#define XB_LAUNCH_CHECK() do { ::xb::ctx().launches++; XB_CUDA(cudaGetLastError()); } while (0)

void set_last_error(const std::string& msg);

struct ProfRecord { std::string name; cudaEvent_t e0, e1; uint64_t launches0, launches1; };
struct ProfTotal { uint64_t scopes = 0, launches = 0; double ms = 0.0; };

// One Context per worker: a worker is a host thread's private {stream, pinned scratch, counters}.  Worker 0 is created
// by xb_init; more are created on demand by xb_worker_select (batches of independent TTs drive one worker per host
// thread so that several sweeps are in flight on the GPU at once).  Device-wide facts and options are copied.
struct Context {
	bool initialised = false;
	int worker = 0;
	double* red_partial = nullptr;   // scratch of the two-stage reductions
	double* red_partial_aux = nullptr;   // the same for reductions issued on the side stream
	std::vector<ProfRecord> prof_pending;
	std::vector<std::pair<std::string, ProfTotal>> prof_totals;
	int device = -1;
	cudaStream_t stream = nullptr;
	cudaStream_t aux = nullptr;      // side stream for work off the critical path of a sweep (explicit Q of a QR), see aux_fork / aux_join
	cudaEvent_t aux_fork_ev = nullptr, aux_join_ev = nullptr;
	bool aux_pending = false;
	cudaMemPool_t pool = nullptr;
	uint64_t launches = 0;
	int num_sms = 148;
	size_t max_smem_optin = 0;
	// pinned scratch for small read-backs (ranks, convergence counters, singular values)
	double* h_scratch = nullptr;   // 4096 doubles
	size_t h_scratch_len = 4096;
	// options
	int svd_max_sweeps = 100;      // graded spectra (kappa ~ 1e14) need ~45 sweeps of the un-preconditioned Jacobi
	int qr_panel = 32;
	int gemm_force_small = 0;
	int qr_defer = 1;              // sweeps: the explicit Q of a QR is formed on the side stream while the main stream carries on with R
	int gemm_big = 1;              // 128 x 128 cp.async GEMM kernel for outputs of about a wave of such tiles or more
	bool profile = false;
	int svd_persistent = 1;
	int tt_svd_polish = 1;         // polish level of the SVDs inside round() / TT-SVD / DMRG splits (see Svd::polish)
	int svd_polish = 1;            // Newton-Schulz re-orthogonalisation of V + recomputed left part after the Jacobi sweeps
	int svd_flip = 1;              // Jacobi on the rows of the triangular factor after a QR reduction (pre-conditioning)
	int svd_square_qr = 1;         // square inputs also go through the QR reduction (needed for svd_flip)
	int qr_cluster = 1;            // QR panels of 128..2048 rows on a thread-block cluster (registers + DSMEM reduction)
	int qr_chol = 1;               // tall QRs of up to 128 columns by Cholesky-QR2 where the input allows it (qr_f64.cu)
	int qr_chol_min_rows = 1024;   // ... from this many rows on (below, the cluster panel kernels are as fast or faster)
	int chol_skip = 0, chol_declines = 0;     // back-off after declined attempts (reset at every C-ABI entry)
	bool chol_off = false;                    // the plan being recorded / captured is Householder only
	int chol_tape_mode = 0;                   // 0 off, 1 record accept / decline per candidate, 2 replay them (plan capture)
	std::vector<char> chol_tape; size_t chol_tape_pos = 0;
	int qr_cluster_min_rows = 64;  // smallest panel height for the cluster panel kernel (below: one CTA, panel in shared memory)
	int svd_fast = 1;              // specialised Jacobi kernel (compile-time row length, 128-bit accesses) up to 512 columns
	int svd_jacc = 1;              // specialised Jacobi kernel: rotations of a block visit accumulated, applied to V once (DMMA)
	int als_cg_merged = 1;         // persistent CG: two grid barriers per iteration (operator applied to r, q by recurrence)
	int als_cg_cluster = 1;        // persistent CG: step 1 shared inside thread-block clusters (partial sums exchanged through DSMEM)
	int als_persistent_cg = 1;     // one-site SPD local problems: a whole CG run in one cooperative launch (spd_cg_kernel)
	int als_graph = 1;             // one-site SPD CG: chunks of 8 iterations replayed as a CUDA graph
	int svd_gram = 0;              // experimental: Jacobi block visits in Gram space (one Gram matrix, 16 x 16 rounds, one DMMA apply
	                               // per visit); correct, but not faster than the column-space kernel on one SM per block pair
	double svd_last_sweep_cos = 1e-7;   // a Jacobi sweep in which no pair had |cos| above this is the last one (quadratic convergence)
	int svd_colsort = 1;           // QR pre-conditioning of the SVD on the columns sorted by descending norm (surrogate of a pivoted QR)
	int svd_dsmem = 1;             // split Jacobi kernel: X workers in one cluster, travelling block handed over through DSMEM (st.async + mbarrier)
	int svd_split = 1;             // Jacobi: separate CTAs apply the rotation products to the accumulated-rotation halves
	int svd_recursive = 1;         // recursive bipartite tournament with point-to-point block flags (power-of-two block counts)
	int svd_max_bw = 0;            // 0 = automatic block width of the Jacobi kernel
	int als_direct_max = 1536;     // local problems up to this size are solved densely (reference semantics), larger ones by CG
	double peer_wait_spins = 268435456.0;   // bound of the polling loops of the fused bond-split exchange (2^28 polls, about a minute)
	int small_kernels = 1;         // min(m,n) <= 32: QR and Jacobi SVD as one single-CTA launch each (small_f64.cu)
	int batch_threads = 0;         // host threads that drive the batch workers (0: min(batch_workers, 2))
	unsigned int* h_flags = nullptr;   // pinned flag words of round-plan replays in flight (tt.cu)
	int batch_workers = 16;        // host threads / library workers of the batched entry points (xb_tt_round_batched, ...)
	int round_plans = 1;           // round(): repeated shapes replay a captured CUDA graph of the whole sweep (speculative ranks, tt.cu)
	// Speculative execution (round plans): rank decisions are not read back; every decision point assumes the outcome the
	// plan was recorded with and raises *spec_flag on the device when the data disagree (the caller then repeats the
	// operation on the synchronising path).
	bool speculate = false;
	unsigned int* spec_flag = nullptr;
	std::vector<struct RoundPlan*> plans;
	// Capture arena: while a plan is being captured every device allocation is a bump allocation from the plan's persistent
	// arena and frees are no-ops, so the graph holds kernel nodes only (graphs with memory nodes do not overlap across streams).
	// alloc_counter measures the arena a shape needs during its first, ordinary run.
	char* arena = nullptr; size_t arena_size = 0, arena_off = 0; bool arena_on = false;
	bool count_allocs = false; size_t alloc_counter = 0;
	uint64_t options_epoch = 0;    // bumped by xb_set_option: plans recorded under other options are not replayed
};
Context& ctx();               // the calling thread's current worker
void release_plans(Context& c);   // destroys the round plans of a worker (tt.cu)
void ensure_init();

// Optional CUDA-event timing of a kernel class (bench/roofline only; no-op unless xb_profile_enable(1)).
struct ProfScope {
	int slot = -1;
	explicit ProfScope(const char* kernel_class);
	~ProfScope();
};

// (group, name, shape) registry of the C-ABI calls (runtime.cu); no-op unless xb_perf_enable(1)
struct PerfScope {
	bool on = false; const char* g = nullptr; const char* n = nullptr; std::string s; double t0 = 0.0;
	PerfScope(const char* group, const char* name, const std::string& shape);
	~PerfScope();
};
inline std::string pa_str(size_t v) { return std::to_string(v); }

// stream-ordered device allocations from the library pool
double* dalloc(size_t n_doubles);
void* dalloc_bytes(size_t bytes);
void dfree(void* p);
struct DBuf {   // RAII device buffer
	double* p = nullptr; size_t n = 0;
	DBuf() {}
	explicit DBuf(size_t n_) : p(n_ ? dalloc(n_) : nullptr), n(n_) {}
	DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
	DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
	DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { reset(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
	~DBuf() { reset(); }
	void reset() { if (p) dfree(p); p = nullptr; n = 0; }
	void resize(size_t n_) { reset(); if (n_) { p = dalloc(n_); n = n_; } }
	operator double*() const { return p; }
};

template <class F> xb_status guard(F&& f) {
	try { Context& c = ctx(); c.chol_skip = 0; c.chol_declines = 0; f(); return XB_OK; }
	catch (const Error& e) { set_last_error(e.what()); return e.code; }
	catch (const std::exception& e) { set_last_error(e.what()); return XB_ERR_INVALID; }
}

// ---- device primitives (all asynchronous on ctx().stream, device pointers) ---------------------------------------
// C (m x n, ldc) = alpha * op(A) * op(B) + beta * C ; batched variant walks `batch` problems with element strides.
void gemm(double* C, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, bool transA, size_t k,
          const double* B, size_t ldb, bool transB, double beta);
void gemm_batched(double* C, size_t ldc, size_t strideC, size_t m, size_t n, double alpha, const double* A, size_t lda,
                  size_t strideA, bool transA, size_t k, const double* B, size_t ldb, size_t strideB, bool transB,
                  double beta, size_t batch);

// C row i -> blk[i / rows_per_block] + (i % rows_per_block) * ldc (beta = 0): the row blocks may be buffers of peer GPUs
// replicate > 0: instead, the whole result is written to blk[0 .. replicate) (same layout in each)
struct GemmScatter { double* blk[8]; size_t rows_per_block; int replicate = 0; };
void gemm_scatter(const GemmScatter& sc, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, bool transA, size_t k,
                  const double* B, size_t ldb, bool transB);

// elementwise / movement
void copy(double* dst, const double* src, size_t n);
void copy2d(double* dst, size_t ldd, const double* src, size_t lds, size_t rows, size_t cols);
void fill(double* dst, double value, size_t n);
void set_identity(double* dst, size_t rows, size_t cols, size_t ld);      // ones on the main diagonal, zero elsewhere
void scale(double* x, double alpha, size_t n);
void axpy(double* y, double alpha, const double* x, size_t n);           // y += alpha x
void transpose(double* out, const double* in, size_t rows, size_t cols); // out (cols x rows) = in^T, both packed
void transpose_reverse(double* out, const double* in, size_t rows, size_t cols); // out(j,i) = in(rows-1-i, cols-1-j)
void permute(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree);
// out(o,q,j) = sum_p W(q,p) in(o,p,j) for small P, Q (one pass over HBM); false if P > 64 or Q > 32 (caller takes the GEMM route)
bool mid_apply(double* out, const double* in, const double* W, size_t outer, size_t P, size_t Q, size_t inner);
void scale_rows(double* A, const double* s, size_t rows, size_t cols, size_t ld);   // A[i,:] *= s[i]
void scale_cols(double* A, const double* s, size_t rows, size_t cols, size_t ld);   // A[:,j] *= s[j]
// reductions: result written to a device double
void dot_dev(double* d_result, const double* x, const double* y, size_t n);
void asum_dev(double* d_result, const double* x, size_t n);
void amax_scale_dev(double* d_scale2, const double* x, size_t n);          // d_scale2[0] = 2^-e with max|x| * 2^-e in [0.5,1), [1] = 2^e
void scale_by_dev(double* dst, const double* src, size_t n, const double* d_factor);
void scale_block_by_dev(double* A, size_t ld, size_t rows, size_t cols, const double* d_factor);
double read_scalar(const double* d_value);                                // D2H + sync
double two_norm(const double* x, size_t n);
double dot(const double* x, const double* y, size_t n);

// Side stream.  aux_fork(): the side stream waits for everything enqueued on the main stream so far.  aux_join(): the main
// stream waits for everything enqueued on the side stream so far (no-op when nothing is pending).  AuxScope redirects
// ctx().stream — and with it every launch, allocation and free of the library — to the side stream for its lifetime.
void aux_fork();
void aux_join();
struct AuxScope {
	cudaStream_t saved = nullptr; bool on = false;
	explicit AuxScope(bool enable);
	~AuxScope();
};

// factorizations
// A (m x n packed, destroyed? no: const) -> Q (m x k packed), R (k x n packed), k = min(m,n).
// defer_q: R is complete on the main stream when this returns, Q is formed on the side stream; the caller calls aux_join()
// before the main stream reads, overwrites or frees Q.
void qr(double* Q, double* R, const double* A, size_t m, size_t n, bool defer_q = false);
// A = L * Q : L m x k packed (lower trapezoidal), Q k x n packed, orthonormal rows
void lq(double* L, double* Q, const double* A, size_t m, size_t n);
// true RQ with LAPACK's convention (R upper-trapezoidal, bottom-right aligned)
void rq(double* R, double* Q, const double* A, size_t m, size_t n);

// thrown by a decision point that cannot run without a host read-back while ctx().speculate is set
struct SpecUnsupported : Error { explicit SpecUnsupported(const std::string& m) : Error(XB_ERR_UNSUPPORTED, m) {} };

// single-CTA kernels for min(m, n) <= 32 (small_f64.cu); G(i, j) = A[i * ars + j * acs], outputs through strides
bool qr_small_fits(size_t m, size_t n);
void qr_small(double* Q, long long qrs, long long qcs, double* R, long long rrs, long long rcs, const double* A, long long ars, long long acs, size_t m, size_t n);
bool chol_inv_fits(size_t n);
void chol_inv(const double* G, size_t nparts, size_t n, double* R, double* W, bool near_identity, double ratio, unsigned int* flag, unsigned int flag_value, bool clear);
void sum_parts(double* out, const double* P, size_t parts, size_t len);
bool svd_small_fits(size_t mw, size_t nw);
void svd_small(const double* A, long long rs, long long cs, size_t mw, size_t nw, double* GT, size_t ld, size_t voff, double* Ssorted, int* perm,
               double* scale2, unsigned int* info, double tol, double last_cos, int max_sweeps, int polish);

struct SvdWork;   // opaque between svd_factor and svd_extract
// Jacobi SVD of A (m x n packed).  Returns singular values (descending) on the host, keeps vectors on device.
// Then extract() writes the first k triplets: U (m x k), Vt (k x n); Sigma optionally folded into U or Vt.
struct Svd {
	size_t m = 0, n = 0, kmax = 0;
	std::vector<double> S;     // host copy, descending
	int sweeps = 0;
	double soft_threshold = 0.0;   // applied to Sigma wherever extract() folds it in
	int polish = 2;                // 2: Newton-Schulz on V + recomputed left part + clean-up sweep (LAPACK-grade orthogonality of both
	                               // factors, per-call layer); 1: without the clean-up sweep (sweep layer: exact projection, left vectors
	                               // orthogonal to ~1e-12); 0: none
	// internals
	bool swapped = false, reduced = false, flipped = false, permuted = false;
	size_t mw = 0, nw = 0, npad = 0, mt = 0, mdot = 0, voff = 0, ld = 0;
	DBuf GT, Qred, Ssorted, perm, colperm, scale;   // scale: [2^-e, 2^e] of the Jacobi input (squares must not overflow)
	bool q_deferred = false;               // Qred is still being formed on the side stream
	~Svd() { if (q_deferred) aux_join(); }
	void factor(const double* A, size_t m, size_t n);
	// rank of the truncation rule (tensor.cpp:1464-1474).  Normally evaluated on the host copy of S; while ctx().speculate is set
	// it returns min(max_rank, kmax) and enqueues a check that raises the speculation flag if the eps rule would cut earlier.
	size_t rank_for(size_t max_rank, double eps);
	void extract(double* U, double* Vt, size_t k, bool scale_u, bool scale_vt, double* dS /* optional device S (k) */);
};

// rank-revealing QC / CQ on device; returns the rank; Q, C are max-size device buffers that come back packed
size_t qc(double* Q, double* C, const double* A, size_t m, size_t n, bool defer_q = false);   // defer_q as in qr()
size_t cq(double* C, double* Q, const double* A, size_t m, size_t n);

// dense solves on device (A n x n packed, destroyed; B n x nrhs packed, overwritten with X). Return false if not SPD.
bool cholesky_solve(double* A, double* B, size_t n, size_t nrhs);
void lu_solve(double* A, double* B, size_t n, size_t nrhs);
void probe_symmetry(const double* A, size_t n, bool& symmetric, bool& definite_diag);

// truncation rule of calculate_svd (reference: src/xerus/tensor.cpp:1464-1474)
inline size_t truncation_rank(const std::vector<double>& S, size_t max_rank, double eps) {
	size_t rank = S.size();
	if (max_rank != 0) rank = std::min(rank, max_rank);
	for (size_t j = 1; j < rank; ++j) if (S[j] <= eps * S[0]) { rank = j; break; }
	return rank;
}

} // namespace xb
