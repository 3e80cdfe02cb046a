/* xb200 — C ABI of the B200-native (sm_100a) tensor-train hot path.
 *
 * This is the drop-in boundary for xerus's dense numerical path.  xerus has no FFI; the seam is the free-function
 * namespace xerus::blasWrapper (reference: include/xerus/blasLapackWrapper.h:37-146, defined in
 * src/xerus/blasLapackWrapper.cpp) whose only library-side caller is src/xerus/tensor.cpp.  Three layers:
 *
 *   1. per-call layer   xb_<name>(...)      host pointers in/out, one call = one blasWrapper function
 *                                           (a ~150-line replacement blasLapackWrapper.cpp forwards to it)
 *   2. device layer     xb_dev_<name>(...)  same operations on device pointers, asynchronous on the library stream
 *   3. sweep layer      xb_tt_* / xb_als_*  device-resident tensor trains: whole TTNetwork::round / move_core /
 *                                           ALS sweeps as single calls (what is benchmarked)
 *
 * Conventions: all matrices row-major, densely packed unless an ld is passed (blasLapackWrapper.h: ld = #cols);
 * all arithmetic FP64 (reference: include/xerus/basic.h:43).  Every function returns 0 on success; on failure a
 * non-zero xb_status and xb_last_error() holds the message (the reference throws xerus::misc::generic_error,
 * misc/check.h:58-65 — the C++ shim turns non-zero into that exception).  There is no CPU fallback: every entry
 * point fails with XB_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef XB200_H
#define XB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int xb_status;
enum {
	XB_OK = 0,
	XB_ERR_INVALID = 1,     /* REQUIRE-style argument error (reference: REQUIRE(...) in blasLapackWrapper.cpp) */
	XB_ERR_CUDA = 2,        /* CUDA runtime error (sticky) */
	XB_ERR_NO_DEVICE = 3,   /* no usable CUDA device: the product path never falls back to the CPU */
	XB_ERR_UNSUPPORTED = 4, /* shape outside the implemented range */
	XB_ERR_NUMERIC = 5      /* factorisation failed (e.g. Cholesky of a non-SPD matrix), Jacobi did not converge */
};

/* ---- runtime ------------------------------------------------------------------------------------------------ */
xb_status   xb_init(int device);                 /* idempotent; creates the library stream + memory pool on `device` */
xb_status   xb_shutdown(void);
const char* xb_last_error(void);                 /* thread-local message of the last failing call */
int         xb_version(void);
xb_status   xb_synchronize(void);                /* waits for the library stream */
/* Workers: a worker is a private {CUDA stream, pinned scratch} bound to the calling host thread (thread-local selection,
 * created on demand, worker 0 by default).  Independent TT operations issued from different host threads on different
 * workers overlap on the GPU (BASELINE config 5).  A device object must be used by one worker at a time; hand it over
 * after xb_synchronize() on the producing worker. */
xb_status   xb_worker_select(int worker);
xb_status   xb_synchronize_all(void);            /* waits for the streams of all workers */
xb_status   xb_get_stream(void** cuda_stream);   /* the cudaStream_t all work is enqueued on (for CUDA-event timing) */
xb_status   xb_kernel_launch_count(uint64_t* n); /* number of xb200 kernels launched so far (bench: gpu_launches) */
xb_status   xb_set_option(const char* key, double value); /* tuning/diagnostic knobs, see DESIGN.md */
/* Optional per-kernel-class timing with CUDA events on the library stream (classes: "svd_jacobi", "svd", "qr",
 * "gemm"); off by default, used by bench.py for the roofline line (DESIGN.md "measurement"). */
xb_status   xb_profile_enable(int on);
xb_status   xb_profile_get(const char* kernel_class, uint64_t* scopes, uint64_t* launches, double* milliseconds);

/* (group, name, shape) call registry — the reference's XERUS_PERFORMANCE_ANALYSIS (misc/performanceAnalysis.h:30-39): every
 * per-call entry point records (calls, microseconds of host wall time) under the reference's own group / name / shape strings
 * (blasLapackWrapper.cpp:83-720, e.g. "Dense BLAS" / "Matrix-Matrix-Multiplication" / "512x256 * 256x256"); the sweep layer
 * records under the group "TT sweep".  Off by default.  Strings returned by xb_perf_entry stay valid until xb_perf_reset. */
xb_status xb_perf_enable(int on);
xb_status xb_perf_reset(void);
xb_status xb_perf_count(size_t* n);
xb_status xb_perf_entry(size_t i, const char** group, const char** name, const char** shape, uint64_t* calls, double* microseconds);

/* ---- memory hooks (back Tensor::denseData, reference: src/xerus/tensor.cpp:58, basic.cpp:31) ------------------ */
xb_status xb_alloc(void** dptr, size_t bytes);
xb_status xb_free(void* dptr);
xb_status xb_alloc_host(void** hptr, size_t bytes);      /* pinned host memory */
xb_status xb_free_host(void* hptr);
xb_status xb_upload(void* dst_dev, const void* src_host, size_t bytes);    /* async on the library stream */
xb_status xb_download(void* dst_host, const void* src_dev, size_t bytes);  /* async + stream synchronise */
/* xb_prefetch pins an existing host range in place (cudaHostRegister) so that the per-call layer's transfers of arrays that
 * live inside it are direct DMA instead of staged copies — for Tensor data that was allocated with new[] and cannot move
 * (the 15 allocation sites of tensor.cpp); xb_release undoes it.  Both are no-ops on a range that is already in that state. */
xb_status xb_prefetch(const void* host_ptr, size_t bytes);
xb_status xb_release(const void* host_ptr);

/* ---- 1. per-call layer: host pointers, mirrors xerus::blasWrapper one to one ----------------------------------- */
/* blasLapackWrapper.h:41-47 */
xb_status xb_one_norm(const double* x, size_t n, double* result);
xb_status xb_two_norm(const double* x, size_t n, double* result);
xb_status xb_dot_product(const double* x, size_t n, const double* y, double* result);
/* :53 x (m) = alpha*op(A)*y (n); A stored m x n, or n x m if transposed   :56 A = alpha*x*y^T */
xb_status xb_matrix_vector_product(double* x, size_t m, double alpha, const double* A, size_t n, int transposed, const double* y);
xb_status xb_dyadic_vector_product(double* A, size_t m, size_t n, double alpha, const double* x, const double* y);
/* :60-70 C = alpha*op(A)*op(B), C is leftDim x rightDim with ldc = rightDim, beta = 0 */
xb_status xb_matrix_matrix_product(double* C, size_t leftDim, size_t rightDim, double alpha, const double* A, size_t lda,
                                   int transposeA, size_t middleDim, const double* B, size_t ldb, int transposeB);
/* :90-93 thin SVD: U m x min, S min, Vt min x n, singular values descending */
xb_status xb_svd(double* U, double* S, double* Vt, const double* A, size_t m, size_t n);
/* :97-107 rank-revealing A = Q*C / A = C*Q.  Caller provides max-size buffers (Q: m x min, C: min x n resp.
 * C: m x min, Q: min x n); on return they are packed for the detected *rank (Q: m x rank, C: rank x n, ...).
 * Rank rule: |d_k| < 16*eps*|d_0| on the rank-revealing factor (the reference compares with the signed R[0,0],
 * blasLapackWrapper.cpp:269 — see DESIGN.md "rank rule"). */
xb_status xb_qc(double* Q, double* C, size_t* rank, const double* A, size_t m, size_t n);
xb_status xb_cq(double* C, double* Q, size_t* rank, const double* A, size_t m, size_t n);
/* :111-131 unpivoted thin QR / RQ: Q m x min, R min x n   |   R m x min, Q min x n */
xb_status xb_qr(double* Q, double* R, const double* A, size_t m, size_t n);
xb_status xb_rq(double* R, double* Q, const double* A, size_t m, size_t n);
/* :135 solve A x = b, A is m x n row-major, b is m x nrhs, x is n x nrhs.  Dispatch as the reference
 * (blasLapackWrapper.cpp:542-651): symmetric + definite diagonal -> Cholesky, otherwise LU with partial pivoting;
 * m != n -> least squares through QR (full column rank). */
xb_status xb_solve(double* x, const double* A, size_t m, size_t n, const double* b, size_t nrhs);
xb_status xb_solve_least_squares(double* x, const double* A, size_t m, size_t n, const double* b, size_t p);

/* Tensor-level helpers that sit directly on the boundary in the reference
 * (reshuffle: src/xerus/indexedTensor_tensor_evaluate.cpp:55-137; out mode shuffle[i] = in mode i). */
xb_status xb_reshuffle(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree);

/* ---- 2. device layer: same semantics on device pointers, asynchronous on the library stream ------------------- */
xb_status xb_dev_gemm(double* C, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, int transA,
                      size_t k, const double* B, size_t ldb, int transB, double beta);
xb_status xb_dev_qr(double* Q, double* R, const double* A, size_t m, size_t n);
xb_status xb_dev_lq(double* L, double* Q, const double* A, size_t m, size_t n);   /* A = L*Q, Q min x n, L m x min (lower) */
xb_status xb_dev_svd(double* U, double* S, double* Vt, const double* A, size_t m, size_t n, size_t k_out,
                     int scale_u, int scale_vt, int* sweeps); /* first k_out triplets; optional Sigma folded into U / Vt */
xb_status xb_dev_reshuffle(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree);
xb_status xb_dev_two_norm(const double* x, size_t n, double* host_result);

/* ---- 3. sweep layer: device-resident tensor trains ------------------------------------------------------------ */
typedef struct xb_tt xb_tt;   /* TTTensor (cores r x n x r') or TTOperator (cores r x m x n x r'), ttNetwork.h:44-519 */

/* dims: d entries for a TTTensor; 2*d entries (m_1..m_d, n_1..n_d) for a TTOperator (reference ordering of
 * TTOperator::dimensions).  ranks: d-1 bond ranks.  Cores are zero-initialised; not canonicalised. */
xb_status xb_tt_create(xb_tt** out, size_t d, const size_t* dims, const size_t* ranks, int is_operator);
xb_status xb_tt_destroy(xb_tt* tt);
xb_status xb_tt_clone(xb_tt** out, const xb_tt* tt);
xb_status xb_tt_degree(const xb_tt* tt, size_t* d, int* is_operator);
xb_status xb_tt_ranks(const xb_tt* tt, size_t* ranks /* d-1 */);
xb_status xb_tt_dims(const xb_tt* tt, size_t* dims /* d or 2d */);
xb_status xb_tt_core_position(const xb_tt* tt, int* canonicalized, size_t* position);
xb_status xb_tt_assume_core_position(xb_tt* tt, size_t position);                 /* ttNetwork.cpp:735-739 */
/* set_component / get_component (ttNetwork.cpp:457-492): host row-major core of shape (rl, n, rr) / (rl, m, n, rr);
 * set_component may change the bond ranks and clears `canonicalized` unless idx is the core. */
xb_status xb_tt_set_component(xb_tt* tt, size_t idx, const double* host_core, size_t rl, size_t rr);
xb_status xb_tt_get_component(const xb_tt* tt, size_t idx, double* host_core);
xb_status xb_tt_component_size(const xb_tt* tt, size_t idx, size_t* rl, size_t* ext, size_t* rr);
/* all d components at once (copies enqueued back to back, one synchronisation): host_cores[i] row-major as above,
 * ranks = the d-1 bond ranks the written components have; clears `canonicalized`. */
xb_status xb_tt_set_components(xb_tt* tt, const double* const* host_cores, const size_t* ranks);
xb_status xb_tt_get_components(const xb_tt* tt, double* const* host_cores);
/* TTNetwork::move_core (ttNetwork.cpp:582-628): keep_rank -> plain QR/LQ, otherwise rank-revealing */
xb_status xb_tt_move_core(xb_tt* tt, size_t position, int keep_rank);
/* TTNetwork::round (ttNetwork.cpp:644-684): max_ranks has d-1 entries (0 = unlimited), 0 <= eps < 1.
 * svals (optional, may be NULL): receives the kept singular values per edge, edge e at svals[e*stride ..]. */
xb_status xb_tt_round(xb_tt* tt, const size_t* max_ranks, double eps);
xb_status xb_tt_round_svals(xb_tt* tt, const size_t* max_ranks, double eps, double* svals, size_t stride);
/* TTNetwork::soft_threshold (ttNetwork.cpp:688-713): the sweep of round() without rank cap and with eps = 0, every singular
 * value replaced by max(0, sigma - tau) (tensorNetwork.cpp:766); ranks do not change.  taus has d-1 entries and — as in the
 * reference, :700 — taus[i] belongs to the i-th edge *from the right*.  prevent_zero is accepted and ignored, as there. */
xb_status xb_tt_soft_threshold(xb_tt* tt, const double* taus, int prevent_zero);
/* Batches of independent items (BASELINE config 5).  The items run concurrently on library workers (CUDA stream + memory pool +
 * plan cache each — option "batch_workers", default 16) driven asynchronously by a few library-owned host threads (option
 * "batch_threads", default 2); the call is stream-ordered with respect to the caller's worker on both sides.  The items of a batch must be distinct objects.  xb_tt_round_batched: tts[b].round(max_rank, eps).  xb_tt_apply_round_batched: the item of
 * config 5, out[b] = round(A x_b) (y(i&0) = A(i/2,j/2) * x(j&0), then TTNetwork::round); the caller destroys out[b]. */
xb_status xb_tt_round_batched(xb_tt** tts, size_t batch, size_t max_rank, double eps);
xb_status xb_tt_apply_round_batched(xb_tt** out, const xb_tt* A, xb_tt* const* xs, size_t batch, size_t max_rank, double eps);
xb_status xb_tt_frob_norm(const xb_tt* tt, double* result);                        /* ttNetwork.cpp:782-789 */
xb_status xb_tt_inner(const xb_tt* a, const xb_tt* b, double* result);             /* a(i&0)*b(i&0) */
xb_status xb_tt_distance(const xb_tt* a, const xb_tt* b, double* result);          /* ||a-b||, cancellation-free */
xb_status xb_tt_scale(xb_tt* tt, double factor);                                   /* operator*= (ttNetwork.cpp:860-868) */
xb_status xb_tt_add(xb_tt** out, const xb_tt* a, const xb_tt* b);                  /* operator+  (ttNetwork.cpp:797-847) */
xb_status xb_tt_apply(xb_tt** out, const xb_tt* A, const xb_tt* x);               /* y(i&0)=A(i/2,j/2)*x(j&0), ttStack.cpp:197-300 */
xb_status xb_tt_from_dense(xb_tt** out, const double* host, size_t d, const size_t* dims, double eps, size_t max_rank); /* ttNetwork.cpp:112-160 */
/* the same constructor with per-bond rank caps (max_ranks: d-1 entries, NULL = unlimited) and for TTOperators: dims then has
 * 2d entries (m_1..m_d, n_1..n_d), the dense operator is reshuffled to (m_1,n_1,m_2,n_2,...) first (ttNetwork.cpp:129-135) */
xb_status xb_tt_from_dense_ex(xb_tt** out, const double* host, size_t d, const size_t* dims, int is_operator, double eps, const size_t* max_ranks);
xb_status xb_tt_to_dense(const xb_tt* tt, double* host);                           /* tensorNetwork.cpp:287-306 */

/* ALS / DMRG (src/xerus/algorithms/als.cpp:483-553).  A may be NULL (projection of b, als.cpp:541-545).
 * sites = 1 (ALS) or 2 (DMRG); assume_spd as ALSVariant::assumeSPD.  The local problems are solved matrix-free on
 * device (conjugate gradients on the three-factor local operator) instead of densifying it
 * (reference: als.cpp:43-48).  Returns the energy the reference returns (als.cpp:548). */
typedef struct {
	uint32_t sites;
	int      assume_spd;
	size_t   num_half_sweeps;        /* 0 = until convergence */
	double   convergence_epsilon;    /* als.h:137 default 1e-6 */
	int      preserve_core_position; /* als.h:120 default true */
	double   local_tolerance;        /* relative residual of the iterative local solves (0 -> 1e-15, i.e. down to the rounding
	                                    floor: the solver stops by itself when the residual stagnates there) */
	size_t   local_max_iterations;   /* 0 -> 4 * local size, capped */
	int      local_solver;           /* 0: ALSVariant::lapack_solver semantics (als.cpp:43-71; dense or matrix-free CG);
	                                    1: ALSVariant::ASD_solver (als.cpp:73-103), one exact-line-search gradient step per site */
} xb_als_options;
xb_status xb_als_default_options(xb_als_options* opt, uint32_t sites, int assume_spd);
xb_status xb_als_solve(const xb_tt* A, xb_tt* x, const xb_tt* b, const xb_als_options* opt, double* energy,
                       size_t* local_iterations /* optional: total CG iterations */);

/* Matrix-free local-operator application y = {L, A_1..A_s, R} v of ALS/DMRG (the un-contracted network that
 * construct_local_operator returns, als.cpp:383-401), SPD environments, on device pointers:
 * L (l, a_left, l), A_p (A_dims[4p..4p+3] = r_l, m, n, r_r), R (r, a_right, r), v (l, n_1..n_s, r), y (l, m_1..m_s, r).
 * [slab_begin, slab_end) restricts the contraction over the right bond r' to a slab: the partial results of disjoint
 * slabs sum to the full application (BASELINE config 4: bond index split across GPUs + one NCCL all-reduce). */
xb_status xb_env_apply(double* y, const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims,
                       size_t sites, const double* R, size_t r, size_t a_right, const double* v, size_t slab_begin, size_t slab_end);

/* Bond-split application fused with its reduction over peer memory (BASELINE config 4, the GPUs of one box; replaces
 * xb_env_apply + ncclAllReduce).  Every rank creates one symmetric buffer of xb_peer_buffer_bytes(rows, cols, world) bytes
 * (rows = l * m_1..m_s, cols = r), exports its CUDA IPC handle (64 bytes) and opens the handles of the other ranks; `sym`
 * holds the world device pointers in rank order (sym[rank] = the rank's own buffer).  The last contraction writes each
 * rank's row block of the partial result straight into that rank's buffer over NVLink, a reduce kernel sums the blocks in
 * rank order and writes the sum into every rank's result area; *y_out points at this rank's copy of the full result, valid
 * in stream order.  `epoch` counts the calls on these buffers from 1 and must agree on all ranks. */
/* The same application split along the LEFT bond index instead (rows of the result): rank g takes the rows [l_begin, l_end) of
 * L and does 1/world of all three stages of the chain; the row blocks of y are disjoint, so the exchange is an all-gather and
 * nothing is summed.  xb_env_apply_rows writes the rank's row block ((l_end - l_begin) x m_1..m_s x r, contiguous) to y_rows
 * (exchange left to the caller, e.g. ncclAllGather); xb_env_apply_rows_fused stores it into the result area of every rank's
 * symmetric buffer from the epilogue of the last GEMM (buffers, epoch and *y_out as for xb_env_apply_fused; the two fused
 * variants use different flag words and may share buffers). */
xb_status xb_env_apply_rows(double* y_rows, const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims,
                            size_t sites, const double* R, size_t r, size_t a_right, const double* v, size_t l_begin, size_t l_end);
xb_status xb_env_apply_rows_fused(const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims, size_t sites,
                                  const double* R, size_t r, size_t a_right, const double* v, size_t l_begin, size_t l_end,
                                  int rank, int world, void* const* sym, unsigned int epoch, double** y_out);
xb_status xb_peer_buffer_bytes(size_t rows, size_t cols, int world, size_t* bytes);
xb_status xb_peer_buffer_create(size_t bytes, void** dptr, unsigned char* handle64);
xb_status xb_peer_buffer_open(const unsigned char* handle64, void** dptr);
xb_status xb_peer_buffer_close(void* dptr);
/* the waits inside xb_env_apply_fused are bounded; this synchronises the calling worker's stream and returns XB_ERR_CUDA if a
 * wait on this rank's own buffer gave up (a peer lagged or died: the result of that call and of all later ones is not valid) */
xb_status xb_peer_buffer_check(void* dptr);
xb_status xb_peer_buffer_destroy(void* dptr);
xb_status xb_env_apply_fused(const double* L, size_t l, size_t a_left, const double* const* A_cores, const size_t* A_dims, size_t sites,
                             const double* R, size_t r, size_t a_right, const double* v, size_t slab_begin, size_t slab_end,
                             int rank, int world, void* const* sym, unsigned int epoch, double** y_out);

/* ---- 4. data files: the reference's own save_to_file / load_from_file format, Binary and TSV ------------------------
 * (misc::save_to_file / load_from_file, include/xerus/misc/fileIO.h:102-164; stream_writer / stream_reader of Tensor,
 * src/xerus/tensor.cpp:1781-1845, of TensorNetwork, src/xerus/tensorNetwork.cpp:1429-1505, and of TTNetwork,
 * src/xerus/ttNetwork.cpp:1455-1488).  Files written by xerus load here and files written here load in xerus.
 * xb_file_* are host-side and need no device; xb_tt_load / xb_tt_save move the cores to / from a device-resident TT. */
typedef struct xb_file xb_file;
xb_status xb_file_open(xb_file** out, const char* filename);     /* parses the whole file; sparse tensors are densified */
xb_status xb_file_close(xb_file* f);
/* kind: 0 = Tensor, 1 = TTTensor, 2 = TTOperator; n_dims = entries of xb_file_dims (Tensor: degree; TT: d or 2d) */
xb_status xb_file_info(const xb_file* f, int* kind, size_t* n_dims, int* canonicalized, size_t* core_position);
xb_status xb_file_dims(const xb_file* f, size_t* dims);
xb_status xb_file_ranks(const xb_file* f, size_t* ranks /* d-1, TT files only */);
/* TT: row-major component idx, (rl, n, rr) / (rl, m, n, rr); Tensor: idx 0 = the dense row-major data */
xb_status xb_file_read_component(const xb_file* f, size_t idx, double* host);
xb_status xb_file_write_tensor(const char* filename, int tsv, const double* data, const size_t* dims, size_t degree);
xb_status xb_file_write_tt(const char* filename, int tsv, size_t d, const size_t* dims /* d or 2d */, const size_t* ranks /* d-1 */,
                           int is_operator, int canonicalized, size_t core_position, const double* const* cores);
xb_status xb_tt_load(xb_tt** out, const char* filename);         /* misc::load_from_file<TTTensor / TTOperator> */
xb_status xb_tt_save(const xb_tt* tt, const char* filename, int tsv); /* misc::save_to_file(tt, filename, format) */

#ifdef __cplusplus
}
#endif
#endif /* XB200_H */
