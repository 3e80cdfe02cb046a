// Device-resident tensor trains: the sweep layer of the C ABI.
//
// A xb_tt owns its cores in HBM for its whole life (row-major (r_left, n, r_right) resp. (r_left, m, n, r_right),
// reference layout: src/xerus/ttNetwork.cpp:96-98, ttNetwork.h:146-149); left/right matricisations are
// reinterpretations of the same buffer (reference: calculate_factorization_sizes, src/xerus/tensor.cpp:1361-1369).
// move_core / round follow TTNetwork::move_core / round (src/xerus/ttNetwork.cpp:582-684) edge by edge through
// transfer_core / round_edge (src/xerus/tensorNetwork.cpp:678-909); nothing leaves the device during a sweep except
// the per-edge rank decision (singular values / one min-max pair), which the host needs to size the next launches.
#include "tt_internal.cuh"
#include <atomic>
#include <mutex>
#include <thread>

using namespace xb;

namespace xb {

static void check_idx(const xb_tt* t, size_t idx) { XB_REQUIRE(t && idx < t->d, "Illegal component index"); }

// shape string of a TT for the call registry: "d=32 n=2 r=256" (largest external dimension and bond rank)
static std::string tt_shape(const xb_tt* t) {
	size_t n = 0, r = 0;
	for (size_t i = 0; i < t->d; ++i) { n = std::max(n, t->ext(i)); r = std::max(r, t->rank[i]); }
	return "d=" + pa_str(t->d) + " n=" + pa_str(n) + " r=" + pa_str(r);
}

// TensorNetwork::transfer_core for a TT chain (tensorNetwork.cpp:821-909).
static void transfer_core(xb_tt* t, size_t from, size_t to, bool allow_rank_reduction) {
	if (to == from + 1) {
		const size_t rows = t->rank[from] * t->ext(from), cols = t->rank[from + 1];
		const size_t kmax = std::min(rows, cols);
		DBuf Q(rows * kmax), R(kmax * cols);
		size_t k = kmax;
		// the explicit Q only becomes the new core: it is formed on the side stream (move_core joins before anything reads it)
		if (allow_rank_reduction) k = qc(Q, R, t->core[from], rows, cols, true);      // :843
		else qr(Q, R, t->core[from], rows, cols, true);                              // :845
		Q.n = rows * k;
		const size_t ncols = t->ext(to) * t->rank[to + 1];
		DBuf nt(k * ncols);
		gemm(nt, ncols, k, ncols, 1.0, R, cols, false, cols, t->core[to], ncols, false, 0.0);   // to = R * to  (:877)
		t->core[from] = std::move(Q);
		t->core[to] = std::move(nt);
		t->rank[from + 1] = k;
	} else if (from == to + 1) {
		const size_t rows = t->rank[from], cols = t->ext(from) * t->rank[from + 1];
		const size_t kmax = std::min(rows, cols);
		DBuf L(rows * kmax), Q(kmax * cols);
		size_t k = kmax;
		if (allow_rank_reduction) k = cq(L, Q, t->core[from], rows, cols);      // :835
		else lq(L, Q, t->core[from], rows, cols);                              // :837 (any orthonormal-row factor serves)
		Q.n = k * cols;
		const size_t nrows = t->rank[to] * t->ext(to);
		DBuf nt(nrows * k);
		gemm(nt, k, nrows, k, 1.0, t->core[to], rows, false, rows, L, k, false, 0.0);           // to = to * L  (:880)
		t->core[from] = std::move(Q);
		t->core[to] = std::move(nt);
		t->rank[from] = k;
	} else {
		throw Error(XB_ERR_INVALID, "transfer_core: nodes are not neighbours");
	}
}

static bool exceeds_maximal_ranks(const xb_tt* t) {   // ttNetwork.cpp:349-359
	for (size_t i = 0; i < t->d; ++i) {
		const size_t e = t->ext(i);
		if (t->rank[i] > e * t->rank[i + 1] || t->rank[i + 1] > e * t->rank[i]) return true;
	}
	return false;
}

void move_core(xb_tt* t, size_t position, bool keep_rank) {   // ttNetwork.cpp:582-628
	XB_REQUIRE(position < t->d, "Illegal core-position chosen for TTNetwork");
	const bool arr = !keep_rank;
	const size_t d = t->d;
	if (t->canonicalized) {
		for (size_t n = t->core_position; n < position; ++n) transfer_core(t, n, n + 1, arr);
		for (size_t n = t->core_position; n > position; --n) transfer_core(t, n, n - 1, arr);
	} else {
		for (size_t n = 0; n < position; ++n) transfer_core(t, n, n + 1, arr);
		for (size_t n = d - 1; n > position; --n) transfer_core(t, n, n - 1, arr);
	}
	aux_join();                                                          // every new core is complete on the main stream from here on
	while (exceeds_maximal_ranks(t)) {                                   // :609-624
		for (size_t n = position; n > 0; --n) transfer_core(t, n, n - 1, arr);
		for (size_t n = 0; n + 1 < d; ++n) transfer_core(t, n, n + 1, arr);
		aux_join();
		for (size_t n = d - 1; n > position; --n) transfer_core(t, n, n - 1, arr);
	}
	t->canonicalized = true;
	t->core_position = position;
}

// TensorNetwork::round_edge as TTNetwork::round uses it (tensorNetwork.cpp:678-818): `from` is the right core,
// `to` = from - 1 the left one; Sigma goes to the left core.  When the left core is known to be left-orthonormal
// (always the case inside round(), which canonicalises to the right end first) its QC (:755) is an identity up to
// signs and is skipped: the SVD of the right core's matricisation alone gives the same truncated pair.  The
// wide-matrix reduction (the reference's CQ of `from`, :749) happens inside Svd::factor as the QR of the transpose.
static void round_edge(xb_tt* t, size_t from, size_t max_rank, double eps, double soft_threshold, bool to_orthonormal,
                       std::vector<double>* svals_out) {
	const size_t to = from - 1;
	const size_t r = t->rank[from];
	const size_t fcols = t->ext(from) * t->rank[from + 1];
	const size_t trows = t->rank[to] * t->ext(to);
	Svd svd;
	svd.soft_threshold = soft_threshold;
	svd.polish = ctx().tt_svd_polish;
	if (to_orthonormal) {
		svd.factor(t->core[from], r, fcols);
		const size_t k = svd.rank_for(max_rank, eps);                         // tensor.cpp:1464-1474
		DBuf US(r * k), nf(k * fcols), nt(trows * k);
		svd.extract(US, nf, k, true, false, nullptr);                         // from = Vt_k ; U_k S_k
		gemm(nt, k, trows, k, 1.0, t->core[to], r, false, r, US, k, false, 0.0);   // to = to * U_k S_k   (:779)
		t->core[from] = std::move(nf);
		t->core[to] = std::move(nt);
		t->rank[from] = k;
		if (svals_out) { XB_REQUIRE(!ctx().speculate, "internal: singular values are not available on the speculative path"); svals_out->assign(svd.S.begin(), svd.S.begin() + k); }
	} else {
		// general case: to = Tq * B, M = B * from, M = U S Vt  ->  to = Tq * U_k S_k, from = Vt_k
		const size_t kb = std::min(trows, r);
		DBuf Tq(trows * kb), B(kb * r), M(kb * fcols);
		qr(Tq, B, t->core[to], trows, r);
		gemm(M, fcols, kb, fcols, 1.0, B, r, false, r, t->core[from], fcols, false, 0.0);
		svd.factor(M, kb, fcols);
		const size_t k = svd.rank_for(max_rank, eps);
		DBuf US(kb * k), nf(k * fcols), nt(trows * k);
		svd.extract(US, nf, k, true, false, nullptr);
		gemm(nt, k, trows, k, 1.0, Tq, kb, false, kb, US, k, false, 0.0);
		t->core[from] = std::move(nf);
		t->core[to] = std::move(nt);
		t->rank[from] = k;
		if (svals_out) svals_out->assign(svd.S.begin(), svd.S.begin() + k);
	}
}

// taus != nullptr: TTNetwork::soft_threshold (ttNetwork.cpp:688-713) — same sweep, no rank cap, eps = 0, taus[i] for the i-th
// edge from the right (:700)
static void round_tt(xb_tt* t, const size_t* max_ranks, double eps, double* svals, size_t stride, const double* taus = nullptr) {   // ttNetwork.cpp:644-665
	XB_REQUIRE(eps >= 0.0 && eps < 1.0, "_eps must be smaller than one.");
	ctx().chol_skip = 0; ctx().chol_declines = 0;                        // a sweep's path is a function of its input alone
	const size_t d = t->d;
	const bool initial_canon = t->canonicalized;
	const size_t initial_core = t->core_position;
	move_core(t, d - 1, false);                                          // canonicalize_right (:654)
	std::vector<double> sv;
	for (size_t i = 0; i + 1 < d; ++i) {                                 // :656-658
		const size_t from = d - 1 - i, edge = from - 1;
		if (taus) round_edge(t, from, 0, 0.0, taus[i], true, nullptr);
		else round_edge(t, from, max_ranks[edge], eps, 0.0, true, svals ? &sv : nullptr);
		if (svals) {
			XB_REQUIRE(sv.size() <= stride, "svals stride too small");
			std::copy(sv.begin(), sv.end(), svals + edge * stride);
			std::fill(svals + edge * stride + sv.size(), svals + (edge + 1) * stride, 0.0);
		}
	}
	t->canonicalized = true;                                             // assume_core_position(0) (:660)
	t->core_position = 0;
	if (initial_canon) move_core(t, initial_core, false);                // :662-664
}


// ---- round plans -----------------------------------------------------------------------------------------------------
// A round() is ~250 (config 1) to ~1 600 (config 3) small launches with a host read-back per edge (the rank decision).  For a
// shape that comes back — batches of equal TTs (config 5), repeated roundings inside an iteration — the whole sweep is
// captured once into a CUDA graph and replayed: every rank decision is *speculated* to come out as it did when the plan was
// recorded (ranks = min(cap, incoming rank), no eps cut, no rank deficiency, Jacobi converged) and checked on the device
// (ctx().spec_flag); one word is read back at the end, and a raised flag sends the call down the ordinary path on the
// untouched input.  Temporaries and result cores of the captured sweep are bump-allocated from an arena the plan owns (sized
// by the first, ordinary run of the shape; never reused inside a sweep, so the side stream needs no extra ordering): the graph
// holds kernel nodes only — with stream-ordered allocations inside the capture the graphs of different workers did not
// overlap on the device.  The caller's cores are copied in and the result cores copied out by one kernel each, so a replay
// is 3 launches + 1 graph launch from the host's point of view.
struct CopyJobs { const double* src[48]; double* dst[48]; unsigned long long n[48]; int count; };
__global__ void copy_many_kernel(const CopyJobs jobs) {
	for (int j = blockIdx.y; j < jobs.count; j += gridDim.y) {
		const double* __restrict__ s = jobs.src[j]; double* __restrict__ d = jobs.dst[j];
		const unsigned long long n = jobs.n[j];
		for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) d[i] = s[i];
	}
}
static void copy_many(const std::vector<const double*>& src, const std::vector<double*>& dst, const std::vector<size_t>& n) {
	for (size_t base = 0; base < src.size(); base += 48) {
		CopyJobs jobs;
		jobs.count = int(std::min<size_t>(48, src.size() - base));
		size_t mx = 1;
		for (int j = 0; j < jobs.count; ++j) { jobs.src[j] = src[base + j]; jobs.dst[j] = dst[base + j]; jobs.n[j] = n[base + j]; mx = std::max(mx, n[base + j]); }
		const unsigned bx = unsigned(std::min<size_t>((mx + 255) / 256, 64));
		copy_many_kernel<<<dim3(bx, unsigned(jobs.count)), 256, 0, ctx().stream>>>(jobs);
		XB_LAUNCH_CHECK();
	}
}

struct RoundPlan {
	std::string key;
	uint64_t uses = 0, stamp = 0;
	bool unplannable = false;
	std::vector<size_t> ranks_out;              // d + 1 entries, as recorded by the first (ordinary) run
	bool canon_out = false; size_t core_out = 0;
	cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
	std::vector<DBuf> staging;                  // the graph's view of the caller's cores (persistent, outside the graph)
	std::vector<double*> out;                   // result cores inside the plan's arena, valid after a launch until the next one
	char* arena = nullptr; size_t arena_size = 0;   // every temporary and result of the captured sweep (bump-allocated, never reused)
	unsigned int* flag = nullptr;
	uint64_t nodes = 0;
	std::vector<char> chol_tape;                // accept / decline of every Cholesky-QR2 candidate of the recording run, in call order
	bool no_chol = false;                       // replays kept failing on a Cholesky-QR2 check: this plan is Householder only
	unsigned chol_failures = 0;
	~RoundPlan() {
		if (exec) cudaGraphExecDestroy(exec);
		if (graph) cudaGraphDestroy(graph);
		if (flag) cudaFree(flag);
		if (arena) cudaFree(arena);
	}
};

}  // namespace xb
void xb::release_plans(Context& c) {
	for (RoundPlan* p : c.plans) delete p;
	c.plans.clear();
}
namespace xb {

static std::string plan_key(const xb_tt* t, const size_t* max_ranks, double eps) {
	std::string k;
	auto put = [&](const void* p, size_t n) { k.append(reinterpret_cast<const char*>(p), n); };
	const size_t head[4] = {t->d, size_t(t->is_operator), size_t(t->canonicalized), t->canonicalized ? t->core_position : 0};
	put(head, sizeof head);
	put(t->dim_m.data(), t->d * sizeof(size_t)); put(t->dim_n.data(), t->d * sizeof(size_t));
	put(t->rank.data(), (t->d + 1) * sizeof(size_t));
	if (t->d > 1) put(max_ranks, (t->d - 1) * sizeof(size_t));
	put(&eps, sizeof eps);
	put(&ctx().options_epoch, sizeof(uint64_t));
	return k;
}

static void capture_plan(RoundPlan& pl, const xb_tt* t, const size_t* max_ranks, double eps) {
	Context& c = ctx();
	const size_t d = t->d;
	pl.staging.resize(d);
	for (size_t i = 0; i < d; ++i) pl.staging[i].resize(t->core_size(i));
	if (cudaMalloc(reinterpret_cast<void**>(&pl.flag), 4 * sizeof(unsigned int)) != cudaSuccess ||
	    cudaMalloc(reinterpret_cast<void**>(&pl.arena), pl.arena_size) != cudaSuccess) {
		cudaGetLastError();                                       // no memory for the arena: this shape stays on the ordinary path
		pl.unplannable = true; pl.staging.clear();
		return;
	}
	XB_CUDA(cudaStreamSynchronize(c.stream));
	xb_tt proto;
	proto.d = d; proto.is_operator = t->is_operator; proto.dim_m = t->dim_m; proto.dim_n = t->dim_n; proto.rank = t->rank;
	proto.canonicalized = t->canonicalized; proto.core_position = t->core_position;
	proto.core.resize(d);
	const uint64_t launches0 = c.launches;
	bool ok = false;
	std::string why;
	XB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
	c.speculate = true; c.spec_flag = pl.flag;
	c.chol_tape = pl.chol_tape; c.chol_tape_pos = 0; c.chol_tape_mode = 2; c.chol_off = pl.no_chol;
	c.arena = pl.arena; c.arena_size = pl.arena_size; c.arena_off = 0; c.arena_on = true;
	try {
		XB_CUDA(cudaMemsetAsync(pl.flag, 0, 4 * sizeof(unsigned int), c.stream));
		std::vector<const double*> src; std::vector<double*> dst; std::vector<size_t> n;
		for (size_t i = 0; i < d; ++i) {
			proto.core[i].resize(t->core_size(i));                       // allocated inside the capture: the sweep may free it
			src.push_back(pl.staging[i].p); dst.push_back(proto.core[i].p); n.push_back(t->core_size(i));
		}
		copy_many(src, dst, n);
		round_tt(&proto, max_ranks, eps, nullptr, 0);
		aux_join();
		ok = true;
	} catch (const std::exception& e) { why = e.what(); }
	// the result cores live in the arena: detach them from their DBufs (arena_on is still set: nothing below reaches cudaFreeAsync)
	pl.out.clear();
	for (size_t i = 0; i < d; ++i) { pl.out.push_back(proto.core[i].p); proto.core[i].p = nullptr; proto.core[i].n = 0; }
	c.speculate = false; c.spec_flag = nullptr;
	c.chol_tape_mode = 0; c.chol_off = false; c.chol_tape.clear();
	c.arena_on = false; c.arena = nullptr;
	cudaGraph_t g = nullptr;
	const cudaError_t e_end = cudaStreamEndCapture(c.stream, &g);
	pl.nodes = c.launches - launches0;
	c.launches = launches0;
	if (!ok || e_end != cudaSuccess || !g) {
		if (g) cudaGraphDestroy(g);
		cudaGetLastError();
		pl.unplannable = true; pl.staging.clear();
		if (getenv("XB_DEBUG_PLAN")) fprintf(stderr, "[plan] capture failed: %s %s\n", why.c_str(), e_end != cudaSuccess ? cudaGetErrorString(e_end) : "");
		return;
	}
	pl.graph = g;
	const cudaError_t e_inst = cudaGraphInstantiate(&pl.exec, pl.graph, 0);
	if (e_inst != cudaSuccess) {
		cudaGetLastError();
		pl.exec = nullptr; pl.unplannable = true; pl.staging.clear();
		if (getenv("XB_DEBUG_PLAN")) fprintf(stderr, "[plan] instantiate failed: %s\n", cudaGetErrorString(e_inst));
		return;
	}
	bool same = proto.rank == pl.ranks_out && proto.canonicalized == pl.canon_out && (!pl.canon_out || proto.core_position == pl.core_out);
	if (!same) { pl.unplannable = true; if (getenv("XB_DEBUG_PLAN")) fprintf(stderr, "[plan] speculated ranks differ from the recorded ones\n"); }
	if (getenv("XB_DEBUG_PLAN")) fprintf(stderr, "[plan] captured: %llu kernels\n", (unsigned long long)pl.nodes);
}

// A replay whose outcome has not been looked at yet: the graph, the read-back of the flag and the copy of the result cores are
// enqueued; finish() runs after the worker's stream has been synchronised.
struct PendingRound {
	xb_tt* t = nullptr;
	std::vector<size_t> ranks_out; bool canon_out = false; size_t core_out = 0;   // copied from the plan (which may be evicted meanwhile)
	std::vector<DBuf> fresh;
	std::string key;
	unsigned int* h_flag = nullptr;
	std::vector<size_t> max_ranks; double eps = 0.0;
};
constexpr size_t XB_PENDING_SLOTS = 512;       // pinned flag words per worker = replays in flight per worker

static RoundPlan* plan_lookup(const xb_tt* t, const size_t* max_ranks, double eps) {
	Context& c = ctx();
	const std::string key = plan_key(t, max_ranks, eps);
	RoundPlan* pl = nullptr;
	static thread_local uint64_t clock = 0;
	for (RoundPlan* q : c.plans) if (q->key == key) pl = q;
	if (!pl) {
		if (c.plans.size() >= 6) {                                    // evict the least recently used plan
			size_t lru = 0;
			for (size_t i = 1; i < c.plans.size(); ++i) if (c.plans[i]->stamp < c.plans[lru]->stamp) lru = i;
			XB_CUDA(cudaStreamSynchronize(c.stream));
			delete c.plans[lru];
			c.plans.erase(c.plans.begin() + lru);
		}
		pl = new RoundPlan();
		pl->key = key;
		c.plans.push_back(pl);
	}
	pl->stamp = ++clock;
	pl->uses += 1;
	return pl;
}

// 0: not plannable (the caller takes the ordinary path); 1: done, synchronously (first sight of the shape: ordinary path that
// records what the plan will speculate on); 2: a replay has been enqueued, `out` must be finished after a synchronisation
static int round_plan_enqueue(xb_tt* t, const size_t* max_ranks, double eps, PendingRound& out, size_t slot) {
	Context& c = ctx();
	if (!c.round_plans || c.profile || c.speculate || t->d < 2 || t->d > 96) return 0;
	RoundPlan* pl = plan_lookup(t, max_ranks, eps);
	if (pl->unplannable) return 0;
	if (pl->uses == 1) {
		// the bytes the ordinary path allocates size the plan's arena
		size_t in_bytes = 0;
		for (size_t i = 0; i < t->d; ++i) in_bytes += (t->core_size(i) * sizeof(double) + 255) / 256 * 256;
		c.count_allocs = true; c.alloc_counter = 0;
		c.chol_tape.clear(); c.chol_tape_mode = 1; c.chol_off = pl->no_chol;
		try { round_tt(t, max_ranks, eps, nullptr, 0); } catch (...) { c.count_allocs = false; c.chol_tape_mode = 0; c.chol_off = false; throw; }
		c.count_allocs = false;
		c.chol_tape_mode = 0; c.chol_off = false;
		pl->chol_tape.swap(c.chol_tape); c.chol_tape.clear();
		pl->arena_size = c.alloc_counter + in_bytes + (1u << 20);
		pl->ranks_out = t->rank; pl->canon_out = t->canonicalized; pl->core_out = t->core_position;
		if (pl->arena_size > (size_t(8) << 30)) pl->unplannable = true;      // sweeps that allocate more than 8 GB stay on the ordinary path
		return 1;
	}
	if (!pl->exec) {
		capture_plan(*pl, t, max_ranks, eps);
		if (pl->unplannable || !pl->exec) return 0;
	}
	const size_t d = t->d;
	{
		std::vector<const double*> src; std::vector<double*> dst; std::vector<size_t> n;
		for (size_t i = 0; i < d; ++i) { src.push_back(t->core[i].p); dst.push_back(pl->staging[i].p); n.push_back(t->core_size(i)); }
		copy_many(src, dst, n);
	}
	XB_CUDA(cudaGraphLaunch(pl->exec, c.stream));
	c.launches += pl->nodes;
	if (!c.h_flags) XB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c.h_flags), XB_PENDING_SLOTS * sizeof(unsigned int)));
	out.h_flag = c.h_flags + (slot % XB_PENDING_SLOTS);
	*out.h_flag = 0xFFFFFFFFu;
	XB_CUDA(cudaMemcpyAsync(out.h_flag, pl->flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream));
	// result cores: sized by the recorded ranks; copied out of the plan's arena (valid until the next launch of this plan, which
	// is ordered behind this copy on the same stream)
	out.fresh.clear(); out.fresh.resize(d);
	{
		std::vector<const double*> src; std::vector<double*> dst; std::vector<size_t> n;
		for (size_t i = 0; i < d; ++i) {
			const size_t sz = pl->ranks_out[i] * t->ext(i) * pl->ranks_out[i + 1];
			out.fresh[i].resize(sz);
			src.push_back(pl->out[i]); dst.push_back(out.fresh[i].p); n.push_back(sz);
		}
		copy_many(src, dst, n);
	}
	out.t = t; out.eps = eps; out.key = pl->key;
	out.ranks_out = pl->ranks_out; out.canon_out = pl->canon_out; out.core_out = pl->core_out;
	out.max_ranks.assign(max_ranks, max_ranks + (d - 1));
	return 2;
}

// after the stream has been synchronised: install the result, or — speculation failed — round the untouched TT the ordinary way
static void round_plan_finish(PendingRound& p) {
	if (*p.h_flag != 0) {
		if (getenv("XB_DEBUG_PLAN")) fprintf(stderr, "[plan] speculation failed (reason %u): ordinary path\n", *p.h_flag);
		if (*p.h_flag == 4u) {
			// a Cholesky-QR2 step met input it had to decline.  The second time this happens the plan is dropped and recorded again
			// without that path (the caller has synchronised the stream: nothing of the plan is in flight)
			Context& c = ctx();
			for (size_t i = 0; i < c.plans.size(); ++i) {
				RoundPlan* pl = c.plans[i];
				if (pl->key != p.key || pl->no_chol) continue;
				if (++pl->chol_failures >= 2) {
					RoundPlan* fresh_plan = new RoundPlan();
					fresh_plan->key = pl->key; fresh_plan->no_chol = true; fresh_plan->stamp = pl->stamp;
					delete pl;
					c.plans[i] = fresh_plan;
				}
				break;
			}
		}
		p.fresh.clear();
		round_tt(p.t, p.max_ranks.data(), p.eps, nullptr, 0);
		return;
	}
	for (size_t i = 0; i < p.t->d; ++i) p.t->core[i] = std::move(p.fresh[i]);
	p.t->rank = p.ranks_out; p.t->canonicalized = p.canon_out; p.t->core_position = p.core_out;
}

// round() through a plan where there is one; false: the caller takes the ordinary path
static bool round_planned(xb_tt* t, const size_t* max_ranks, double eps) {
	PendingRound p;
	const int how = round_plan_enqueue(t, max_ranks, eps, p, 0);
	if (how == 0) return false;
	if (how == 2) {
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
		round_plan_finish(p);
	}
	return true;
}

// ---- batches of independent items (BASELINE config 5) -------------------------------------------------------------------
// The items of a batch run on W library workers (CUDA stream, memory pool, plan cache each; option "batch_workers") that are
// driven by a few library-owned host threads (option "batch_threads", default min(W, 2); measured: 1 to 8 threads give the same
// 1 290 items/s at config 5): no interpreter and no caller-side
// threading on the hot path.  Item b goes to worker b mod W; a host thread owns every T-th worker and deals with its items in
// order.  Repeated shapes replay their round plan *asynchronously*: copy-in, graph, flag read-back and copy-out are enqueued
// and the thread moves on to the next item (of another worker), so the number of items in flight on the GPU is W, not the
// number of host threads — on a box with 4 cores per GPU 16 threads that each block in a synchronisation starve each other.
// The flags are looked at once per worker at the end; an item whose speculation failed is redone on the ordinary path.
// The workers' streams first wait for everything the caller has enqueued (the items were produced on the caller's stream)
// and the caller's stream waits for all of them at the end, so the call composes with stream-ordered code on either side.
constexpr int XB_BATCH_WORKER_BASE = 40;     // workers 40 .. 40 + batch_workers - 1 belong to the batch executor
struct BatchItem { xb_tt* t; size_t max_rank; double eps; };
// make(b) produces the TT of item b on the calling thread's current worker (may enqueue work), the executor rounds it
template <class F> static void run_batch(size_t batch, size_t max_rank, double eps, F&& make) {
	if (batch == 0) return;
	Context& caller = ctx();
	const int W = int(std::min<size_t>(batch, size_t(std::max(1, std::min(caller.batch_workers, 24)))));
	const int T = std::max(1, std::min(W, caller.batch_threads > 0 ? caller.batch_threads : 2));
	cudaEvent_t start;
	XB_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
	XB_CUDA(cudaEventRecord(start, caller.stream));
	std::vector<cudaEvent_t> done(W, nullptr);
	std::mutex err_mutex;
	std::string err; xb_status err_code = XB_OK;
	auto body = [&](int th) {
		try {
			std::vector<std::vector<PendingRound>> pending(W);
			std::vector<char> started(W, 0);
			auto drain = [&](int w) {
				if (xb_worker_select(XB_BATCH_WORKER_BASE + w) != XB_OK) throw Error(XB_ERR_CUDA, xb_last_error());
				XB_CUDA(cudaStreamSynchronize(ctx().stream));
				for (PendingRound& p : pending[w]) round_plan_finish(p);
				pending[w].clear();
			};
			for (size_t b = 0; b < batch; ++b) {
				const int w = int(b % size_t(W));
				if (w % T != th) continue;
				if (xb_worker_select(XB_BATCH_WORKER_BASE + w) != XB_OK) throw Error(XB_ERR_CUDA, xb_last_error());
				Context& c = ctx();
				if (!started[w]) { XB_CUDA(cudaSetDevice(c.device)); XB_CUDA(cudaStreamWaitEvent(c.stream, start, 0)); started[w] = 1; }
				if (pending[w].size() >= XB_PENDING_SLOTS) drain(w);
				xb_tt* t = make(b);
				std::vector<size_t> mr(t->d > 1 ? t->d - 1 : 1, max_rank);
				PendingRound p;
				const int how = round_plan_enqueue(t, mr.data(), eps, p, pending[w].size());
				if (how == 0) round_tt(t, mr.data(), eps, nullptr, 0);
				else if (how == 2) pending[w].push_back(std::move(p));
			}
			for (int w = th; w < W; w += T) {
				if (!started[w]) continue;
				drain(w);
				aux_join();
				XB_CUDA(cudaEventCreateWithFlags(&done[w], cudaEventDisableTiming));
				XB_CUDA(cudaEventRecord(done[w], ctx().stream));
			}
		} catch (const Error& e) {
			std::lock_guard<std::mutex> lock(err_mutex);
			if (err.empty()) { err = e.what(); err_code = e.code; }
		} catch (const std::exception& e) {
			std::lock_guard<std::mutex> lock(err_mutex);
			if (err.empty()) { err = e.what(); err_code = XB_ERR_INVALID; }
		}
	};
	std::vector<std::thread> threads;
	for (int th = 0; th < T; ++th) threads.emplace_back(body, th);
	for (auto& th : threads) th.join();
	if (!err.empty()) xb_synchronize_all();      // nothing of a failed batch may still be running when the caller gets its items back
	for (int w = 0; w < W; ++w) if (done[w]) { cudaStreamWaitEvent(caller.stream, done[w], 0); cudaEventDestroy(done[w]); }
	cudaEventDestroy(start);
	if (!err.empty()) throw Error(err_code, err);
}

double tt_inner(const xb_tt* a, const xb_tt* b) {
	XB_REQUIRE(a->d == b->d && a->is_operator == b->is_operator, "TT inner product: formats differ");
	for (size_t i = 0; i < a->d; ++i) XB_REQUIRE(a->ext(i) == b->ext(i), "TT inner product: dimensions differ");
	DBuf E(1);
	fill(E, 1.0, 1);
	for (size_t i = 0; i < a->d; ++i) {
		const size_t ra = a->rank[i], rb = b->rank[i], ra2 = a->rank[i + 1], rb2 = b->rank[i + 1], e = a->ext(i);
		DBuf tmp(rb * e * ra2), E2(ra2 * rb2);
		// tmp (rb x e*ra2) = E^T (rb x ra) * A_i (ra x e*ra2)
		gemm(tmp, e * ra2, rb, e * ra2, 1.0, E, rb, true, ra, a->core[i], e * ra2, false, 0.0);
		// E2 (ra2 x rb2) = tmp^T ((rb*e) x ra2)^T * B_i ((rb*e) x rb2)
		gemm(E2, rb2, ra2, rb2, 1.0, tmp, ra2, true, rb * e, b->core[i], rb2, false, 0.0);
		E = std::move(E2);
	}
	return read_scalar(E);
}

double tt_frob_norm(const xb_tt* t) {   // ttNetwork.cpp:782-789
	if (t->canonicalized) return two_norm(t->core[t->core_position], t->core_size(t->core_position));
	return std::sqrt(std::max(0.0, tt_inner(t, t)));
}

xb_tt* tt_clone(const xb_tt* t) {
	xb_tt* c = new xb_tt();
	c->d = t->d; c->is_operator = t->is_operator; c->dim_m = t->dim_m; c->dim_n = t->dim_n; c->rank = t->rank;
	c->canonicalized = t->canonicalized; c->core_position = t->core_position;
	c->core.resize(t->d);
	for (size_t i = 0; i < t->d; ++i) { c->core[i].resize(t->core_size(i)); copy(c->core[i], t->core[i], t->core_size(i)); }
	return c;
}

// TTNetwork::operator+= (ttNetwork.cpp:797-847): block stacking, then back to the caller's core position
static xb_tt* tt_add(const xb_tt* a, const xb_tt* b, double beta_scale, bool recanonicalize) {
	XB_REQUIRE(a->d == b->d && a->is_operator == b->is_operator, "The dimensions in TT sum must coincide");
	for (size_t i = 0; i < a->d; ++i) XB_REQUIRE(a->ext(i) == b->ext(i), "The dimensions in TT sum must coincide");
	const size_t d = a->d;
	xb_tt* c = new xb_tt();
	c->d = d; c->is_operator = a->is_operator; c->dim_m = a->dim_m; c->dim_n = a->dim_n;
	c->rank.assign(d + 1, 1);
	for (size_t i = 1; i < d; ++i) c->rank[i] = a->rank[i] + b->rank[i];
	c->core.resize(d);
	for (size_t i = 0; i < d; ++i) {
		const size_t e = a->ext(i), la = a->rank[i], ra = a->rank[i + 1], lb = b->rank[i], rb = b->rank[i + 1];
		const size_t rr = c->rank[i + 1];
		c->core[i].resize(c->core_size(i));
		if (d == 1) {
			copy(c->core[i], a->core[i], e);
			axpy(c->core[i], beta_scale, b->core[i], e);
			continue;
		}
		fill(c->core[i], 0.0, c->core_size(i));
		const size_t row_off = (i == 0) ? 0 : la * e, col_off = (i == d - 1) ? 0 : ra;
		copy2d(c->core[i], rr, a->core[i], ra, la * e, ra);
		copy2d(c->core[i].p + row_off * rr + col_off, rr, b->core[i], rb, lb * e, rb);
	}
	if (beta_scale != 1.0 && d > 1) {
		// scale b's contribution once: its block of the first core
		const size_t e = a->ext(0), ra = a->rank[1], rb = b->rank[1], rr = c->rank[1];
		DBuf s(rr);
		fill(s, 1.0, ra);
		fill(s.p + ra, beta_scale, rb);
		scale_cols(c->core[0], s, e, rr, rr);
	}
	c->canonicalized = false;
	if (recanonicalize && a->canonicalized) move_core(c, a->core_position, false);
	return c;
}

// y(i&0) = A(i/2, j/2) * x(j&0): per site ((a r), m, (b s)) = sum_n A(a, m, n, b) x(r, n, s)   (ttStack.cpp:197-300)
static xb_tt* tt_apply(const xb_tt* A, const xb_tt* x) {
	XB_REQUIRE(A->is_operator && !x->is_operator && A->d == x->d, "TT apply needs a TTOperator and a TTTensor of the same order");
	const size_t d = A->d;
	xb_tt* y = new xb_tt();
	y->d = d; y->is_operator = false; y->dim_m.resize(d); y->dim_n.assign(d, 1);
	y->rank.assign(d + 1, 1);
	y->core.resize(d);
	for (size_t i = 0; i < d; ++i) {
		XB_REQUIRE(A->dim_n[i] == x->dim_m[i], "TT apply: dimensions do not coincide");
		const size_t a = A->rank[i], b = A->rank[i + 1], m = A->dim_m[i], n = A->dim_n[i], r = x->rank[i], s = x->rank[i + 1];
		y->dim_m[i] = m;
		y->rank[i + 1] = (i + 1 == d) ? 1 : b * s;
		// A (a,m,n,b) -> (a,m,b,n) ; x (r,n,s) -> (n,r,s) ; P ((a m b) x (r s)) ; (a,m,b,r,s) -> (a,r,m,b,s)
		DBuf Ap(a * m * n * b), xp(n * r * s), P(a * m * b * r * s);
		{ const size_t dims[4] = {a, m, n, b}, sh[4] = {0, 1, 3, 2}; permute(Ap, A->core[i], dims, sh, 4); }
		{ const size_t dims[3] = {r, n, s}, sh[3] = {1, 0, 2}; permute(xp, x->core[i], dims, sh, 3); }
		gemm(P, r * s, a * m * b, r * s, 1.0, Ap, n, false, n, xp, r * s, false, 0.0);
		y->core[i].resize(a * r * m * b * s);
		{ const size_t dims[5] = {a, m, b, r, s}, sh[5] = {0, 2, 3, 1, 4}; permute(y->core[i], P, dims, sh, 5); }
	}
	y->canonicalized = false;
	return y;
}

static void tt_to_dense(const xb_tt* t, DBuf& out) {   // tensorNetwork.cpp:287-306
	const size_t d = t->d;
	DBuf cur(t->core_size(0));
	copy(cur, t->core[0], t->core_size(0));
	size_t rows = t->ext(0);
	for (size_t i = 1; i < d; ++i) {
		const size_t r = t->rank[i], cols = t->ext(i) * t->rank[i + 1];
		DBuf nxt(rows * cols);
		gemm(nxt, cols, rows, cols, 1.0, cur, r, false, r, t->core[i], cols, false, 0.0);
		cur = std::move(nxt);
		rows *= t->ext(i);
	}
	if (t->is_operator && d > 1) {
		// (m1,n1,m2,n2,...) -> (m1..md, n1..nd)
		std::vector<size_t> dims(2 * d), sh(2 * d);
		for (size_t i = 0; i < d; ++i) { dims[2 * i] = t->dim_m[i]; dims[2 * i + 1] = t->dim_n[i]; sh[2 * i] = i; sh[2 * i + 1] = d + i; }
		XB_REQUIRE(2 * d <= 16, "operator too long to densify (at most 8 sites)");
		DBuf p(rows);
		permute(p, cur, dims.data(), sh.data(), 2 * d);
		cur = std::move(p);
	}
	out = std::move(cur);
}

} // namespace xb

// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

xb_status xb_tt_create(xb_tt** out, size_t d, const size_t* dims, const size_t* ranks, int is_operator) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(out && dims && d > 0, "xb_tt_create: bad arguments");
		XB_REQUIRE(d == 1 || ranks, "xb_tt_create: ranks missing");
		xb_tt* t = new xb_tt();
		t->d = d; t->is_operator = is_operator != 0;
		t->dim_m.assign(dims, dims + d);
		if (is_operator) t->dim_n.assign(dims + d, dims + 2 * d); else t->dim_n.assign(d, 1);
		t->rank.assign(d + 1, 1);
		for (size_t i = 0; i + 1 < d; ++i) { XB_REQUIRE(ranks[i] > 0, "rank 0 is illegal"); t->rank[i + 1] = ranks[i]; }
		for (size_t i = 0; i < d; ++i) XB_REQUIRE(t->dim_m[i] > 0 && t->dim_n[i] > 0, "dimension 0 is not possible");
		t->core.resize(d);
		for (size_t i = 0; i < d; ++i) { t->core[i].resize(t->core_size(i)); fill(t->core[i], 0.0, t->core_size(i)); }
		*out = t;
	});
}

xb_status xb_tt_destroy(xb_tt* tt) { return guard([&] { if (tt) { ensure_init(); delete tt; } }); }

xb_status xb_tt_clone(xb_tt** out, const xb_tt* tt) {
	return guard([&] { ensure_init(); XB_REQUIRE(out && tt, "null"); *out = tt_clone(tt); });
}

xb_status xb_tt_degree(const xb_tt* tt, size_t* d, int* is_operator) {
	return guard([&] { XB_REQUIRE(tt, "null"); if (d) *d = tt->d; if (is_operator) *is_operator = tt->is_operator; });
}
xb_status xb_tt_ranks(const xb_tt* tt, size_t* ranks) {
	return guard([&] { XB_REQUIRE(tt && (ranks || tt->d == 1), "null"); for (size_t i = 0; i + 1 < tt->d; ++i) ranks[i] = tt->rank[i + 1]; });
}
xb_status xb_tt_dims(const xb_tt* tt, size_t* dims) {
	return guard([&] {
		XB_REQUIRE(tt && dims, "null");
		for (size_t i = 0; i < tt->d; ++i) { dims[i] = tt->dim_m[i]; if (tt->is_operator) dims[tt->d + i] = tt->dim_n[i]; }
	});
}
xb_status xb_tt_core_position(const xb_tt* tt, int* canonicalized, size_t* position) {
	return guard([&] { XB_REQUIRE(tt, "null"); if (canonicalized) *canonicalized = tt->canonicalized; if (position) *position = tt->core_position; });
}
xb_status xb_tt_assume_core_position(xb_tt* tt, size_t position) {
	return guard([&] { XB_REQUIRE(tt && position < tt->d, "Illegal core position"); tt->canonicalized = true; tt->core_position = position; });
}

xb_status xb_tt_set_component(xb_tt* tt, size_t idx, const double* host_core, size_t rl, size_t rr) {
	return guard([&] {
		ensure_init();
		check_idx(tt, idx);
		XB_REQUIRE(host_core && rl > 0 && rr > 0, "set_component: bad arguments");
		XB_REQUIRE(idx > 0 || rl == 1, "first component must have left rank 1");
		XB_REQUIRE(idx + 1 < tt->d || rr == 1, "last component must have right rank 1");
		tt->rank[idx] = rl; tt->rank[idx + 1] = rr;      // link dimensions follow the written component (ttNetwork.cpp:480-488)
		tt->core[idx].resize(tt->core_size(idx));
		XB_CUDA(cudaMemcpyAsync(tt->core[idx].p, host_core, tt->core_size(idx) * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
		tt->canonicalized = tt->canonicalized && (tt->core_position == idx);   // ttNetwork.cpp:491
	});
}

xb_status xb_tt_get_component(const xb_tt* tt, size_t idx, double* host_core) {
	return guard([&] {
		ensure_init();
		check_idx(tt, idx);
		XB_REQUIRE(host_core, "null");
		XB_REQUIRE(tt->core[idx].n == tt->core_size(idx), "component dimensions are inconsistent with the bond ranks (set all components first)");
		XB_CUDA(cudaMemcpyAsync(host_core, tt->core[idx].p, tt->core_size(idx) * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
	});
}

// All components in one call: the copies are enqueued back to back and the stream is synchronised once (the per-component
// entry points above synchronise per component, which at 32 components is milliseconds of host latency).
xb_status xb_tt_set_components(xb_tt* tt, const double* const* host_cores, const size_t* ranks) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(tt && host_cores && (ranks || tt->d == 1), "null");
		for (size_t i = 0; i + 1 < tt->d; ++i) XB_REQUIRE(ranks[i] > 0, "set_components: bond ranks must be positive");
		for (size_t i = 0; i < tt->d; ++i) {
			XB_REQUIRE(host_cores[i], "null component");
			tt->rank[i] = (i == 0) ? 1 : ranks[i - 1];
			tt->rank[i + 1] = (i + 1 == tt->d) ? 1 : ranks[i];
		}
		for (size_t i = 0; i < tt->d; ++i) {
			tt->core[i].resize(tt->core_size(i));
			XB_CUDA(cudaMemcpyAsync(tt->core[i].p, host_cores[i], tt->core_size(i) * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
		}
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
		tt->canonicalized = false;                                             // ttNetwork.cpp:491 (more than the core was written)
	});
}

xb_status xb_tt_get_components(const xb_tt* tt, double* const* host_cores) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(tt && host_cores, "null");
		require_correct_format(tt);
		for (size_t i = 0; i < tt->d; ++i) {
			XB_REQUIRE(host_cores[i], "null component");
			XB_CUDA(cudaMemcpyAsync(host_cores[i], tt->core[i].p, tt->core_size(i) * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
		}
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
	});
}

xb_status xb_tt_component_size(const xb_tt* tt, size_t idx, size_t* rl, size_t* ext, size_t* rr) {
	return guard([&] { check_idx(tt, idx); if (rl) *rl = tt->rank[idx]; if (ext) *ext = tt->ext(idx); if (rr) *rr = tt->rank[idx + 1]; });
}

void require_correct_format(const xb_tt* tt) {   // ttNetwork.cpp:290-341 (the structural part that can fail here)
	XB_REQUIRE(tt, "null TT");
	for (size_t i = 0; i < tt->d; ++i) XB_REQUIRE(tt->core[i].n == tt->core_size(i), "TT is not in correct format: component size does not match its bond ranks");
}

xb_status xb_tt_move_core(xb_tt* tt, size_t position, int keep_rank) {
	return guard([&] { ensure_init(); require_correct_format(tt); move_core(tt, position, keep_rank != 0); });
}

xb_status xb_tt_round_svals(xb_tt* tt, const size_t* max_ranks, double eps, double* svals, size_t stride) {
	return guard([&] {
		ensure_init();
		require_correct_format(tt);
		XB_REQUIRE(max_ranks || tt->d == 1, "There must be exactly degree-1 maxRanks");
		XB_REQUIRE(eps >= 0.0 && eps < 1.0, "_eps must be smaller than one.");
		size_t cap = 0;
		for (size_t i = 0; i + 1 < tt->d; ++i) cap = std::max(cap, max_ranks[i]);
		PerfScope pa("TT sweep", "round", tt_shape(tt) + " -> " + (cap ? pa_str(cap) : std::string("eps")));
		if (!svals && round_planned(tt, max_ranks, eps)) return;
		round_tt(tt, max_ranks, eps, svals, stride);
	});
}
xb_status xb_tt_round(xb_tt* tt, const size_t* max_ranks, double eps) { return xb_tt_round_svals(tt, max_ranks, eps, nullptr, 0); }

xb_status xb_tt_soft_threshold(xb_tt* tt, const double* taus, int /*prevent_zero*/) {
	return guard([&] {
		ensure_init();
		require_correct_format(tt);
		XB_REQUIRE(taus || tt->d == 1, "There must be exactly degree/N-1 taus.");
		static const double none = 0.0;                                  // d == 1: no edge, nothing is read
		round_tt(tt, nullptr, 0.0, nullptr, 0, taus ? taus : &none);
	});
}

xb_status xb_tt_round_batched(xb_tt** tts, size_t batch, size_t max_rank, double eps) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(tts || batch == 0, "null");
		XB_REQUIRE(eps >= 0.0 && eps < 1.0, "_eps must be smaller than one.");
		for (size_t b = 0; b < batch; ++b) require_correct_format(tts[b]);
		run_batch(batch, max_rank, eps, [&](size_t b) { return tts[b]; });
	});
}

xb_status xb_tt_apply_round_batched(xb_tt** out, const xb_tt* A, xb_tt* const* xs, size_t batch, size_t max_rank, double eps) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE((out && xs) || batch == 0, "null");
		XB_REQUIRE(eps >= 0.0 && eps < 1.0, "_eps must be smaller than one.");
		require_correct_format(A);
		for (size_t b = 0; b < batch; ++b) { require_correct_format(xs[b]); out[b] = nullptr; }
		try {
			run_batch(batch, max_rank, eps, [&](size_t b) { out[b] = tt_apply(A, xs[b]); return out[b]; });
		} catch (...) {
			for (size_t b = 0; b < batch; ++b) { delete out[b]; out[b] = nullptr; }
			throw;
		}
	});
}

xb_status xb_tt_frob_norm(const xb_tt* tt, double* result) {
	return guard([&] { ensure_init(); require_correct_format(tt); XB_REQUIRE(result, "null"); *result = tt_frob_norm(tt); });
}
xb_status xb_tt_inner(const xb_tt* a, const xb_tt* b, double* result) {
	return guard([&] { ensure_init(); require_correct_format(a); require_correct_format(b); XB_REQUIRE(result, "null"); *result = tt_inner(a, b); });
}
xb_status xb_tt_distance(const xb_tt* a, const xb_tt* b, double* result) {
	return guard([&] {
		ensure_init(); require_correct_format(a); require_correct_format(b); XB_REQUIRE(result, "null");
		xb_tt* diff = tt_add(a, b, -1.0, false);
		try {
			if (diff->d > 1) move_core(diff, diff->d - 1, true);     // QR sweep: ||a-b|| = ||last core||, no cancellation
			*result = two_norm(diff->core[diff->d - 1], diff->core_size(diff->d - 1));
		} catch (...) { delete diff; throw; }
		delete diff;
	});
}
xb_status xb_tt_scale(xb_tt* tt, double factor) {   // ttNetwork.cpp:860-868: scales the core (or component 0)
	return guard([&] {
		ensure_init(); require_correct_format(tt);
		const size_t i = tt->canonicalized ? tt->core_position : 0;
		scale(tt->core[i], factor, tt->core_size(i));
	});
}
xb_status xb_tt_add(xb_tt** out, const xb_tt* a, const xb_tt* b) {
	return guard([&] { ensure_init(); require_correct_format(a); require_correct_format(b); XB_REQUIRE(out, "null"); *out = tt_add(a, b, 1.0, true); });
}
xb_status xb_tt_apply(xb_tt** out, const xb_tt* A, const xb_tt* x) {
	return guard([&] { ensure_init(); require_correct_format(A); require_correct_format(x); XB_REQUIRE(out, "null"); *out = tt_apply(A, x); });
}

xb_status xb_tt_to_dense(const xb_tt* tt, double* host) {
	return guard([&] {
		ensure_init(); require_correct_format(tt); XB_REQUIRE(host, "null");
		DBuf dense;
		tt_to_dense(tt, dense);
		size_t n = 1;
		for (size_t i = 0; i < tt->d; ++i) n *= tt->ext(i);
		XB_CUDA(cudaMemcpyAsync(host, dense.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
	});
}

// TT-SVD constructor TTNetwork(Tensor, eps, maxRanks) (ttNetwork.cpp:112-160): successive SVDs from the right, Sigma
// pushed into the left remainder (:151-155).  Result is canonicalised with the core at position 0.
xb_status xb_tt_from_dense_ex(xb_tt** out, const double* host, size_t d, const size_t* dims, int is_operator, double eps, const size_t* max_ranks) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(out && host && dims && d > 0, "xb_tt_from_dense: bad arguments");
		XB_REQUIRE(eps >= 0.0 && eps < 1.0, "_eps must be positive and smaller than one.");       // ttNetwork.cpp:114
		const bool op = is_operator != 0;
		std::vector<size_t> ext(d);
		size_t total = 1;
		for (size_t i = 0; i < d; ++i) {
			XB_REQUIRE(dims[i] > 0 && (!op || dims[d + i] > 0), "dimension 0");
			ext[i] = dims[i] * (op ? dims[d + i] : 1);
			total *= ext[i];
		}
		for (size_t i = 0; max_ranks && i + 1 < d; ++i) XB_REQUIRE(max_ranks[i] > 0, "Maximal ranks must be strictly positive.");   // :116
		xb_tt* t = new xb_tt();
		try {
			t->d = d; t->is_operator = op; t->dim_m.assign(dims, dims + d);
			if (op) t->dim_n.assign(dims + d, dims + 2 * d); else t->dim_n.assign(d, 1);
			t->rank.assign(d + 1, 1);
			t->core.resize(d);
			DBuf remains(total);
			XB_CUDA(cudaMemcpyAsync(remains.p, host, total * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
			if (op && d > 1) {                                 // (m_1..m_d, n_1..n_d) -> (m_1,n_1,m_2,n_2,...)   (:129-135)
				XB_REQUIRE(2 * d <= 16, "operator too long to reshuffle (at most 8 sites)");
				std::vector<size_t> sh(2 * d);
				for (size_t i = 0; i < d; ++i) { sh[i] = 2 * i; sh[d + i] = 2 * i + 1; }
				DBuf p(total);
				permute(p, remains, dims, sh.data(), 2 * d);
				remains = std::move(p);
			}
			std::vector<size_t> prefix(d + 1, 1);
			for (size_t i = 0; i < d; ++i) prefix[i + 1] = prefix[i] * ext[i];
			size_t r = 1;                                     // remains is (e_0...e_pos-1) x (e_pos * r)
			for (size_t pos = d - 1; pos > 0; --pos) {
				const size_t lrows = prefix[pos], cols = ext[pos] * r;
				Svd svd;
				svd.polish = ctx().tt_svd_polish;
				svd.factor(remains, lrows, cols);
				const size_t k = truncation_rank(svd.S, max_ranks ? max_ranks[pos - 1] : 0, eps);
				DBuf US(lrows * k), Vt(k * cols);
				svd.extract(US, Vt, k, true, false, nullptr);
				t->core[pos] = std::move(Vt);
				t->rank[pos] = k;
				remains = std::move(US);
				r = k;
			}
			t->core[0] = std::move(remains);
			t->canonicalized = true; t->core_position = 0;
			for (size_t i = 0; i < d; ++i) XB_REQUIRE(t->core[i].n == t->core_size(i), "internal: TT-SVD core size");
		} catch (...) { delete t; throw; }
		*out = t;
	});
}

xb_status xb_tt_from_dense(xb_tt** out, const double* host, size_t d, const size_t* dims, double eps, size_t max_rank) {
	std::vector<size_t> mr(d > 1 ? d - 1 : 1, max_rank);
	return xb_tt_from_dense_ex(out, host, d, dims, 0, eps, max_rank ? mr.data() : nullptr);
}

} // extern "C"
