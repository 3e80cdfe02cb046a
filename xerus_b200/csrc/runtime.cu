// xb200 runtime: device context, stream, stream-ordered memory pool, error state, memory hooks of the C ABI.
#include "xb_internal.cuh"
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <cstdlib>

namespace xb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

constexpr int XB_MAX_WORKERS = 64;
static Context g_ctx;                              // worker 0
static Context* g_workers[XB_MAX_WORKERS] = {&g_ctx};
static int g_num_workers = 1;
static thread_local int tl_worker = 0;
static std::mutex g_ctx_mutex;
static std::atomic<bool> g_ready{false};        // set once worker 0 is fully initialised (read without the mutex by every entry point)
Context& ctx() { return *g_workers[tl_worker]; }

static void apply_option(Context& c, const std::string& k, double value) {
	c.options_epoch += 1;
	if (k == "svd_max_sweeps") c.svd_max_sweeps = int(value);
	else if (k == "gemm_force_small") c.gemm_force_small = int(value);
	else if (k == "gemm_big") c.gemm_big = int(value);
	else if (k == "qr_defer") c.qr_defer = int(value);
	else if (k == "svd_persistent") c.svd_persistent = int(value);
	else if (k == "svd_max_bw") c.svd_max_bw = int(value);
	else if (k == "svd_recursive") c.svd_recursive = int(value);
	else if (k == "svd_flip") c.svd_flip = int(value);
	else if (k == "svd_split") c.svd_split = int(value);
	else if (k == "svd_dsmem") c.svd_dsmem = int(value);
	else if (k == "svd_colsort") c.svd_colsort = int(value);
	else if (k == "svd_last_sweep_cos") c.svd_last_sweep_cos = value;
	else if (k == "svd_gram") c.svd_gram = int(value);
	else if (k == "als_graph") c.als_graph = int(value);
	else if (k == "als_persistent_cg") c.als_persistent_cg = int(value);
	else if (k == "als_cg_cluster") c.als_cg_cluster = int(value);
	else if (k == "als_cg_merged") c.als_cg_merged = int(value);
	else if (k == "svd_jacc") c.svd_jacc = int(value);
	else if (k == "svd_fast") c.svd_fast = int(value);
	else if (k == "qr_cluster") c.qr_cluster = int(value);
	else if (k == "qr_chol") c.qr_chol = int(value);
	else if (k == "qr_chol_min_rows") c.qr_chol_min_rows = int(value);
	else if (k == "qr_cluster_min_rows") c.qr_cluster_min_rows = int(value);
	else if (k == "svd_square_qr") c.svd_square_qr = int(value);
	else if (k == "svd_polish") c.svd_polish = int(value);
	else if (k == "tt_svd_polish") c.tt_svd_polish = int(value);
	else if (k == "als_direct_max") c.als_direct_max = int(value);
	else if (k == "round_plans") c.round_plans = int(value);
	else if (k == "small_kernels") c.small_kernels = int(value);
	else if (k == "peer_wait_spins") c.peer_wait_spins = value;
	else if (k == "batch_workers") c.batch_workers = int(value);
	else if (k == "batch_threads") c.batch_threads = int(value);
	else throw Error(XB_ERR_INVALID, "xb_set_option: unknown key " + k);
}

static void init_locked(int device) {
	if (g_ctx.initialised) return;
	// hardware work queues of the context: read by the driver when the context is created, so this only has an effect when the
	// library is the first CUDA user of the process (the Python package and bench.py set it before torch creates the context).
	// 8 queues (the default) serialise the worker streams of the batched entry points behind each other: config 5 runs at
	// 400 items/s with 8 queues and at 1200 with 32.
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		throw Error(XB_ERR_NO_DEVICE, std::string("xb200: no usable CUDA device (") + cudaGetErrorString(e) +
		            "); there is no CPU fallback for this library");
	}
	if (device < 0 || device >= count) throw Error(XB_ERR_INVALID, "xb_init: device index out of range");
	XB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	XB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		throw Error(XB_ERR_NO_DEVICE, std::string("xb200 is built for sm_100a only; found ") + prop.name);
	}
	g_ctx.device = device;
	g_ctx.num_sms = prop.multiProcessorCount;
	g_ctx.max_smem_optin = prop.sharedMemPerBlockOptin;
	XB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
	XB_CUDA(cudaDeviceGetDefaultMemPool(&g_ctx.pool, device));
	uint64_t threshold = UINT64_MAX;   // keep freed blocks cached: sweeps must not hit cudaMalloc
	XB_CUDA(cudaMemPoolSetAttribute(g_ctx.pool, cudaMemPoolAttrReleaseThreshold, &threshold));
	XB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g_ctx.h_scratch), g_ctx.h_scratch_len * sizeof(double)));
	g_ctx.initialised = true;
	g_ready.store(true, std::memory_order_release);
	// XB_OPTIONS="key=value,key=value": the knobs of xb_set_option from the environment (A/B runs of any driver)
	if (const char* env = getenv("XB_OPTIONS")) {
		std::string all(env);
		size_t pos = 0;
		while (pos < all.size()) {
			size_t end = all.find(',', pos);
			if (end == std::string::npos) end = all.size();
			const std::string item = all.substr(pos, end - pos);
			const size_t eq = item.find('=');
			if (eq != std::string::npos) apply_option(g_ctx, item.substr(0, eq), atof(item.c_str() + eq + 1));
			pos = end + 1;
		}
	}
}

void ensure_init() {
	if (g_ready.load(std::memory_order_acquire)) {
		// the caller may be a different host thread (ctypes); make the library device current
		cudaSetDevice(g_ctx.device);
		return;
	}
	std::lock_guard<std::mutex> lock(g_ctx_mutex);
	init_locked(0);
}

static void select_worker(int w) {
	XB_REQUIRE(w >= 0 && w < XB_MAX_WORKERS, "worker index out of range");
	ensure_init();
	std::lock_guard<std::mutex> lock(g_ctx_mutex);
	for (int i = g_num_workers; i <= w; ++i) {
		Context* c = new Context();
		c->initialised = true; c->worker = i; c->device = g_ctx.device; c->pool = g_ctx.pool;
		c->num_sms = g_ctx.num_sms; c->max_smem_optin = g_ctx.max_smem_optin;
		c->svd_max_sweeps = g_ctx.svd_max_sweeps; c->gemm_force_small = g_ctx.gemm_force_small; c->gemm_big = g_ctx.gemm_big; c->qr_defer = g_ctx.qr_defer;
		c->svd_persistent = g_ctx.svd_persistent; c->svd_polish = g_ctx.svd_polish; c->tt_svd_polish = g_ctx.tt_svd_polish; c->svd_recursive = g_ctx.svd_recursive; c->svd_flip = g_ctx.svd_flip; c->svd_split = g_ctx.svd_split; c->svd_dsmem = g_ctx.svd_dsmem; c->svd_colsort = g_ctx.svd_colsort; c->svd_last_sweep_cos = g_ctx.svd_last_sweep_cos; c->svd_gram = g_ctx.svd_gram; c->als_graph = g_ctx.als_graph; c->als_persistent_cg = g_ctx.als_persistent_cg; c->als_cg_cluster = g_ctx.als_cg_cluster; c->als_cg_merged = g_ctx.als_cg_merged; c->svd_jacc = g_ctx.svd_jacc; c->svd_fast = g_ctx.svd_fast; c->qr_cluster = g_ctx.qr_cluster; c->qr_cluster_min_rows = g_ctx.qr_cluster_min_rows; c->qr_chol = g_ctx.qr_chol; c->qr_chol_min_rows = g_ctx.qr_chol_min_rows; c->svd_square_qr = g_ctx.svd_square_qr;
		c->svd_max_bw = g_ctx.svd_max_bw; c->als_direct_max = g_ctx.als_direct_max; c->round_plans = g_ctx.round_plans; c->small_kernels = g_ctx.small_kernels; c->peer_wait_spins = g_ctx.peer_wait_spins; c->batch_workers = g_ctx.batch_workers; c->batch_threads = g_ctx.batch_threads;
		XB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		XB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_scratch), c->h_scratch_len * sizeof(double)));
		{
			// a memory pool per worker: with one shared pool a block freed on one worker's stream and reused on another's makes the
			// second stream wait for the first (the pool inserts the dependency), which serialises workers that should overlap
			cudaMemPoolProps props = {};
			props.allocType = cudaMemAllocationTypePinned;
			props.handleTypes = cudaMemHandleTypeNone;
			props.location.type = cudaMemLocationTypeDevice;
			props.location.id = g_ctx.device;
			XB_CUDA(cudaMemPoolCreate(&c->pool, &props));
			uint64_t threshold = UINT64_MAX;
			XB_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &threshold));
		}
		g_workers[i] = c;
		g_num_workers = i + 1;
	}
	tl_worker = w;
}

void* dalloc_bytes(size_t bytes) {
	void* p = nullptr;
	if (bytes == 0) bytes = 8;
	Context& c = ctx();
	if (c.arena_on) {
		const size_t b = (bytes + 255) / 256 * 256;
		if (c.arena_off + b > c.arena_size) throw SpecUnsupported("capture arena exhausted");
		p = c.arena + c.arena_off;
		c.arena_off += b;
		return p;
	}
	if (c.count_allocs) c.alloc_counter += (bytes + 255) / 256 * 256;
	XB_CUDA(cudaMallocFromPoolAsync(&p, bytes, c.pool, c.stream));
	return p;
}
static void ensure_aux(Context& c) {
	if (c.aux) return;
	XB_CUDA(cudaStreamCreateWithFlags(&c.aux, cudaStreamNonBlocking));
	XB_CUDA(cudaEventCreateWithFlags(&c.aux_fork_ev, cudaEventDisableTiming));
	XB_CUDA(cudaEventCreateWithFlags(&c.aux_join_ev, cudaEventDisableTiming));
}
void aux_fork() {
	Context& c = ctx();
	ensure_aux(c);
	XB_CUDA(cudaEventRecord(c.aux_fork_ev, c.stream));
	XB_CUDA(cudaStreamWaitEvent(c.aux, c.aux_fork_ev, 0));
}
void aux_join() {
	Context& c = ctx();
	if (!c.aux_pending) return;
	XB_CUDA(cudaEventRecord(c.aux_join_ev, c.aux));
	XB_CUDA(cudaStreamWaitEvent(c.stream, c.aux_join_ev, 0));
	c.aux_pending = false;
}
AuxScope::AuxScope(bool enable) : on(enable) {
	if (!on) return;
	Context& c = ctx();
	ensure_aux(c);
	saved = c.stream;
	c.stream = c.aux;
}
AuxScope::~AuxScope() {
	if (!on) return;
	Context& c = ctx();
	c.stream = saved;
	c.aux_pending = true;
}

double* dalloc(size_t n) { return static_cast<double*>(dalloc_bytes(n * sizeof(double))); }
void dfree(void* p) {
	if (!p) return;
	Context& c = ctx();
	if (c.arena_on) return;                       // arena memory lives as long as the plan
	cudaFreeAsync(p, c.stream);
}

double read_scalar(const double* d_value) {
	Context& c = ctx();
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, d_value, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	return c.h_scratch[0];
}

// ---- (group, name, shape) call registry -------------------------------------------------------------------------------
// The reference brackets every blasWrapper call with XERUS_PA_START / XERUS_PA_END(group, name, shape) and keeps
// (calls, microseconds) per shape (misc/performanceAnalysis.h:30-39, blasLapackWrapper.cpp:83-720).  The same registry at the
// C ABI, with the reference's own group / name / shape strings for the per-call layer, so shape catalogues can be compared
// one to one; the sweep layer adds the group "TT sweep".  Off by default (xb_perf_enable); host wall time of the call.
static std::atomic<bool> g_perf_on{false};
static std::mutex g_perf_mutex;
struct PerfEntry { std::string group, name, shape; uint64_t calls = 0; double us = 0.0; };
static std::map<std::string, size_t> g_perf_index;
static std::vector<PerfEntry> g_perf_entries;

PerfScope::PerfScope(const char* group, const char* name, const std::string& shape) {
	if (!g_perf_on.load(std::memory_order_relaxed)) return;
	on = true; g = group; n = name; s = shape;
	t0 = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
PerfScope::~PerfScope() {
	if (!on) return;
	const double t1 = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
	std::lock_guard<std::mutex> lock(g_perf_mutex);
	const std::string key = std::string(g) + "\x1f" + n + "\x1f" + s;
	auto it = g_perf_index.find(key);
	if (it == g_perf_index.end()) {
		it = g_perf_index.emplace(key, g_perf_entries.size()).first;
		PerfEntry e; e.group = g; e.name = n; e.shape = s;
		g_perf_entries.push_back(e);
	}
	g_perf_entries[it->second].calls += 1;
	g_perf_entries[it->second].us += t1 - t0;
}

// ---- profiling ----------------------------------------------------------------------------------------------------
#define g_prof_pending (ctx().prof_pending)
#define g_prof_totals (ctx().prof_totals)

ProfScope::ProfScope(const char* kernel_class) {
	Context& c = ctx();
	if (!c.profile) return;
	ProfRecord r;
	r.name = kernel_class;
	cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
	r.launches0 = c.launches; r.launches1 = 0;
	cudaEventRecord(r.e0, c.stream);
	slot = int(g_prof_pending.size());
	g_prof_pending.push_back(r);
}
ProfScope::~ProfScope() {
	if (slot < 0) return;
	Context& c = ctx();
	if (size_t(slot) >= g_prof_pending.size()) return;
	cudaEventRecord(g_prof_pending[slot].e1, c.stream);
	g_prof_pending[slot].launches1 = c.launches;
}

static void prof_collect() {
	Context& c = ctx();
	if (g_prof_pending.empty()) return;
	cudaStreamSynchronize(c.stream);
	for (ProfRecord& r : g_prof_pending) {
		float ms = 0.f;
		if (r.launches1 >= r.launches0 && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
			ProfTotal* t = nullptr;
			for (auto& kv : g_prof_totals) if (kv.first == r.name) t = &kv.second;
			if (!t) { g_prof_totals.emplace_back(r.name, ProfTotal()); t = &g_prof_totals.back().second; }
			t->scopes += 1; t->launches += r.launches1 - r.launches0; t->ms += ms;
		}
		cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
	}
	g_prof_pending.clear();
}

} // namespace xb

using namespace xb;

extern "C" {

xb_status xb_init(int device) {
	return guard([&] {
		std::lock_guard<std::mutex> lock(g_ctx_mutex);
		if (g_ctx.initialised) {
			XB_REQUIRE(device == g_ctx.device, "xb_init: already initialised on another device");
			return;
		}
		init_locked(device);
	});
}

xb_status xb_shutdown(void) {
	return guard([&] {
		std::lock_guard<std::mutex> lock(g_ctx_mutex);
		if (!g_ctx.initialised) return;
		g_ready.store(false, std::memory_order_release);
		// every worker: streams, events, pinned scratch, reduction scratch, its memory pool (worker 0 uses the device's default pool)
		for (int w = 0; w < g_num_workers; ++w) {
			Context* c = g_workers[w];
			cudaStreamSynchronize(c->stream);
			if (c->aux) { cudaStreamSynchronize(c->aux); cudaStreamDestroy(c->aux); }
			if (c->aux_fork_ev) cudaEventDestroy(c->aux_fork_ev);
			if (c->aux_join_ev) cudaEventDestroy(c->aux_join_ev);
			release_plans(*c);
			if (c->red_partial) cudaFree(c->red_partial);
			if (c->red_partial_aux) cudaFree(c->red_partial_aux);
			if (c->h_scratch) cudaFreeHost(c->h_scratch);
			if (c->h_flags) cudaFreeHost(c->h_flags);
			cudaStreamDestroy(c->stream);
			if (w > 0) { if (c->pool) cudaMemPoolDestroy(c->pool); delete c; g_workers[w] = nullptr; }
		}
		g_num_workers = 1;
		tl_worker = 0;
		g_ctx = Context();
	});
}

const char* xb_last_error(void) { return g_last_error.c_str(); }
int xb_version(void) { return 100; }

xb_status xb_synchronize(void) { return guard([&] { ensure_init(); XB_CUDA(cudaStreamSynchronize(ctx().stream)); }); }

xb_status xb_get_stream(void** s) { return guard([&] { ensure_init(); XB_REQUIRE(s, "null"); *s = ctx().stream; }); }

xb_status xb_kernel_launch_count(uint64_t* n) {
	return guard([&] {
		XB_REQUIRE(n, "null");
		uint64_t total = 0;
		for (int w = 0; w < g_num_workers; ++w) total += g_workers[w]->launches;
		*n = total;
	});
}

xb_status xb_worker_select(int worker) { return guard([&] { select_worker(worker); }); }

xb_status xb_synchronize_all(void) {
	return guard([&] { ensure_init(); for (int w = 0; w < g_num_workers; ++w) XB_CUDA(cudaStreamSynchronize(g_workers[w]->stream)); });
}

xb_status xb_set_option(const char* key, double value) {
	return guard([&] {
		XB_REQUIRE(key, "null key");
		const std::string k(key);
		std::lock_guard<std::mutex> lock(g_ctx_mutex);
		for (int w = 0; w < g_num_workers; ++w) apply_option(*g_workers[w], k, value);
	});
}

xb_status xb_profile_enable(int on) {
	return guard([&] {
		ensure_init();
		prof_collect();
		g_prof_totals.clear();
		ctx().profile = on != 0;
	});
}

xb_status xb_profile_get(const char* kernel_class, uint64_t* scopes, uint64_t* launches, double* milliseconds) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(kernel_class, "null");
		prof_collect();
		ProfTotal t;
		for (auto& kv : g_prof_totals) if (kv.first == kernel_class) t = kv.second;
		if (scopes) *scopes = t.scopes;
		if (launches) *launches = t.launches;
		if (milliseconds) *milliseconds = t.ms;
	});
}

xb_status xb_perf_enable(int on) { return guard([&] { g_perf_on.store(on != 0); }); }
xb_status xb_perf_reset(void) {
	return guard([&] { std::lock_guard<std::mutex> lock(g_perf_mutex); g_perf_index.clear(); g_perf_entries.clear(); });
}
xb_status xb_perf_count(size_t* n) {
	return guard([&] { XB_REQUIRE(n, "null"); std::lock_guard<std::mutex> lock(g_perf_mutex); *n = g_perf_entries.size(); });
}
xb_status xb_perf_entry(size_t i, const char** group, const char** name, const char** shape, uint64_t* calls, double* microseconds) {
	return guard([&] {
		std::lock_guard<std::mutex> lock(g_perf_mutex);
		XB_REQUIRE(i < g_perf_entries.size(), "perf entry index out of range");
		const PerfEntry& e = g_perf_entries[i];            // the strings stay valid until xb_perf_reset
		if (group) *group = e.group.c_str();
		if (name) *name = e.name.c_str();
		if (shape) *shape = e.shape.c_str();
		if (calls) *calls = e.calls;
		if (microseconds) *microseconds = e.us;
	});
}

xb_status xb_alloc(void** dptr, size_t bytes) {
	return guard([&] { ensure_init(); XB_REQUIRE(dptr, "null"); *dptr = dalloc_bytes(bytes); });
}
xb_status xb_free(void* dptr) { return guard([&] { ensure_init(); dfree(dptr); }); }
xb_status xb_alloc_host(void** hptr, size_t bytes) {
	return guard([&] { ensure_init(); XB_REQUIRE(hptr, "null"); XB_CUDA(cudaMallocHost(hptr, bytes ? bytes : 8)); });
}
xb_status xb_free_host(void* hptr) { return guard([&] { if (hptr) XB_CUDA(cudaFreeHost(hptr)); }); }
xb_status xb_prefetch(const void* host_ptr, size_t bytes) {
	return guard([&] {
		ensure_init();
		XB_REQUIRE(host_ptr && bytes > 0, "null");
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, host_ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return;      // already pinned
		cudaGetLastError();
		const cudaError_t e = cudaHostRegister(const_cast<void*>(host_ptr), bytes, cudaHostRegisterDefault);
		if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return; }
		XB_CUDA(e);
	});
}
xb_status xb_release(const void* host_ptr) {
	return guard([&] {
		if (!host_ptr) return;
		const cudaError_t e = cudaHostUnregister(const_cast<void*>(host_ptr));
		if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return; }
		XB_CUDA(e);
	});
}
xb_status xb_upload(void* dst, const void* src, size_t bytes) {
	return guard([&] { ensure_init(); if (bytes) XB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().stream)); });
}
xb_status xb_download(void* dst, const void* src, size_t bytes) {
	return guard([&] {
		ensure_init();
		if (bytes) XB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
	});
}

} // extern "C"
