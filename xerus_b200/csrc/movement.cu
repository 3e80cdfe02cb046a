// Memory-bound kernels of the TT hot path: copies, fills, transposes, general mode permutation (xerus::reshuffle,
// reference: src/xerus/indexedTensor_tensor_evaluate.cpp:55-137), diagonal scalings (the sparse-diagonal S * Vt
// product of round_edge, src/xerus/tensorNetwork.cpp:769), axpy/scal (misc/basicArraySupport.h:53-111) and the
// level-1 reductions (cblas_dnrm2/ddot/dasum behind blasLapackWrapper.cpp:76-112).  All HBM/L2-bound: coalesced,
// grid-stride, grids sized in multiples of the SM count; reductions are deterministic (fixed two-stage tree).
#include "xb_internal.cuh"

namespace xb {

static inline unsigned grid_for(size_t n, unsigned threads, unsigned per_sm = 8) {
	size_t blocks = (n + threads - 1) / threads;
	const size_t cap = size_t(ctx().num_sms) * per_sm;
	if (blocks > cap) blocks = cap;
	if (blocks == 0) blocks = 1;
	return unsigned(blocks);
}

__global__ void copy_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void copy2d_kernel(double* __restrict__ dst, size_t ldd, const double* __restrict__ src, size_t lds, size_t rows, size_t cols) {
	const size_t n = rows * cols;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / cols, c = i % cols;
		dst[r * ldd + c] = src[r * lds + c];
	}
}
__global__ void fill_kernel(double* __restrict__ dst, double v, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void identity_kernel(double* __restrict__ dst, size_t rows, size_t cols, size_t ld) {
	const size_t n = rows * cols;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / cols, c = i % cols;
		dst[r * ld + c] = (r == c) ? 1.0 : 0.0;
	}
}
__global__ void scale_kernel(double* __restrict__ x, double a, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= a;
}
__global__ void axpy_kernel(double* __restrict__ y, double a, const double* __restrict__ x, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void scale_rows_kernel(double* __restrict__ A, const double* __restrict__ s, size_t rows, size_t cols, size_t ld) {
	const size_t n = rows * cols;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / cols, c = i % cols;
		A[r * ld + c] *= s[r];
	}
}
__global__ void scale_cols_kernel(double* __restrict__ A, const double* __restrict__ s, size_t rows, size_t cols, size_t ld) {
	const size_t n = rows * cols;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const size_t r = i / cols, c = i % cols;
		A[r * ld + c] *= s[c];
	}
}

// out (cols x rows) = in^T ; 32x32 tiles through shared memory, both sides coalesced.  REVERSE additionally flips
// both index ranges (used to express RQ through QR: out(j,i) = in(rows-1-i, cols-1-j)).
template <bool REVERSE>
__global__ void transpose_kernel(double* __restrict__ out, const double* __restrict__ in, size_t rows, size_t cols, const int swap_xy) {
	__shared__ double tile[32][33];
	// swap_xy: the row tiles of a very tall input sit on grid.x (limit 2^31 - 1; grid.y stops at 65535)
	const size_t c0 = (size_t)(swap_xy ? blockIdx.y : blockIdx.x) * 32, r0 = (size_t)(swap_xy ? blockIdx.x : blockIdx.y) * 32;
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		const size_t r = r0 + dy, c = c0 + threadIdx.x;
		if (r < rows && c < cols) tile[dy][threadIdx.x] = in[r * cols + c];
	}
	__syncthreads();
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		const size_t c = c0 + dy, r = r0 + threadIdx.x;   // out element (c, r)
		if (r < rows && c < cols) {
			if (REVERSE) out[(cols - 1 - c) * rows + (rows - 1 - r)] = tile[threadIdx.x][dy];
			else out[c * rows + r] = tile[threadIdx.x][dy];
		}
	}
}

struct PermuteArgs {
	int degree;
	unsigned long long out_dims[16];     // dimensions of the output tensor
	unsigned long long in_strides[16];   // stride in the input of the mode that became output mode i
};
// General mode permutation as a gather: consecutive threads write consecutive output elements.
__global__ void permute_kernel(double* __restrict__ out, const double* __restrict__ in, const PermuteArgs a, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		size_t rem = i, off = 0;
#pragma unroll 1
		for (int d = a.degree - 1; d >= 0; --d) {
			const size_t idx = rem % a.out_dims[d];
			rem /= a.out_dims[d];
			off += idx * a.in_strides[d];
		}
		out[i] = in[off];
	}
}

// ---- deterministic reductions -----------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 296;

// MODE 0: sum x*y   1: sum |x|   2: max |x|   3: sum (x * *y)^2  (y points to a device scalar: overflow-safe 2-norm)
template <int MODE>
__global__ void reduce_kernel(double* __restrict__ result, double* __restrict__ partial, unsigned int* __restrict__ counter,
                              const double* __restrict__ x, const double* __restrict__ y, size_t n) {
	__shared__ double sh[RED_THREADS / 32];
	__shared__ bool last;
	double acc = 0.0;
	const double sc = (MODE == 3) ? y[0] : 1.0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		if (MODE == 0) acc += x[i] * y[i];
		else if (MODE == 1) acc += fabs(x[i]);
		else if (MODE == 2) acc = fmax(acc, fabs(x[i]));
		else { const double v = x[i] * sc; acc += v * v; }
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const double other = __shfl_down_sync(0xffffffffu, acc, o);
		acc = (MODE == 2) ? fmax(acc, other) : acc + other;
	}
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int w = 0; w < RED_THREADS / 32; ++w) s = (MODE == 2) ? fmax(s, sh[w]) : s + sh[w];
		partial[blockIdx.x] = s;
		// the last block to arrive finishes the reduction (one launch instead of two); the counter wraps back to zero
		__threadfence();
		last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
	}
	__syncthreads();
	if (last && threadIdx.x == 0) {
		__threadfence();
		// fixed order over the block partials: deterministic.  MODE 2: the result is turned into an exact power-of-two scaling
		// pair  result[0] = s with amax * s in [0.5, 1)   result[1] = 1 / s          (s = 1 for an all-zero input)
		double s = 0.0;
		for (unsigned i = 0; i < gridDim.x; ++i) { const double v = __ldcg(partial + i); s = (MODE == 2) ? fmax(s, v) : s + v; }
		if (MODE == 2) {
			int e = 0;
			double scl = 1.0, inv = 1.0;
			if (s > 0.0 && s < HUGE_VAL) { frexp(s, &e); scl = ldexp(1.0, -e); inv = ldexp(1.0, e); }
			result[0] = scl; result[1] = inv;
		} else {
			*result = s;
		}
	}
}
__global__ void scale_by_dev_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n, const double* __restrict__ factor) {
	const double f = *factor;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * f;
}
__global__ void scale_block_by_dev_kernel(double* __restrict__ A, size_t ld, size_t rows, size_t cols, const double* __restrict__ factor) {
	const double f = *factor;
	const size_t n = rows * cols;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) A[(i / cols) * ld + (i % cols)] *= f;
}

static void reduce(double* d_result, const double* x, const double* y, size_t n, int mode) {
	// the side stream (AuxScope) has its own scratch: block partials and the arrival counter of two reductions in flight on the
	// two streams of a worker must not be shared
	Context& c = ctx();
	double*& partial = (c.aux && c.stream == c.aux) ? c.red_partial_aux : c.red_partial;
	if (!partial) {
		if (c.arena_on) throw SpecUnsupported("reduction scratch must exist before a plan is captured");
		partial = dalloc(RED_MAX_BLOCKS + 2);                 // block partials | arrival counter
		XB_CUDA(cudaMemsetAsync(partial + RED_MAX_BLOCKS, 0, 2 * sizeof(double), ctx().stream));
	}
	unsigned int* counter = reinterpret_cast<unsigned int*>(partial + RED_MAX_BLOCKS);
	unsigned blocks = unsigned(std::min<size_t>(RED_MAX_BLOCKS, std::max<size_t>(1, (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4))));
	if (mode == 0) reduce_kernel<0><<<blocks, RED_THREADS, 0, ctx().stream>>>(d_result, partial, counter, x, y, n);
	else if (mode == 1) reduce_kernel<1><<<blocks, RED_THREADS, 0, ctx().stream>>>(d_result, partial, counter, x, y, n);
	else if (mode == 2) reduce_kernel<2><<<blocks, RED_THREADS, 0, ctx().stream>>>(d_result, partial, counter, x, y, n);
	else reduce_kernel<3><<<blocks, RED_THREADS, 0, ctx().stream>>>(d_result, partial, counter, x, y, n);
	XB_LAUNCH_CHECK();
}

// ---- host wrappers ----------------------------------------------------------------------------------------------
void copy(double* dst, const double* src, size_t n) {
	if (!n || dst == src) return;
	XB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
}
void copy2d(double* dst, size_t ldd, const double* src, size_t lds, size_t rows, size_t cols) {
	if (!rows || !cols) return;
	copy2d_kernel<<<grid_for(rows * cols, 256), 256, 0, ctx().stream>>>(dst, ldd, src, lds, rows, cols);
	XB_LAUNCH_CHECK();
}
void fill(double* dst, double v, size_t n) {
	if (!n) return;
	fill_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(dst, v, n);
	XB_LAUNCH_CHECK();
}
void set_identity(double* dst, size_t rows, size_t cols, size_t ld) {
	if (!rows || !cols) return;
	identity_kernel<<<grid_for(rows * cols, 256), 256, 0, ctx().stream>>>(dst, rows, cols, ld);
	XB_LAUNCH_CHECK();
}
void scale(double* x, double a, size_t n) {
	if (!n || a == 1.0) return;
	scale_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(x, a, n);
	XB_LAUNCH_CHECK();
}
void axpy(double* y, double a, const double* x, size_t n) {
	if (!n) return;
	axpy_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(y, a, x, n);
	XB_LAUNCH_CHECK();
}
void scale_rows(double* A, const double* s, size_t rows, size_t cols, size_t ld) {
	if (!rows || !cols) return;
	scale_rows_kernel<<<grid_for(rows * cols, 256), 256, 0, ctx().stream>>>(A, s, rows, cols, ld);
	XB_LAUNCH_CHECK();
}
void scale_cols(double* A, const double* s, size_t rows, size_t cols, size_t ld) {
	if (!rows || !cols) return;
	scale_cols_kernel<<<grid_for(rows * cols, 256), 256, 0, ctx().stream>>>(A, s, rows, cols, ld);
	XB_LAUNCH_CHECK();
}
void transpose(double* out, const double* in, size_t rows, size_t cols) {
	if (!rows || !cols) return;
	if (rows == 1 || cols == 1) { copy(out, in, rows * cols); return; }
	const size_t tc = (cols + 31) / 32, tr = (rows + 31) / 32;
	const int swap_xy = tr > 65535 ? 1 : 0;
	XB_REQUIRE((swap_xy ? tc : tr) <= 65535, "both dimensions are too large for the transpose grid");
	dim3 grid(unsigned(swap_xy ? tr : tc), unsigned(swap_xy ? tc : tr)), block(32, 8);
	transpose_kernel<false><<<grid, block, 0, ctx().stream>>>(out, in, rows, cols, swap_xy);
	XB_LAUNCH_CHECK();
}
void transpose_reverse(double* out, const double* in, size_t rows, size_t cols) {
	if (!rows || !cols) return;
	const size_t tc = (cols + 31) / 32, tr = (rows + 31) / 32;
	const int swap_xy = tr > 65535 ? 1 : 0;
	XB_REQUIRE((swap_xy ? tc : tr) <= 65535, "both dimensions are too large for the transpose grid");
	dim3 grid(unsigned(swap_xy ? tr : tc), unsigned(swap_xy ? tc : tr)), block(32, 8);
	transpose_kernel<true><<<grid, block, 0, ctx().stream>>>(out, in, rows, cols, swap_xy);
	XB_LAUNCH_CHECK();
}

// out(o, q, j) = sum_p W(q, p) in(o, p, j): a small dense matrix applied to the middle mode of a three-mode view.  This is the
// operator-core step of the matrix-free local apply (the un-contracted network of als.cpp:383-401): o = left bond and the sites
// already applied, p = (a, n), q = (m, b), j = the remaining sites and the right bond.  P and Q are a handful (r_A * n), so the
// step is pure HBM traffic: every input element is read once and every output element written once, with
// 16-byte accesses along j; W sits in shared memory and is read as a broadcast.
template <int QMAX, int VEC>
__global__ void __launch_bounds__(256) mid_apply_kernel(double* __restrict__ out, const double* __restrict__ in, const double* __restrict__ W,
                                                         const size_t outer, const int P, const int Q, const size_t inner) {
	__shared__ double Ws[QMAX * 64];
	for (int e = threadIdx.x; e < P * Q; e += blockDim.x) Ws[e] = W[e];
	__syncthreads();
	const size_t jv = inner / VEC, total = outer * jv;
	for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
		const size_t o = e / jv, j = (e % jv) * VEC;
		const double* src = in + o * P * inner + j;
		double acc[QMAX][VEC];
#pragma unroll
		for (int q = 0; q < QMAX; ++q)
#pragma unroll
			for (int v = 0; v < VEC; ++v) acc[q][v] = 0.0;
#pragma unroll 4
		for (int p = 0; p < P; ++p) {
			double x[VEC];
			if (VEC == 2) { const double2 t = __ldcs(reinterpret_cast<const double2*>(src + (size_t)p * inner)); x[0] = t.x; x[VEC - 1] = t.y; }
			else x[0] = __ldcs(src + (size_t)p * inner);
#pragma unroll
			for (int q = 0; q < QMAX; ++q) {
				if (q < Q) {
					const double w = Ws[q * P + p];
#pragma unroll
					for (int v = 0; v < VEC; ++v) acc[q][v] += w * x[v];
				}
			}
		}
		double* dst = out + o * Q * inner + j;
#pragma unroll
		for (int q = 0; q < QMAX; ++q) {
			if (q < Q) {
				if (VEC == 2) *reinterpret_cast<double2*>(dst + (size_t)q * inner) = make_double2(acc[q][0], acc[q][VEC - 1]);
				else dst[(size_t)q * inner] = acc[q][0];
			}
		}
	}
}

bool mid_apply(double* out, const double* in, const double* W, size_t outer, size_t P, size_t Q, size_t inner) {
	if (P == 0 || Q == 0 || P > 64 || Q > 32) return false;
	if (outer == 0 || inner == 0) return true;
	ProfScope prof("mid_apply");
	const bool vec = inner % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
	const size_t work = outer * (vec ? inner / 2 : inner);
	const unsigned grid = unsigned(std::min<size_t>((work + 255) / 256, size_t(ctx().num_sms) * 16));
	cudaStream_t st = ctx().stream;
	const int p = int(P), q = int(Q);
	if (vec) {
		if (Q <= 8) mid_apply_kernel<8, 2><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
		else if (Q <= 16) mid_apply_kernel<16, 2><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
		else mid_apply_kernel<32, 2><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
	} else {
		if (Q <= 8) mid_apply_kernel<8, 1><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
		else if (Q <= 16) mid_apply_kernel<16, 1><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
		else mid_apply_kernel<32, 1><<<grid, 256, 0, st>>>(out, in, W, outer, p, q, inner);
	}
	XB_LAUNCH_CHECK();
	return true;
}

void permute(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree) {
	XB_REQUIRE(out != in, "in-place reshuffle is not supported at this level");
	// validate: shuffle must be a permutation (reference: indexedTensor_tensor_evaluate.cpp:57-71)
	std::vector<char> seen(degree, 0);
	size_t n = 1;
	for (size_t i = 0; i < degree; ++i) {
		XB_REQUIRE(shuffle[i] < degree && !seen[shuffle[i]], "shuffle is not a permutation");
		seen[shuffle[i]] = 1;
		n *= dims[i];
	}
	if (n == 0) return;
	// merge modes that stay adjacent (in order) so that e.g. (a,b,c)->(a,c,b) with many modes still fits 8 slots
	std::vector<size_t> in_stride(degree), inv(degree);
	for (size_t i = 0, s = 1; i < degree; ++i) { in_stride[degree - 1 - i] = s; s *= dims[degree - 1 - i]; }
	for (size_t i = 0; i < degree; ++i) inv[shuffle[i]] = i;        // output mode o came from input mode inv[o]
	std::vector<unsigned long long> od, os;
	for (size_t o = 0; o < degree; ++o) {
		const size_t im = inv[o];
		if (dims[im] == 1) continue;
		if (!od.empty() && o > 0 && inv[o - 1] + 1 == im && dims[inv[o - 1]] != 1 && os.back() == in_stride[inv[o - 1]]) {
			od.back() *= dims[im]; os.back() = in_stride[im];
		} else {
			od.push_back(dims[im]); os.push_back(in_stride[im]);
		}
	}
	// identity (possibly after dropping size-1 modes): plain copy (reference :74-77)
	bool identity = true;
	{
		unsigned long long expect = 1;
		for (size_t i = od.size(); i-- > 0;) { if (os[i] != expect) { identity = false; break; } expect *= od[i]; }
	}
	if (identity) { copy(out, in, n); return; }
	if (od.size() == 2 && os[1] == od[0] && os[0] == 1) { transpose(out, in, od[1], od[0]); return; }
	XB_REQUIRE(od.size() <= 16, "reshuffle supports at most 16 non-mergeable modes");
	PermuteArgs a;
	a.degree = int(od.size());
	for (size_t i = 0; i < od.size(); ++i) { a.out_dims[i] = od[i]; a.in_strides[i] = os[i]; }
	permute_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(out, in, a, n);
	XB_LAUNCH_CHECK();
}

void dot_dev(double* d_result, const double* x, const double* y, size_t n) { reduce(d_result, x, y, n, 0); }
void asum_dev(double* d_result, const double* x, size_t n) { reduce(d_result, x, x, n, 1); }

void amax_scale_dev(double* d_scale2, const double* x, size_t n) { reduce(d_scale2, x, x, n, 2); }
void scale_by_dev(double* dst, const double* src, size_t n, const double* d_factor) {
	if (!n) return;
	scale_by_dev_kernel<<<grid_for(n, 256), 256, 0, ctx().stream>>>(dst, src, n, d_factor);
	XB_LAUNCH_CHECK();
}
void scale_block_by_dev(double* A, size_t ld, size_t rows, size_t cols, const double* d_factor) {
	if (!rows || !cols) return;
	scale_block_by_dev_kernel<<<grid_for(rows * cols, 256), 256, 0, ctx().stream>>>(A, ld, rows, cols, d_factor);
	XB_LAUNCH_CHECK();
}

double dot(const double* x, const double* y, size_t n) {
	DBuf r(1);
	dot_dev(r, x, y, n);
	return read_scalar(r);
}
double two_norm(const double* x, size_t n) {
	// overflow/underflow safe (TT cores carry norms like 1e33 .. 1e150): scale by an exact power of two first
	if (!n) return 0.0;
	DBuf sc(2), r(1);
	amax_scale_dev(sc, x, n);
	reduce(r, x, sc.p, n, 3);
	Context& c = ctx();
	XB_CUDA(cudaMemcpyAsync(c.h_scratch, r.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaMemcpyAsync(c.h_scratch + 1, sc.p + 1, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	XB_CUDA(cudaStreamSynchronize(c.stream));
	return std::sqrt(std::max(0.0, c.h_scratch[0])) * c.h_scratch[1];
}

} // namespace xb
