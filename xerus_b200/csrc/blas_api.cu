// Per-call layer of the C ABI: one function per xerus::blasWrapper entry point, host pointers in and out
// (reference: include/xerus/blasLapackWrapper.h:37-146).  Each call stages its operands into HBM, runs the same
// device kernels the sweep layer uses, and copies the results back; it exists for drop-in parity,
// the measured path is the sweep layer.  Also the device-pointer (xb_dev_*) variants.
#include "xb_internal.cuh"

using namespace xb;

namespace {

struct Staged {   // host -> device copy of an input operand
	DBuf d;
	Staged(const double* h, size_t n) : d(n) {
		if (n) XB_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
	}
	operator double*() const { return d.p; }
};

void to_host(double* h, const double* d, size_t n) {
	if (n) XB_CUDA(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
}
void sync() { XB_CUDA(cudaStreamSynchronize(ctx().stream)); }

size_t span(size_t rows, size_t cols, size_t ld) { return rows == 0 ? 0 : (rows - 1) * ld + cols; }

} // namespace

extern "C" {

xb_status xb_one_norm(const double* x, size_t n, double* result) {
	return guard([&] {
		PerfScope pa("Dense BLAS", "One Norm", pa_str(n));
		ensure_init(); XB_REQUIRE(result && (x || !n), "null");
		if (!n) { *result = 0.0; return; }
		Staged dx(x, n); DBuf r(1);
		asum_dev(r, dx, n);
		*result = read_scalar(r);
	});
}

xb_status xb_two_norm(const double* x, size_t n, double* result) {
	return guard([&] {
		PerfScope pa("Dense BLAS", "Two Norm", pa_str(n));
		ensure_init(); XB_REQUIRE(result && (x || !n), "null");
		if (!n) { *result = 0.0; return; }
		Staged dx(x, n);
		*result = two_norm(dx, n);
	});
}

xb_status xb_dot_product(const double* x, size_t n, const double* y, double* result) {
	return guard([&] {
		PerfScope pa("Dense BLAS", "Dot Product", pa_str(n) + "*" + pa_str(n));
		ensure_init(); XB_REQUIRE(result && ((x && y) || !n), "null");
		if (!n) { *result = 0.0; return; }
		Staged dx(x, n), dy(y, n);
		*result = dot(dx, dy, n);
	});
}

xb_status xb_matrix_vector_product(double* x, size_t m, double alpha, const double* A, size_t n, int transposed, const double* y) {
	// x (m entries) = alpha * op(A) * y (n entries); A is stored m x n, or n x m when `transposed`
	// (blasLapackWrapper.cpp:114-131: dgemv NoTrans(m, n, lda = n) resp. Trans(n, m, lda = m))
	return guard([&] {
		PerfScope pa("Dense BLAS", "Matrix Vector Product", pa_str(m) + "x" + pa_str(n) + " * " + pa_str(n));
		ensure_init(); XB_REQUIRE(x && A && y, "null");
		Staged dA(A, m * n), dy(y, n); DBuf dx(m);
		gemm(dx, 1, m, 1, alpha, dA, transposed ? m : n, transposed != 0, n, dy, 1, false, 0.0);
		to_host(x, dx, m); sync();
	});
}

xb_status xb_dyadic_vector_product(double* A, size_t m, size_t n, double alpha, const double* x, const double* y) {
	return guard([&] {
		PerfScope pa("Dense BLAS", "Dyadic Vector Product", pa_str(m) + " o " + pa_str(n));
		ensure_init(); XB_REQUIRE(A && x && y, "null");
		Staged dx(x, m), dy(y, n); DBuf dA(m * n);
		gemm(dA, n, m, n, alpha, dx, 1, false, 1, dy, n, false, 0.0);
		to_host(A, dA, m * n); sync();
	});
}

xb_status xb_matrix_matrix_product(double* C, size_t leftDim, size_t rightDim, double alpha, const double* A, size_t lda,
                                   int transposeA, size_t middleDim, const double* B, size_t ldb, int transposeB) {
	return guard([&] {
		PerfScope pa("Dense BLAS", "Matrix-Matrix-Multiplication", pa_str(leftDim) + "x" + pa_str(middleDim) + " * " + pa_str(middleDim) + "x" + pa_str(rightDim));
		ensure_init(); XB_REQUIRE(C && A && B, "null");
		XB_REQUIRE(leftDim <= 0x7fffffffULL && rightDim <= 0x7fffffffULL && middleDim <= 0x7fffffffULL, "Dimension to large for BLAS/Lapack");
		const size_t a_rows = transposeA ? middleDim : leftDim, a_cols = transposeA ? leftDim : middleDim;
		const size_t b_rows = transposeB ? rightDim : middleDim, b_cols = transposeB ? middleDim : rightDim;
		XB_REQUIRE(lda >= a_cols && ldb >= b_cols, "leading dimension too small");
		Staged dA(A, span(a_rows, a_cols, lda)), dB(B, span(b_rows, b_cols, ldb));
		DBuf dC(leftDim * rightDim);
		gemm(dC, rightDim, leftDim, rightDim, alpha, dA, lda, transposeA != 0, middleDim, dB, ldb, transposeB != 0, 0.0);
		to_host(C, dC, leftDim * rightDim); sync();
	});
}

xb_status xb_svd(double* U, double* S, double* Vt, const double* A, size_t m, size_t n) {
	return guard([&] {
		PerfScope pa("Dense LAPACK", "Singular Value Decomposition", pa_str(m) + "x" + pa_str(n));
		ensure_init(); XB_REQUIRE(U && S && Vt && A, "null");
		XB_REQUIRE(m <= 0x7fffffffULL && n <= 0x7fffffffULL, "Dimension to large for BLAS/Lapack");
		const size_t k = std::min(m, n);
		Staged dA(A, m * n); DBuf dU(m * k), dVt(k * n);
		Svd s; s.factor(dA, m, n);
		s.extract(dU, dVt, k, false, false, nullptr);
		to_host(U, dU, m * k); to_host(Vt, dVt, k * n); sync();
		std::copy(s.S.begin(), s.S.end(), S);
	});
}

xb_status xb_qc(double* Q, double* C, size_t* rank, const double* A, size_t m, size_t n) {
	return guard([&] {
		ensure_init(); XB_REQUIRE(Q && C && rank && A, "null");
		XB_REQUIRE(m > 0 && n > 0, "Dimension m and n must be larger than zero");
		const size_t k = std::min(m, n);
		Staged dA(A, m * n); DBuf dQ(m * k), dC(k * n);
		PerfScope pa("Dense LAPACK", "QRP Factorisation", "");
		const size_t r = qc(dQ, dC, dA, m, n);
		to_host(Q, dQ, m * r); to_host(C, dC, r * n); sync();
		*rank = r;
		pa.s = pa_str(m) + "x" + pa_str(r) + " * " + pa_str(r) + "x" + pa_str(n);      // blasLapackWrapper.cpp:302
	});
}

xb_status xb_cq(double* C, double* Q, size_t* rank, const double* A, size_t m, size_t n) {
	return guard([&] {
		ensure_init(); XB_REQUIRE(Q && C && rank && A, "null");
		XB_REQUIRE(m > 0 && n > 0, "Dimension m and n must be larger than zero");
		const size_t k = std::min(m, n);
		Staged dA(A, m * n); DBuf dC(m * k), dQ(k * n);
		PerfScope pa("Dense LAPACK", "QRP Factorisation", "");
		const size_t r = cq(dC, dQ, dA, m, n);
		to_host(C, dC, m * r); to_host(Q, dQ, r * n); sync();
		*rank = r;
		pa.s = pa_str(n) + "x" + pa_str(r) + " * " + pa_str(r) + "x" + pa_str(m);      // blasLapackWrapper.cpp:368
	});
}

xb_status xb_qr(double* Q, double* R, const double* A, size_t m, size_t n) {
	return guard([&] {
		PerfScope pa("Dense LAPACK", "QR Factorisation", pa_str(m) + "x" + pa_str(n));
		ensure_init(); XB_REQUIRE(Q && R && A, "QR decomposition must not be called with null pointers");
		XB_REQUIRE(A != R, "_A and _R must be different, otherwise qr call will fail.");     // blasLapackWrapper.cpp:396
		XB_REQUIRE(m > 0 && n > 0, "Dimension m and n must be larger than zero");
		const size_t k = std::min(m, n);
		Staged dA(A, m * n); DBuf dQ(m * k), dR(k * n);
		qr(dQ, dR, dA, m, n);
		to_host(Q, dQ, m * k); to_host(R, dR, k * n); sync();
	});
}

xb_status xb_rq(double* R, double* Q, const double* A, size_t m, size_t n) {
	return guard([&] {
		PerfScope pa("Dense LAPACK", "RQ Factorisation", pa_str(m) + "x" + pa_str(n));
		ensure_init(); XB_REQUIRE(Q && R && A, "QR decomposition must not be called with null pointers");
		XB_REQUIRE(A != R, "_A and _R must be different, otherwise qr call will fail.");     // :463
		XB_REQUIRE(m > 0 && n > 0, "Dimension m and n must be larger than zero");
		const size_t k = std::min(m, n);
		Staged dA(A, m * n); DBuf dR(m * k), dQ(k * n);
		rq(dR, dQ, dA, m, n);
		to_host(R, dR, m * k); to_host(Q, dQ, k * n); sync();
	});
}

static void least_squares(double* dX, const double* dA, size_t m, size_t n, const double* dB, size_t p) {
	// minimum-norm least squares through the SVD with the reference's cut-off rcond = EPSILON (dgelsd, :701-712):
	// X = V_r S_r^-1 U_r^T B
	const double EPSILON = 8 * 2.220446049250313e-16;
	Svd s; s.factor(dA, m, n);
	size_t r = 0;
	while (r < s.S.size() && s.S[r] > EPSILON * s.S[0]) ++r;
	if (r == 0) { fill(dX, 0.0, n * p); return; }
	DBuf U(m * r), Vt(r * n), Y(r * p), invS(r);
	s.extract(U, Vt, r, false, false, nullptr);
	std::vector<double> inv(r);
	for (size_t i = 0; i < r; ++i) inv[i] = 1.0 / s.S[i];
	XB_CUDA(cudaMemcpyAsync(invS.p, inv.data(), r * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
	XB_CUDA(cudaStreamSynchronize(ctx().stream));
	gemm(Y, p, r, p, 1.0, U, r, true, m, dB, p, false, 0.0);
	scale_rows(Y, invS, r, p, p);
	gemm(dX, p, n, p, 1.0, Vt, n, true, r, Y, p, false, 0.0);
}

xb_status xb_solve_least_squares(double* x, const double* A, size_t m, size_t n, const double* b, size_t p) {
	return guard([&] {
		PerfScope pa("Dense LAPACK", "Solve Least Squares", pa_str(m) + "x" + pa_str(n) + " * " + pa_str(p));
		ensure_init(); XB_REQUIRE(x && A && b, "null");
		Staged dA(A, m * n), dB(b, m * p); DBuf dX(n * p);
		least_squares(dX, dA, m, n, dB, p);
		to_host(x, dX, n * p); sync();
	});
}

xb_status xb_solve(double* x, const double* A, size_t m, size_t n, const double* b, size_t nrhs) {
	return guard([&] {
		ensure_init(); XB_REQUIRE(x && A && b, "null");
		XB_REQUIRE(m <= 0x7fffffffULL && n <= 0x7fffffffULL && nrhs <= 0x7fffffffULL, "Dimension to large for BLAS/Lapack");
		PerfScope pa("Dense LAPACK", "Solve (PLU)", pa_str(n) + "x" + pa_str(n) + "x" + pa_str(nrhs));      // :582 / :622
		Staged dA(A, m * n), dB(b, m * nrhs);
		if (m != n) {                                             // :553-559
			pa.n = "Solve Least Squares"; pa.s = pa_str(m) + "x" + pa_str(n) + " * " + pa_str(nrhs);
			DBuf dX(n * nrhs);
			least_squares(dX, dA, m, n, dB, nrhs);
			to_host(x, dX, n * nrhs); sync();
			return;
		}
		bool symmetric = false, definite = false;
		probe_symmetry(dA, n, symmetric, definite);              // :562, :590
		bool done = false;
		if (symmetric && definite) {
			DBuf Ac(n * n), Bc(n * nrhs);
			copy(Ac, dA, n * n); copy(Bc, dB, n * nrhs);
			if (cholesky_solve(Ac, Bc, n, nrhs)) { to_host(x, Bc, n * nrhs); sync(); done = true; pa.n = "Solve (Cholesky)"; }
		}
		if (!done) {   // LU with partial pivoting: the reference's dgesv branch, also standing in for dsysv (:638)
			lu_solve(dA, dB, n, nrhs);
			to_host(x, dB, n * nrhs); sync();
		}
	});
}

xb_status xb_reshuffle(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree) {
	return guard([&] {
		ensure_init(); XB_REQUIRE(out && in && (degree == 0 || (dims && shuffle)), "null");
		size_t n = 1;
		for (size_t i = 0; i < degree; ++i) n *= dims[i];
		Staged dIn(in, n); DBuf dOut(n);
		permute(dOut, dIn, dims, shuffle, degree);
		to_host(out, dOut, n); sync();
	});
}

// ---- device layer ---------------------------------------------------------------------------------------------------
xb_status xb_dev_gemm(double* C, size_t ldc, size_t m, size_t n, double alpha, const double* A, size_t lda, int transA,
                      size_t k, const double* B, size_t ldb, int transB, double beta) {
	return guard([&] { ensure_init(); gemm(C, ldc, m, n, alpha, A, lda, transA != 0, k, B, ldb, transB != 0, beta); });
}
xb_status xb_dev_qr(double* Q, double* R, const double* A, size_t m, size_t n) {
	return guard([&] { ensure_init(); qr(Q, R, A, m, n); });
}
xb_status xb_dev_lq(double* L, double* Q, const double* A, size_t m, size_t n) {
	return guard([&] { ensure_init(); lq(L, Q, A, m, n); });
}
xb_status xb_dev_svd(double* U, double* S, double* Vt, const double* A, size_t m, size_t n, size_t k_out, int scale_u,
                     int scale_vt, int* sweeps) {
	return guard([&] {
		ensure_init();
		Svd s; s.factor(A, m, n);
		const size_t k = k_out ? std::min(k_out, s.kmax) : s.kmax;
		s.extract(U, Vt, k, scale_u != 0, scale_vt != 0, S);
		if (sweeps) *sweeps = s.sweeps;
	});
}
xb_status xb_dev_reshuffle(double* out, const double* in, const size_t* dims, const size_t* shuffle, size_t degree) {
	return guard([&] { ensure_init(); permute(out, in, dims, shuffle, degree); });
}
xb_status xb_dev_two_norm(const double* x, size_t n, double* host_result) {
	return guard([&] { ensure_init(); XB_REQUIRE(host_result, "null"); *host_result = two_norm(x, n); });
}

} // extern "C"
