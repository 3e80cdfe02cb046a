// Drop-in replacement for the reference's src/xerus/blasLapackWrapper.cpp: the same xerus::blasWrapper entry points
// (include/xerus/blasLapackWrapper.h:37-146), forwarding to the xb200 C ABI (include/xb200.h) instead of CBLAS/LAPACKE.
// Compile this file INSTEAD of blasLapackWrapper.cpp and link libxb200.so; no OpenBLAS/LAPACKE is needed any more.
// Error convention: a non-zero xb_status becomes the reference's own exception type (misc/exceptions.h:37-73).
#include <memory>
#include <tuple>
#include <cstring>

#include <xerus/misc/standard.h>
#include <xerus/misc/check.h>
#include <xerus/misc/exceptions.h>
#include <xerus/misc/stringUtilities.h>
#include <xerus/basic.h>
#include <xerus/blasLapackWrapper.h>
#include <xerus/misc/internal.h>

#include "../include/xb200.h"

namespace xerus {
	namespace blasWrapper {

		static void ok(const xb_status _status, const char* const _what) {
			if (_status != XB_OK) {
				XERUS_THROW(misc::generic_error() << _what << " failed in libxb200 (status " << _status << "): " << xb_last_error());
			}
		}

		//----------------------------------------------- LEVEL I BLAS ----------------------------------------------------------
		double one_norm(const double* const _x, const size_t _n) {
			double r = 0.0; ok(xb_one_norm(_x, _n, &r), "one_norm"); return r;
		}

		double two_norm(const double* const _x, const size_t _n) {
			double r = 0.0; ok(xb_two_norm(_x, _n, &r), "two_norm"); return r;
		}

		double dot_product(const double* const _x, const size_t _n, const double* const _y) {
			double r = 0.0; ok(xb_dot_product(_x, _n, _y, &r), "dot_product"); return r;
		}

		//----------------------------------------------- LEVEL II BLAS ---------------------------------------------------------
		void matrix_vector_product(double* const _x, const size_t _m, const double _alpha, const double* const _A, const size_t _n, const bool _transposed, const double* const _y) {
			ok(xb_matrix_vector_product(_x, _m, _alpha, _A, _n, _transposed ? 1 : 0, _y), "matrix_vector_product");
		}

		void dyadic_vector_product(double* _A, const size_t _m, const size_t _n, const double _alpha, const double* const _x, const double* const _y) {
			ok(xb_dyadic_vector_product(_A, _m, _n, _alpha, _x, _y), "dyadic_vector_product");
		}

		//----------------------------------------------- LEVEL III BLAS --------------------------------------------------------
		void matrix_matrix_product(double* const _C, const size_t _leftDim, const size_t _rightDim, const double _alpha,
				const double* const _A, const size_t _lda, const bool _transposeA, const size_t _middleDim,
				const double* const _B, const size_t _ldb, const bool _transposeB) {
			ok(xb_matrix_matrix_product(_C, _leftDim, _rightDim, _alpha, _A, _lda, _transposeA ? 1 : 0, _middleDim, _B, _ldb, _transposeB ? 1 : 0),
			   "matrix_matrix_product");
		}

		//----------------------------------------------- LAPACK ----------------------------------------------------------------
		void svd(double* const _U, double* const _S, double* const _Vt, const double* const _A, const size_t _m, const size_t _n) {
			ok(xb_svd(_U, _S, _Vt, _A, _m, _n), "svd");
		}

		void svd_destructive(double* const _U, double* const _S, double* const _Vt, double* const _A, const size_t _m, const size_t _n) {
			ok(xb_svd(_U, _S, _Vt, _A, _m, _n), "svd");
		}

		std::tuple<std::unique_ptr<double[]>, std::unique_ptr<double[]>, size_t> qc(const double* const _A, const size_t _m, const size_t _n) {
			REQUIRE(_n > 0, "Dimension n must be larger than zero");
			REQUIRE(_m > 0, "Dimension m must be larger than zero");
			const size_t maxRank = std::min(_m, _n);
			std::unique_ptr<double[]> Qmax(new double[_m*maxRank]), Cmax(new double[maxRank*_n]);
			size_t rank = 0;
			ok(xb_qc(Qmax.get(), Cmax.get(), &rank, _A, _m, _n), "qc");
			if (rank == maxRank) { return std::make_tuple(std::move(Qmax), std::move(Cmax), rank); }
			// the callee owns exactly rank-sized arrays in the reference (blasLapackWrapper.cpp:276-301)
			std::unique_ptr<double[]> Q(new double[_m*rank]), C(new double[rank*_n]);
			std::memcpy(Q.get(), Qmax.get(), _m*rank*sizeof(double));
			std::memcpy(C.get(), Cmax.get(), rank*_n*sizeof(double));
			return std::make_tuple(std::move(Q), std::move(C), rank);
		}

		std::tuple<std::unique_ptr<double[]>, std::unique_ptr<double[]>, size_t> qc_destructive(double* const _A, const size_t _m, const size_t _n) {
			return qc(_A, _m, _n);
		}

		std::tuple<std::unique_ptr<double[]>, std::unique_ptr<double[]>, size_t> cq(const double* const _A, const size_t _m, const size_t _n) {
			REQUIRE(_n > 0, "Dimension n must be larger than zero");
			REQUIRE(_m > 0, "Dimension m must be larger than zero");
			const size_t maxRank = std::min(_m, _n);
			std::unique_ptr<double[]> Cmax(new double[_m*maxRank]), Qmax(new double[maxRank*_n]);
			size_t rank = 0;
			ok(xb_cq(Cmax.get(), Qmax.get(), &rank, _A, _m, _n), "cq");
			if (rank == maxRank) { return std::make_tuple(std::move(Cmax), std::move(Qmax), rank); }
			std::unique_ptr<double[]> C(new double[_m*rank]), Q(new double[rank*_n]);
			std::memcpy(C.get(), Cmax.get(), _m*rank*sizeof(double));
			std::memcpy(Q.get(), Qmax.get(), rank*_n*sizeof(double));
			return std::make_tuple(std::move(C), std::move(Q), rank);
		}

		std::tuple<std::unique_ptr<double[]>, std::unique_ptr<double[]>, size_t> cq_destructive(double* const _A, const size_t _m, const size_t _n) {
			return cq(_A, _m, _n);
		}

		void qr(double* const _Q, double* const _R, const double* const _A, const size_t _m, const size_t _n) {
			ok(xb_qr(_Q, _R, _A, _m, _n), "qr");
		}

		void inplace_qr(double* const _AtoQ, double* const _R, const size_t _m, const size_t _n) {
			// the C ABI stages its inputs before it writes any output, so A == Q aliasing is safe
			ok(xb_qr(_AtoQ, _R, _AtoQ, _m, _n), "inplace_qr");
		}

		void qr_destructive(double* const _Q, double* const _R, double* const _A, const size_t _m, const size_t _n) {
			REQUIRE(_A != _R, "_A and _R must be different, otherwise qr call will fail.");
			ok(xb_qr(_Q, _R, _A, _m, _n), "qr");
		}

		void rq(double* const _R, double* const _Q, const double* const _A, const size_t _m, const size_t _n) {
			ok(xb_rq(_R, _Q, _A, _m, _n), "rq");
		}

		void inplace_rq(double* const _R, double* const _AtoQ, const size_t _m, const size_t _n) {
			ok(xb_rq(_R, _AtoQ, _AtoQ, _m, _n), "inplace_rq");
		}

		void rq_destructive(double* const _R, double* const _Q, double* const _A, const size_t _m, const size_t _n) {
			REQUIRE(_A != _R, "_A and _R must be different, otherwise qr call will fail.");
			ok(xb_rq(_R, _Q, _A, _m, _n), "rq");
		}

		void solve(double* const _x, const double* const _A, const size_t _m, const size_t _n, const double* const _b, const size_t _nrhs) {
			ok(xb_solve(_x, _A, _m, _n, _b, _nrhs), "solve");
		}

		void solve_least_squares(double* const _x, const double* const _A, const size_t _m, const size_t _n, const double* const _b, const size_t _p) {
			ok(xb_solve_least_squares(_x, _A, _m, _n, _b, _p), "solve_least_squares");
		}

		void solve_least_squares_destructive(double* const _x, double* const _A, const size_t _m, const size_t _n, double* const _b, const size_t _p) {
			ok(xb_solve_least_squares(_x, _A, _m, _n, _b, _p), "solve_least_squares");
		}

	} // namespace blasWrapper
} // namespace xerus
