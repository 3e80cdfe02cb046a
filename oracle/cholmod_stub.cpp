// TEST INFRASTRUCTURE ONLY (oracle build). The reference's dense path references five static
// CholmodSparse members from tensor.cpp (declared in include/xerus/cholmod_wrapper.h:86-138);
// SuiteSparse is not in this image, so they are defined here to throw: any test that reaches the
// sparse representation aborts loudly instead of silently computing something else.
#include <stdexcept>
#include <xerus/cholmod_wrapper.h>

namespace xerus { namespace internal {
	static void no_sparse() { throw std::runtime_error("oracle build: SuiteSparse (sparse representation) is not available"); }

	void CholmodSparse::matrix_matrix_product(std::map<size_t, double>&, const size_t, const size_t, const double,
			const std::map<size_t, double>&, const bool, const size_t, const std::map<size_t, double>&, const bool) { no_sparse(); }

	void CholmodSparse::solve_sparse_rhs(std::map<size_t, double>&, size_t, const std::map<size_t, double>&, const bool,
			const std::map<size_t, double>&, size_t) { no_sparse(); }

	void CholmodSparse::solve_dense_rhs(double*, size_t, const std::map<size_t, double>&, const bool, const double*, size_t) { no_sparse(); }

	std::tuple<std::map<size_t, double>, std::map<size_t, double>, size_t> CholmodSparse::qc(
			const std::map<size_t, double>&, const bool, size_t, size_t, bool) { no_sparse(); return {}; }

	std::tuple<std::map<size_t, double>, std::map<size_t, double>, size_t> CholmodSparse::cq(
			const std::map<size_t, double>&, const bool, size_t, size_t, bool) { no_sparse(); return {}; }
}}
