// Resident front-end: the sweeps of xerus::TTNetwork / ALSVariant carried out by the sweep layer of libxb200 (include/xb200.h)
// behind the reference's own API.  integration/patch_reference.py inserts one call to each of these hooks into a transient build-time
// copy of the reference's ttNetwork.cpp / als.cpp (the reference tree itself is never modified and none of its text is stored here):
//
//   TTNetwork<isOperator>::round(maxRanks, eps)      (src/xerus/ttNetwork.cpp:644)   -> xb200_resident::round
//   TTNetwork<isOperator>::move_core(pos, keepRank)  (src/xerus/ttNetwork.cpp:582)   -> xb200_resident::move_core
//   TTNetwork<isOperator>::soft_threshold(taus, .)   (src/xerus/ttNetwork.cpp:688)   -> xb200_resident::soft_threshold
//   TTNetwork<isOperator>::operator+=(other)         (src/xerus/ttNetwork.cpp:797)   -> xb200_resident::add
//   ALSVariant::solve(A, x, b, halfSweeps, eps, .)   (src/xerus/algorithms/als.cpp:483) -> xb200_resident::als_solve
//
// A hook returns true when it has done the work; false sends the call down the reference's own code (hooks disabled with
// XB200_RESIDENT=0, TTs below XB200_RESIDENT_MIN doubles, custom local solvers, zero-degree networks).  Inside a hook the cores
// go to the device once, the whole sweep runs there (xb_tt_round / xb_tt_move_core / xb_als_solve: no per-BLAS-call transfers),
// and the result comes back once.
#pragma once
#include <xerus.h>

namespace xb200_resident {
	bool enabled(size_t total_doubles);
	template<bool isOperator> bool round(xerus::TTNetwork<isOperator>& tt, const std::vector<size_t>& maxRanks, double eps);
	template<bool isOperator> bool move_core(xerus::TTNetwork<isOperator>& tt, size_t position, bool keepRank);
	template<bool isOperator> bool soft_threshold(xerus::TTNetwork<isOperator>& tt, const std::vector<double>& taus);
	template<bool isOperator> bool add(xerus::TTNetwork<isOperator>& tt, const xerus::TTNetwork<isOperator>& other);
	bool als_solve(const xerus::ALSVariant& variant, const xerus::TTOperator* A, xerus::TTTensor& x, const xerus::TTTensor& b,
	               size_t numHalfSweeps, double convergenceEpsilon, double& energy);
	// statistics for the tests: how many calls each hook has served
	struct Counters { size_t round = 0, move_core = 0, soft_threshold = 0, add = 0, als = 0, h2d_bytes = 0, d2h_bytes = 0; };
	const Counters& counters();
}
