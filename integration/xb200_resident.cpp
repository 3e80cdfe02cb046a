// See xb200_resident.h.  Conversion between xerus::TTNetwork and the device-resident xb_tt:
// component(i) is a dense row-major Tensor (r_l, n, r_r) / (r_l, m, n, r_r) (src/xerus/ttNetwork.cpp:96-98) whose lazy scalar
// factor (tensor.h:105) is applied on the way up; on the way down every component is set with TTNetwork::set_component
// (:457-492) and the canonicalisation state with assume_core_position (:735-739).
#include "xb200_resident.h"
#include "../include/xb200.h"
#include <cstdlib>
#include <cstring>
#include <memory>

namespace xb200_resident {

static Counters g_counters;
const Counters& counters() { return g_counters; }

static void check(xb_status s) {
	if (s != XB_OK) XERUS_THROW(xerus::misc::generic_error() << "xb200: " << xb_last_error());
}

bool enabled(size_t total_doubles) {
	static const bool on = [] { const char* e = std::getenv("XB200_RESIDENT"); return !(e && e[0] == '0'); }();
	static const size_t min_size = [] { const char* e = std::getenv("XB200_RESIDENT_MIN"); return e ? size_t(std::atoll(e)) : size_t(0); }();
	return on && total_doubles >= min_size;
}

struct Handle {
	xb_tt* h = nullptr;
	~Handle() { if (h) xb_tt_destroy(h); }
};

template<bool isOperator> static size_t total_size(const xerus::TTNetwork<isOperator>& tt) {
	const size_t d = tt.degree() / (isOperator ? 2 : 1);
	size_t n = 0;
	for (size_t i = 0; i < d; ++i) n += tt.get_component(i).size;
	return n;
}

template<bool isOperator> static void to_device(Handle& H, const xerus::TTNetwork<isOperator>& tt) {
	const size_t N = isOperator ? 2 : 1, d = tt.degree() / N;
	std::vector<size_t> ranks = tt.ranks();
	check(xb_tt_create(&H.h, d, tt.dimensions.data(), ranks.data(), isOperator ? 1 : 0));
	std::vector<xerus::Tensor> keep(d);                      // dense, factor applied; kept alive until the upload has completed
	std::vector<const double*> ptrs(d);
	for (size_t i = 0; i < d; ++i) {
		keep[i] = tt.get_component(i);
		keep[i].use_dense_representation();
		keep[i].apply_factor();
		ptrs[i] = keep[i].get_dense_data();
		g_counters.h2d_bytes += keep[i].size * sizeof(double);
	}
	check(xb_tt_set_components(H.h, ptrs.data(), ranks.data()));
	if (tt.canonicalized) check(xb_tt_assume_core_position(H.h, tt.corePosition));
}

template<bool isOperator> static void from_device(xerus::TTNetwork<isOperator>& tt, const Handle& H) {
	const size_t N = isOperator ? 2 : 1, d = tt.degree() / N;
	std::vector<xerus::Tensor> comps(d);
	std::vector<double*> ptrs(d);
	for (size_t i = 0; i < d; ++i) {
		size_t rl, ext, rr;
		check(xb_tt_component_size(H.h, i, &rl, &ext, &rr));
		std::vector<size_t> dims;
		dims.push_back(rl);
		dims.push_back(tt.dimensions[i]);
		if (isOperator) dims.push_back(tt.dimensions[d + i]);
		dims.push_back(rr);
		comps[i] = xerus::Tensor(dims, xerus::Tensor::Representation::Dense, xerus::Tensor::Initialisation::None);
		ptrs[i] = comps[i].get_unsanitized_dense_data();
		g_counters.d2h_bytes += comps[i].size * sizeof(double);
	}
	check(xb_tt_get_components(H.h, ptrs.data()));
	for (size_t i = 0; i < d; ++i) tt.set_component(i, std::move(comps[i]));
	int canon = 0; size_t pos = 0;
	check(xb_tt_core_position(H.h, &canon, &pos));
	if (canon) tt.assume_core_position(pos); else tt.canonicalized = false;
}

template<bool isOperator> bool round(xerus::TTNetwork<isOperator>& tt, const std::vector<size_t>& maxRanks, double eps) {
	const size_t d = tt.degree() / (isOperator ? 2 : 1);
	if (d < 2 || !enabled(total_size(tt))) return false;
	Handle H;
	to_device(H, tt);
	std::vector<size_t> caps(maxRanks);
	for (size_t& c : caps) if (c > (size_t(1) << 40)) c = 0;          // "no cap" is size_t max in the reference (:682-684), 0 here
	check(xb_tt_round(H.h, caps.data(), eps));
	from_device(tt, H);
	g_counters.round += 1;
	return true;
}

template<bool isOperator> bool move_core(xerus::TTNetwork<isOperator>& tt, size_t position, bool keepRank) {
	const size_t d = tt.degree() / (isOperator ? 2 : 1);
	if (d < 2 || !enabled(total_size(tt))) return false;
	if (tt.canonicalized && tt.corePosition == position) return false;      // nothing to move: the reference's loop is empty too
	Handle H;
	to_device(H, tt);
	check(xb_tt_move_core(H.h, position, keepRank ? 1 : 0));
	from_device(tt, H);
	g_counters.move_core += 1;
	return true;
}

template<bool isOperator> bool soft_threshold(xerus::TTNetwork<isOperator>& tt, const std::vector<double>& taus) {
	const size_t d = tt.degree() / (isOperator ? 2 : 1);
	if (d < 2 || !enabled(total_size(tt))) return false;
	Handle H;
	to_device(H, tt);
	check(xb_tt_soft_threshold(H.h, taus.data(), 0));
	from_device(tt, H);
	g_counters.soft_threshold += 1;
	return true;
}

// operator+= (:797-847): block stacking of the components and, for a canonicalised left operand, the move_core that restores
// its core position — stacking and sweep in one device call (xb_tt_add)
template<bool isOperator> bool add(xerus::TTNetwork<isOperator>& tt, const xerus::TTNetwork<isOperator>& other) {
	const size_t d = tt.degree() / (isOperator ? 2 : 1);
	if (d < 2 || !enabled(total_size(tt) + total_size(other))) return false;
	Handle Ha, Hb, Hc;
	to_device(Ha, tt);
	to_device(Hb, other);
	check(xb_tt_add(&Hc.h, Ha.h, Hb.h));
	from_device(tt, Hc);
	g_counters.add += 1;
	return true;
}

bool als_solve(const xerus::ALSVariant& variant, const xerus::TTOperator* A, xerus::TTTensor& x, const xerus::TTTensor& b,
               size_t numHalfSweeps, double convergenceEpsilon, double& energy) {
	if (x.degree() < 1 || !enabled(total_size(x))) return false;
	if (variant.sites != 1 && variant.sites != 2) return false;
	// only the reference's own local solvers have a device counterpart (als.h:126-128); anything else stays on the reference's path
	using Fn = decltype(&xerus::ALSVariant::lapack_solver);       // the nested ALSAlgorithmicData type is protected: name it through the public static
	const Fn* target = variant.localSolver.template target<Fn>();
	int solver = -1;
	if (target && *target == &xerus::ALSVariant::lapack_solver) solver = 0;
	if (target && *target == &xerus::ALSVariant::ASD_solver) solver = 1;
	if (solver < 0 || variant.useResidualForEndCriterion) return false;
	if (solver == 1 && variant.sites != 1) return false;
	Handle HA, Hx, Hb;
	if (A) to_device(HA, *A);
	to_device(Hx, x);
	to_device(Hb, b);
	xb_als_options opt;
	check(xb_als_default_options(&opt, variant.sites, variant.assumeSPD ? 1 : 0));
	opt.num_half_sweeps = numHalfSweeps;
	opt.convergence_epsilon = convergenceEpsilon;
	opt.preserve_core_position = variant.preserveCorePosition ? 1 : 0;
	opt.local_solver = solver;
	check(xb_als_solve(A ? HA.h : nullptr, Hx.h, Hb.h, &opt, &energy, nullptr));
	from_device(x, Hx);
	g_counters.als += 1;
	return true;
}

template bool round<false>(xerus::TTNetwork<false>&, const std::vector<size_t>&, double);
template bool round<true>(xerus::TTNetwork<true>&, const std::vector<size_t>&, double);
template bool move_core<false>(xerus::TTNetwork<false>&, size_t, bool);
template bool move_core<true>(xerus::TTNetwork<true>&, size_t, bool);
template bool soft_threshold<false>(xerus::TTNetwork<false>&, const std::vector<double>&);
template bool soft_threshold<true>(xerus::TTNetwork<true>&, const std::vector<double>&);
template bool add<false>(xerus::TTNetwork<false>&, const xerus::TTNetwork<false>&);
template bool add<true>(xerus::TTNetwork<true>&, const xerus::TTNetwork<true>&);

} // namespace xb200_resident
