// Shared between tt.cu and als.cu: the device-resident tensor train behind the opaque xb_tt handle.
#pragma once
#include "xb_internal.cuh"

struct xb_tt {
	size_t d = 0;
	bool is_operator = false;
	std::vector<size_t> dim_m, dim_n;   // external dims per site (dim_n unused for tensors)
	std::vector<size_t> rank;           // d + 1 entries, rank[0] = rank[d] = 1
	std::vector<xb::DBuf> core;
	bool canonicalized = false;
	size_t core_position = 0;

	size_t ext(size_t i) const { return is_operator ? dim_m[i] * dim_n[i] : dim_m[i]; }
	size_t core_size(size_t i) const { return rank[i] * ext(i) * rank[i + 1]; }
};

namespace xb {
void move_core(xb_tt* t, size_t position, bool keep_rank);     // TTNetwork::move_core (ttNetwork.cpp:582-628)
double tt_frob_norm(const xb_tt* t);
double tt_inner(const xb_tt* a, const xb_tt* b);
xb_tt* tt_clone(const xb_tt* t);
}
extern "C" void require_correct_format(const xb_tt* tt);
