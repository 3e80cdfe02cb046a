// TEST INFRASTRUCTURE ONLY (oracle drivers): helpers shared by ref_golden / ref_bench.
// They call the unmodified reference library through its public API only.
#pragma once
#include <xerus.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

namespace drv {
using namespace xerus;

// Record container: magic "XBGOLD01", then records
//   u64 name_len, name bytes, u64 ndims, u64 dims[ndims], f64 data[prod(dims)]
struct Writer {
	FILE* f;
	explicit Writer(const std::string& path) : f(std::fopen(path.c_str(), "wb")) {
		if (!f) { std::perror(path.c_str()); std::exit(2); }
		std::fwrite("XBGOLD01", 1, 8, f);
	}
	~Writer() { if (f) std::fclose(f); }
	void raw(const std::string& name, const std::vector<size_t>& dims, const double* data) {
		uint64_t l = name.size(); std::fwrite(&l, 8, 1, f); std::fwrite(name.data(), 1, l, f);
		uint64_t nd = dims.size(); std::fwrite(&nd, 8, 1, f);
		size_t n = 1;
		for (size_t d : dims) { uint64_t v = d; std::fwrite(&v, 8, 1, f); n *= d; }
		std::fwrite(data, 8, n, f);
	}
	void scalar(const std::string& name, double v) { raw(name, {}, &v); }
	void vec(const std::string& name, const std::vector<double>& v) { raw(name, {v.size()}, v.data()); }
	void sizes(const std::string& name, const std::vector<size_t>& v) {
		std::vector<double> d(v.begin(), v.end()); vec(name, d);
	}
	// Dense tensor with its lazy scalar factor applied (reference: tensor.h:105, tensor.cpp:1186).
	void tensor(const std::string& name, Tensor t) {
		t.use_dense_representation();
		t.apply_factor();
		raw(name, t.dimensions, t.get_dense_data());
	}
	template<bool isOp> void tt(const std::string& name, const TTNetwork<isOp>& t) {
		const size_t d = t.degree() / (isOp ? 2 : 1);
		scalar(name + ".d", double(d));
		scalar(name + ".core", t.canonicalized ? double(t.corePosition) : -1.0);
		for (size_t i = 0; i < d; ++i) tensor(name + ".c" + std::to_string(i), t.get_component(i));
	}
};

// Laplace-like rank-2 TT operator: sum_k I x ... x L x ... x I with L = tridiag(-1,2,-1)
// (built through the reference's public set_component).
inline TTOperator laplace_operator(size_t d, size_t n) {
	std::vector<size_t> dims(2 * d, n);
	TTOperator A(dims);
	for (size_t k = 0; k < d; ++k) {
		const size_t rl = (k == 0) ? 1 : 2, rr = (k + 1 == d) ? 1 : 2;
		Tensor c({rl, n, n, rr});
		auto L = [&](size_t i, size_t j) { return i == j ? 2.0 : ((i + 1 == j || j + 1 == i) ? -1.0 : 0.0); };
		auto I = [&](size_t i, size_t j) { return i == j ? 1.0 : 0.0; };
		for (size_t a = 0; a < rl; ++a) for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t b = 0; b < rr; ++b) {
			double v;
			if (d == 1) v = L(i, j);
			else if (k == 0) v = (b == 0) ? L(i, j) : I(i, j);                 // [L I]
			else if (k + 1 == d) v = (a == 0) ? I(i, j) : L(i, j);            // [I; L]
			else v = (a == 0 && b == 0) ? I(i, j) : (a == 1 && b == 0) ? L(i, j) : (a == 1 && b == 1) ? I(i, j) : 0.0; // [[I 0],[L I]]
			c[{a, i, j, b}] = v;
		}
		A.set_component(k, c);
	}
	return A;
}

inline double now_ms() {
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
} // namespace drv
