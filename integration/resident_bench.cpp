// BASELINE configs through the reference's own API on the resident front-end (integration/Makefile: resident_bench):
//   resident_bench round <d> <n> <r> <maxRank> <reps>   -> xerus::TTTensor::random({n}^d, r); x.round(maxRank)
//   resident_bench als   <d> <n> <r> <halfSweeps> <reps> -> xerus::ALS_SPD(Laplace, random rank-r x, ones, halfSweeps)
// Same inputs as oracle/drivers/ref_bench.cpp (seed 0xBAADF00D), so the printed summary values are directly comparable with the
// reference's.  Wall time with steady_clock around the xerus call: host TT in, host TT out (uploads and downloads inside).
#include <xerus.h>
#include "xb200_resident.h"
#include "../include/xb200.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <string>

using namespace xerus;

static double now_ms() {
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
static double inner(const TTTensor& a, const TTTensor& b) { Index i; Tensor r; r() = a(i&0) * b(i&0); return r[0]; }

static TTOperator laplace_operator(size_t d, size_t n) {       // as oracle/drivers/common.h
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
			else if (k == 0) v = (b == 0) ? L(i, j) : I(i, j);
			else if (k + 1 == d) v = (a == 0) ? I(i, j) : L(i, j);
			else v = (a == 0 && b == 0) ? I(i, j) : (a == 1 && b == 0) ? L(i, j) : (a == 1 && b == 1) ? I(i, j) : 0.0;
			c[{a, i, j, b}] = v;
		}
		A.set_component(k, c);
	}
	return A;
}

int main(int argc, char** argv) {
	if (argc < 7) { std::fprintf(stderr, "usage: see header\n"); return 2; }
	const std::string mode = argv[1];
	const size_t d = std::stoul(argv[2]), n = std::stoul(argv[3]), r = std::stoul(argv[4]), p = std::stoul(argv[5]), reps = std::stoul(argv[6]);
	misc::randomEngine.seed(0xBAADF00D);
	const std::vector<size_t> dims(d, n);
	std::vector<double> times;
	std::printf("{\"mode\": \"%s\", \"d\": %zu, \"n\": %zu, \"r\": %zu, \"param\": %zu, ", mode.c_str(), d, n, r, p);
	if (mode == "round") {
		const TTTensor A = TTTensor::random(dims, std::vector<size_t>(d - 1, r));
		TTTensor R;
		for (size_t rep = 0; rep < reps; ++rep) {
			R = A;
			const double t0 = now_ms();
			R.round(p);
			times.push_back(now_ms() - t0);
		}
		std::printf("\"norm_in\": %.17g, \"norm_out\": %.17g, \"inner\": %.17g, ", frob_norm(A), frob_norm(R), inner(A, R));
	} else if (mode == "als") {
		const TTOperator A = laplace_operator(d, n);
		const TTTensor b = TTTensor::ones(dims);
		const TTTensor x0 = TTTensor::random(dims, std::vector<size_t>(d - 1, r));
		TTTensor x; double energy = 0;
		for (size_t rep = 0; rep < reps; ++rep) {
			x = x0;
			const double t0 = now_ms();
			energy = ALS_SPD(A, x, b, p);
			times.push_back(now_ms() - t0);
		}
		Index i, j;
		std::printf("\"energy\": %.17g, \"residual\": %.17g, ", energy, frob_norm(A(i/2, j/2) * x(j&0) - b(i&0)) / frob_norm(b));
	} else { std::fprintf(stderr, "unknown mode\n"); return 2; }
	std::printf("\"times_ms\": [");
	for (size_t i = 0; i < times.size(); ++i) std::printf("%s%.6f", i ? ", " : "", times[i]);
	std::vector<double> s = times; std::sort(s.begin(), s.end());
	const xb200_resident::Counters& c = xb200_resident::counters();
	uint64_t launches = 0; xb_kernel_launch_count(&launches);
	std::printf("], \"best_ms\": %.6f, \"median_ms\": %.6f, \"hooks\": {\"round\": %zu, \"move_core\": %zu, \"als\": %zu}, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"gpu_launches\": %llu}\n",
	            s.front(), s[s.size() / 2], c.round, c.move_core, c.als, c.h2d_bytes, c.d2h_bytes, (unsigned long long)launches);
	return 0;
}
