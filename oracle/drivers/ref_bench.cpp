// TEST INFRASTRUCTURE ONLY.  Times the unmodified reference CPU path (OpenBLAS/LAPACKE) for the bench harness:
//   ref_bench round <d> <n> <r> <maxRank> <reps> [dump.bin]   -> TTTensor::random({n}^d, r).round(maxRank)
//   ref_bench als   <d> <n> <r> <halfSweeps> <reps> [dump.bin]-> ALS_SPD(Laplace, random rank-r x, ones)
//   ref_bench dmrg  <d> <n> <r> <halfSweeps> <reps> [dump.bin]-> DMRG_SPD(Laplace, random rank-r x, ones); only one increasing
//                                                                 half-sweep is runnable in the reference
//   ref_bench matvec_round <d> <n> <r> <maxRank> <reps> [dump.bin [items]]
//                                                              -> y = A*x ; y.round(maxRank)   (config 5 item); with a dump path,
//                                                                 `items` consecutive random x_i and their y_i are written
// Prints one JSON line: per-rep wall times (steady_clock) and summary values used for parity.
// With a dump path it also writes inputs/outputs in the golden container so the GPU run can use the very same TT.
#include "common.h"
#include <algorithm>
#include <thread>

using namespace xerus;
using namespace drv;

static double inner(const TTTensor& a, const TTTensor& b) {
	Index i; Tensor r; r() = a(i&0) * b(i&0); return r[0];
}

static void print_times(const std::vector<double>& t) {
	std::printf("\"times_ms\": [");
	for (size_t i = 0; i < t.size(); ++i) std::printf("%s%.6f", i ? ", " : "", t[i]);
	std::printf("], \"best_ms\": %.6f, \"median_ms\": %.6f", *std::min_element(t.begin(), t.end()),
	            [&] { auto s = t; std::sort(s.begin(), s.end()); return s[s.size() / 2]; }());
}

int main(int argc, char** argv) {
	if (argc < 7) { std::fprintf(stderr, "usage: see header\n"); return 2; }
	const std::string mode = argv[1];
	const size_t d = std::stoul(argv[2]), n = std::stoul(argv[3]), r = std::stoul(argv[4]), p = std::stoul(argv[5]);
	const size_t reps = std::stoul(argv[6]);
	const std::string dump = argc > 7 ? argv[7] : "";
	misc::randomEngine.seed(0xBAADF00D);
	const std::vector<size_t> dims(d, n);
	std::vector<double> times;
	std::printf("{\"mode\": \"%s\", \"d\": %zu, \"n\": %zu, \"r\": %zu, \"param\": %zu, \"hw_threads\": %u, ",
	            mode.c_str(), d, n, r, p, std::thread::hardware_concurrency());
	if (mode == "round") {
		const TTTensor A = TTTensor::random(dims, std::vector<size_t>(d - 1, r));
		TTTensor R;
		for (size_t rep = 0; rep < reps; ++rep) {
			R = A;
			const double t0 = now_ms();
			R.round(p);
			times.push_back(now_ms() - t0);
		}
		print_times(times);
		std::printf(", \"norm_in\": %.17g, \"norm_out\": %.17g, \"inner\": %.17g", frob_norm(A), frob_norm(R), inner(A, R));
		if (!dump.empty()) { Writer w(dump); w.tt("in", A); w.tt("out", R); w.sizes("out.ranks", R.ranks()); }
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
		print_times(times);
		Index i, j;
		std::printf(", \"energy\": %.17g, \"residual\": %.17g", energy, frob_norm(A(i/2, j/2) * x(j&0) - b(i&0)) / frob_norm(b));
		if (!dump.empty()) { Writer w(dump); w.tt("A", A); w.tt("b", b); w.tt("x0", x0); w.tt("x", x); w.scalar("energy", energy); }
	} else if (mode == "dmrg") {
		const TTOperator A = laplace_operator(d, n);
		const TTTensor b = TTTensor::ones(dims);
		const TTTensor x0 = TTTensor::random(dims, std::vector<size_t>(d - 1, r));
		TTTensor x; double energy = 0;
		for (size_t rep = 0; rep < reps; ++rep) {
			x = x0;
			const double t0 = now_ms();
			energy = DMRG_SPD(A, x, b, p);
			times.push_back(now_ms() - t0);
		}
		print_times(times);
		std::printf(", \"energy\": %.17g", energy);
		if (!dump.empty()) { Writer w(dump); w.tt("A", A); w.tt("b", b); w.tt("x0", x0); w.tt("x", x); w.scalar("energy", energy); w.sizes("x.ranks", x.ranks()); }
	} else if (mode == "matvec_round") {
		const TTOperator A = laplace_operator(d, n);
		const size_t items = argc > 8 ? std::stoul(argv[8]) : 1;
		Writer* w = dump.empty() ? nullptr : new Writer(dump);
		if (w) w->tt("A", A);
		TTTensor y;
		for (size_t it = 0; it < items; ++it) {
			const TTTensor x = TTTensor::random(dims, std::vector<size_t>(d - 1, r));
			for (size_t rep = 0; rep < reps; ++rep) {
				Index i, j;
				const double t0 = now_ms();
				y(i&0) = A(i/2, j/2) * x(j&0);
				y.round(p);
				times.push_back(now_ms() - t0);
			}
			if (w) { w->tt("x" + std::to_string(it), x); w->tt("y" + std::to_string(it), y); w->sizes("y" + std::to_string(it) + ".ranks", y.ranks()); }
		}
		delete w;
		print_times(times);
		std::printf(", \"items\": %zu, \"norm_out\": %.17g", items, frob_norm(y));
	} else {
		std::fprintf(stderr, "unknown mode\n"); return 2;
	}
	std::printf("}\n");
	return 0;
}
