// TEST INFRASTRUCTURE ONLY.  Runs the unmodified reference (through its public API) on small seeded inputs
// and dumps inputs + outputs as golden vectors.  `python oracle/make_golden.py` converts the dump into the
// committed fixtures under tests/golden/.  Seed and RNG: misc::randomEngine.seed(0xBAADF00D), the seed the
// reference's own unit-test runner uses (src/xerus/test/test.cpp:105).
#include "common.h"
#include <xerus/blasLapackWrapper.h>

using namespace xerus;
using namespace drv;

static Tensor randn(const std::vector<size_t>& dims) { return Tensor::random(dims); }

static std::vector<double> svals_at_core(TTTensor x, size_t pos) {
	// singular values of the unfolding left of component `pos` (core moved there first)
	x.move_core(pos, true);
	Tensor U, S, Vt;
	calculate_svd(U, S, Vt, x.get_component(pos), 1, 0, 0.0);
	S.use_dense_representation();
	std::vector<double> s(S.dimensions[0]);
	for (size_t i = 0; i < s.size(); ++i) s[i] = S[{i, i}];
	return s;
}

static double inner(const TTTensor& a, const TTTensor& b) {
	Index i;
	Tensor r;
	r() = a(i&0) * b(i&0);
	return r[0];
}

static void blas_section(Writer& w) {
	misc::randomEngine.seed(0xBAADF00D);
	// --- matrix_matrix_product, all four transposition cases (blasLapackWrapper.cpp:149-195)
	const size_t m = 37, k = 23, n = 19;
	Tensor A = randn({m, k}), B = randn({k, n}), At = randn({k, m}), Bt = randn({n, k});
	w.tensor("mm.A", A); w.tensor("mm.B", B); w.tensor("mm.At", At); w.tensor("mm.Bt", Bt);
	std::vector<double> C(m * n);
	blasWrapper::matrix_matrix_product(C.data(), m, n, 1.5, A.get_dense_data(), false, k, B.get_dense_data(), false);
	w.raw("mm.C_nn", {m, n}, C.data());
	blasWrapper::matrix_matrix_product(C.data(), m, n, 1.5, At.get_dense_data(), true, k, B.get_dense_data(), false);
	w.raw("mm.C_tn", {m, n}, C.data());
	blasWrapper::matrix_matrix_product(C.data(), m, n, 1.5, A.get_dense_data(), false, k, Bt.get_dense_data(), true);
	w.raw("mm.C_nt", {m, n}, C.data());
	blasWrapper::matrix_matrix_product(C.data(), m, n, 1.5, At.get_dense_data(), true, k, Bt.get_dense_data(), true);
	w.raw("mm.C_tt", {m, n}, C.data());
	// degenerate shapes: gemv / ger paths (:160-166)
	{
		Tensor a = randn({1, k}), b = randn({k, 1}), col = randn({m, 1}), row = randn({1, n});
		w.tensor("mm.row1", a); w.tensor("mm.col1", b); w.tensor("mm.colm", col); w.tensor("mm.rown", row);
		std::vector<double> c1(n), c2(m), c3(m * n);
		blasWrapper::matrix_matrix_product(c1.data(), 1, n, 2.0, a.get_dense_data(), false, k, B.get_dense_data(), false);
		w.raw("mm.C_left1", {1, n}, c1.data());
		blasWrapper::matrix_matrix_product(c2.data(), m, 1, 2.0, A.get_dense_data(), false, k, b.get_dense_data(), false);
		w.raw("mm.C_right1", {m, 1}, c2.data());
		blasWrapper::matrix_matrix_product(c3.data(), m, n, 2.0, col.get_dense_data(), false, 1, row.get_dense_data(), false);
		w.raw("mm.C_mid1", {m, n}, c3.data());
	}
	// --- norms / dot
	{
		Tensor x = randn({1001}), y = randn({1001});
		w.tensor("l1.x", x); w.tensor("l1.y", y);
		w.scalar("l1.one_norm", blasWrapper::one_norm(x.get_dense_data(), 1001));
		w.scalar("l1.two_norm", blasWrapper::two_norm(x.get_dense_data(), 1001));
		w.scalar("l1.dot", blasWrapper::dot_product(x.get_dense_data(), 1001, y.get_dense_data()));
	}
	// --- QR / RQ, tall and wide (blasLapackWrapper.cpp:374-498)
	for (int wide = 0; wide < 2; ++wide) {
		const size_t qm = wide ? 12 : 40, qn = wide ? 40 : 12, r = std::min(qm, qn);
		const std::string tag = wide ? "wide" : "tall";
		Tensor M = randn({qm, qn});
		w.tensor("qr." + tag + ".A", M);
		std::vector<double> Q(qm * r), R(r * qn);
		blasWrapper::qr(Q.data(), R.data(), M.get_dense_data(), qm, qn);
		w.raw("qr." + tag + ".Q", {qm, r}, Q.data()); w.raw("qr." + tag + ".R", {r, qn}, R.data());
		std::vector<double> R2(qm * r), Q2(r * qn);
		blasWrapper::rq(R2.data(), Q2.data(), M.get_dense_data(), qm, qn);
		w.raw("rq." + tag + ".R", {qm, r}, R2.data()); w.raw("rq." + tag + ".Q", {r, qn}, Q2.data());
	}
	// --- QC / CQ with rank detection (blasLapackWrapper.cpp:235-371): full rank and rank-7 inputs.
	// NOTE the reference's rank test is input-sign dependent (uses signed R[0,0], :269/:343); both signs are dumped.
	{
		Tensor F = randn({30, 20});
		Tensor L = randn({30, 7}), Rr = randn({7, 20});
		Tensor D; contract(D, L, false, Rr, false, 1);
		Tensor Dneg = -1.0 * D; Dneg.apply_factor();
		int idx = 0;
		for (Tensor* M : {&F, &D, &Dneg}) {
			M->use_dense_representation();
			const std::string tag = "qc" + std::to_string(idx++);
			w.tensor(tag + ".A", *M);
			auto qc = blasWrapper::qc(M->get_dense_data(), 30, 20);
			const size_t rank = std::get<2>(qc);
			w.scalar(tag + ".rank", double(rank));
			w.raw(tag + ".Q", {30, rank}, std::get<0>(qc).get()); w.raw(tag + ".C", {rank, 20}, std::get<1>(qc).get());
			auto cq = blasWrapper::cq(M->get_dense_data(), 30, 20);
			const size_t rank2 = std::get<2>(cq);
			w.scalar(tag + ".cq_rank", double(rank2));
			w.raw(tag + ".cq_C", {30, rank2}, std::get<0>(cq).get()); w.raw(tag + ".cq_Q", {rank2, 20}, std::get<1>(cq).get());
		}
	}
	// --- SVD (blasLapackWrapper.cpp:201-232)
	for (int wide = 0; wide < 2; ++wide) {
		const size_t sm = wide ? 21 : 33, sn = wide ? 33 : 21, r = std::min(sm, sn);
		const std::string tag = wide ? "svd.wide" : "svd.tall";
		Tensor M = randn({sm, sn});
		w.tensor(tag + ".A", M);
		std::vector<double> U(sm * r), S(r), Vt(r * sn);
		blasWrapper::svd(U.data(), S.data(), Vt.data(), M.get_dense_data(), sm, sn);
		w.raw(tag + ".U", {sm, r}, U.data()); w.raw(tag + ".S", {r}, S.data()); w.raw(tag + ".Vt", {r, sn}, Vt.data());
	}
	// --- solve: SPD (Cholesky branch), non-symmetric (LU branch), indefinite symmetric (LDL branch)
	{
		const size_t sn = 24, nrhs = 1;   // the reference copies only _n entries of b (:563,:598,:628): nrhs > 1 is defective there
		Tensor G = randn({sn, sn});
		Tensor spd; contract(spd, G, true, G, false, 1);
		for (size_t i = 0; i < sn; ++i) spd[{i, i}] += 1.0;
		Tensor sym = G; { Tensor Gt; reshuffle(Gt, G, {1, 0}); sym += Gt; }
		Tensor rhs = randn({sn, nrhs});
		w.tensor("solve.rhs", rhs);
		int idx = 0;
		for (const Tensor* M : {&spd, &G, &sym}) {
			const std::string tag = "solve" + std::to_string(idx++);
			Tensor Md = *M; Md.use_dense_representation(); Md.apply_factor();
			w.tensor(tag + ".A", Md);
			std::vector<double> x(sn * nrhs);
			blasWrapper::solve(x.data(), Md.get_dense_data(), sn, sn, rhs.get_dense_data(), nrhs);
			w.raw(tag + ".x", {sn, nrhs}, x.data());
		}
	}
}

static void tensor_section(Writer& w) {
	misc::randomEngine.seed(0xBAADF00D);
	// contract over trailing/leading modes with T flags (tensor.cpp:1252-1352)
	Tensor A = randn({3, 4, 5, 6}), B = randn({5, 6, 7}), Bt = randn({7, 5, 6}), At = randn({5, 6, 3, 4});
	w.tensor("ct.A", A); w.tensor("ct.B", B); w.tensor("ct.Bt", Bt); w.tensor("ct.At", At);
	Tensor C;
	contract(C, A, false, B, false, 2);  w.tensor("ct.C_nn", C);
	contract(C, A, false, Bt, true, 2);  w.tensor("ct.C_nt", C);
	contract(C, At, true, B, false, 2);  w.tensor("ct.C_tn", C);
	contract(C, At, true, Bt, true, 2);  w.tensor("ct.C_tt", C);
	// scalar factors folded into alpha (:1310)
	Tensor A2 = 2.5 * A, B2 = -0.5 * B;
	contract(C, A2, false, B2, false, 2); w.tensor("ct.C_factor", C);
	// reshuffle (indexedTensor_tensor_evaluate.cpp:55-137): _shuffle[old] = new
	const std::vector<std::vector<size_t>> perms = {{0,1,2,3}, {1,0,2,3}, {0,2,1,3}, {3,2,1,0}, {1,2,3,0}, {3,0,1,2}, {0,1,3,2}, {2,3,0,1}};
	for (size_t p = 0; p < perms.size(); ++p) {
		Tensor R; reshuffle(R, A, perms[p]);
		w.sizes("rs.perm" + std::to_string(p), perms[p]);
		w.tensor("rs.out" + std::to_string(p), R);
	}
	// index-notation contraction from the README: A(i,j) = B(i,k,l) * C(k,j,l)
	{
		Index i, j, k, l;
		Tensor Bx = randn({6, 5, 4}), Cx = randn({5, 7, 4}), Ax;
		Ax(i, j) = Bx(i, k, l) * Cx(k, j, l);
		w.tensor("idx.B", Bx); w.tensor("idx.C", Cx); w.tensor("idx.A", Ax);
	}
	// calculate_svd truncation rule (tensor.cpp:1464-1474): maxRank and eps
	{
		Tensor L = randn({12, 5}), R = randn({5, 14});
		Tensor M; contract(M, L, false, R, false, 1);     // exact rank 5
		w.tensor("tsvd.A", M);
		Tensor U, S, Vt;
		calculate_svd(U, S, Vt, M, 1, 0, EPSILON);  w.scalar("tsvd.rank_eps", double(S.dimensions[0]));
		calculate_svd(U, S, Vt, M, 1, 3, EPSILON);  w.scalar("tsvd.rank_max3", double(S.dimensions[0]));
		S.use_dense_representation(); w.tensor("tsvd.S3", S); w.tensor("tsvd.U3", U); w.tensor("tsvd.Vt3", Vt);
		calculate_svd(U, S, Vt, M, 1, 0, 0.5);      w.scalar("tsvd.rank_eps05", double(S.dimensions[0]));
	}
}

static void tt_section(Writer& w) {
	// ---------- config 1: TTTensor::random({4}x8, 32).round(16)   
	misc::randomEngine.seed(0xBAADF00D);
	TTTensor A = TTTensor::random(std::vector<size_t>(8, 4), std::vector<size_t>(7, 32));
	w.tt("c1.in", A);
	w.sizes("c1.in.ranks", A.ranks());
	w.scalar("c1.in.norm", frob_norm(A));
	w.vec("c1.in.svals_bond3", svals_at_core(A, 4));
	TTTensor R = A;
	R.round(size_t(16));
	w.tt("c1.round16", R);
	w.sizes("c1.round16.ranks", R.ranks());
	w.scalar("c1.round16.norm", frob_norm(R));
	w.scalar("c1.round16.inner", inner(A, R));
	w.scalar("c1.round16.relerr", frob_norm(A - R) / frob_norm(A));
	// eps-rounding: all singular values below eps*sigma_0 cut, no rank cap (ttNetwork.cpp:682-684)
	TTTensor E = A;
	E.round(0.35);
	w.tt("c1.roundeps", E);
	w.sizes("c1.roundeps.ranks", E.ranks());
	w.scalar("c1.roundeps.relerr", frob_norm(A - E) / frob_norm(A));
	// per-edge rank vector (ttNetwork.cpp:644)
	TTTensor V = A;
	V.round(std::vector<size_t>{3, 9, 20, 32, 11, 7, 2}, EPSILON);
	w.tt("c1.roundvec", V);
	w.sizes("c1.roundvec.ranks", V.ranks());
	// move_core, rank-revealing and keepRank (ttNetwork.cpp:582-628)
	TTTensor M = A; M.move_core(5);
	w.tt("c1.core5", M);
	TTTensor M2 = A; M2.move_core(3, true);
	w.tt("c1.core3keep", M2);
	// rounding restores the caller's core position (ttNetwork.cpp:662-664)
	TTTensor M3 = M; M3.round(size_t(8));
	w.tt("c1.core5.round8", M3);
	w.sizes("c1.core5.round8.ranks", M3.ranks());

	// ---------- non-canonicalised, ragged dims, rank-deficient sum: y = x + x, then round
	misc::randomEngine.seed(0xBAADF00D);
	TTTensor x = TTTensor::random({3, 4, 2, 5, 3}, {3, 5, 6, 3});
	TTTensor y = x + x;
	w.tt("sum.x", x); w.tt("sum.y", y);
	w.sizes("sum.y.ranks", y.ranks());
	TTTensor yr = y; yr.round(1e-12);
	w.tt("sum.y.round", yr);
	w.sizes("sum.y.round.ranks", yr.ranks());
	w.tensor("sum.y.dense", Tensor(y));
	// raw (un-canonicalised) random cores through set_component + round
	TTTensor raw(std::vector<size_t>{3, 4, 2, 5, 3});
	{
		const std::vector<size_t> rk = {1, 3, 7, 6, 3, 1}, nn = {3, 4, 2, 5, 3};
		for (size_t i = 0; i < 5; ++i) raw.set_component(i, randn({rk[i], nn[i], rk[i + 1]}));
	}
	w.tt("raw.in", raw);
	TTTensor rawr = raw; rawr.round(size_t(4));
	w.tt("raw.round4", rawr);
	w.sizes("raw.round4.ranks", rawr.ranks());
	w.tensor("raw.round4.dense", Tensor(rawr));

	// ---------- TT-SVD constructor (ttNetwork.cpp:112-160)
	misc::randomEngine.seed(0xBAADF00D);
	Tensor full = randn({4, 3, 5, 2, 4});
	w.tensor("ttsvd.full", full);
	TTTensor ts(full, 1e-14);
	w.tt("ttsvd.tt", ts);
	w.sizes("ttsvd.ranks", ts.ranks());
	TTTensor ts3(full, 0.0, 3);
	w.tt("ttsvd.tt3", ts3);
	w.tensor("ttsvd.tt3.dense", Tensor(ts3));

	// ---------- operator application and TT sum feeding a round (config 5 item, reduced)
	misc::randomEngine.seed(0xBAADF00D);
	const size_t d5 = 6, n5 = 4;
	TTOperator Lap = laplace_operator(d5, n5);
	w.tt("mv.A", Lap);
	TTTensor x5 = TTTensor::random(std::vector<size_t>(d5, n5), std::vector<size_t>(d5 - 1, 8));
	w.tt("mv.x", x5);
	Index i, j;
	TTTensor y5;
	y5(i&0) = Lap(i/2, j/2) * x5(j&0);
	w.tt("mv.y", y5);
	w.sizes("mv.y.ranks", y5.ranks());
	w.tensor("mv.y.dense", Tensor(y5));
	TTTensor y5r = y5; y5r.round(size_t(8));
	w.tt("mv.y.round8", y5r);
	w.sizes("mv.y.round8.ranks", y5r.ranks());
	w.scalar("mv.y.round8.relerr", frob_norm(y5 - y5r) / frob_norm(y5));
}

static void als_section(Writer& w) {
	// ---------- ALS_SPD, Laplace-like operator, b = ones (config 2, reduced)
	struct Case { const char* tag; size_t d, n, r; };
	for (const Case& c : {Case{"als_small", 6, 4, 3}, Case{"als_mid", 8, 5, 6}}) {
		misc::randomEngine.seed(0xBAADF00D);
		const std::string tag = c.tag;
		TTOperator A = laplace_operator(c.d, c.n);
		TTTensor b = TTTensor::ones(std::vector<size_t>(c.d, c.n));
		TTTensor x0 = TTTensor::random(std::vector<size_t>(c.d, c.n), std::vector<size_t>(c.d - 1, c.r));
		w.tt(tag + ".A", A); w.tt(tag + ".b", b); w.tt(tag + ".x0", x0);
		for (size_t hs : {size_t(1), size_t(2), size_t(4)}) {
			TTTensor x = x0;
			const double energy = ALS_SPD(A, x, b, hs);
			const std::string t2 = tag + ".spd_hs" + std::to_string(hs);
			w.tt(t2 + ".x", x);
			w.scalar(t2 + ".energy", energy);
			Index i, j;
			w.scalar(t2 + ".residual", frob_norm(A(i/2, j/2) * x(j&0) - b(i&0)) / frob_norm(b));
		}
		{	// general (non-SPD) ALS: normal equations A^T A (als.cpp:189-193)
			TTTensor x = x0;
			const double res = ALS(A, x, b, size_t(2));
			w.tt(tag + ".gen_hs2.x", x);
			w.scalar(tag + ".gen_hs2.energy", res);
		}
		{	// two-site DMRG: only one increasing half-sweep is runnable in the reference
			TTTensor x = x0;
			const double energy = DMRG_SPD(A, x, b, size_t(1));
			w.tt(tag + ".dmrg_hs1.x", x);
			w.sizes(tag + ".dmrg_hs1.ranks", x.ranks());
			w.scalar(tag + ".dmrg_hs1.energy", energy);
		}
	}
	// ---------- ALS without operator: projection of b onto the rank manifold (als.cxx:88-103)
	{
		misc::randomEngine.seed(0xBAADF00D);
		TTTensor B = TTTensor::random({4, 4, 4, 4, 4}, {4, 8, 8, 4});
		TTTensor X = B; X.round(size_t(3));
		w.tt("proj.b", B); w.tt("proj.x0", X);
		const double roundNorm = frob_norm(X - B);
		ALS_SPD(X, B, 1e-4);
		w.tt("proj.x", X);
		w.scalar("proj.roundNorm", roundNorm);
		w.scalar("proj.projNorm", frob_norm(X - B));
	}
}

// Second container (tests/golden/xerus_ref_v2.npz, `ref_golden <out.bin> v2`): rows a11 / a12 / a19 of SURVEY.md 8 that the first
// file does not cover — soft_threshold, the TT-SVD constructor for TTOperators and with per-bond rank caps, ASD / ASD_SPD.
static void v2_section(Writer& w) {
	// ---------- TTNetwork::soft_threshold (ttNetwork.cpp:688-713): scalar tau and per-edge taus (taus[i]: i-th edge from the right)
	misc::randomEngine.seed(0xBAADF00D);
	TTTensor A = TTTensor::random(std::vector<size_t>(8, 4), std::vector<size_t>(7, 32));
	w.tt("st.in", A);
	{ TTTensor T = A; T.soft_threshold(2.0e4); w.tt("st.scalar", T); w.sizes("st.scalar.ranks", T.ranks()); w.scalar("st.scalar.norm", frob_norm(T)); }
	{
		const std::vector<double> taus = {1.0e4, 3.0e4, 0.0, 6.0e4, 2.5e4, 5.0e3, 4.0e4};
		w.vec("st.taus", taus);
		TTTensor T = A; T.soft_threshold(taus);
		w.tt("st.vector", T); w.sizes("st.vector.ranks", T.ranks()); w.scalar("st.vector.norm", frob_norm(T));
	}
	{ TTTensor T = A; T.move_core(5); T.soft_threshold(1.5e6); w.tt("st.core5", T); w.sizes("st.core5.ranks", T.ranks()); }   // thresholds whole spectra to 0
	// ---------- TT-SVD constructor: TTOperator (reshuffle branch, ttNetwork.cpp:129-135) and per-bond rank caps
	misc::randomEngine.seed(0xBAADF00D);
	{
		Tensor full = randn({3, 4, 2, 3, 2, 5, 3, 2});         // (m_1..m_4, n_1..n_4)
		w.tensor("opsvd.full", full);
		TTOperator T(full, 1e-14);
		w.tt("opsvd.tt", T); w.sizes("opsvd.ranks", T.ranks());
		TTOperator T2(full, 0.0, std::vector<size_t>{4, 7, 3});
		w.tt("opsvd.tt_caps", T2); w.sizes("opsvd.caps.ranks", T2.ranks()); w.tensor("opsvd.tt_caps.dense", Tensor(T2));
		Tensor f2 = randn({4, 3, 5, 2, 4});
		w.tensor("ttsvd2.full", f2);
		TTTensor T3(f2, 0.0, std::vector<size_t>{2, 5, 6, 3});
		w.tt("ttsvd2.tt_caps", T3); w.sizes("ttsvd2.caps.ranks", T3.ranks()); w.tensor("ttsvd2.tt_caps.dense", Tensor(T3));
	}
	// ---------- ASD / ASD_SPD (als.cpp:73-103, :562-563): fixed numbers of half-sweeps from the same start as the ALS goldens
	struct Case { const char* tag; size_t d, n, r; };
	for (const Case& c : {Case{"asd_small", 6, 4, 3}, Case{"asd_mid", 8, 5, 6}}) {
		misc::randomEngine.seed(0xBAADF00D);
		const std::string tag = c.tag;
		TTOperator Aop = laplace_operator(c.d, c.n);
		TTTensor b = TTTensor::ones(std::vector<size_t>(c.d, c.n));
		TTTensor x0 = TTTensor::random(std::vector<size_t>(c.d, c.n), std::vector<size_t>(c.d - 1, c.r));
		x0 /= frob_norm(x0);                                   // a start of norm one: gradient steps are sensitive to scale
		w.tt(tag + ".x0", x0);
		for (size_t hs : {size_t(1), size_t(2), size_t(6)}) {
			{ TTTensor x = x0; const double e = ASD_SPD(Aop, x, b, hs); w.tt(tag + ".spd_hs" + std::to_string(hs) + ".x", x); w.scalar(tag + ".spd_hs" + std::to_string(hs) + ".energy", e); }
			// the reference's non-SPD step (ratio of norms, als.cpp:97) diverges on this operator: only the first steps are comparable
			if (hs <= 2) { TTTensor x = x0; const double e = ASD(Aop, x, b, hs); w.tt(tag + ".gen_hs" + std::to_string(hs) + ".x", x); w.scalar(tag + ".gen_hs" + std::to_string(hs) + ".energy", e); }
		}
	}
}

int main(int argc, char** argv) {
	if (argc < 2) { std::fprintf(stderr, "usage: %s <out.bin> [v2]\n", argv[0]); return 2; }
	Writer w(argv[1]);
	if (argc > 2 && std::string(argv[2]) == "v2") { v2_section(w); return 0; }
	blas_section(w);
	tensor_section(w);
	tt_section(w);
	als_section(w);
	return 0;
}
