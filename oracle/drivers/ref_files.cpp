// TEST INFRASTRUCTURE ONLY.  Data files written and read by the unmodified reference (misc::save_to_file /
// misc::load_from_file, include/xerus/misc/fileIO.h:102-164), for the file-format bridge of libxb200 (csrc/fileio.cu):
//
//   ref_files write <dir>        small seeded objects saved by the reference in both formats  -> tests/golden/files/
//   ref_files read  <file>       loads a file (e.g. one written by libxb200) with the reference's own reader and prints
//                                kind, dimensions, ranks, canonicalisation and the Frobenius norm (17 digits)
//   ref_files copy  <in> <out>   reference load followed by reference binary save (round trip through the reference)
#include "common.h"
#include <fstream>
using namespace drv;

static std::string kind_of(const std::string& path) {
	std::ifstream in(path);
	std::string first;
	std::getline(in, first);
	if (first.find("TTNetwork<false>") != std::string::npos) return "TTTensor";
	if (first.find("TTNetwork<true>") != std::string::npos) return "TTOperator";
	if (first.find("xerus::Tensor") != std::string::npos) return "Tensor";
	return "?";
}

template <class T> static void describe(const T& t, const char* kind) {
	std::printf("kind %s\ndims", kind);
	for (size_t d : t.dimensions) std::printf(" %zu", d);
	std::printf("\nranks");
	for (size_t r : t.ranks()) std::printf(" %zu", r);
	std::printf("\ncore %d %zu\nnorm %.17g\n", t.canonicalized ? 1 : 0, t.corePosition, double(frob_norm(t)));
}

int main(int argc, char** argv) {
	if (argc >= 3 && std::string(argv[1]) == "write") {
		const std::string dir = argv[2];
		misc::randomEngine.seed(0xBAADF00D);
		TTTensor x = TTTensor::random({3, 4, 2, 5}, {3, 6, 4});                      // canonicalised, core at 0
		misc::save_to_file(x, dir + "/tttensor.bin", misc::FileFormat::BINARY);
		misc::save_to_file(x, dir + "/tttensor.tsv", misc::FileFormat::TSV);
		const TTTensor x0 = x;
		x.move_core(2);
		misc::save_to_file(x, dir + "/tttensor_core2.bin", misc::FileFormat::BINARY);
		TTTensor y = x + x;                                                           // not canonicalised, ranks doubled
		misc::save_to_file(y, dir + "/tttensor_sum.bin", misc::FileFormat::BINARY);
		TTOperator A = TTOperator::random({3, 2, 4, 2, 3, 3}, {2, 5});
		misc::save_to_file(A, dir + "/ttoperator.bin", misc::FileFormat::BINARY);
		misc::save_to_file(A, dir + "/ttoperator.tsv", misc::FileFormat::TSV);
		misc::save_to_file(laplace_operator(4, 3), dir + "/laplace.bin", misc::FileFormat::BINARY);
		Tensor T = Tensor::random({4, 3, 5});
		T *= -2.5;                                                                    // lazy factor: stored applied (tensor.cpp:1795)
		misc::save_to_file(T, dir + "/tensor_dense.bin", misc::FileFormat::BINARY);
		misc::save_to_file(T, dir + "/tensor_dense.tsv", misc::FileFormat::TSV);
		Tensor S = Tensor::random({6, 7}, 9);                                         // sparse representation, 9 entries
		S *= 3.0;
		misc::save_to_file(S, dir + "/tensor_sparse.bin", misc::FileFormat::BINARY);
		misc::save_to_file(S, dir + "/tensor_sparse.tsv", misc::FileFormat::TSV);
		// what the files hold, through the record container of the other goldens
		Writer w(dir + "/contents.bin");
		w.tt("tttensor_core2", x);
		w.tt("tttensor", x0);
		w.tt("tttensor_sum", y);
		w.tt("ttoperator", A);
		w.tt("laplace", laplace_operator(4, 3));
		w.tensor("tensor_dense", T);
		w.tensor("tensor_sparse", S);
		return 0;
	}
	if (argc >= 3 && std::string(argv[1]) == "read") {
		const std::string k = kind_of(argv[2]);
		if (k == "TTTensor") describe(misc::load_from_file<TTTensor>(argv[2]), "TTTensor");
		else if (k == "TTOperator") describe(misc::load_from_file<TTOperator>(argv[2]), "TTOperator");
		else if (k == "Tensor") {
			Tensor t = misc::load_from_file<Tensor>(argv[2]);
			std::printf("kind Tensor\ndims");
			for (size_t d : t.dimensions) std::printf(" %zu", d);
			std::printf("\nnorm %.17g\n", double(frob_norm(t)));
		} else return 3;
		return 0;
	}
	if (argc >= 4 && std::string(argv[1]) == "copy") {
		const std::string k = kind_of(argv[2]);
		if (k == "TTTensor") misc::save_to_file(misc::load_from_file<TTTensor>(argv[2]), argv[3]);
		else if (k == "TTOperator") misc::save_to_file(misc::load_from_file<TTOperator>(argv[2]), argv[3]);
		else if (k == "Tensor") misc::save_to_file(misc::load_from_file<Tensor>(argv[2]), argv[3]);
		else return 3;
		return 0;
	}
	std::fprintf(stderr, "usage: ref_files write <dir> | read <file> | copy <in> <out>\n");
	return 2;
}
