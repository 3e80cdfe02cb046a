// Data-file bridge: reads and writes the reference's own save_to_file / load_from_file format, so existing xerus data files
// go straight into (and out of) the device-resident tensor trains and goldens can be exchanged without a dump tool.
//
// Format (reference: include/xerus/misc/fileIO.h:102-164 for the header and the scalar / vector encoding; Tensor:
// src/xerus/tensor.cpp:1781-1845; TensorNetwork: src/xerus/tensorNetwork.cpp:1429-1505; TTNetwork: src/xerus/ttNetwork.cpp:1455-1488):
//   "Xerus <demangled type> datafile.\nFormat: Binary|TSV\n"   then, Binary = raw little-endian size_t / bool (1 byte) / double,
//   TSV = the same values as text separated by whitespace (doubles with 16 significant digits).
//   Tensor      : version 1 | dims (count, entries) | representation 1 dense: size doubles / 2 sparse: count, (position, value)*
//   TensorNetwork: version 1 | dims | per external index (node, position in that node, dimension) | #nodes |
//                  per node: #links, per link (external, other, indexPosition, dimension) | the node tensors
//   TTNetwork   : version 1 | canonicalized | corePosition | TensorNetwork with nodes [ones{1}, core_0 .. core_{d-1}, ones{1}]
// The node / link layout of a TTNetwork is the one its constructor builds (ttNetwork.cpp:70-108); the reader checks it the way
// require_correct_format does, the writer reproduces it byte for byte.
// Everything here is host code: no device is needed (or touched) except in xb_tt_load / xb_tt_save.
#include "tt_internal.cuh"
#include <fstream>
#include <iomanip>
#include <limits>
#include <memory>
#include <sstream>

using namespace xb;

struct xb_file {
	int kind = 0;                       // 0 Tensor, 1 TTTensor, 2 TTOperator
	std::vector<size_t> dims;           // Tensor: its dimensions; TT: d or 2d external dimensions (reference ordering)
	std::vector<size_t> ranks;          // TT: d - 1 bond ranks
	bool canonicalized = false;
	size_t core_position = 0;
	std::vector<std::vector<double>> data;   // Tensor: one dense array; TT: the d cores, row-major
};

namespace {

const char* const kind_name[3] = {"xerus::Tensor", "xerus::TTNetwork<false>", "xerus::TTNetwork<true>"};

struct Reader {
	std::istream& in;
	bool tsv;
	template <class T> T get() {
		T v{};
		if (tsv) in >> v; else in.read(reinterpret_cast<char*>(&v), sizeof(T));
		XB_REQUIRE(bool(in), "Unexpected end of stream in the data file");
		return v;
	}
	size_t size() { return get<size_t>(); }
	std::vector<size_t> sizes() {
		const size_t n = size();
		XB_REQUIRE(n < (size_t(1) << 32), "implausible vector length in the data file");
		std::vector<size_t> v(n);
		for (size_t i = 0; i < n; ++i) v[i] = size();
		return v;
	}
};

struct Writer {
	std::ostream& out;
	bool tsv;
	template <class T> void put(const T v) {
		if (tsv) out << v << '\t'; else out.write(reinterpret_cast<const char*>(&v), std::streamsize(sizeof(T)));
	}
	void size(const size_t v) { put<size_t>(v); }
	void sizes(const std::vector<size_t>& v) { size(v.size()); for (size_t e : v) size(e); }
};

size_t product(const std::vector<size_t>& d) { size_t s = 1; for (size_t e : d) s *= e; return s; }

// misc::stream_reader(Tensor) : dense data as stored; the sparse representation is densified (the hot path is dense)
std::vector<double> read_tensor(Reader& r, std::vector<size_t>& dims) {
	XB_REQUIRE(r.size() == 1, "Unknown stream version to open");                      // tensor.cpp:1808
	dims = r.sizes();
	const size_t rep = r.size(), n = product(dims);
	XB_REQUIRE(n < (size_t(1) << 40), "implausible tensor size in the data file");
	std::vector<double> v(n, 0.0);
	if (rep == 1) {
		if (r.tsv) { for (size_t i = 0; i < n; ++i) v[i] = r.get<double>(); }
		else { r.in.read(reinterpret_cast<char*>(v.data()), std::streamsize(n * sizeof(double))); XB_REQUIRE(bool(r.in), "Unexpected end of stream in reading dense Tensor."); }
	} else {
		XB_REQUIRE(rep == 2, "Unknown tensor representation in stream");              // tensor.cpp:1829
		const size_t num = r.size();
		for (size_t i = 0; i < num; ++i) {
			const size_t pos = r.size();
			const double val = r.get<double>();
			XB_REQUIRE(pos < n, "sparse entry outside of the tensor");
			v[pos] = val;
		}
	}
	return v;
}

void write_tensor(Writer& w, const double* data, const std::vector<size_t>& dims) {
	if (w.tsv) w.out << std::setprecision(std::numeric_limits<double>::digits10 + 1);
	w.size(1);
	w.sizes(dims);
	w.size(1);                                                                        // dense
	const size_t n = product(dims);
	if (w.tsv) { for (size_t i = 0; i < n; ++i) w.put<double>(data[i]); }
	else w.out.write(reinterpret_cast<const char*>(data), std::streamsize(n * sizeof(double)));
}

struct Link { bool external; size_t other, position, dimension; };

void read_tt(Reader& r, xb_file& f) {
	const size_t N = f.kind == 2 ? 2 : 1;
	XB_REQUIRE(r.size() == 1, "Unknown stream version to open");                      // ttNetwork.cpp:1477
	f.canonicalized = r.get<bool>();
	f.core_position = r.size();
	XB_REQUIRE(r.size() == 1, "Unknown stream version to open");                      // tensorNetwork.cpp:1471
	f.dims = r.sizes();
	XB_REQUIRE(f.dims.size() % N == 0, "Illegal degree for TTOperator.");
	const size_t d = f.dims.size() / N;
	XB_REQUIRE(d >= 1, "degree-0 tensor trains carry no cores");
	for (size_t i = 0; i < f.dims.size(); ++i) {
		const size_t other = r.size(), pos = r.size(), dim = r.size();
		// require_correct_format (ttNetwork.cpp:230-236): external index i sits on node (i mod d) + 1 at position 1 (+1 for the column index)
		XB_REQUIRE(other == (i % d) + 1 && pos == 1 + i / d && dim == f.dims[i], "external links of the stored network are not those of a TTNetwork");
	}
	const size_t nodes = r.size();
	XB_REQUIRE(nodes == d + 2, "Wrong number of nodes for a TTNetwork in the data file");
	std::vector<std::vector<Link>> links(nodes);
	for (auto& nl : links) {
		nl.resize(r.size());
		XB_REQUIRE(nl.size() <= 4, "a TTNetwork node has at most four links");
		for (Link& l : nl) { l.external = r.get<bool>(); l.other = r.size(); l.position = r.size(); l.dimension = r.size(); }
	}
	// virtual end nodes and the chain structure (ttNetwork.cpp:238-281)
	XB_REQUIRE(links[0].size() == 1 && !links[0][0].external && links[0][0].other == 1 && links[0][0].dimension == 1, "illegal left virtual node");
	XB_REQUIRE(links[d + 1].size() == 1 && !links[d + 1][0].external && links[d + 1][0].other == d && links[d + 1][0].dimension == 1, "illegal right virtual node");
	f.ranks.assign(d > 0 ? d - 1 : 0, 1);
	for (size_t i = 0; i < d; ++i) {
		const std::vector<Link>& nl = links[i + 1];
		XB_REQUIRE(nl.size() == N + 2, "Wrong degree of a TTNetwork component in the data file");
		XB_REQUIRE(!nl[0].external && nl[0].other == i && !nl[N + 1].external && nl[N + 1].other == i + 2, "components of the stored network do not form a chain");
		XB_REQUIRE(nl[1].external && nl[1].position == i && nl[1].dimension == f.dims[i], "illegal external link of a component");
		if (N == 2) XB_REQUIRE(nl[2].external && nl[2].position == d + i && nl[2].dimension == f.dims[d + i], "illegal external link of a component");
		if (i + 1 < d) f.ranks[i] = nl[N + 1].dimension;
		XB_REQUIRE(nl[0].dimension == (i == 0 ? 1 : f.ranks[i - 1]), "bond dimensions of neighbouring components do not coincide");
	}
	XB_REQUIRE(links[d][N + 1].dimension == 1, "the last component must end in a bond of dimension one");
	XB_REQUIRE(!f.canonicalized || f.core_position < d, "core position outside of the tensor train");
	f.data.resize(d);
	for (size_t node = 0; node < nodes; ++node) {
		std::vector<size_t> td;
		std::vector<double> v = read_tensor(r, td);
		if (node == 0 || node == d + 1) {
			XB_REQUIRE(td.size() == 1 && td[0] == 1 && v[0] == 1.0, "virtual nodes of a TTNetwork hold the scalar one");   // ttNetwork.cpp:244
			continue;
		}
		const size_t i = node - 1;
		std::vector<size_t> want = {i == 0 ? 1 : f.ranks[i - 1], f.dims[i]};
		if (N == 2) want.push_back(f.dims[d + i]);
		want.push_back(i + 1 < d ? f.ranks[i] : 1);
		XB_REQUIRE(td == want, "component dimensions do not match the links of the stored network");
		f.data[i] = std::move(v);
	}
}

void write_tt(Writer& w, const size_t d, const std::vector<size_t>& dims, const std::vector<size_t>& ranks, const bool is_operator,
              const bool canonicalized, const size_t core_position, const double* const* cores) {
	const size_t N = is_operator ? 2 : 1;
	const size_t none = size_t(-1);
	auto rk = [&](size_t i) { return (i == 0 || i == d) ? size_t(1) : ranks[i - 1]; };   // bond to the left of component i
	if (w.tsv) w.out << std::setprecision(std::numeric_limits<double>::digits10 + 1);
	w.size(1); w.put<bool>(canonicalized); w.size(core_position);                     // ttNetwork.cpp:1461-1465
	if (w.tsv) w.out << std::setprecision(std::numeric_limits<double>::digits10 + 1);
	w.size(1);
	w.sizes(dims);
	if (w.tsv) w.out << '\n';
	for (size_t i = 0; i < dims.size(); ++i) { w.size((i % d) + 1); w.size(1 + i / d); w.size(dims[i]); }
	if (w.tsv) w.out << "\n\n";
	w.size(d + 2);
	if (w.tsv) w.out << '\n';
	auto link = [&](bool ext, size_t other, size_t pos, size_t dim) { w.put<bool>(ext); w.size(other); w.size(pos); w.size(dim); };
	w.size(1); link(false, 1, 0, 1);
	for (size_t i = 0; i < d; ++i) {
		w.size(N + 2);
		link(false, i, i == 0 ? 0 : N + 1, rk(i));
		link(true, none, i, dims[i]);
		if (is_operator) link(true, none, d + i, dims[d + i]);
		link(false, i + 2, 0, rk(i + 1));
	}
	w.size(1); link(false, d, N + 1, 1);
	if (w.tsv) w.out << '\n';
	const double one = 1.0;
	for (size_t node = 0; node < d + 2; ++node) {
		if (node == 0 || node == d + 1) write_tensor(w, &one, {1});
		else {
			const size_t i = node - 1;
			std::vector<size_t> td = {rk(i), dims[i]};
			if (is_operator) td.push_back(dims[d + i]);
			td.push_back(rk(i + 1));
			write_tensor(w, cores[i], td);
		}
		if (w.tsv) w.out << '\n';
	}
}

void open_for_write(std::ofstream& out, const char* filename, const int kind, const bool tsv) {
	out.open(filename, tsv ? std::ofstream::out : (std::ofstream::out | std::ofstream::binary));
	XB_REQUIRE(out.good(), std::string("cannot open ") + filename + " for writing");
	const std::string header = std::string("Xerus ") + kind_name[kind] + " datafile.\nFormat: " + (tsv ? "TSV" : "Binary") + "\n";   // fileIO.h:106-114
	out.write(header.c_str(), std::streamsize(header.size()));
}

void check_tt_description(size_t d, const size_t* dims, const size_t* ranks, int is_operator, const double* const* cores) {
	XB_REQUIRE(d >= 1 && dims && cores && (d == 1 || ranks), "null");
	for (size_t i = 0; i < d * (is_operator ? 2 : 1); ++i) XB_REQUIRE(dims[i] > 0, "Zero is no valid dimension.");
	for (size_t i = 0; i + 1 < d; ++i) XB_REQUIRE(ranks[i] > 0, "Zero is no valid rank.");
	for (size_t i = 0; i < d; ++i) XB_REQUIRE(cores[i], "null component");
}

} // namespace

extern "C" {

xb_status xb_file_open(xb_file** out, const char* filename) {
	return guard([&] {
		XB_REQUIRE(out && filename, "null");
		*out = nullptr;
		std::ifstream in(filename, std::ifstream::in | std::ifstream::binary);
		XB_REQUIRE(in.good(), std::string("cannot open ") + filename);
		std::string first, second;
		std::getline(in, first);
		std::getline(in, second);
		XB_REQUIRE(bool(in), "Unexpected end of stream in load_from_file().");        // fileIO.h:133
		std::unique_ptr<xb_file> f(new xb_file());
		f->kind = -1;
		for (int k = 0; k < 3; ++k) if (first == std::string("Xerus ") + kind_name[k] + " datafile.") f->kind = k;
		XB_REQUIRE(f->kind >= 0, std::string("Invalid input file ") + filename + ": " + first);
		bool tsv;
		if (second == "Format: TSV") tsv = true;
		else if (second == "Format: Binary") tsv = false;
		else throw Error(XB_ERR_INVALID, "Invalid value for format detected: " + second);   // fileIO.h:153
		Reader r{in, tsv};
		if (f->kind == 0) { f->data.resize(1); f->data[0] = read_tensor(r, f->dims); }
		else read_tt(r, *f);
		*out = f.release();
	});
}

xb_status xb_file_close(xb_file* f) { return guard([&] { delete f; }); }

xb_status xb_file_info(const xb_file* f, int* kind, size_t* n_dims, int* canonicalized, size_t* core_position) {
	return guard([&] {
		XB_REQUIRE(f, "null");
		if (kind) *kind = f->kind;
		if (n_dims) *n_dims = f->dims.size();
		if (canonicalized) *canonicalized = f->canonicalized ? 1 : 0;
		if (core_position) *core_position = f->core_position;
	});
}

xb_status xb_file_dims(const xb_file* f, size_t* dims) {
	return guard([&] { XB_REQUIRE(f && dims, "null"); std::copy(f->dims.begin(), f->dims.end(), dims); });
}

xb_status xb_file_ranks(const xb_file* f, size_t* ranks) {
	return guard([&] { XB_REQUIRE(f && (ranks || f->ranks.empty()), "null"); std::copy(f->ranks.begin(), f->ranks.end(), ranks); });
}

xb_status xb_file_read_component(const xb_file* f, size_t idx, double* host) {
	return guard([&] {
		XB_REQUIRE(f && host, "null");
		XB_REQUIRE(idx < f->data.size(), "Illegal index in xb_file_read_component");
		std::copy(f->data[idx].begin(), f->data[idx].end(), host);
	});
}

xb_status xb_file_write_tensor(const char* filename, int tsv, const double* data, const size_t* dims, size_t degree) {
	return guard([&] {
		XB_REQUIRE(filename && (dims || degree == 0), "null");
		std::vector<size_t> d(dims, dims + degree);
		XB_REQUIRE(data || product(d) == 0, "null");
		std::ofstream out;
		open_for_write(out, filename, 0, tsv != 0);
		Writer w{out, tsv != 0};
		write_tensor(w, data, d);
		out.close();
		XB_REQUIRE(!out.fail(), std::string("error occured while writing to file ") + filename);
	});
}

xb_status xb_file_write_tt(const char* filename, int tsv, size_t d, const size_t* dims, const size_t* ranks, int is_operator,
                           int canonicalized, size_t core_position, const double* const* cores) {
	return guard([&] {
		XB_REQUIRE(filename, "null");
		check_tt_description(d, dims, ranks, is_operator, cores);
		XB_REQUIRE(!canonicalized || core_position < d, "core position outside of the tensor train");
		std::ofstream out;
		open_for_write(out, filename, is_operator ? 2 : 1, tsv != 0);
		Writer w{out, tsv != 0};
		write_tt(w, d, std::vector<size_t>(dims, dims + d * (is_operator ? 2 : 1)), std::vector<size_t>(ranks, ranks + (d - 1)), is_operator != 0,
		         canonicalized != 0, core_position, cores);
		out.close();
		XB_REQUIRE(!out.fail(), std::string("error occured while writing to file ") + filename);
	});
}

// file -> device-resident tensor train (cores uploaded on the library stream)
xb_status xb_tt_load(xb_tt** out, const char* filename) {
	return guard([&] {
		XB_REQUIRE(out, "null");
		*out = nullptr;
		xb_file* raw = nullptr;
		const xb_status st = xb_file_open(&raw, filename);
		if (st != XB_OK) throw Error(st, xb_last_error());
		std::unique_ptr<xb_file> f(raw);
		XB_REQUIRE(f->kind == 1 || f->kind == 2, "the data file holds a Tensor, not a TTNetwork");
		ensure_init();
		const size_t d = f->dims.size() / (f->kind == 2 ? 2 : 1);
		xb_tt* t = nullptr;
		xb_status s2 = xb_tt_create(&t, d, f->dims.data(), f->ranks.data(), f->kind == 2);
		if (s2 != XB_OK) throw Error(s2, xb_last_error());
		for (size_t i = 0; i < d; ++i) {
			XB_REQUIRE(f->data[i].size() == t->core_size(i), "internal: component size");
			XB_CUDA(cudaMemcpyAsync(t->core[i].p, f->data[i].data(), f->data[i].size() * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
		}
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
		t->canonicalized = f->canonicalized;
		t->core_position = f->core_position;
		*out = t;
	});
}

xb_status xb_tt_save(const xb_tt* tt, const char* filename, int tsv) {
	return guard([&] {
		XB_REQUIRE(tt && filename, "null");
		ensure_init();
		std::vector<std::vector<double>> host(tt->d);
		std::vector<const double*> ptr(tt->d);
		for (size_t i = 0; i < tt->d; ++i) {
			host[i].resize(tt->core_size(i));
			XB_CUDA(cudaMemcpyAsync(host[i].data(), tt->core[i].p, host[i].size() * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
			ptr[i] = host[i].data();
		}
		XB_CUDA(cudaStreamSynchronize(ctx().stream));
		std::vector<size_t> dims(tt->dim_m.begin(), tt->dim_m.begin() + tt->d);
		if (tt->is_operator) dims.insert(dims.end(), tt->dim_n.begin(), tt->dim_n.begin() + tt->d);
		std::vector<size_t> ranks(tt->rank.begin() + 1, tt->rank.begin() + tt->d);
		const xb_status st = xb_file_write_tt(filename, tsv, tt->d, dims.data(), ranks.data(), tt->is_operator, tt->canonicalized,
		                                      tt->core_position, ptr.data());
		if (st != XB_OK) throw Error(st, xb_last_error());
	});
}

} // extern "C"
