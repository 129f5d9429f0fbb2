// sitewindow_binding.cpp -- TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md §1,
// compiled against the UNMODIFIED reference source (oracle/Makefile, target `binding`; needs
// /root/reference, output oracle/_ref/{fstWindow,hetWindow}_pgt).
//
// The reference translation unit is included as it lies (-I$(REFERENCE)): its info() and parseArgs()
// (fstWindow.cpp:23-67, hetWindow.cpp:20-64) stay in charge of argv, usage text, messages and exit
// codes.  Its line loop + calcWindow (fstWindow.cpp:69-155, hetWindow.cpp:66-153) are renamed out of
// the way by the two macros below and replaced by calcFst / calcHeterozygosity as a maintainer would
// write them: the same `ss >> chr >> pos >> ...` parsing into column vectors, then one pgt_plan_create
// + pgt_scan_* call through the C ABI (include/pgt_scan.h) and the reference's own print statement.
// main() repeats the reference's eight lines (fstWindow.cpp:158-177, hetWindow.cpp:156-175).
#include <cstdint>

#include "pgt_scan.h"

#if defined(BIND_FST)
#define calcFst reference_calcFst
#define main reference_main
#include "fstWindow.cpp"
#undef calcFst
#undef main
#elif defined(BIND_HET)
#define calcHeterozygosity reference_calcHeterozygosity
#define main reference_main
#include "hetWindow.cpp"
#undef calcHeterozygosity
#undef main
#else
#error "define BIND_FST or BIND_HET"
#endif

#include <algorithm>

namespace {

struct Sites {
	std::vector<unsigned int> pos;
	std::vector<double> a, b;
	std::vector<int8_t> geno;
	std::vector<std::string> names;  // contig names in order of appearance
	std::vector<uint64_t> off{0};    // cumulative sites per contig
	void contig(const std::string& chr) {
		if (names.empty() || chr != names.back()) {
			names.push_back(chr);
			off.push_back(off.back());
		}
		++off.back();
	}
};

int fail() {
	std::cerr << pgt_last_error() << "\n";
	return -1;
}

// replaces the flush triggers (fstWindow.cpp:132-138,150-152) and calcWindow (:69-107)
int scanAndPrint(Sites& s, pgt_stat stat, unsigned int winsize, unsigned int stepsize) {
	if (s.pos.empty()) return 0;
	pgt_plan* plan = NULL;
	if (pgt_plan_create(&plan, PGT_MODE_SITES, s.off.data(), (uint32_t)s.names.size(), winsize, stepsize, 0) != PGT_OK) return fail();
	const uint64_t nwin = pgt_plan_num_windows(plan);
	std::vector<uint32_t> label(nwin), start(nwin), end(nwin), mid(nwin), n(nwin);
	std::vector<double> value(nwin);
	pgt_columns cols = {};
	cols.pos = s.pos.data();
	pgt_windows out = {};
	out.label = label.data();
	out.start_pos = start.data();
	out.end_pos = end.data();
	out.mid_pos = mid.data();
	if (stat == PGT_STAT_FST) {
		cols.a = s.a.data();
		cols.b = s.b.data();
		out.fst = value.data();
		out.nsites = n.data();
	} else {
		cols.geno = s.geno.data();
		out.het = value.data();
		out.nonmissing = n.data();
	}
	if (nwin) {
		void* ws = NULL;
		const size_t wsb = pgt_scan_workspace_bytes(plan, NULL, stat, PGT_MEM_HOST);
		if (pgt_device_alloc(&ws, wsb) != PGT_OK) return fail();
		if (pgt_scan(plan, NULL, stat, &cols, 1, NULL, &out, ws, wsb, PGT_MEM_HOST, NULL) != PGT_OK) return fail();
		pgt_device_free(ws);
	}
	for (uint64_t w = 0; w < nwin; ++w)  // fstWindow.cpp:88, hetWindow.cpp:87
		std::cout << s.names[label[w]] << "\t" << start[w] << "\t" << end[w] << "\t" << mid[w] << "\t" << value[w] << "\t" << n[w] << "\n";
	pgt_plan_destroy(plan);
	return 0;
}

}  // namespace

#if defined(BIND_FST)
int calcFst(std::fstream& varcomp, unsigned int winsize, unsigned int stepsize) {
	Sites s;
	std::string line, chr;
	std::stringstream ss;
	unsigned int p = 0;
	double x = 0, y = 0;
	getline(varcomp, line);
	while (!line.empty()) {  // fstWindow.cpp:125: the first empty line ends the input
		ss.str(std::string());
		ss.clear();
		ss.str(line);
		ss >> chr >> p >> x >> y;  // fstWindow.cpp:130,141
		s.contig(chr);
		s.pos.push_back(p);
		s.a.push_back(x);
		s.b.push_back(y);
		getline(varcomp, line);
	}
	return scanAndPrint(s, PGT_STAT_FST, winsize, stepsize);
}
#else
int calcHeterozygosity(std::fstream& genofile, unsigned int winsize, unsigned int stepsize) {
	Sites s;
	std::string line, chr;
	std::stringstream ss;
	unsigned int p = 0;
	int g = 0;
	getline(genofile, line);
	while (!line.empty()) {  // hetWindow.cpp:123
		ss.str(std::string());
		ss.clear();
		ss.str(line);
		ss >> chr >> p >> g;  // hetWindow.cpp:128,139
		s.contig(chr);
		s.pos.push_back(p);
		s.geno.push_back((int8_t)(g < 0 ? -1 : std::min(g, 127)));  // hetWindow.cpp:78-80 tests g >= 0 and g == 1 only
		getline(genofile, line);
	}
	return scanAndPrint(s, PGT_STAT_HET, winsize, stepsize);
}
#endif

int main(int argc, char** argv) {
	int rv = 0;
	std::fstream f;
	unsigned int winsize = 1;
	unsigned int stepsize = 1;
	if (argc < 2) {
		info(winsize, stepsize);
		return 0;
	}
	if ((rv = parseArgs(argc, argv, f, winsize, stepsize)) != 0) return rv;
#if defined(BIND_FST)
	rv = calcFst(f, winsize, stepsize);
#else
	rv = calcHeterozygosity(f, winsize, stepsize);
#endif
	f.close();
	return rv;
}
