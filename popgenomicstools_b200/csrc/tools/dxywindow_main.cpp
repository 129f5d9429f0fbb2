// dxywindow_main.cpp -- drop-in `dxyWindow`.
//
// Same options, defaults, help text, messages, stdout/stderr split and exit codes as
// /root/reference/dxyWindow.cpp:34-139,518-550.  maf2dxy (dxyWindow.cpp:253-436) is replaced by:
//   1. read both MAFs (plain or gzip, magic sniff as dxyWindow.cpp:82-83), parse them into
//      columns on several threads, skipping the header line (dxyWindow.cpp:284,290);
//   2. the two-file position sync state machine of dxyWindow.cpp:315-331, restated over the
//      parsed arrays (position-only catch-up, silent stop);
//   3. closed-form window plan + CUDA scan in libpgtscan (site windows for -fixedsite 1, sparse
//      bp windows for -fixedsite 0: no per-bp filler entries are ever materialised);
//   4. rows to stdout (-skip_missing filters neffective == 0, dxyWindow.cpp:189), the global
//      line to stderr, or to stdout for -winsize 0 (dxyWindow.cpp:429-433).
//
// Documented deviations (reference behaviour is undefined or a crash there):
//   * -winsize 0 with -fixedsite 0 segfaults in the reference; here: error, exit 255.
//   * -stepsize > -winsize: error, exit 255.
//   * bp mode needs positions strictly increasing within a chromosome and <= its sizefile
//     length (the reference silently produces misplaced entries otherwise): error, exit 255.
//   * MAF files without data lines: error instead of reading an empty record.
// Binary columnar cache (pgt_colfile.h): PGT_PACK=<pop1.pgtc> PGT_PACK2=<pop2.pgtc> parse the two MAFs,
// write one cache per population and exit without touching the GPU; a `.pgtc` file given in place of
// a MAF (magic sniff, next to the gzip sniff of dxyWindow.cpp:82-83) skips the parsing of that file.
// PGT_PACK_SYNCED=<out.pgtc> writes the SYNCED site list of the two files (kind 5) and exits: the result of the
// two-file sync state machine of dxyWindow.cpp:315-331, for inspection and for the CPU fuzz of that logic.
#include <map>

#include "pgt_cli.h"
#include "pgt_colfile.h"

using namespace pgtcli;

static void help_info(unsigned winsize, unsigned stepsize, int minind, int fixedsite, int skip_missing) {
	// dxyWindow.cpp:34-61 (setw(14) / setw(8), left aligned)
	printf("\ndxyWindow [options] <pop1 maf file> <pop2 maf file>\n\nOptions:\n");
	printf("%-14s%-8sWindow size in base pairs (0 for global calculation) [%u]\n", "-winsize", "INT", winsize);
	printf("%-14s%-8sNumber of base pairs to progress window [%u]\n", "-stepsize", "INT", stepsize);
	printf("%-14s%-8sMinimum number of individuals in each population with data [%d]\n", "-minind", "INT", minind);
	printf("%-14s%-8s(1) Use fixed number of sites from MAF input for each window (window sizes may vary) or (0) constant window size [%d]\n",
	       "-fixedsite", "INT", fixedsite);
	printf("%-14s%-8sTwo-column TSV file with each row having (1) chromsome name (2) chromosome size in base pairs\n", "-sizefile", "FILE");
	printf("%-14s%-8sDo not print windows with zero effective sites if INT=1 [%d]\n", "-skip_missing", "INT", skip_missing);
	printf("\nNotes:\n"
	       "* -winsize 1 -stepsize 1 calculates per site dxy\n"
	       "* -sizefile is REQUIRED(!) with -fixedsite 0 (the default)\n"
	       "* Both input MAF files need to have the same chromosomes in the same order\n"
	       "* Assumes SNPs are biallelic across populations\n"
	       "* For global Dxy calculations only columns 4, 5, and 6 below are printed\n"
	       "* Input MAF files can contain all sites (including monomorphic sites) or just variable sites\n"
	       "* -fixedsite 1 -winsize 500 would for example ensure that all windows contain 500 SNPs\n"
	       "\nOutput:\n"
	       "(1) chromosome\n"
	       "(2) Window start\n"
	       "(3) Window end\n"
	       "(4) dxy\n"
	       "(5) number sites in MAF input that were analyzed\n"
	       "(6) number of sites in MAF input that were skipped due to too few individuals\n\n");
}

struct Maf {
	std::vector<uint32_t> chr;  // id into names (per file)
	std::vector<uint32_t> pos;
	std::vector<double> freq;
	std::vector<int32_t> nind;
	std::vector<std::string> names;  // run names in order of appearance (a name may repeat)
};

struct MafChunk {
	size_t begin, end;
	std::vector<uint32_t> run;  // chunk-local run index per line
	std::vector<uint32_t> pos;
	std::vector<double> freq;
	std::vector<int32_t> nind;
	std::vector<std::string> names;
	long bad_line = -1;
};

// tokenizeStr, dxyWindow.cpp:141-153: chr, unsigned pos, three single chars, double, int
static bool parse_maf_line(const char* p, const char* le, const char** name_b, const char** name_e, uint32_t* pos, double* f, int32_t* n) {
	const char* s = skip_ws(p, le);
	const char* t = token_end(s, le);
	if (t == s) return false;
	*name_b = s;
	*name_e = t;
	const char* cur = t;
	if (!parse_u32(cur, le, pos)) return false;
	for (int k = 0; k < 3; ++k) {  // `ss >> char` takes ONE non-blank character
		cur = skip_ws(cur, le);
		if (cur >= le) return false;
		++cur;
	}
	if (!parse_f64(cur, le, f)) return false;
	if (!parse_i32(cur, le, n)) return false;
	return true;
}

// The common MAF line `chr <b> pos <b> X <b> Y <b> Z <b> freq <b> nInd` with plain decimals, in one pass
// (pgt_cli.h fast_decimal: exact by Clinger's fast path); anything else returns false without side
// effects and goes through parse_maf_line.
static bool fast_maf_line(const char* p, const char* le, const char** name_b, const char** name_e, uint32_t* pos, double* f, int32_t* n) {
	const char* t = p;
	while (t < le && *t != ' ' && *t != '\t' && *t != '\r' && *t != '\v' && *t != '\f') ++t;
	if (t == p || t >= le) return false;
	const char* q = skip_blanks(t, le);
	uint64_t x = 0;
	int nd = 0;
	while (q < le && (unsigned)(*q - '0') <= 9u) {
		x = x * 10u + (unsigned)(*q - '0');
		++q;
		++nd;
	}
	if (nd == 0 || nd > 9) return false;
	for (int k = 0; k < 3; ++k) {  // three single-character columns
		if (q >= le || (*q != ' ' && *q != '\t')) return false;
		q = skip_blanks(q, le);
		if (q >= le || *q == '\r') return false;
		++q;
	}
	if (q >= le || (*q != ' ' && *q != '\t')) return false;
	q = skip_blanks(q, le);
	double fv;
	if (!fast_decimal(q, le, &fv)) return false;
	if (q >= le || (*q != ' ' && *q != '\t')) return false;
	q = skip_blanks(q, le);
	bool neg = false;
	if (q < le && *q == '-') {
		neg = true;
		++q;
	}
	uint32_t g = 0;
	int gd = 0;
	while (q < le && (unsigned)(*q - '0') <= 9u) {
		g = g * 10u + (unsigned)(*q - '0');
		++q;
		++gd;
	}
	if (gd == 0 || gd > 9) return false;
	if (q < le && *q != ' ' && *q != '\t' && *q != '\r') return false;
	*name_b = p;
	*name_e = t;
	*pos = (uint32_t)x;
	*f = fv;
	*n = neg ? -(int32_t)g : (int32_t)g;
	return true;
}

static int load_maf(const char* path, Maf* m, const char* which, Input* keep) {
	if (read_input(path, keep, true) != 0) {
		fprintf(stderr, "Unable to open %s MAF file: %s\n", which, path);
		return -1;
	}
	return 0;
}

static int parse_maf(const Input& in, Maf* m, const char* path) {
	if (pgtcol::is_colfile(in.data, in.size)) {  // cached columns of one population
		pgtcol::View v;
		std::string err;
		if (pgtcol::open_view(in.data, in.size, &v, &err) != 0 || v.kind != pgtcol::KIND_MAF) {
			fprintf(stderr, "dxyWindow: %s: %s\n", path, err.empty() ? "columnar file of another tool" : err.c_str());
			return -1;
		}
		const uint64_t n = v.nsites;
		m->pos.assign((const uint32_t*)v.col[0], (const uint32_t*)v.col[0] + n);
		m->freq.assign((const double*)v.col[1], (const double*)v.col[1] + n);
		m->nind.assign((const int32_t*)v.col[2], (const int32_t*)v.col[2] + n);
		m->chr.reserve(n);
		for (const ContigRun& r : v.runs) {
			m->names.push_back(r.name);
			m->chr.insert(m->chr.end(), r.count, (uint32_t)m->names.size() - 1);
		}
		return 0;
	}
	// skip the header line (dxyWindow.cpp:284), stop at the first empty line (:313)
	const char* nl = (const char*)memchr(in.data, '\n', in.size);
	size_t begin = nl ? (size_t)(nl + 1 - in.data) : in.size;
	size_t n_eff = begin + effective_size(in.data + begin, in.size - begin);
	const unsigned nt = (n_eff - begin) < parallel_min_bytes() ? 1 : parse_threads();
	std::vector<size_t> starts = chunk_starts(in.data, begin, n_eff, nt);
	std::vector<MafChunk> chunks(starts.size() - 1);
	for (size_t i = 0; i + 1 < starts.size(); ++i) {
		chunks[i].begin = starts[i];
		chunks[i].end = starts[i + 1];
	}
	auto work = [&](MafChunk& c) {
		const char* p = in.data + c.begin;
		const char* e = in.data + c.end;
		const char* prev = nullptr;
		size_t prev_len = 0;
		long li = 0;
		while (p < e) {
			const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
			const char* le = q ? q : e;
			const char *nb = nullptr, *ne = nullptr;
			uint32_t ps = 0;
			double f = 0;
			int32_t ni = 0;
			if (!fast_maf_line(p, le, &nb, &ne, &ps, &f, &ni) && !parse_maf_line(p, le, &nb, &ne, &ps, &f, &ni)) {
				if (c.bad_line < 0) c.bad_line = li;
				nb = ne = p;
			}
			if (!prev || (size_t)(ne - nb) != prev_len || memcmp(prev, nb, prev_len) != 0) {
				c.names.emplace_back(nb, ne);
				prev = nb;
				prev_len = (size_t)(ne - nb);
			}
			c.run.push_back((uint32_t)c.names.size() - 1);
			c.pos.push_back(ps);
			c.freq.push_back(f);
			c.nind.push_back(ni);
			++li;
			if (!q) break;
			p = q + 1;
		}
	};
	{
		std::vector<std::thread> th;
		for (MafChunk& c : chunks) th.emplace_back(work, std::ref(c));
		for (auto& x : th) x.join();
	}
	uint64_t line0 = 2;
	for (MafChunk& c : chunks) {
		if (c.bad_line >= 0) {
			fprintf(stderr, "dxyWindow: cannot parse line %llu of %s\n", (unsigned long long)(line0 + c.bad_line), path);
			return -1;
		}
		line0 += c.pos.size();
		uint32_t base = (uint32_t)m->names.size();
		bool merge = !m->names.empty() && !c.names.empty() && m->names.back() == c.names.front();
		if (merge) base -= 1;
		for (size_t i = merge ? 1 : 0; i < c.names.size(); ++i) m->names.push_back(c.names[i]);
		for (uint32_t r : c.run) m->chr.push_back(base + r);
		m->pos.insert(m->pos.end(), c.pos.begin(), c.pos.end());
		m->freq.insert(m->freq.end(), c.freq.begin(), c.freq.end());
		m->nind.insert(m->nind.end(), c.nind.begin(), c.nind.end());
	}
	return 0;
}

int main(int argc, char** argv) {
	unsigned winsize = 0, stepsize = 0;  // dxyWindow.cpp:529-534
	int minind = 1, fixedsite = 0, skip_missing = 0;
	if (argc < 3) {
		help_info(winsize, stepsize, minind, fixedsite, skip_missing);
		return 0;  // parseArgs returns 1, main maps it to 0 (dxyWindow.cpp:537-538)
	}
	const double t_start = now_ms();
	Timing tm;
	tm.threads = parse_threads();
	// files are opened before the options are looked at (dxyWindow.cpp:73-95)
	Input in1, in2;
	Maf m1, m2;
	if (load_maf(argv[argc - 2], &m1, "Pop1", &in1) != 0) return -1;
	if (load_maf(argv[argc - 1], &m2, "Pop2", &in2) != 0) return -1;
	const char* sizefname = nullptr;
	bool have_sizefile = false;
	Input sizein;
	for (int argpos = 1; argpos < argc - 2; argpos += 2) {  // dxyWindow.cpp:97-126
		const char* o = argv[argpos];
		const char* v = argv[argpos + 1];
		if (strcmp(o, "-winsize") == 0) {
			int x = atoi(v);
			if (x < 0) {
				fprintf(stderr, "-winsize must not be negative\n");
				return -1;
			}
			winsize = (unsigned)x;
		} else if (strcmp(o, "-stepsize") == 0) {
			int x = atoi(v);
			stepsize = x < 0 ? 0u : (unsigned)x;
		} else if (strcmp(o, "-minind") == 0) {
			minind = atoi(v);
			if (minind <= 0) {
				fprintf(stderr, "-minind must be at least 1\n");
				return -1;
			}
		} else if (strcmp(o, "-sizefile") == 0) {
			sizefname = v;
			if (read_input(sizefname, &sizein, false) != 0) {
				fprintf(stderr, "Unable to open sizefile: %s\n", sizefname);
				return -1;
			}
			have_sizefile = true;
		} else if (strcmp(o, "-fixedsite") == 0) {
			fixedsite = atoi(v);
		} else if (strcmp(o, "-skip_missing") == 0) {
			skip_missing = atoi(v);
		} else {
			fprintf(stderr, "Unknown command: %s\n", o);
			return -1;
		}
	}
	if (winsize > 0 && stepsize < 1) {
		fprintf(stderr, "Must specify a -stepsize > 0 when -winsize is > 0\n");
		return -1;
	}
	// dxyWindow.cpp:133-136 tests `!sizefile` on a never-opened ifstream, which is false: the
	// "Must supply size file unless -fixedsite 1" branch is unreachable in the reference.  Without
	// -sizefile the size map is simply empty and the run ends with "Unable to determine size for X"
	// at the first chromosome end, which is what happens here too.
	(void)have_sizefile;
	if (winsize > 0 && stepsize > winsize) {
		fprintf(stderr, "-stepsize must not exceed -winsize\n");
		return -1;
	}
	if (winsize == 0 && !fixedsite) {
		fprintf(stderr, "-winsize 0 (global calculation) requires -fixedsite 1\n");
		return -1;
	}

	// parseSizes, dxyWindow.cpp:155-170 (first entry of a name wins: std::map::insert)
	std::map<std::string, unsigned> chrsize;
	if (!fixedsite) {
		const char* p = sizein.data;
		const char* e = sizein.data + sizein.size;
		while (p < e) {
			const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
			const char* le = q ? q : e;
			const char* s = skip_ws(p, le);
			const char* t = token_end(s, le);
			const char* cur = t;
			uint32_t len = 0;
			bool ok = t > s && parse_u32(cur, le, &len) && len > 0;
			if (!ok) {
				fprintf(stderr, "Unable to correctly parse chromosome size file\n");
				return -1;
			}
			chrsize.insert(std::make_pair(std::string(s, t), len));
			if (!q) break;
			p = q + 1;
		}
	}

	DeviceWarmup warm;
	const char* pack1 = getenv("PGT_PACK");
	const char* pack2 = getenv("PGT_PACK2");
	if (!pack1 && !pack2 && !getenv("PGT_PACK_SYNCED")) warm.start();
	if (parse_maf(in1, &m1, argv[argc - 2]) != 0) return -1;
	if (parse_maf(in2, &m2, argv[argc - 1]) != 0) return -1;
	if (pack1 || pack2) {  // write the binary columnar caches and stop: no GPU involved
		auto pack = [&](const char* path, const Maf& m) -> int {
			std::vector<ContigRun> runs;
			for (uint64_t i = 0; i < m.pos.size(); ++i) {
				if (runs.empty() || m.chr[i] != m.chr[i - 1]) runs.push_back(ContigRun{m.names[m.chr[i]], 0});
				runs.back().count++;
			}
			const void* cols[3] = {m.pos.data(), m.freq.data(), m.nind.data()};
			if (pgtcol::write_file(path, pgtcol::KIND_MAF, runs, m.pos.size(), cols) != 0) {
				fprintf(stderr, "dxyWindow: cannot write %s\n", path);
				return -1;
			}
			return 0;
		};
		if (pack1 && pack(pack1, m1) != 0) return -1;
		if (pack2 && pack(pack2, m2) != 0) return -1;
		return 0;
	}
	const uint64_t n1 = m1.pos.size(), n2 = m2.pos.size();
	if (n1 == 0 || n2 == 0) {
		fprintf(stderr, "dxyWindow: MAF file without data lines\n");
		return -1;
	}
	if (m1.names[m1.chr[0]] != m2.names[m2.chr[0]]) {  // dxyWindow.cpp:294-298
		fprintf(stderr, "Chromosomes in MAF files differ\n");
		return -1;
	}

	// ---- two-file sync, dxyWindow.cpp:313-331,399-403 ------------------------------------------
	std::vector<uint32_t> pos;
	std::vector<double> f1, f2;
	std::vector<int32_t> ni1, ni2;
	std::vector<ContigRun> runs;
	{
		pos.reserve(std::min(n1, n2));
		f1.reserve(std::min(n1, n2));
		f2.reserve(std::min(n1, n2));
		ni1.reserve(std::min(n1, n2));
		ni2.reserve(std::min(n1, n2));
		uint64_t i1 = 0, i2 = 0;
		const std::string* chr = &m1.names[m1.chr[0]];
		auto name1 = [&](uint64_t i) -> const std::string& { return m1.names[m1.chr[i]]; };
		auto name2 = [&](uint64_t i) -> const std::string& { return m2.names[m2.chr[i]]; };
		for (;;) {
			if (m1.pos[i1] != m2.pos[i2] || name1(i1) != name2(i2)) {
				const bool samechr = name1(i1) == name2(i2);
				if ((samechr && m1.pos[i1] < m2.pos[i2]) || (!samechr && name2(i2) != *chr)) {
					while (m1.pos[i1] != m2.pos[i2]) {  // catch maf1 up to maf2 (position-only compare)
						if (i1 + 1 >= n1) break;
						++i1;
					}
					if (m1.pos[i1] != m2.pos[i2]) break;
				} else {
					while (m2.pos[i2] < m1.pos[i1]) {
						if (i2 + 1 >= n2) break;
						++i2;
					}
					if (m1.pos[i1] != m2.pos[i2]) break;
				}
			}
			chr = &name1(i1);
			if (runs.empty() || runs.back().name != *chr) runs.push_back(ContigRun{*chr, 0});
			runs.back().count++;
			pos.push_back(m1.pos[i1]);
			f1.push_back(m1.freq[i1]);
			f2.push_back(m2.freq[i2]);
			ni1.push_back(m1.nind[i1]);
			ni2.push_back(m2.nind[i2]);
			if (++i1 >= n1) break;
			if (++i2 >= n2) break;
		}
	}
	const uint64_t n = pos.size();
	tm.sites = n;
	tm.parse_ms = now_ms() - t_start;
	if (const char* synced = getenv("PGT_PACK_SYNCED")) {  // the synced site list as a kind-5 .pgtc, then stop (no GPU)
		const void* cols[5] = {pos.data(), f1.data(), f2.data(), ni1.data(), ni2.data()};
		if (pgtcol::write_file(synced, pgtcol::KIND_DXY, runs, n, cols) != 0) {
			fprintf(stderr, "dxyWindow: cannot write %s\n", synced);
			return -1;
		}
		return 0;
	}

	// ---- plan ---------------------------------------------------------------------------------
	const double t_scan = now_ms();
	std::vector<uint64_t> soff(runs.size() + 1, 0);
	for (size_t i = 0; i < runs.size(); ++i) soff[i + 1] = soff[i] + runs[i].count;
	uint32_t nchr = (uint32_t)runs.size();
	std::vector<uint64_t> axis = soff;  // plan axis: sites (-fixedsite 1) or bp (-fixedsite 0)
	const char* nosize = nullptr;       // first chromosome (data order) without a sizefile entry
	uint64_t print_limit = UINT64_MAX;  // bp mode with a missing size: windows whose last entry < limit
	uint64_t n_used = n;
	if (!fixedsite) {
		axis.assign(1, 0);
		for (uint32_t c = 0; c < nchr; ++c) {
			auto it = chrsize.find(runs[c].name);
			uint64_t L;
			if (it == chrsize.end()) {
				// dxyWindow.cpp:338-343,408-413: the reference has printed every window that flushed
				// before this chromosome ended, then fails.  Those are the windows completed by the
				// arrival of a later entry: last entry index <= (#entries so far) - 2.
				nosize = runs[c].name.c_str();
				L = pos[soff[c + 1] - 1];
				axis.push_back(axis.back() + L);
				print_limit = axis.back() - 1;
				nchr = c + 1;
				n_used = soff[c + 1];
				break;
			}
			L = it->second;
			axis.push_back(axis.back() + L);
		}
		for (uint32_t c = 0; c < nchr; ++c) {
			const uint64_t L = axis[c + 1] - axis[c];
			for (uint64_t i = soff[c]; i < soff[c + 1]; ++i) {
				if (pos[i] < 1 || pos[i] > L || (i > soff[c] && pos[i] <= pos[i - 1])) {
					fprintf(stderr, "dxyWindow: %s position %u (site %llu) is not strictly increasing within 1..%llu; bp windows need sorted positions inside the sizefile length\n",
					        runs[c].name.c_str(), pos[i], (unsigned long long)(i + 1), (unsigned long long)L);
					return -1;
				}
			}
		}
	}
	const unsigned W = winsize ? winsize : 1, S = winsize ? stepsize : 1;  // -winsize 0: only the global line is used
	pgt_plan* plan = nullptr;
	// bp mode: a reduction unit is a bp range; on sparse data (SNP-only MAFs) 256-bp units hold a handful
	// of sites each and the per-unit cost dominates, so the unit grows with the bp-per-site ratio
	// (measured on B200, 1e8 sites over 3e9 bp: 6.3 ms at 256 bp, 0.65 ms at 4096 bp).  The unit size
	// only changes the summation order.
	uint32_t unit_sites = 0;
	if (!fixedsite && n_used > 0) {
		const double bp_per_site = (double)axis[nchr] / (double)n_used;
		if (bp_per_site >= 1.5) {
			const double u = 341.0 * bp_per_site;
			unit_sites = u >= 4096.0 ? 4096u : ((uint32_t)u + 31u) / 32u * 32u;
		}
	}
	if (pgt_plan_create(&plan, fixedsite ? PGT_MODE_SITES : PGT_MODE_BP, axis.data(), nchr, W, S, unit_sites) != PGT_OK) {
		fprintf(stderr, "%s\n", pgt_last_error());
		return -1;
	}
	const uint64_t nwin = pgt_plan_num_windows(plan);
	std::vector<uint32_t> label(nwin), startp(nwin), endp(nwin), neff(nwin), nskip(nwin);
	std::vector<double> dxy(nwin);
	double global[3] = {0, 0, 0};
	{
		if (warm.finish() != 0) return -1;
		tm.cuda_init_ms = warm.ms;
		pgt_columns cols;
		memset(&cols, 0, sizeof(cols));
		cols.pos = pos.data();
		cols.f1 = f1.data();
		cols.f2 = f2.data();
		cols.n1 = ni1.data();
		cols.n2 = ni2.data();
		pgt_windows out;
		memset(&out, 0, sizeof(out));
		out.label = label.data();
		out.start_pos = startp.data();
		out.end_pos = endp.data();
		out.dxy = dxy.data();
		out.neffective = neff.data();
		out.nskip = nskip.data();
		out.dxy_global = global;
		pgt_range range;
		memset(&range, 0, sizeof(range));
		range.w_hi = nwin;
		range.site_count = n_used;
		std::vector<uint64_t> soff_used(soff.begin(), soff.begin() + nchr + 1);
		DeviceWorkspace ws;
		ws.bytes = pgt_scan_workspace_bytes(plan, &range, PGT_STAT_DXY, PGT_MEM_HOST);
		if (pgt_device_alloc(&ws.p, ws.bytes) != PGT_OK ||
		    pgt_scan(plan, &range, PGT_STAT_DXY, &cols, minind, fixedsite ? nullptr : soff_used.data(), &out, ws.p, ws.bytes, PGT_MEM_HOST, nullptr) != PGT_OK) {
			fprintf(stderr, "dxyWindow: %s\n", pgt_last_error());
			return -1;
		}
	}
	tm.windows = nwin;
	tm.scan_ms = now_ms() - t_scan;

	// ---- print --------------------------------------------------------------------------------
	const double t_fmt = now_ms();
	static char obuf[1 << 20];
	setvbuf(stdout, obuf, _IOFBF, sizeof(obuf));
	if (winsize > 0) {
		std::vector<uint64_t> last_entry;
		if (nosize) {
			last_entry.resize(nwin);
			pgt_plan_windows(plan, nullptr, last_entry.data(), nullptr);
		}
		size_t maxname = 0;
		for (const ContigRun& r : runs) maxname = std::max(maxname, r.name.size());
		write_rows(stdout, nwin, maxname + 96, [&](char* p, uint64_t w) {
			if (skip_missing && neff[w] == 0) return p;        // dxyWindow.cpp:189
			if (nosize && last_entry[w] >= print_limit) return p;  // never flushed before the failure
			const std::string& nm = runs[label[w]].name;
			memcpy(p, nm.data(), nm.size());
			p += nm.size();
			*p++ = '\t';
			p = put_i32(p, (int32_t)startp[w]);
			*p++ = '\t';
			p = put_i32(p, (int32_t)endp[w]);
			*p++ = '\t';
			p = put_g(p, dxy[w]);
			*p++ = '\t';
			p = put_u32(p, neff[w]);
			*p++ = '\t';
			p = put_u32(p, nskip[w]);
			*p++ = '\n';
			return p;
		});
	}
	fflush(stdout);
	int rv = 0;
	if (nosize) {
		fprintf(stderr, "Unable to determine size for %s\n", nosize);
		rv = -1;
	} else {
		// dxyWindow.cpp:429-433
		char line[128];
		char* p = put_g(line, global[0]);
		*p++ = '\t';
		p = put_u32(p, (uint32_t)global[1]);
		*p++ = '\t';
		p = put_u32(p, (uint32_t)global[2]);
		*p++ = '\n';
		fwrite(line, 1, (size_t)(p - line), winsize == 0 ? stdout : stderr);
		fflush(stdout);
	}
	tm.format_ms = now_ms() - t_fmt;
	tm.total_ms = now_ms() - t_start;
	tm.report("dxyWindow");
	pgt_plan_destroy(plan);
	return rv;
}
