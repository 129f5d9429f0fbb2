// pgt_plan.cpp -- closed-form window enumeration (host, O(#contigs), no data touched).
//
// Replaces the flush-trigger state machines of the reference
//   calcFst              /root/reference/fstWindow.cpp:125-152
//   calcHeterozygosity   /root/reference/hetWindow.cpp:123-150
//   maf2dxy              /root/reference/dxyWindow.cpp:334-361,376-378,407-426
// by the segment model of SURVEY.md Appendix A.1/A.2:
//   * contigs are concatenated into segments; after contig i (not the last) the segment
//     CONTINUES iff the reference's buffer is exactly full there, N >= W && (N-W) % S == 0
//     (calcWindow then takes its "same chromosome" branch, fstWindow.cpp:92-99), or -- bp mode
//     only -- iff N <= W-S (dxyWindow.cpp:353-355 neither flushes nor clears: stale carry);
//   * a segment of N sites has K = (N>=W ? (N-W)/S+1 : 0) full windows [kS, kS+W) and, if
//     N < W or (N-W) % S != 0, a trailing partial window [KS, N), which is dropped when it has
//     <= W-S sites at EOF (fstWindow.cpp:150-152; dxyWindow.cpp:353,424 apply the same rule at
//     every chromosome end in bp mode);
//   * label = contig of the window's last site.
#include <algorithm>
#include <cstring>
#include <new>

#include "pgt_internal.h"

static thread_local std::string g_last_error;

int pgt_set_error(int code, const std::string& msg) {
	g_last_error = msg;
	return code;
}

extern "C" const char* pgt_last_error(void) { return g_last_error.c_str(); }
extern "C" int pgt_abi_version(void) { return PGT_ABI_VERSION; }

extern "C" int pgt_plan_create(pgt_plan** out, pgt_mode mode, const uint64_t* contig_offsets, uint32_t ncontig,
                               uint32_t W, uint32_t S, uint32_t unit_sites) {
	if (!out) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_create: plan pointer is NULL");
	*out = nullptr;
	if (mode != PGT_MODE_SITES && mode != PGT_MODE_BP) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_create: unknown mode");
	if (W < 1) return pgt_set_error(PGT_ERR_ARGS, "Window size must be a positive integer");
	if (S < 1) return pgt_set_error(PGT_ERR_ARGS, "Step size must be a positive integer");
	if (S > W) return pgt_set_error(PGT_ERR_ARGS, "Step size must not exceed the window size");
	if (!contig_offsets && ncontig) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_create: contig_offsets is NULL");
	uint32_t u = unit_sites ? unit_sites : 256u;
	if (u % 32u != 0 || u > 4096u) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_create: unit_sites must be a multiple of 32, <= 4096");
	for (uint32_t c = 0; c < ncontig; ++c)
		if (contig_offsets[c + 1] < contig_offsets[c]) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_create: contig_offsets must be non-decreasing");

	pgt_plan* p = new (std::nothrow) pgt_plan();
	if (!p) return pgt_set_error(PGT_ERR_NOMEM, "pgt_plan_create: out of memory");
	p->mode = mode;
	p->g = pgt_make_geom(W, S, u);
	if (ncontig) p->off.assign(contig_offsets, contig_offsets + ncontig + 1);
	else p->off.assign(1, 0);
	const uint64_t origin = p->off[0];
	p->nsites = p->off[ncontig] - origin;
	p->nwin = 0;
	p->nunits = 0;
	p->nblocks = 0;
	p->window_units_total = 0;
	p->scan_units_total = 0;

	uint32_t last_nonempty = ncontig;  // index of the last non-empty contig
	for (uint32_t c = ncontig; c-- > 0;)
		if (p->off[c + 1] > p->off[c]) {
			last_nonempty = c;
			break;
		}

	uint64_t N = 0, base = origin;
	uint32_t firstc = 0;
	for (uint32_t c = 0; c < ncontig; ++c) {
		uint64_t len = p->off[c + 1] - p->off[c];
		if (len == 0) continue;
		if (N == 0) firstc = c;
		N += len;
		const bool eof = (c == last_nonempty);
		bool cont = false;
		if (!eof) {
			if (N >= W && (N - W) % S == 0) cont = true;             // exactly-full buffer: carry
			else if (mode == PGT_MODE_BP && N <= (uint64_t)(W - S)) cont = true;  // stale carry
		}
		if (cont) continue;
		pgt_seg sg;
		std::memset(&sg, 0, sizeof(sg));
		sg.site_base = base;
		sg.nsites = N;
		sg.first_contig = firstc;
		sg.ncontig = c - firstc + 1;
		sg.nfull = N >= W ? (N - W) / S + 1 : 0;
		const bool has_partial = N < W || (N - W) % S != 0;
		const uint64_t rp = N - sg.nfull * S;
		const bool drop = (eof || mode == PGT_MODE_BP) && rp <= (uint64_t)(W - S);
		sg.nwin = sg.nfull + ((has_partial && !drop) ? 1 : 0);
		sg.nunits = pgt_seg_nunits(p->g, N);
		sg.win_base = p->nwin;
		sg.unit_base = p->nunits;
		sg.blk_base = p->nblocks;
		p->nwin += sg.nwin;
		p->nunits += sg.nunits;
		// partials a direct level 2 reads: wunits per full window, the rest of the segment for a trailing partial
		p->window_units_total += sg.nfull * (uint64_t)p->g.wunits + (sg.nwin > sg.nfull ? sg.nunits - sg.nfull * (uint64_t)p->g.upp : 0);
		p->scan_units_total += (sg.nunits + p->g.wunits - 1) / p->g.wunits * (uint64_t)p->g.wunits;
		p->nblocks += (sg.nunits + p->g.wunits - 1) / p->g.wunits;
		p->segs.push_back(sg);
		base += N;
		N = 0;
	}
	*out = p;
	return PGT_OK;
}

extern "C" void pgt_plan_destroy(pgt_plan* plan) { delete plan; }

extern "C" uint64_t pgt_plan_num_windows(const pgt_plan* p) { return p ? p->nwin : 0; }
extern "C" uint64_t pgt_plan_num_units(const pgt_plan* p) { return p ? p->nunits : 0; }
extern "C" uint32_t pgt_plan_num_segments(const pgt_plan* p) { return p ? (uint32_t)p->segs.size() : 0; }
extern "C" uint64_t pgt_plan_num_sites(const pgt_plan* p) { return p ? p->nsites : 0; }

uint32_t pgt_plan_seg_of_window(const pgt_plan* p, uint64_t w) {
	// last segment with win_base <= w (segments without windows share win_base with their successor)
	uint32_t lo = 0, hi = (uint32_t)p->segs.size();
	while (hi - lo > 1) {
		uint32_t mid = lo + (hi - lo) / 2;
		if (p->segs[mid].win_base <= w) lo = mid;
		else hi = mid;
	}
	return lo;
}

uint32_t pgt_plan_seg_of_unit(const pgt_plan* p, uint64_t j) {
	uint32_t lo = 0, hi = (uint32_t)p->segs.size();
	while (hi - lo > 1) {
		uint32_t mid = lo + (hi - lo) / 2;
		if (p->segs[mid].unit_base <= j) lo = mid;
		else hi = mid;
	}
	return lo;
}

static uint32_t contig_of(const pgt_plan* p, uint64_t x) {
	// contig c with off[c] <= x < off[c+1]
	size_t c = std::upper_bound(p->off.begin(), p->off.end(), x) - p->off.begin();
	return (uint32_t)(c - 1);
}
uint32_t pgt_plan_contig_of(const pgt_plan* p, uint64_t x) { return contig_of(p, x); }

uint64_t pgt_plan_unit_containing(const pgt_plan* p, uint64_t x) {
	uint32_t lo = 0, hi = (uint32_t)p->segs.size();
	while (hi - lo > 1) {
		uint32_t mid = lo + (hi - lo) / 2;
		if (p->segs[mid].site_base <= x) lo = mid;
		else hi = mid;
	}
	const pgt_seg& sg = p->segs[lo];
	return sg.unit_base + pgt_unit_of_site(p->g, x - sg.site_base);
}

uint64_t pgt_plan_unit_start(const pgt_plan* p, uint64_t j) {
	if (j >= p->nunits) return p->off[0] + p->nsites;
	const pgt_seg& sg = p->segs[pgt_plan_seg_of_unit(p, j)];
	uint64_t st;
	pgt_unit_range(p->g, sg.nsites, j - sg.unit_base, &st);
	return sg.site_base + st;
}

// Largest wb in (w, w_hi] such that every window of [w, wb) ends at or before global site `limit`
// (exclusive end); at least w + 1.  Window ends are non-decreasing in the window index.  O(#segments walked).
uint64_t pgt_plan_windows_within(const pgt_plan* p, uint64_t w, uint64_t w_hi, uint64_t limit) {
	uint64_t wb = w;
	for (uint32_t si = pgt_plan_seg_of_window(p, w); si < p->segs.size() && wb < w_hi; ++si) {
		const pgt_seg& sg = p->segs[si];
		if (sg.nwin == 0 || sg.win_base + sg.nwin <= wb) continue;
		if (sg.site_base + sg.nsites <= limit) {  // the whole segment fits
			wb = sg.win_base + sg.nwin;
			continue;
		}
		// full windows k with site_base + k*S + W <= limit; a trailing partial window ends at the segment end (> limit)
		uint64_t nfit = 0;
		if (limit >= sg.site_base + p->g.W) nfit = std::min<uint64_t>((limit - sg.site_base - p->g.W) / p->g.S + 1, sg.nfull);
		wb = std::max(wb, sg.win_base + nfit);
		break;
	}
	if (wb > w_hi) wb = w_hi;
	return wb > w ? wb : w + 1;
}

extern "C" int pgt_plan_window(const pgt_plan* p, uint64_t w, uint64_t* first, uint64_t* last, uint32_t* label) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_window: plan is NULL");
	if (w >= p->nwin) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_window: window index out of range");
	const pgt_seg& sg = p->segs[pgt_plan_seg_of_window(p, w)];
	uint64_t f;
	uint32_t cnt = pgt_window_sites(p->g, sg, w - sg.win_base, &f);
	uint64_t gf = sg.site_base + f, gl = gf + cnt - 1;
	if (first) *first = gf;
	if (last) *last = gl;
	if (label) *label = contig_of(p, gl);
	return PGT_OK;
}

extern "C" int pgt_plan_windows(const pgt_plan* p, uint64_t* first, uint64_t* last, uint32_t* label) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_windows: plan is NULL");
	for (const pgt_seg& sg : p->segs) {
		uint32_t c = sg.first_contig;
		for (uint64_t k = 0; k < sg.nwin; ++k) {
			uint64_t f;
			uint32_t cnt = pgt_window_sites(p->g, sg, k, &f);
			uint64_t gf = sg.site_base + f, gl = gf + cnt - 1;
			uint64_t w = sg.win_base + k;
			if (first) first[w] = gf;
			if (last) last[w] = gl;
			if (label) {
				while (p->off[c + 1] <= gl) ++c;  // last sites are non-decreasing in k
				label[w] = c;
			}
		}
	}
	return PGT_OK;
}

extern "C" int pgt_plan_unit(const pgt_plan* p, uint64_t j, uint64_t* start, uint32_t* len) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_unit: plan is NULL");
	if (j >= p->nunits) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_unit: unit index out of range");
	const pgt_seg& sg = p->segs[pgt_plan_seg_of_unit(p, j)];
	uint64_t st;
	uint32_t l = pgt_unit_range(p->g, sg.nsites, j - sg.unit_base, &st);
	if (start) *start = sg.site_base + st;
	if (len) *len = l;
	return PGT_OK;
}

extern "C" int pgt_plan_window_units(const pgt_plan* p, uint64_t w, uint64_t* first_unit, uint64_t* count) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_window_units: plan is NULL");
	if (w >= p->nwin) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_window_units: window index out of range");
	const pgt_seg& sg = p->segs[pgt_plan_seg_of_window(p, w)];
	uint64_t f;
	uint64_t cnt = pgt_window_units(p->g, sg, w - sg.win_base, &f);
	if (first_unit) *first_unit = sg.unit_base + f;
	if (count) *count = cnt;
	return PGT_OK;
}

// first emitted window whose first site is >= x (global site index); nwin if none
static uint64_t first_window_at_or_after(const pgt_plan* p, uint64_t x) {
	for (const pgt_seg& sg : p->segs) {  // O(#segments); called nshards times
		if (sg.nwin == 0) continue;
		if (x <= sg.site_base) return sg.win_base;
		uint64_t k = (x - sg.site_base + p->g.S - 1) / p->g.S;
		if (k < sg.nwin) return sg.win_base + k;
	}
	return p->nwin;
}

extern "C" int pgt_plan_shard(const pgt_plan* p, uint32_t shard, uint32_t nshards, uint64_t* w_lo, uint64_t* w_hi,
                              uint64_t* site_lo, uint64_t* site_hi) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_shard: plan is NULL");
	if (nshards == 0 || shard >= nshards) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_shard: shard out of range");
	const uint64_t origin = p->off[0];
	auto cut = [&](uint32_t i) -> uint64_t {
		if (i == 0) return 0;
		if (i >= nshards) return p->nwin;
		// balance by sites read: target site boundary i/nshards of the axis
		unsigned __int128 t = (unsigned __int128)p->nsites * i / nshards;
		return first_window_at_or_after(p, origin + (uint64_t)t);
	};
	uint64_t lo = cut(shard), hi = cut(shard + 1);
	if (hi < lo) hi = lo;
	// Site range.  The shard holding window 0 also owns the axis head, the shard holding the last
	// window also owns the tail outside any window (dropped EOF partial): those sites count
	// towards dxyWindow's global line (dxyWindow.cpp:382-385).  Shards without windows hold nothing,
	// except shard 0 of a plan without any window (it owns the whole axis).
	const uint64_t end = origin + p->nsites;
	uint64_t slo = end, shi = end;
	if (hi > lo) {
		uint64_t f, l;
		pgt_plan_window(p, lo, &f, nullptr, nullptr);
		pgt_plan_window(p, hi - 1, nullptr, &l, nullptr);
		slo = lo == 0 ? origin : f;
		shi = hi == p->nwin ? end : l + 1;
	} else if (p->nwin == 0 && shard == 0) {
		slo = origin;
	}
	if (w_lo) *w_lo = lo;
	if (w_hi) *w_hi = hi;
	if (site_lo) *site_lo = slo;
	if (site_hi) *site_hi = shi;
	return PGT_OK;
}
