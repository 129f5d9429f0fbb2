// pgt_geom.h -- reduction geometry shared by the host planner and the CUDA kernels.
//
// The reference re-sums a W-entry buffer per window and slides it by copy
// (/root/reference/fstWindow.cpp:80-83,92-99; hetWindow.cpp:77-82,90-97;
// dxyWindow.cpp:179-186,194-201).  Here every site is read once: the site axis of a
// *segment* (a maximal run of contigs the reference's buffer never fully clears, SURVEY.md
// Appendix A.1) is cut at every window boundary -- window starts k*S and window ends k*S+W,
// i.e. offsets 0 and r = W mod S inside each step "pair" [k*S,(k+1)*S) -- into *pieces*
// A = [k*S, k*S+r) and B = [k*S+r, (k+1)*S); pieces are cut again every `u` sites from the
// piece start into *units* (<= u sites).  Every window is then an exact union of consecutive
// units:  window k = units [k*upp, k*upp + q*upp + upa),  q = W div S.
// The geometry depends only on (W, S, u) and the segment origin, never on tiles, CTAs or
// shards, which makes results bit-identical for any GPU count.
#ifndef PGT_GEOM_H
#define PGT_GEOM_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PGT_GEOM_HD __host__ __device__ __forceinline__
#else
#define PGT_GEOM_HD static inline
#endif

struct pgt_geom {
	uint32_t W, S, u;
	uint32_t r;       // W % S   : size of piece A
	uint32_t q;       // W / S
	uint32_t upa;     // units in piece A = ceil(r / u)   (0 when r == 0)
	uint32_t upb;     // units in piece B = ceil((S - r) / u)
	uint32_t upp;     // units per pair   = upa + upb
	uint32_t wunits;  // units per full window = q * upp + upa
	uint32_t ueff;    // longest unit = min(u, max(r, S - r))
	uint32_t gw;      // lanes that reduce one unit: min(32, pow2ceil(ceil(ueff / 8)))
	uint32_t pad;
};

struct pgt_seg {
	uint64_t site_base;  // global site (entry) index of the segment's first site
	uint64_t nsites;     // N
	uint64_t unit_base;  // global index of its first unit
	uint64_t nunits;
	uint64_t win_base;   // global index of its first emitted window
	uint64_t nwin;       // emitted windows (full + emitted trailing partial)
	uint64_t nfull;      // K = full windows
	uint64_t blk_base;   // global index of its first scan block (blocks of g.wunits units, see k_block_scan)
	uint32_t first_contig;
	uint32_t ncontig;
};

PGT_GEOM_HD pgt_geom pgt_make_geom(uint32_t W, uint32_t S, uint32_t u) {
	pgt_geom g;
	g.W = W;
	g.S = S;
	g.u = u;
	g.r = W % S;
	g.q = W / S;
	g.upa = (g.r + u - 1) / u;
	g.upb = (S - g.r + u - 1) / u;
	g.upp = g.upa + g.upb;
	g.wunits = g.q * g.upp + g.upa;
	const uint32_t piece = g.r > S - g.r ? g.r : S - g.r;
	g.ueff = piece < u ? piece : u;
	uint32_t need = (g.ueff + 7) / 8, gw = 1;
	while (gw < need && gw < 32) gw <<= 1;
	g.gw = gw;
	g.pad = 0;
	return g;
}

// Unit j (segment-local) of a segment with N sites: [*start, *start + len) segment-local.
PGT_GEOM_HD uint32_t pgt_unit_range(const pgt_geom& g, uint64_t N, uint64_t j, uint64_t* start) {
	uint64_t k;
	uint32_t s;
	if (g.upp == 1) {  // every piece is one unit (r == 0 and S <= u): no division
		k = j;
		s = 0;
	} else if ((j >> 32) == 0) {
		const uint32_t jj = (uint32_t)j;
		const uint32_t kk = jj / g.upp;
		k = kk;
		s = jj - kk * g.upp;
	} else {
		k = j / g.upp;
		s = (uint32_t)(j - k * g.upp);
	}
	uint64_t st, bound;
	if (s < g.upa) {
		st = k * g.S + (uint64_t)s * g.u;
		bound = k * g.S + g.r;
	} else {
		st = k * g.S + g.r + (uint64_t)(s - g.upa) * g.u;
		bound = (k + 1) * g.S;
	}
	uint64_t en = st + g.u;
	if (en > bound) en = bound;
	if (en > N) en = N;
	*start = st;
	return en > st ? (uint32_t)(en - st) : 0u;
}

// Index (segment-local) of the unit containing segment-local site x.
PGT_GEOM_HD uint64_t pgt_unit_of_site(const pgt_geom& g, uint64_t x) {
	uint64_t k = x / g.S;
	uint32_t off = (uint32_t)(x - k * g.S);
	uint32_t s = off < g.r ? off / g.u : g.upa + (off - g.r) / g.u;
	return k * g.upp + s;
}

PGT_GEOM_HD uint64_t pgt_seg_nunits(const pgt_geom& g, uint64_t N) {
	return N == 0 ? 0 : pgt_unit_of_site(g, N - 1) + 1;
}

// Units of emitted window k of a segment: [*first, *first + count) segment-local.
PGT_GEOM_HD uint64_t pgt_window_units(const pgt_geom& g, const pgt_seg& sg, uint64_t k, uint64_t* first) {
	*first = k * g.upp;
	return k < sg.nfull ? (uint64_t)g.wunits : sg.nunits - k * g.upp;
}

// Sites of emitted window k: first site (segment-local) k*S, count min(W, N - k*S).
PGT_GEOM_HD uint32_t pgt_window_sites(const pgt_geom& g, const pgt_seg& sg, uint64_t k, uint64_t* first) {
	*first = k * g.S;
	uint64_t rem = sg.nsites - k * g.S;
	return rem < g.W ? (uint32_t)rem : g.W;
}

#endif  // PGT_GEOM_H
