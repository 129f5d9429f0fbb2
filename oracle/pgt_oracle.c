/*
 * pgt_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed
 * by the product path (popgenomicstools_b200/, the C-ABI library, the CLIs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * CPU restatement, over columnar arrays, of the reference's windowed
 * site-statistic scan:
 *   fst : calcFst + calcWindow            /root/reference/fstWindow.cpp:69-155
 *   het : calcHeterozygosity + calcWindow /root/reference/hetWindow.cpp:66-153
 *   dxy : maf2dxy + calcWindow            /root/reference/dxyWindow.cpp:172-209,300-433
 * It keeps the reference's OPERATIONAL form on purpose: one W-entry buffer,
 * strictly sequential double sums per flush, slide-by-copy of the W-S overlap,
 * the three flush triggers (contig change / buffer full / EOF rule) and, for
 * dxy bp mode, the dense per-bp filler stream.  That makes it (a) the FP and
 * enumeration parity oracle for the CUDA path and (b) a compute-only CPU
 * baseline with the reference's O(n*W/S) cost.
 *
 * Parity PINNED: tests/test_oracle_vs_reference.py checks this file's output
 * row-for-row against stdout of the unmodified reference binaries
 * (oracle/_ref/, built by oracle/Makefile from /root/reference) and against the
 * committed transcripts in tests/golden/ that those binaries produced.
 *
 * Inputs are the parsed columns (text parsing is not restated here; the
 * reference binaries themselves are the oracle for the CLI layer).  `chr` is a
 * per-site contig id; ids only need to differ where the contig name differs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PGT_ORACLE_ERR_ARGS (-1)
#define PGT_ORACLE_ERR_NOSIZE (-2)
#define PGT_ORACLE_ERR_NOMEM (-3)

/* ------------------------------------------------------------------ fst --- */

typedef struct {
	uint32_t pos;
	double a;
	double b;
	uint64_t idx;
} fst_rec;

typedef struct {
	uint64_t cap, nrow;
	uint32_t *label, *start, *end, *mid, *n;
	double *asum, *bsum, *fst;
	uint64_t *first, *last;
	volatile double sink;
} fst_out;

/* fstWindow.cpp:69-107 */
static uint32_t fst_flush(fst_rec* buf, uint32_t label, uint32_t W, uint32_t S, uint32_t nsites, fst_out* o) {
	uint32_t startpos = buf[0].pos;
	uint32_t endpos = buf[nsites - 1].pos;
	uint32_t mid = (startpos + endpos) / 2; /* uint32 wrap, fstWindow.cpp:73 */
	double asum = 0, bsum = 0;
	for (uint32_t i = 0; i < nsites; ++i) { /* fstWindow.cpp:80-83 */
		asum += buf[i].a;
		bsum += buf[i].b;
	}
	double fst = bsum != 0.0 ? asum / bsum : 0.0; /* fstWindow.cpp:85 */
	if (o->nrow < o->cap) {
		uint64_t r = o->nrow;
		if (o->label) o->label[r] = label;
		if (o->start) o->start[r] = startpos;
		if (o->end) o->end[r] = endpos;
		if (o->mid) o->mid[r] = mid;
		if (o->asum) o->asum[r] = asum;
		if (o->bsum) o->bsum[r] = bsum;
		if (o->fst) o->fst[r] = fst;
		if (o->n) o->n[r] = nsites;
		if (o->first) o->first[r] = buf[0].idx;
		if (o->last) o->last[r] = buf[nsites - 1].idx;
	}
	o->sink += fst;
	o->nrow++;
	if (nsites == W) { /* fstWindow.cpp:92-99: slide (also taken at a contig change) */
		uint32_t nnew = W - S;
		for (uint32_t i = 0; i < nnew; ++i) buf[i] = buf[S + i];
		return nnew;
	}
	return 0; /* fstWindow.cpp:100-103 */
}

int64_t pgt_oracle_fst(const uint32_t* chr, const uint32_t* pos, const double* a, const double* b, uint64_t n,
                       uint32_t W, uint32_t S, uint64_t cap, uint32_t* o_label, uint32_t* o_start, uint32_t* o_end,
                       uint32_t* o_mid, double* o_asum, double* o_bsum, double* o_fst, uint32_t* o_n,
                       uint64_t* o_first, uint64_t* o_last) {
	if (W < 1 || S < 1 || S > W) return PGT_ORACLE_ERR_ARGS; /* reference has UB outside 1<=S<=W */
	fst_rec* buf = (fst_rec*)malloc((size_t)W * sizeof(fst_rec));
	if (!buf) return PGT_ORACLE_ERR_NOMEM;
	fst_out o = {cap, 0, o_label, o_start, o_end, o_mid, o_n, o_asum, o_bsum, o_fst, o_first, o_last, 0.0};
	uint32_t nsites = 0, oldchr = 0, c = 0;
	for (uint64_t i = 0; i < n; ++i) { /* fstWindow.cpp:125-147 */
		c = chr[i];
		if (nsites > 0 && c != oldchr) nsites = fst_flush(buf, oldchr, W, S, nsites, &o);
		else if (nsites == W) nsites = fst_flush(buf, c, W, S, nsites, &o);
		buf[nsites].pos = pos[i];
		buf[nsites].a = a[i];
		buf[nsites].b = b[i];
		buf[nsites].idx = i;
		++nsites;
		oldchr = c;
	}
	if (nsites > (W - S) && nsites <= W) nsites = fst_flush(buf, c, W, S, nsites, &o); /* fstWindow.cpp:150-152 */
	free(buf);
	return (int64_t)o.nrow;
}

/* ------------------------------------------------------------------ het --- */

typedef struct {
	uint32_t pos;
	int32_t g;
	uint64_t idx;
} het_rec;

typedef struct {
	uint64_t cap, nrow;
	uint32_t *label, *start, *end, *mid, *nhet, *nonmissing;
	double* h;
	uint64_t *first, *last;
	volatile double sink;
} het_out;

/* hetWindow.cpp:66-105 */
static uint32_t het_flush(het_rec* buf, uint32_t label, uint32_t W, uint32_t S, uint32_t nsites, het_out* o) {
	uint32_t startpos = buf[0].pos;
	uint32_t lastpos = buf[nsites - 1].pos;
	uint32_t mid = (startpos + lastpos) / 2;
	uint32_t nonmissing = 0, nhet = 0;
	for (uint32_t i = 0; i < nsites; ++i) { /* hetWindow.cpp:77-82 */
		if (buf[i].g >= 0) {
			++nonmissing;
			if (buf[i].g == 1) ++nhet;
		}
	}
	double h = nonmissing != 0 ? (double)nhet / nonmissing : 0.0; /* hetWindow.cpp:84 */
	if (o->nrow < o->cap) {
		uint64_t r = o->nrow;
		if (o->label) o->label[r] = label;
		if (o->start) o->start[r] = startpos;
		if (o->end) o->end[r] = lastpos;
		if (o->mid) o->mid[r] = mid;
		if (o->nhet) o->nhet[r] = nhet;
		if (o->nonmissing) o->nonmissing[r] = nonmissing;
		if (o->h) o->h[r] = h;
		if (o->first) o->first[r] = buf[0].idx;
		if (o->last) o->last[r] = buf[nsites - 1].idx;
	}
	o->sink += h;
	o->nrow++;
	if (nsites == W) { /* hetWindow.cpp:90-97 */
		uint32_t nnew = W - S;
		for (uint32_t i = 0; i < nnew; ++i) buf[i] = buf[S + i];
		return nnew;
	}
	return 0;
}

int64_t pgt_oracle_het(const uint32_t* chr, const uint32_t* pos, const int8_t* geno, uint64_t n, uint32_t W,
                       uint32_t S, uint64_t cap, uint32_t* o_label, uint32_t* o_start, uint32_t* o_end,
                       uint32_t* o_mid, uint32_t* o_nhet, uint32_t* o_nonmissing, double* o_h, uint64_t* o_first,
                       uint64_t* o_last) {
	if (W < 1 || S < 1 || S > W) return PGT_ORACLE_ERR_ARGS;
	het_rec* buf = (het_rec*)malloc((size_t)W * sizeof(het_rec));
	if (!buf) return PGT_ORACLE_ERR_NOMEM;
	het_out o = {cap, 0, o_label, o_start, o_end, o_mid, o_nhet, o_nonmissing, o_h, o_first, o_last, 0.0};
	uint32_t nsites = 0, oldchr = 0, c = 0;
	for (uint64_t i = 0; i < n; ++i) { /* hetWindow.cpp:123-145 */
		c = chr[i];
		if (nsites > 0 && c != oldchr) nsites = het_flush(buf, oldchr, W, S, nsites, &o);
		else if (nsites == W) nsites = het_flush(buf, c, W, S, nsites, &o);
		buf[nsites].pos = pos[i];
		buf[nsites].g = geno[i];
		buf[nsites].idx = i;
		++nsites;
		oldchr = c;
	}
	if (nsites > (W - S) && nsites <= W) nsites = het_flush(buf, c, W, S, nsites, &o); /* hetWindow.cpp:148-150 */
	free(buf);
	return (int64_t)o.nrow;
}

/* ------------------------------------------------------------------ dxy --- */

typedef struct {
	int32_t pos;
	double v;
	int64_t entry; /* running index in the entry stream (sites + fillers) */
} dxy_rec;

typedef struct {
	uint64_t cap, nrow;
	uint32_t* label;
	int32_t *start, *end;
	double* dxy;
	uint32_t *neff, *nskip;
	int64_t *first, *last;
	int skip_missing;
	volatile double sink;
} dxy_out;

/* dxyWindow.cpp:172-209 */
static uint32_t dxy_flush(dxy_rec* buf, uint32_t label, uint32_t W, uint32_t S, uint32_t nsites, dxy_out* o) {
	double dxy = 0;
	uint32_t neffective = 0;
	int nskip = 0;
	for (uint32_t i = 0; i < nsites; ++i) { /* dxyWindow.cpp:179-186 */
		if (buf[i].v >= 0) {
			dxy += buf[i].v;
			++neffective;
		} else if (buf[i].v == -9) {
			++nskip;
		}
	}
	if (neffective > 0 || !o->skip_missing) { /* dxyWindow.cpp:189-191 */
		if (o->nrow < o->cap) {
			uint64_t r = o->nrow;
			if (o->label) o->label[r] = label;
			if (o->start) o->start[r] = buf[0].pos;
			if (o->end) o->end[r] = buf[nsites - 1].pos;
			if (o->dxy) o->dxy[r] = dxy;
			if (o->neff) o->neff[r] = neffective;
			if (o->nskip) o->nskip[r] = (uint32_t)nskip;
			if (o->first) o->first[r] = buf[0].entry;
			if (o->last) o->last[r] = buf[nsites - 1].entry;
		}
		o->nrow++;
	}
	o->sink += dxy;
	if (nsites == W) { /* dxyWindow.cpp:194-201 */
		uint32_t nnew = W - S;
		for (uint32_t i = 0; i < nnew; ++i) buf[i] = buf[S + i];
		return nnew;
	}
	return 0;
}

/* dxyWindow.cpp:381 -- plain double mul/sub/add, compiled with -ffp-contract=off */
static double dxy_site(double f1, double f2, int n1, int n2, int minind) {
	return (n1 >= minind && n2 >= minind) ? f1 * (1.0 - f2) + f2 * (1.0 - f1) : -9;
}

/*
 * maf2dxy main loop after the two-file sync (dxyWindow.cpp:300-433).
 * Columns hold the SYNCED sites (both files agree on chr/pos).
 *   fixedsite != 0 : windows of W sites  (dxyWindow.cpp:357-359,376-378)
 *   fixedsite == 0 : windows of W bp over the dense per-bp entry stream; chr_len[id]
 *                    is the sizefile length, 0 = not listed -> PGT_ORACLE_ERR_NOSIZE
 *                    at the point the reference would fail (rows before it are kept,
 *                    *o_rows_before_error gets their count).
 *   W == 0         : global only (requires fixedsite != 0; the reference segfaults otherwise).
 * global[3] = {dxy_global, neffective_global, nskip_global} (dxyWindow.cpp:382-385,429-433).
 */
int64_t pgt_oracle_dxy(const uint32_t* chr, const uint32_t* pos, const double* f1, const double* f2,
                       const int32_t* n1, const int32_t* n2, uint64_t n, int minind, uint32_t W, uint32_t S,
                       int fixedsite, int skip_missing, const uint32_t* chr_len, uint32_t n_chr_len, uint64_t cap,
                       uint32_t* o_label, int32_t* o_start, int32_t* o_end, double* o_dxy, uint32_t* o_neff,
                       uint32_t* o_nskip, int64_t* o_first, int64_t* o_last, double* global,
                       uint64_t* o_rows_before_error) {
	if (W > 0 && (S < 1 || S > W)) return PGT_ORACLE_ERR_ARGS;
	if (W == 0 && !fixedsite) return PGT_ORACLE_ERR_ARGS;
	if (n == 0) return PGT_ORACLE_ERR_ARGS; /* reference reads garbage on empty MAFs */
	dxy_rec* buf = (dxy_rec*)malloc((size_t)(W ? W : 1) * sizeof(dxy_rec));
	if (!buf) return PGT_ORACLE_ERR_NOMEM;
	dxy_out o = {cap, 0, o_label, o_start, o_end, o_dxy, o_neff, o_nskip, o_first, o_last, skip_missing, 0.0};
	uint32_t nsites = 0;
	uint32_t c = chr[0], prevchr = chr[0];
	uint32_t positer = 1, lastpos;
	int64_t entry = 0;
	double dxy_global = 0;
	uint32_t neffective_global = 0, nskip_g = 0;
	int64_t rc = 0;

#define PGT_FILL_TO(LIMIT_EXPR, LABEL, INCLUSIVE)                                               \
	while ((INCLUSIVE) ? (positer <= (LIMIT_EXPR)) : (positer < (LIMIT_EXPR))) {                  \
		if (nsites == W) nsites = dxy_flush(buf, (LABEL), W, S, nsites, &o);                      \
		buf[nsites].pos = (int32_t)positer;                                                       \
		buf[nsites].v = -7;                                                                       \
		buf[nsites].entry = entry++;                                                              \
		++nsites;                                                                                 \
		++positer;                                                                                \
	}

	for (uint64_t i = 0; i < n; ++i) {
		c = chr[i];
		if (W > 0 && c != prevchr) { /* dxyWindow.cpp:334-361 */
			if (!fixedsite) {
				if (prevchr < n_chr_len && chr_len[prevchr] > 0) lastpos = chr_len[prevchr];
				else { rc = PGT_ORACLE_ERR_NOSIZE; goto done; }
				PGT_FILL_TO(lastpos, prevchr, 1)
				if (nsites > (W - S)) nsites = dxy_flush(buf, prevchr, W, S, nsites, &o);
			} else {
				if (nsites > 0) nsites = dxy_flush(buf, prevchr, W, S, nsites, &o);
			}
			positer = 1;
		}
		if (W > 0 && !fixedsite) { /* dxyWindow.cpp:363-373 */
			PGT_FILL_TO(pos[i], c, 0)
		}
		if (W > 0 && nsites == W) nsites = dxy_flush(buf, c, W, S, nsites, &o); /* dxyWindow.cpp:376-378 */
		double dxy = dxy_site(f1[i], f2[i], n1[i], n2[i], minind);
		if (dxy != -9) { /* dxyWindow.cpp:382-385 */
			dxy_global += dxy;
			++neffective_global;
		} else ++nskip_g;
		if (W > 0) { /* dxyWindow.cpp:388-394 */
			buf[nsites].pos = (int32_t)pos[i];
			buf[nsites].v = dxy;
			buf[nsites].entry = entry++;
			++nsites;
			++positer;
		}
		prevchr = c;
	}
	if (!fixedsite) { /* dxyWindow.cpp:407-423 */
		if (c < n_chr_len && chr_len[c] > 0) lastpos = chr_len[c];
		else { rc = PGT_ORACLE_ERR_NOSIZE; goto done; }
		PGT_FILL_TO(lastpos, c, 1)
	}
	if (nsites > (W - S) && nsites <= W) nsites = dxy_flush(buf, c, W, S, nsites, &o); /* dxyWindow.cpp:424-426 */
done:
#undef PGT_FILL_TO
	if (global) {
		global[0] = dxy_global;
		global[1] = (double)neffective_global;
		global[2] = (double)nskip_g;
	}
	if (o_rows_before_error) *o_rows_before_error = o.nrow;
	free(buf);
	return rc < 0 ? rc : (int64_t)o.nrow;
}
