/*
 * pgt_oracle_extreme.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * CPU restatement, over columnar arrays, of the bp-window "most extreme score" scans
 *   ihs   : calciHSWindows + updateMax + printWindow   /root/reference/ihsWindow.cpp:69-187
 *   xpehh : calcXpehhWindows + updateOutlier           /root/reference/xpehhWindow.cpp:59-193
 * in the reference's OPERATIONAL form: one pass over the sites carrying (winstart, winend, nsites,
 * nbig, best) exactly as the reference's line loop does, incl. its quirks (all probed on the
 * compiled binaries):
 *   - a site flushes the current window when pos >= winend, but empty windows are only skipped
 *     while pos > winend (ihsWindow.cpp:145,151): a site AT a window end lands in that window
 *     when arriving from an earlier one, in the next window when that window is already open;
 *   - the first site of every chromosome but the first is put into window [1, W] whatever its
 *     position (the else-if at :145 is not evaluated after a chromosome change, :130-144);
 *   - the first window of the first chromosome is not clipped to the chromosome length (:99),
 *     every later window is (:142,148,154);
 *   - strict comparisons keep the FIRST extreme on ties (:166, xpehhWindow.cpp:172,175);
 *   - unsigned 32-bit window arithmetic.
 * Parity PINNED: tests/test_extreme_oracle.py checks this file row-for-row against the
 * unmodified reference binaries (oracle/_ref/{ihsWindow,xpehhWindow}, built by oracle/Makefile)
 * and against the committed transcripts tests/golden/ref_transcripts_extreme.json.
 *
 * `chr` is a per-site name id (equal ids <=> equal names); chrlen[id] is the -chrlen length of
 * that name, 0 when absent (lenmap lookup at :125,139-141).
 */
#include <math.h>
#include <stdint.h>

#define PGT_ORACLE_ERR_ARGS (-1)
#define PGT_ORACLE_ERR_LOOP (-4) /* the reference would print empty windows forever (pos > chrlen) */

typedef struct {
	uint64_t cap, nrow;
	uint32_t *label, *start, *end, *n, *nbig, *extpos;
	double *ext, *prop;
	uint64_t* first;
} x_out;

/* printWindow, ihsWindow.cpp:75-84 */
static void x_row(x_out* o, uint32_t label, uint32_t ws, uint32_t we, int nbig, uint32_t nsites, const double* best, int which,
                  uint64_t first) {
	if (o->nrow < o->cap) {
		uint64_t r = o->nrow;
		if (o->label) o->label[r] = label;
		if (o->start) o->start[r] = ws;
		if (o->end) o->end[r] = we;
		if (o->n) o->n[r] = nsites;
		if (o->first) o->first[r] = first;
		if (nsites > 0) {
			if (o->ext) o->ext[r] = best[which];
			if (o->extpos) o->extpos[r] = (uint32_t)best[2];
			if (o->nbig) o->nbig[r] = (uint32_t)nbig;
			if (o->prop) o->prop[r] = (double)nbig / nsites;
		} else {
			if (o->ext) o->ext[r] = NAN;
			if (o->extpos) o->extpos[r] = 0;
			if (o->nbig) o->nbig[r] = 0;
			if (o->prop) o->prop[r] = NAN;
		}
	}
	o->nrow++;
}

/* mode 0 = ihsWindow (key |v|, count |v| > cutoff); mode 1 = xpehhWindow (cutoff < 0: most
 * negative and count v < cutoff; else most positive and count v > cutoff).
 * Returns the number of rows (may exceed cap), or a negative error. */
int64_t pgt_oracle_extreme(int mode, const uint32_t* chr, const uint32_t* pos, const double* val, uint64_t n,
                           const uint32_t* chrlen, uint32_t nchrlen, uint32_t W, double cutoff, uint64_t cap, uint32_t* label,
                           uint32_t* start, uint32_t* end, double* ext, uint32_t* extpos, uint32_t* nbig_out, double* prop,
                           uint32_t* nsites_out, uint64_t* first) {
	if (W == 0) return PGT_ORACLE_ERR_ARGS;
	x_out o = {cap, 0, label, start, end, nsites_out, nbig_out, extpos, ext, prop, first};
	uint32_t winstart = 1;
	uint32_t winend = winstart + (W - 1); /* ihsWindow.cpp:98-99: not clipped */
	double best[3] = {0, 0, 0};          /* ihs: |v|, v, pos ; xpehh: v, (unused), pos */
	const int which = mode == 0 ? 1 : 0;  /* printed member: maxihs[1] / outlierstat[0] */
	uint32_t nsites = 0;
	int nbig = 0;
	uint32_t chrlen_cur = 0;
	uint32_t cur = 0; /* id of `chr` */
	uint64_t wfirst = 0;
#define LEN_OF(id) ((chrlen && (id) < nchrlen) ? chrlen[(id)] : 0u)
	for (uint64_t i = 0; i < n; ++i) {
		const uint32_t newchr = chr[i];
		const uint32_t p = pos[i];
		if (i == 0) { /* chr.empty(), :123-126 */
			cur = newchr;
			chrlen_cur = LEN_OF(newchr);
		}
		if (newchr != cur) { /* :130-144 */
			x_row(&o, cur, winstart, winend, nbig, nsites, best, which, wfirst);
			while (winend < chrlen_cur) {
				winstart = winend + 1;
				winend = winstart + (W - 1);
				if (chrlen_cur && winend > chrlen_cur) winend = chrlen_cur;
				x_row(&o, cur, winstart, winend, 0, 0, best, which, i);
			}
			nsites = 0;
			chrlen_cur = LEN_OF(newchr);
			winstart = 1;
			winend = winstart + (W - 1);
			if (chrlen_cur && winend > chrlen_cur) winend = chrlen_cur;
		} else if (p >= winend) { /* :145-157 */
			x_row(&o, cur, winstart, winend, nbig, nsites, best, which, wfirst);
			winstart = winend + 1;
			winend = winstart + (W - 1);
			if (chrlen_cur && winend > chrlen_cur) winend = chrlen_cur;
			nsites = 0;
			while (p > winend) {
				if (chrlen_cur && winend >= chrlen_cur) return PGT_ORACLE_ERR_LOOP;
				x_row(&o, cur, winstart, winend, 0, 0, best, which, i);
				winstart = winend + 1;
				winend = winstart + (W - 1);
				if (chrlen_cur && winend > chrlen_cur) winend = chrlen_cur;
			}
		}
		/* update window, :159-177 / xpehhWindow.cpp:164-181 */
		const double v = val[i];
		if (nsites == 0) {
			nbig = 0;
			wfirst = i;
		}
		if (mode == 0) {
			const double av = fabs(v);
			if (nsites == 0 || av > best[0]) {
				best[0] = av;
				best[1] = v;
				best[2] = p;
			}
			if (av > cutoff) ++nbig;
		} else {
			if (nsites == 0) {
				best[0] = v;
				best[2] = p;
			}
			if (cutoff < 0) {
				if (v < best[0]) {
					best[0] = v;
					best[2] = p;
				}
				if (v < cutoff) ++nbig;
			} else {
				if (v > best[0]) {
					best[0] = v;
					best[2] = p;
				}
				if (v > cutoff) ++nbig;
			}
		}
		cur = newchr;
		++nsites;
	}
	/* last windows, :179-186 (also printed for an input without sites, with an empty name) */
	x_row(&o, n ? cur : 0xffffffffu, winstart, winend, nbig, nsites, best, which, wfirst);
	while (winend < chrlen_cur) {
		winstart = winend + 1;
		winend = winstart + (W - 1);
		if (winend > chrlen_cur) winend = chrlen_cur;
		x_row(&o, cur, winstart, winend, 0, 0, best, which, n);
	}
#undef LEN_OF
	return (int64_t)o.nrow;
}
