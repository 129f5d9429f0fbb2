/*
 * pgt_synth_cpu.c -- TEST INFRASTRUCTURE ONLY (see pgt_oracle.c header).
 *
 * CPU twin of the on-device synthetic generator: fills columnar arrays from
 * include/pgt_synth.h (bit-identical to popgenomicstools_b200/csrc/pgt_synth.cu)
 * and writes the same sites as the text formats the reference binaries parse:
 *   fst  "chr pos a b"                         /root/reference/fstWindow.cpp:141
 *   het  "chr pos genotype"                    /root/reference/hetWindow.cpp:139
 *   mafs "chromo position major minor ref knownEM nInd" + header line
 *                                              /root/reference/dxyWindow.cpp:146-152,284
 * Values are printed from their integer micro-units with exactly 6 decimals, so
 * strtod(text) == (double)k / 1e6 on every platform.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../include/pgt_synth.h"

void pgt_oracle_synth_fst(uint64_t seed, uint64_t site0, uint64_t n, double* a, double* b) {
	for (uint64_t i = 0; i < n; ++i) {
		a[i] = pgt_synth_fst_a(seed, site0 + i);
		b[i] = pgt_synth_fst_b(seed, site0 + i);
	}
}

void pgt_oracle_synth_het(uint64_t seed, uint64_t site0, uint64_t n, int8_t* g) {
	for (uint64_t i = 0; i < n; ++i) g[i] = (int8_t)pgt_synth_het_g(seed, site0 + i);
}

void pgt_oracle_synth_dxy(uint64_t seed, uint64_t site0, uint64_t n, double* f1, double* f2, int32_t* n1, int32_t* n2) {
	for (uint64_t i = 0; i < n; ++i) {
		f1[i] = pgt_synth_dxy_f1(seed, site0 + i);
		f2[i] = pgt_synth_dxy_f2(seed, site0 + i);
		n1[i] = pgt_synth_dxy_n1(seed, site0 + i);
		n2[i] = pgt_synth_dxy_n2(seed, site0 + i);
	}
}

/* positions of sites [site0, site0+n) of one contig whose first site has global index contig_site0 */
void pgt_oracle_synth_score(uint64_t seed, uint64_t site0, uint64_t n, double* score) {
	for (uint64_t i = 0; i < n; ++i) score[i] = pgt_synth_score(seed, site0 + i);
}

void pgt_oracle_synth_pos(uint64_t seed, uint64_t site0, uint64_t n, uint64_t contig_site0, uint32_t density, uint32_t* pos) {
	for (uint64_t i = 0; i < n; ++i) pos[i] = pgt_synth_pos(seed, site0 + i, site0 + i - contig_site0, density);
}

static char* put_micro(char* p, int64_t k) {
	if (k < 0) {
		*p++ = '-';
		k = -k;
	}
	return p + sprintf(p, "%lld.%06lld", (long long)(k / 1000000), (long long)(k % 1000000));
}

/* Appends n sites of contig `name` (global site indices site0.., contig-local index local0..) */
int pgt_oracle_write_fst_text(const char* path, int append, const char* name, uint64_t seed, uint64_t site0,
                              uint64_t local0, uint64_t n, uint32_t density) {
	FILE* f = fopen(path, append ? "a" : "w");
	if (!f) return -1;
	static char big[1 << 20];
	setvbuf(f, big, _IOFBF, sizeof(big));
	char line[256];
	for (uint64_t i = 0; i < n; ++i) {
		char* p = line;
		p += sprintf(p, "%s\t%u\t", name, pgt_synth_pos(seed, site0 + i, local0 + i, density));
		p = put_micro(p, pgt_synth_fst_a_micro(seed, site0 + i));
		*p++ = '\t';
		p = put_micro(p, pgt_synth_fst_b_micro(seed, site0 + i));
		*p++ = '\n';
		fwrite(line, 1, (size_t)(p - line), f);
	}
	fclose(f);
	return 0;
}

int pgt_oracle_write_het_text(const char* path, int append, const char* name, uint64_t seed, uint64_t site0,
                              uint64_t local0, uint64_t n, uint32_t density) {
	FILE* f = fopen(path, append ? "a" : "w");
	if (!f) return -1;
	static char big[1 << 20];
	setvbuf(f, big, _IOFBF, sizeof(big));
	for (uint64_t i = 0; i < n; ++i)
		fprintf(f, "%s\t%u\t%d\n", name, pgt_synth_pos(seed, site0 + i, local0 + i, density), pgt_synth_het_g(seed, site0 + i));
	fclose(f);
	return 0;
}

/* pop = 1 or 2; header written when !append */
int pgt_oracle_write_maf_text(const char* path, int append, const char* name, uint64_t seed, uint64_t site0,
                              uint64_t local0, uint64_t n, uint32_t density, int pop) {
	FILE* f = fopen(path, append ? "a" : "w");
	if (!f) return -1;
	static char big[1 << 20];
	setvbuf(f, big, _IOFBF, sizeof(big));
	if (!append) fputs("chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n", f);
	char line[256];
	for (uint64_t i = 0; i < n; ++i) {
		char* p = line;
		p += sprintf(p, "%s\t%u\tA\tC\tA\t", name, pgt_synth_pos(seed, site0 + i, local0 + i, density));
		p = put_micro(p, pop == 1 ? pgt_synth_dxy_f1_micro(seed, site0 + i) : pgt_synth_dxy_f2_micro(seed, site0 + i));
		p += sprintf(p, "\t%d\n", pop == 1 ? pgt_synth_dxy_n1(seed, site0 + i) : pgt_synth_dxy_n2(seed, site0 + i));
		fwrite(line, 1, (size_t)(p - line), f);
	}
	fclose(f);
	return 0;
}
