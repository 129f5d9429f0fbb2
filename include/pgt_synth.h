/*
 * pgt_synth.h -- counter-based synthetic site generator.
 *
 * Every value is a pure function of (seed, column, site index), computed with
 * integer arithmetic plus ONE correctly rounded IEEE division (k / 1e6), so the
 * CUDA generator kernels (popgenomicstools_b200/csrc/pgt_synth.cu), the CPU
 * twin used by the oracle/tests, and the 6-decimal text the reference binaries
 * parse (libstdc++ num_get -> strtod, /root/reference/fstWindow.cpp:141) all
 * yield bit-identical doubles.  3e9-site inputs are generated in HBM and never
 * touch disk.
 *
 * Columns mirror the reference's input formats:
 *   fst : chr pos a b                      (/root/reference/fstWindow.cpp:17-21,141)
 *   het : chr pos genotype                 (/root/reference/hetWindow.cpp:18,139)
 *   dxy : chromo position .. freq nInd x2  (/root/reference/dxyWindow.cpp:24-32,146-152)
 *   score : normalised iHS / XP-EHH column   (/root/reference/ihsWindow.cpp:160, xpehhWindow.cpp:165)
 */
#ifndef PGT_SYNTH_H
#define PGT_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PGT_HD __host__ __device__ __forceinline__
#else
#define PGT_HD static inline
#endif

enum {
	PGT_COL_FST_A = 1,
	PGT_COL_FST_B = 2,
	PGT_COL_HET_G = 3,
	PGT_COL_DXY_F = 4,
	PGT_COL_DXY_N = 5,
	PGT_COL_POS = 6,
	PGT_COL_SCORE = 7
};

PGT_HD uint64_t pgt_splitmix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

PGT_HD uint64_t pgt_site_hash(uint64_t seed, uint32_t column, uint64_t site) {
	return pgt_splitmix64(pgt_splitmix64(seed ^ ((uint64_t)column << 56)) ^ (site * 0xD1342543DE82EF95ull));
}

/* Irwin-Hall(4) of 16-bit uniforms, centred: range [-131070, 131070], sd ~ 37837. */
PGT_HD int64_t pgt_ih4(uint64_t h) {
	int64_t s = (int64_t)(h & 0xFFFF) + (int64_t)((h >> 16) & 0xFFFF) + (int64_t)((h >> 32) & 0xFFFF) + (int64_t)(h >> 48);
	return s - 131070;
}

/* integer micro-units (value * 1e6) ------------------------------------------------ */

/* a ~ N(0.01, 0.02), signed (ANGSD numerators can be negative) */
PGT_HD int64_t pgt_synth_fst_a_micro(uint64_t seed, uint64_t site) {
	return 10000 + pgt_ih4(pgt_site_hash(seed, PGT_COL_FST_A, site)) * 20000 / 37837;
}
/* b = |N(0.1, 0.03)| + 1e-3 */
PGT_HD int64_t pgt_synth_fst_b_micro(uint64_t seed, uint64_t site) {
	int64_t k = 100000 + pgt_ih4(pgt_site_hash(seed, PGT_COL_FST_B, site)) * 30000 / 37837;
	return (k < 0 ? -k : k) + 1000;
}
PGT_HD double pgt_micro_to_double(int64_t k) {
	return (double)k / 1000000.0;
}
PGT_HD double pgt_synth_fst_a(uint64_t seed, uint64_t site) { return pgt_micro_to_double(pgt_synth_fst_a_micro(seed, site)); }
PGT_HD double pgt_synth_fst_b(uint64_t seed, uint64_t site) { return pgt_micro_to_double(pgt_synth_fst_b_micro(seed, site)); }

/* genotype in {-1,0,1,2} with P = (.05,.60,.25,.10) */
PGT_HD int pgt_synth_het_g(uint64_t seed, uint64_t site) {
	uint32_t t = (uint32_t)(pgt_site_hash(seed, PGT_COL_HET_G, site) % 100u);
	return t < 5 ? -1 : (t < 65 ? 0 : (t < 90 ? 1 : 2));
}

/* f1: cube-law skew towards 0 on [0,1); f2 = clip(f1 + N(0,0.1), 0, 1) */
PGT_HD int64_t pgt_synth_dxy_f1_micro(uint64_t seed, uint64_t site) {
	uint64_t h = pgt_site_hash(seed, PGT_COL_DXY_F, site);
	uint64_t v = h & 0xFFFFF;
	uint64_t c = (((v * v) >> 20) * v) >> 20;
	return (int64_t)((c * 1000000ull) >> 20);
}
PGT_HD int64_t pgt_synth_dxy_f2_micro(uint64_t seed, uint64_t site) {
	uint64_t h = pgt_site_hash(seed, PGT_COL_DXY_F, site);
	int64_t k = pgt_synth_dxy_f1_micro(seed, site) + pgt_ih4(pgt_splitmix64(h)) * 100000 / 37837;
	return k < 0 ? 0 : (k > 1000000 ? 1000000 : k);
}
PGT_HD double pgt_synth_dxy_f1(uint64_t seed, uint64_t site) { return pgt_micro_to_double(pgt_synth_dxy_f1_micro(seed, site)); }
PGT_HD double pgt_synth_dxy_f2(uint64_t seed, uint64_t site) { return pgt_micro_to_double(pgt_synth_dxy_f2_micro(seed, site)); }
/* nInd ~ U{0..20} per population */
PGT_HD int pgt_synth_dxy_n1(uint64_t seed, uint64_t site) { return (int)(pgt_site_hash(seed, PGT_COL_DXY_N, site) % 21u); }
PGT_HD int pgt_synth_dxy_n2(uint64_t seed, uint64_t site) { return (int)((pgt_site_hash(seed, PGT_COL_DXY_N, site) >> 32) % 21u); }

/* normalised selection score ~ N(0,1) at 1e-6 resolution (ties across a window stay possible: the
 * reference keeps the first extreme, ihsWindow.cpp:166) */
PGT_HD int64_t pgt_synth_score_micro(uint64_t seed, uint64_t site) {
	uint64_t h = pgt_site_hash(seed, PGT_COL_SCORE, site);
	return pgt_ih4(h) * 26 + (int64_t)(pgt_splitmix64(h) % 27u) - 13;
}
PGT_HD double pgt_synth_score(uint64_t seed, uint64_t site) { return pgt_micro_to_double(pgt_synth_score_micro(seed, site)); }

/* position of the i-th (0-based) site of a contig.
 * density 1 : every bp is a site (pos = i+1);
 * density d>1: one site per d bp at a hashed offset (strictly increasing). */
PGT_HD uint32_t pgt_synth_pos(uint64_t seed, uint64_t site, uint64_t local_index, uint32_t density) {
	if (density <= 1) return (uint32_t)(local_index + 1);
	return (uint32_t)(local_index * density + 1 + pgt_site_hash(seed, PGT_COL_POS, site) % density);
}

#endif /* PGT_SYNTH_H */
