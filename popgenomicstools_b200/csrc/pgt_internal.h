// pgt_internal.h -- shared between the host planner (pgt_plan.cpp) and the CUDA side (pgt_scan.cu).
#ifndef PGT_INTERNAL_H
#define PGT_INTERNAL_H

#include <string>
#include <vector>

#include "../../include/pgt_scan.h"
#include "pgt_geom.h"

struct pgt_plan {
	pgt_mode mode;
	pgt_geom g;
	std::vector<uint64_t> off;  // contig offsets, ncontig+1
	std::vector<pgt_seg> segs;
	uint64_t nwin;
	uint64_t nunits;
	uint64_t nsites;
	uint64_t nblocks;  // scan blocks of g.wunits units over all segments
	uint64_t window_units_total;  // unit partials a direct level 2 reads over all windows (short segments have short windows)
	uint64_t scan_units_total;    // units the block scans of scan mode touch (whole blocks, at least one per segment)
	// optional caller-owned device copy of segs | off (pgt_plan_bind_device); nullptr = upload per scan
	unsigned char* d_tables = nullptr;
};

int pgt_set_error(int code, const std::string& msg);

// shared by the .cu files: launch counter and the optional per-kernel event timing (pgt_scan.cu)
void pgt_count_launch();
bool pgt_profile_enabled();
void pgt_profile_push(int kind, void* ev_a, void* ev_b);  // cudaEvent_t pair, recorded by the caller

// segment index containing global window w / global unit j / global site x
uint32_t pgt_plan_seg_of_window(const pgt_plan* p, uint64_t w);
uint32_t pgt_plan_seg_of_unit(const pgt_plan* p, uint64_t j);
// global index of the unit containing global site/entry x (x inside the axis)
uint64_t pgt_plan_unit_containing(const pgt_plan* p, uint64_t x);
// global site/entry index where global unit j starts; j == nunits gives the end of the axis
uint64_t pgt_plan_unit_start(const pgt_plan* p, uint64_t j);
// largest wb in (w, w_hi] such that all windows [w, wb) end at or before global site `limit` (>= w + 1)
uint64_t pgt_plan_windows_within(const pgt_plan* p, uint64_t w, uint64_t w_hi, uint64_t limit);
// contig c with off[c] <= x < off[c+1]
uint32_t pgt_plan_contig_of(const pgt_plan* p, uint64_t x);

#endif
