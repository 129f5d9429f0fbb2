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
};

int pgt_set_error(int code, const std::string& msg);

// segment index containing global window w / global unit j / global site x
uint32_t pgt_plan_seg_of_window(const pgt_plan* p, uint64_t w);
uint32_t pgt_plan_seg_of_unit(const pgt_plan* p, uint64_t j);

#endif
