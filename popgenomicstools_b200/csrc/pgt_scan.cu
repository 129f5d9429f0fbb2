// pgt_scan.cu -- CUDA (sm_100a) implementation of the windowed site-statistic scan + C ABI.
//
// Replaces calcWindow of the reference (/root/reference/fstWindow.cpp:69-107,
// hetWindow.cpp:66-105, dxyWindow.cpp:172-209) -- a sequential re-sum of a W-entry buffer per
// window followed by a slide-by-copy -- with a one-pass reduction over columnar site arrays:
//
//   level 1  k_units<Stat>   : every site is read from HBM exactly once; the per-site statistic
//                              is evaluated in registers and reduced (lane-strided partial sums +
//                              warp-shuffle butterfly) into one partial per *unit* (pgt_geom.h).
//   level 2  k_windows<Stat> : every window is the sum of its consecutive unit partials (the
//                              carry across overlapping windows: W/S-fold overlap costs re-reads
//                              of small partials from L2, never of sites), plus the epilogue
//                              (ratio, position gather at the two window edges, label lookup).
//   (dxy)    k_global_*      : dxyWindow's global line from the same unit partials (two stages).
//   (bp)     k_bp_bounds     : dxyWindow -fixedsite 0: the reference materialises one buffer entry
//                              per bp (dxyWindow.cpp:365-372); here units live on the bp axis and
//                              each unit's site range is found by binary search in `pos` (sparse,
//                              O(#sites) instead of O(chromosome length)).
//
//   (fine)   k_slide<Stat>   : fine steps under long windows (W = 1000, S = 1): windows formed straight
//                              from the sites in shared memory, no unit array (pgt_slide.cuh).
//
// Device code lives in pgt_kernels_common.cuh (plan, statistics), pgt_level1.cuh, pgt_level2.cuh and
// pgt_slide.cuh, all included below into this one translation unit; this file holds the host side: layout of
// the caller-owned workspace, kernel selection and launches, host-memory mode, the C ABI entry points.
// There is no CPU fallback: every entry point fails with PGT_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/pgt_extreme.h"
#include "pgt_internal.h"

// ----------------------------------------------------------------------------- utilities

static std::atomic<uint64_t> g_launches{0};
extern "C" uint64_t pgt_kernel_launch_count(void) { return g_launches.load(); }
void pgt_count_launch() { g_launches++; }

static int cuda_fail(cudaError_t e, const char* what) {
	return pgt_set_error(PGT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define PGT_CUDA(call)                                         \
	do {                                                       \
		cudaError_t e__ = (call);                              \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call);  \
	} while (0)
#define PGT_TRY(call)                    \
	do {                                 \
		int rc__ = (call);               \
		if (rc__ != PGT_OK) return rc__; \
	} while (0)

extern "C" int pgt_device_count(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
	return n;
}
extern "C" int pgt_set_device(int device) {
	PGT_CUDA(cudaSetDevice(device));
	return PGT_OK;
}
extern "C" int pgt_host_alloc(void** p, size_t bytes) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_host_alloc: NULL");
	PGT_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocPortable));  // page-locked for every device of the process (pgt_scan_sharded)
	return PGT_OK;
}
extern "C" int pgt_host_free(void* p) {
	PGT_CUDA(cudaFreeHost(p));
	return PGT_OK;
}
extern "C" int pgt_device_alloc(void** p, size_t bytes) {
	if (!p) return pgt_set_error(PGT_ERR_ARGS, "pgt_device_alloc: NULL");
	PGT_CUDA(cudaMalloc(p, bytes ? bytes : 1));
	return PGT_OK;
}
extern "C" int pgt_device_free(void* p) {
	PGT_CUDA(cudaFree(p));
	return PGT_OK;
}
// ---- result tables shared between the processes of one box (one process per GPU): rank 0 exports
// its table (pgt_device_alloc memory), the others map it and pass pointers into it as their `out` arrays, so the
// window kernels write each shard's rows straight into rank 0's HBM over NVLink -- no gather step at all.
extern "C" int pgt_ipc_export(const void* devptr, void* handle, size_t handle_bytes) {
	if (!devptr || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return pgt_set_error(PGT_ERR_ARGS, "pgt_ipc_export: need a device pointer and 64 bytes for the handle");
	cudaIpcMemHandle_t h;
	PGT_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(devptr)));
	memcpy(handle, &h, sizeof(h));
	return PGT_OK;
}
extern "C" int pgt_ipc_open(const void* handle, void** devptr) {
	if (!handle || !devptr) return pgt_set_error(PGT_ERR_ARGS, "pgt_ipc_open: NULL");
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	PGT_CUDA(cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess));
	return PGT_OK;
}
extern "C" int pgt_ipc_close(void* devptr) {
	PGT_CUDA(cudaIpcCloseMemHandle(devptr));
	return PGT_OK;
}
extern "C" int pgt_host_register(void* p, size_t bytes) {
	PGT_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
	return PGT_OK;
}
extern "C" int pgt_host_unregister(void* p) {
	PGT_CUDA(cudaHostUnregister(p));
	return PGT_OK;
}

// ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream
// kinds: 0 = level 1 (k_units*), 1 = level 2 (k_windows*), 2 / 3 = level 1 / 2 of the extreme scan
struct ProfEvent {
	cudaEvent_t a, b;
	int kind;
};
static std::atomic<bool> g_profile{false};
static std::mutex g_prof_mutex;  // scans may run from several host threads while profiling is on
static std::vector<ProfEvent> g_prof_events;
static double g_prof_ms[4] = {0, 0, 0, 0};
static uint64_t g_prof_n[4] = {0, 0, 0, 0};

bool pgt_profile_enabled() { return g_profile; }
void pgt_profile_push(int kind, void* ev_a, void* ev_b) {
	std::lock_guard<std::mutex> lk(g_prof_mutex);
	g_prof_events.push_back(ProfEvent{(cudaEvent_t)ev_a, (cudaEvent_t)ev_b, kind});
}

struct ProfScope {
	cudaStream_t st;
	ProfEvent ev;
	bool on;
	ProfScope(int kind, cudaStream_t s) : st(s), on(g_profile) {
		if (!on) return;
		ev.kind = kind;
		if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) {
			on = false;
			return;
		}
		cudaEventRecord(ev.a, st);
	}
	~ProfScope() {
		if (!on) return;
		cudaEventRecord(ev.b, st);
		std::lock_guard<std::mutex> lk(g_prof_mutex);
		g_prof_events.push_back(ev);
	}
};

extern "C" int pgt_profile(int enable) {
	g_profile = enable != 0;
	return PGT_OK;
}

// synchronise all recorded events and fold them into the per-kind totals
static int prof_drain() {
	std::lock_guard<std::mutex> lk(g_prof_mutex);
	for (ProfEvent& e : g_prof_events) {
		float t = 0;
		cudaError_t err = cudaEventSynchronize(e.b);
		if (err == cudaSuccess) err = cudaEventElapsedTime(&t, e.a, e.b);
		cudaEventDestroy(e.a);
		cudaEventDestroy(e.b);
		if (err != cudaSuccess) {
			g_prof_events.clear();
			return cuda_fail(err, "pgt_profile_read");
		}
		g_prof_ms[e.kind & 3] += t;
		g_prof_n[e.kind & 3]++;
	}
	g_prof_events.clear();
	return PGT_OK;
}
static int prof_read(int k0, double* units_ms, uint64_t* units_launches, double* windows_ms, uint64_t* windows_launches) {
	PGT_TRY(prof_drain());
	if (units_ms) *units_ms = g_prof_ms[k0];
	if (units_launches) *units_launches = g_prof_n[k0];
	if (windows_ms) *windows_ms = g_prof_ms[k0 + 1];
	if (windows_launches) *windows_launches = g_prof_n[k0 + 1];
	g_prof_ms[k0] = g_prof_ms[k0 + 1] = 0;
	g_prof_n[k0] = g_prof_n[k0 + 1] = 0;
	return PGT_OK;
}
extern "C" int pgt_profile_read(double* units_ms, uint64_t* units_launches, double* windows_ms, uint64_t* windows_launches) {
	return prof_read(0, units_ms, units_launches, windows_ms, windows_launches);
}
extern "C" int pgt_profile_read_extreme(double* units_ms, uint64_t* units_launches, double* windows_ms, uint64_t* windows_launches) {
	return prof_read(2, units_ms, units_launches, windows_ms, windows_launches);
}

// tuning knobs (tests / experiments, pgt_tune):
//   level1: 0 auto, 1 force the direct kernel (k_units; only valid when pgt_geom.gw == 32), 2 force the tiled kernel
//   level2: 0 auto, 1 force warp-per-window, 2 force scan mode (k_block_scan + k_windows_hgw)
static int g_tune_level1 = 0;
static int g_tune_level2 = 0;
static int g_tune_stages = 2;      // tiled kernel: shared-memory stages (2..4)
static int g_tune_stage_kb = 110;  // tiled kernel: KB per stage (stages * stage_kb <= 224)
extern int g_tune_xgroup;          // extreme scan (pgt_extreme.cu): lanes per unit, 0 = auto
extern int g_tune_xsmall;          // extreme scan: 0 auto, 1 never / 2 always (if the longest window fits) the small-window kernel
//   unittable: 0 auto (unit-start table for plans of more than 32 segments, device mode), 1 never, 2 always
static int g_tune_unittable = 0;
static constexpr size_t kUnitTableMinSegs = 33;  // measured (profiles/r02o_*): better from 1e3 contigs on, 3.8x at 1e6; few-contig genomes keep the closed form
//   fused2: 0 auto (fused sliding tile with one block per step: two crews, k_slide_fused2), 1 the single-crew k_slide<FusedStat>
static int g_tune_fused2 = 0;
//   persite: 0 auto (W = S = 1: four windows per thread with 128-bit accesses when everything is aligned), 1 always the scalar kernel
static int g_tune_persite = 0;
//   slide: 0 auto, 1 never use the sliding-tile kernel (k_slide), 2 use it whenever it fits shared memory (any W <= 1048)
static int g_tune_slide = 0;
// 0: the sliding tile also takes dxyWindow's global line when its runs own exactly the line's sites; 1: always a separate pass
static int g_tune_slideglobal = 0;

static int num_sms() {
	int dev = 0, n = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 0;
	cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
	return n;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Grid of a grid-stride kernel: the blocks that are resident at once (one wave), or fewer when the work is small.
// (A fixed cap of 8 blocks per SM launched the per-site kernel -- 6 resident blocks at 40 registers -- as one full wave
// plus a third of one: ncu "waves per SM 1.33", profiles/r02h_persite_ncu_details.txt.)
template <class K>
static unsigned one_wave_grid(K kern, int threads, uint64_t want_blocks) {
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	const uint64_t cap = (uint64_t)num_sms() * per_sm;
	return (unsigned)(want_blocks < cap ? (want_blocks ? want_blocks : 1) : cap);
}

#include "pgt_kernels_common.cuh"
#include "pgt_level1.cuh"
#include "pgt_level2.cuh"
#include "pgt_slide.cuh"

// ----------------------------------------------------------------------------- host side of a scan

static const uint64_t kSlabSites = 1ull << 22;  // PGT_MEM_HOST: sites (entries) staged per slab
static const uint64_t kSlabSlack = 8192;         // >= 2 units

struct ColDesc {
	const void* ptr;        // host pointer (PGT_MEM_HOST) or device pointer
	size_t elem;            // bytes per site
	size_t offset_in_cols;  // byte offset of the pointer member inside Cols
};

// the columns a statistic reads, in staging order; `c` may be NULL (sizing only)
static int stat_columns(pgt_stat stat, pgt_mode mode, const pgt_columns* c, ColDesc* d) {
	int n = 0;
	auto put = [&](const void* p, size_t e, size_t off) { d[n++] = ColDesc{p, e, off}; };
	if (mode == PGT_MODE_BP) put(c ? c->pos : nullptr, 4, offsetof(Cols, pos));
	if (stat == PGT_STAT_FST || stat == PGT_STAT_FUSED) {
		put(c ? c->a : nullptr, 8, offsetof(Cols, a));
		put(c ? c->b : nullptr, 8, offsetof(Cols, b));
	}
	if (stat == PGT_STAT_DXY || stat == PGT_STAT_FUSED) {
		put(c ? c->f1 : nullptr, 8, offsetof(Cols, f1));
		put(c ? c->f2 : nullptr, 8, offsetof(Cols, f2));
		put(c ? c->n1 : nullptr, 4, offsetof(Cols, n1));
		put(c ? c->n2 : nullptr, 4, offsetof(Cols, n2));
	}
	if (stat == PGT_STAT_HET || stat == PGT_STAT_FUSED) put(c ? c->geno : nullptr, 1, offsetof(Cols, g));
	return n;
}

static size_t acc_bytes(pgt_stat stat) {
	switch (stat) {
		case PGT_STAT_FST: return sizeof(FstStat::Acc);
		case PGT_STAT_HET: return sizeof(HetStat::Acc);
		case PGT_STAT_DXY: return sizeof(DxyStat::Acc);
		default: return sizeof(FusedStat::Acc);
	}
}

struct Layout {
	uint64_t w_lo, w_hi, u_lo, u_hi, g_hi, origin;
	size_t segs_off, off_off, siteoff_off, gpart_off, units_off, pre_off, bounds_off, stage_off, outs_off, tileseg_off, total;
	uint64_t tileseg_cap;   // entries of the tile -> segment table (0 = not used: few segments)
	bool hgw;               // level 2 in scan mode
	bool unit_table;        // very many segments: level 1 reads unit starts from a table (k_unit_starts) instead of the segment records
	bool slide;             // fine steps: windows straight from the sites (k_slide), no unit array
	bool persite;           // W = S = 1: elementwise (k_windows_persite), no unit array
	uint32_t slide_stages;
	size_t slide_smem;
	uint64_t gsite_lo, gsite_hi;  // global sites this scan owns for dxyWindow's global line
	size_t gslab_off;             // host mode without units: one global-line triple per slab
	uint64_t out_slab_rows;       // host mode without units: rows of each of the two slab-sized output tables
	uint64_t gslab_cap;
	uint64_t blk_lo, blk_hi;  // scan blocks covering [u_lo, u_hi)
	size_t stage_col_bytes[8];
	uint64_t slab_sites;
};

// Sliding-tile scan (k_slide): shared-memory shape, and whether the geometry takes that path at all.
// The choice fixes the summation order, so it is a pure function of (W, S, unit, statistic) and the
// tuning knob -- never of the input size, the window range or the shard.
struct SlideShape {
	uint32_t E, wpb, G, wp, nstages, stage_bytes, sf_off, pr_off, pos_off, stage_off, ncol;
	uint32_t col_off[kMaxStageCols], col_cap[kMaxStageCols];
	size_t smem;
};
// (sized with the position column staged, whether or not a scan passes positions: the choice of path must
// not depend on it)
static void slide_footprint(const pgt_geom& g, pgt_stat stat, SlideShape* sh) {
	const SlideTeam tm = slide_team(g.W);
	sh->E = tm.E;
	sh->wpb = tm.wpb;
	sh->G = tm.G;
	sh->wp = (uint32_t)align_up(g.W, 32);
	ColDesc d[8];
	const int nc = stat_columns(stat, PGT_MODE_SITES, nullptr, d);
	uint32_t o = 0;
	for (int c = 0; c <= nc; ++c) {  // column nc = pos (uint32); a stage holds the G blocks of a step
		sh->col_off[c] = o;
		sh->col_cap[c] = (uint32_t)align_up((size_t)tm.G * g.W * (c < nc ? d[c].elem : 4) + 48, 16);
		o += sh->col_cap[c];
	}
	sh->ncol = (uint32_t)nc;
	sh->stage_bytes = (uint32_t)align_up(o, 128);
	const uint32_t accb = (uint32_t)align_up((size_t)tm.G * sh->wp * acc_bytes(stat), 128);
	sh->sf_off = kTileCtlBytes + kSlideWtBytes;
	sh->pr_off = sh->sf_off + 2 * accb;              // SUF: two halves (step parity)
	sh->pos_off = sh->pr_off + accb;
	sh->stage_off = sh->pos_off + 2 * tm.G * sh->wp * 4;  // positions: two halves
}
static bool slide_shape(const pgt_geom& g, pgt_mode mode, pgt_stat stat, SlideShape* sh) {
	if (mode != PGT_MODE_SITES || g_tune_slide == 1) return false;
	if (g.W > (uint32_t)kSlideConsumers * kSlideMaxE) return false;
	if (g_tune_slide != 2) {
		// auto: no piece of a step reaches a sector of doubles (units of < 32 sites) under windows of many
		// units -- where level 2 would otherwise run its block scans over a unit array as large as the input
		if (g.ueff >= 32 || g.wunits <= 32) return false;
	}
	const size_t half = 113u * 1024u, full = 226u * 1024u;
	// One rule for all statistics (the fused scan must equal the three single scans bit for bit): the block
	// of the widest one -- fused, 41 B/site staged twice + two 40-byte accumulators per site -- has to fit,
	// which bounds W at 1048 sites.
	slide_footprint(g, PGT_STAT_FUSED, sh);
	if (sh->stage_off + 2 * (size_t)sh->stage_bytes > full) return false;
	slide_footprint(g, stat, sh);
	// two CTAs per SM when three (or two) stages fit half the shared memory, else one CTA with up to four
	for (uint32_t ns : {3u, 2u}) {
		if (sh->stage_off + (size_t)ns * sh->stage_bytes <= half) {
			sh->nstages = ns;
			sh->smem = sh->stage_off + (size_t)ns * sh->stage_bytes;
			return true;
		}
	}
	for (uint32_t ns : {4u, 3u, 2u}) {
		if (sh->stage_off + (size_t)ns * sh->stage_bytes <= full) {
			sh->nstages = ns;
			sh->smem = sh->stage_off + (size_t)ns * sh->stage_bytes;
			return true;
		}
	}
	return false;
}

static int make_layout(const pgt_plan* plan, const pgt_range* range, pgt_stat stat, pgt_mem mem, Layout* L) {
	if (!plan) return pgt_set_error(PGT_ERR_ARGS, "plan is NULL");
	if ((int)stat < 0 || (int)stat > (int)PGT_STAT_FUSED) return pgt_set_error(PGT_ERR_ARGS, "unknown statistic");
	uint64_t lo = 0, hi = plan->nwin, org = plan->off[0];
	if (range) {
		if (range->w_hi != 0 || range->w_lo != 0) {
			lo = range->w_lo;
			hi = range->w_hi;
		}
		org = range->site_origin;
	}
	if (lo > hi || hi > plan->nwin) return pgt_set_error(PGT_ERR_ARGS, "window range out of bounds");
	memset(L, 0, sizeof(*L));
	L->w_lo = lo;
	L->w_hi = hi;
	L->origin = org;
	uint64_t f, c;
	// units reduced: from the first unit of window w_lo (axis head for w_lo == 0) to the last unit
	// of window w_hi-1 (axis tail for w_hi == nwin); units owned for the global line end where
	// the next shard's first window starts.
	if (hi == lo && plan->nwin != 0) {
		L->u_lo = L->u_hi = L->g_hi = 0;
	} else {
		if (lo == 0) L->u_lo = 0;
		else {
			pgt_plan_window_units(plan, lo, &f, &c);
			L->u_lo = f;
		}
		if (hi == plan->nwin) L->u_hi = L->g_hi = plan->nunits;
		else {
			pgt_plan_window_units(plan, hi - 1, &f, &c);
			L->u_hi = f + c;
			pgt_plan_window_units(plan, hi, &f, &c);
			L->g_hi = f;
		}
	}
	// sites owned for the global line (same boundaries as the units [u_lo, g_hi))
	L->gsite_lo = pgt_plan_unit_start(plan, L->u_lo);
	L->gsite_hi = pgt_plan_unit_start(plan, L->g_hi);
	{
		SlideShape sh;
		L->slide = slide_shape(plan->g, plan->mode, stat, &sh);
		if (L->slide) {
			L->slide_stages = sh.nstages;
			L->slide_smem = sh.smem;
		}
		L->persite = !L->slide && plan->mode == PGT_MODE_SITES && plan->g.W == 1 && plan->g.S == 1 && g_tune_level2 != 1;
	}
	const bool nounits = L->slide || L->persite;  // windows come straight from the sites: no unit array, no level 2
	const uint64_t nunits = nounits ? 0 : L->u_hi - L->u_lo;
	size_t o = 0;
	L->segs_off = o;
	o += align_up(plan->segs.size() * sizeof(pgt_seg) + 8, 256);
	L->off_off = o;
	o += align_up(plan->off.size() * sizeof(uint64_t), 256);
	L->siteoff_off = o;
	if (plan->mode == PGT_MODE_BP) o += align_up(plan->off.size() * sizeof(uint64_t), 256);
	L->gpart_off = o;
	o += align_up((size_t)kGlobalMaxPartials * 3 * sizeof(double), 256);
	L->units_off = o;
	o += align_up((size_t)nunits * acc_bytes(stat) + 8, 256);
	// level 2 in scan mode when summing wunits partials per window would dominate (fine steps, long windows)
	{
		// costs in partial reads/writes; the totals are exact for the whole plan (genomes of many short
		// contigs have windows far shorter than wunits and at least one scan block per segment) and
		// scaled to the window range of a shard
		const double frac = plan->nwin ? (double)(hi - lo) / (double)plan->nwin : 0.0;
		const double direct = (double)plan->window_units_total * frac;
		const double scan = 4.0 * (3.0 * (double)plan->scan_units_total * frac + 2.0 * (double)(hi - lo));
		L->hgw = plan->g.wunits > 32 && direct > scan;
		if (g_tune_level2 == 1) L->hgw = false;
		if (g_tune_level2 == 2 && plan->g.wunits >= 2) L->hgw = true;
	}
	if (nounits) L->hgw = false;
	L->pre_off = o;
	L->blk_lo = L->blk_hi = 0;
	// PRE array: reserved whenever scan mode is possible, so the workspace size does not depend on the
	// (tunable) kernel choice
	if (plan->g.wunits >= 2 && nunits) o += align_up((size_t)nunits * acc_bytes(stat) + 8, 256);
	if (L->hgw && nunits) {
		const pgt_seg& s0 = plan->segs[pgt_plan_seg_of_unit(plan, L->u_lo)];
		const pgt_seg& s1 = plan->segs[pgt_plan_seg_of_unit(plan, L->u_hi - 1)];
		L->blk_lo = s0.blk_base + (L->u_lo - s0.unit_base) / plan->g.wunits;
		L->blk_hi = s1.blk_base + (L->u_hi - 1 - s1.unit_base) / plan->g.wunits + 1;
	}
	L->bounds_off = o;
	L->unit_table = plan->mode == PGT_MODE_SITES && mem == PGT_MEM_DEVICE && nunits > 0 &&
	                (g_tune_unittable == 2 || (g_tune_unittable == 0 && plan->segs.size() >= kUnitTableMinSegs));
	if (plan->mode == PGT_MODE_BP || L->unit_table) o += align_up((size_t)(nunits + 1) * sizeof(uint64_t), 256);
	L->stage_off = o;
	L->outs_off = o;
	if (mem == PGT_MEM_HOST) {
		L->slab_sites = std::min<uint64_t>(kSlabSites, std::max<uint64_t>(plan->nsites, 1)) + kSlabSlack;
		ColDesc d[8];
		const int nc = stat_columns(stat, plan->mode, nullptr, d);
		for (int i = 0; i < nc; ++i) L->stage_col_bytes[i] = align_up((size_t)L->slab_sites * d[i].elem, 256);
		for (int s = 0; s < 2; ++s)
			for (int i = 0; i < nc; ++i) o += L->stage_col_bytes[i];
		L->outs_off = o;
		// staged per-window outputs + global[3]: the whole table, or (no unit array: the windows are produced
		// slab by slab) two tables of at most one slab's windows -- a window starts at its own site
		L->out_slab_rows = std::min<uint64_t>(hi - lo, L->slab_sites);
		if (nounits) o += 2 * 12 * align_up((size_t)L->out_slab_rows * 8 + 8, 256);
		else o += 12 * align_up((size_t)(hi - lo) * 8 + 8, 256);
		L->gslab_off = o;
		L->gslab_cap = nounits ? plan->nsites / (L->slab_sites - kSlabSlack - std::min<uint64_t>(plan->g.W, L->slab_sites - kSlabSlack - 1)) + 16 : 0;
		o += align_up((size_t)L->gslab_cap * 3 * sizeof(double), 256);
	}
	// tile -> first segment table of the tiled level-1 kernel (genomes of many contigs, see k_tile_segs)
	L->tileseg_off = o;
	L->tileseg_cap = plan->segs.size() > kTileSegTableMin ? nunits / 8 + 1024 : 0;
	o += align_up((size_t)L->tileseg_cap * sizeof(uint32_t), 256);
	L->total = o + 256;
	return PGT_OK;
}

extern "C" int pgt_plan_scan_path(const pgt_plan* plan, pgt_stat stat) {
	if (!plan) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_scan_path: plan is NULL");
	if ((int)stat < 0 || (int)stat > (int)PGT_STAT_FUSED) return pgt_set_error(PGT_ERR_ARGS, "unknown statistic");
	SlideShape sh;
	if (slide_shape(plan->g, plan->mode, stat, &sh)) return PGT_PATH_SLIDE;
	if (plan->mode == PGT_MODE_SITES && plan->g.W == 1 && plan->g.S == 1 && g_tune_level2 != 1) return PGT_PATH_PERSITE;
	return PGT_PATH_UNITS;
}

extern "C" size_t pgt_plan_device_bytes(const pgt_plan* plan) {
	if (!plan) return 0;
	return align_up(plan->segs.size() * sizeof(pgt_seg), 256) + align_up(plan->off.size() * sizeof(uint64_t), 256) + 256;
}

extern "C" int pgt_plan_bind_device(pgt_plan* plan, void* buffer, size_t bytes, void* stream) {
	if (!plan) return pgt_set_error(PGT_ERR_ARGS, "pgt_plan_bind_device: plan is NULL");
	if (!buffer) {
		plan->d_tables = nullptr;
		return PGT_OK;
	}
	if (bytes < pgt_plan_device_bytes(plan)) return pgt_set_error(PGT_ERR_NOMEM, "pgt_plan_bind_device: buffer too small");
	cudaStream_t st = (cudaStream_t)stream;
	unsigned char* b = (unsigned char*)(((uintptr_t)buffer + 255) / 256 * 256);
	const size_t seg_bytes = plan->segs.size() * sizeof(pgt_seg);
	if (seg_bytes) PGT_CUDA(cudaMemcpyAsync(b, plan->segs.data(), seg_bytes, cudaMemcpyHostToDevice, st));
	PGT_CUDA(cudaMemcpyAsync(b + align_up(seg_bytes, 256), plan->off.data(), plan->off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
	PGT_CUDA(cudaStreamSynchronize(st));
	plan->d_tables = b;
	return PGT_OK;
}

extern "C" size_t pgt_scan_workspace_bytes(const pgt_plan* plan, const pgt_range* range, pgt_stat stat, pgt_mem mem) {
	Layout L;
	if (make_layout(plan, range, stat, mem, &L) != PGT_OK) return 0;
	return L.total;
}

extern "C" int pgt_tune(const char* key, int value) {
	if (key && strcmp(key, "level1") == 0) g_tune_level1 = value;
	else if (key && strcmp(key, "level2") == 0 && value >= 0 && value <= 2) g_tune_level2 = value;
	else if (key && strcmp(key, "slide") == 0 && value >= 0 && value <= 2) g_tune_slide = value;
	else if (key && strcmp(key, "fused2") == 0 && value >= 0 && value <= 1) g_tune_fused2 = value;
	else if (key && strcmp(key, "slideglobal") == 0 && value >= 0 && value <= 1) g_tune_slideglobal = value;
	else if (key && strcmp(key, "persite") == 0 && value >= 0 && value <= 1) g_tune_persite = value;
	else if (key && strcmp(key, "unittable") == 0 && value >= 0 && value <= 2) g_tune_unittable = value;
	else if (key && strcmp(key, "stages") == 0 && value >= 2 && value <= kTileMaxStages) g_tune_stages = value;
	else if (key && strcmp(key, "stage_kb") == 0 && value >= 8 && value <= 110) g_tune_stage_kb = value;
	else if (key && strcmp(key, "xsmall") == 0 && value >= 0 && value <= 2) g_tune_xsmall = value;
	else if (key && strcmp(key, "xgroup") == 0 && (value == 0 || value == 4 || value == 8 || value == 16 || value == 32)) g_tune_xgroup = value;
	else return pgt_set_error(PGT_ERR_ARGS, "pgt_tune: unknown key");
	return PGT_OK;
}

struct StatCols {
	const void* ptr[kMaxTileCols];
	uint32_t elem[kMaxTileCols];
	uint32_t n;
};
template <class Stat>
static StatCols tile_columns(const Cols& c);
template <>
StatCols tile_columns<FstStat>(const Cols& c) { return StatCols{{c.a, c.b}, {8, 8}, 2}; }
template <>
StatCols tile_columns<HetStat>(const Cols& c) { return StatCols{{c.g}, {1}, 1}; }
template <>
StatCols tile_columns<DxyStat>(const Cols& c) { return StatCols{{c.f1, c.f2, c.n1, c.n2}, {8, 8, 4, 4}, 4}; }
template <>
StatCols tile_columns<FusedStat>(const Cols& c) { return StatCols{{c.a, c.b, c.f1, c.f2, c.n1, c.n2, c.g}, {8, 8, 8, 8, 4, 4, 1}, 7}; }

template <class Stat, bool INDIRECT>
static int launch_units_tiled(const DevPlan& P, const Cols& cols, typename Stat::Acc* units, const uint64_t* bounds, uint64_t valid_elems,
                              uint32_t* tileseg, uint64_t tileseg_cap, cudaStream_t st) {
	const uint64_t nunits = P.unit_hi - P.unit_lo;
	const StatCols sc = tile_columns<Stat>(cols);
	uint32_t bps = 0;
	for (uint32_t c = 0; c < sc.n; ++c) bps += sc.elem[c];
	const int nsm = num_sms();
	// `nstages` stages of <= `stage_kb` KB; tiles of whole units; at least ~4 tiles per SM when the input allows
	uint32_t nstages = (uint32_t)g_tune_stages;
	uint32_t stage_budget = (uint32_t)g_tune_stage_kb * 1024u;
	if (nstages * stage_budget > 224u * 1024u) stage_budget = 224u * 1024u / nstages;
	uint32_t tsites = (stage_budget - 48u * sc.n) / bps;
	if (tsites > 16384u && bps > 1) tsites = 16384u;  // (the 1-byte genotype column alone fills a stage with ~110 K sites)
	uint32_t m = tsites / P.g.ueff;
	if (m < 1) m = 1;
	{
		// A tile's units are dealt to NG consumer groups; a tile of m units takes ceil(m / NG) rounds.
		// When the last round would be mostly idle (measured: dxy, m = 18 over 15 warps, 6.29 TB/s;
		// m = 15, 6.80 TB/s) a smaller tile of whole rounds wins; a nearly full last round (fst,
		// m = 27 of 30) is better left alone: the larger tile keeps more bytes in flight.
		const uint32_t ng = (uint32_t)kTileConsumerWarps * (32u / P.g.gw);
		if (m > ng && m % ng != 0 && (double)m / (double)((m + ng - 1) / ng * ng) < 0.8) m = m / ng * ng;
	}
	const uint32_t cap_sites = m * P.g.ueff;  // sites a stage can hold (what the shared-memory regions are sized for)
	if (INDIRECT && nunits) {
		// bp mode: a unit is a bp range and holds anything from 0 to ueff sites.  Sizing tiles for the
		// dense worst case makes them nearly empty on sparse data (SNP-only MAFs: ~3 % full, and the
		// per-tile cost dominated: 5.6 ms for 1e8 sites over 3e9 bp).  Size them for the average
		// density instead; a tile whose slice does not fit is read straight from global memory by
		// the consumers (`staged` in the kernel), so dense stretches stay correct.
		const double avg = (double)valid_elems / (double)nunits;  // sites per unit
		const double want = 0.75 * (double)cap_sites / (avg > 1e-3 ? avg : 1e-3);
		const uint32_t m_sparse = want > 65536.0 ? 65536u : (uint32_t)want;
		if (m_sparse > m) m = m_sparse;
	}
	const uint64_t want_tiles = (uint64_t)nsm * 4;
	if ((nunits + m - 1) / m < want_tiles) {
		uint64_t mm = (nunits + want_tiles - 1) / want_tiles;
		m = (uint32_t)(mm < 1 ? 1 : mm);
	}
	TileCfg tc;
	memset(&tc, 0, sizeof(tc));
	tc.ncol = sc.n;
	tc.m = m;
	tc.nstages = nstages;
	tc.minind = cols.minind;
	tc.valid_elems = valid_elems;
	uint32_t o = 0;
	for (uint32_t c = 0; c < sc.n; ++c) {
		tc.gcol[c] = (const char*)sc.ptr[c];
		tc.elem[c] = sc.elem[c];
		tc.col_off[c] = o;
		tc.col_cap[c] = (uint32_t)align_up((size_t)cap_sites * sc.elem[c] + 48, 16);
		o += tc.col_cap[c];
	}
	tc.stage_bytes = (uint32_t)align_up(o, 128);
	const size_t smem = kTileCtlBytes + nstages * (size_t)tc.stage_bytes;
	void (*kern)(DevPlan, TileCfg, typename Stat::Acc*, const uint64_t*);
	switch (P.g.gw) {
		case 1: kern = k_units_tiled<Stat, 1, INDIRECT>; break;
		case 2: kern = k_units_tiled<Stat, 2, INDIRECT>; break;
		case 4: kern = k_units_tiled<Stat, 4, INDIRECT>; break;
		case 8: kern = k_units_tiled<Stat, 8, INDIRECT>; break;
		case 16: kern = k_units_tiled<Stat, 16, INDIRECT>; break;
		default: kern = k_units_tiled<Stat, 32, INDIRECT>; break;
	}
	PGT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const uint64_t ntiles = (nunits + m - 1) / m;
	const unsigned grid = (unsigned)(ntiles < (uint64_t)nsm ? ntiles : (uint64_t)nsm);
	{
		ProfScope prof(0, st);
		if (!INDIRECT && tileseg && ntiles <= tileseg_cap && ntiles > grid) {  // many segments: tile -> segment table first
			k_tile_segs<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(P, m, ntiles, tileseg);
			g_launches++;
			tc.tile_seg = tileseg;
		}
		kern<<<grid, kTileThreads, smem, st>>>(P, tc, units, bounds);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// hetWindow: the vectorised byte kernel replaces the generic direct kernel
template <class Stat>
static bool launch_het_vec(const DevPlan&, const Cols&, typename Stat::Acc*, const uint64_t*, uint64_t, cudaStream_t) {
	return false;
}
template <>
bool launch_het_vec<HetStat>(const DevPlan& P, const Cols& cols, HetStat::Acc* units, const uint64_t* bounds, uint64_t want, cudaStream_t st) {
	const uint64_t cap = (uint64_t)num_sms() * 8;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	ProfScope prof(0, st);
	if (bounds) k_units_het_vec<true><<<grid, 256, 0, st>>>(P, cols.g, units, bounds);
	else k_units_het_vec<false><<<grid, 256, 0, st>>>(P, cols.g, units, bounds);
	return true;
}

// valid_elems: number of elements every column is known to hold from element 0
template <class Stat>
static int launch_units(const DevPlan& P, const Cols& cols, typename Stat::Acc* units, const uint64_t* bounds, uint64_t valid_elems,
                        uint32_t* tileseg, uint64_t tileseg_cap, cudaStream_t st) {
	const uint64_t nunits = P.unit_hi - P.unit_lo;
	if (nunits == 0) return PGT_OK;
	// Kernel choice.  The direct kernel reduces every unit with a full warp, so it implements the
	// summation order only when gw == 32 (long units); short units always take the tiled kernel.
	// For long units both kernels give bit-identical results and the choice is purely a speed
	// matter (measured on B200, profiles/): the TMA-staged tiled kernel wins on big inputs
	// (no sector over-fetch, bytes in flight without registers), the direct kernel on small ones
	// (no pipeline fill) and for the 1-byte genotype column.
	uint32_t bps = 0;
	{
		const StatCols sc = tile_columns<Stat>(cols);
		for (uint32_t c = 0; c < sc.n; ++c) bps += sc.elem[c];
	}
	const uint64_t approx_bytes = nunits * (uint64_t)P.g.ueff * bps;
	// (hetWindow's 1-byte column stays on the vectorised direct kernel: measured at 3e9 sites, 4096-site units,
	// 6.14 TB/s direct against 5.44 TB/s through the ring -- both were issue-bound on the byte counting, 4.5 and
	// 4.3 TB/s, until it went from POPC to dp4a; the ring's het consumer remains for forced / short-unit runs)
	bool tiled = P.g.gw != 32 || (bps > 1 && approx_bytes >= (256ull << 20));
	if (g_tune_level1 == 2) tiled = true;
	if (g_tune_level1 == 1 && P.g.gw == 32) tiled = false;
	{
		// a tile holds whole units: very long units (unit_sites up to 4096) may not fit a stage;
		// they always have gw == 32, so the direct kernel applies
		uint32_t nst = (uint32_t)g_tune_stages, budget = (uint32_t)g_tune_stage_kb * 1024u;
		if (nst * budget > 224u * 1024u) budget = 224u * 1024u / nst;
		if ((uint64_t)P.g.ueff * bps + 64ull * kMaxTileCols > budget) tiled = false;
	}
	if (tiled)
		return bounds ? launch_units_tiled<Stat, true>(P, cols, units, bounds, valid_elems, nullptr, 0, st)
		              : launch_units_tiled<Stat, false>(P, cols, units, bounds, valid_elems, tileseg, tileseg_cap, st);
	const int threads = 256;
	const uint64_t want = (nunits + 7) / 8;  // one warp per unit, 8 warps per block
	if (launch_het_vec<Stat>(P, cols, units, bounds, want, st)) {
		g_launches++;
		PGT_CUDA(cudaGetLastError());
		return PGT_OK;
	}
	const uint32_t upl = P.g.u / 32u;
	void (*kern)(DevPlan, Cols, typename Stat::Acc*, const uint64_t*);
	if (bounds) kern = upl == 8 ? k_units<Stat, 8, true> : k_units<Stat, 0, true>;
	else kern = upl == 8 ? k_units<Stat, 8, false> : (upl == 4 ? k_units<Stat, 4, false> : k_units<Stat, 0, false>);
	int per_sm = 0;
	PGT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
	const uint64_t cap = (uint64_t)num_sms() * (per_sm > 0 ? per_sm : 1);
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	{
		ProfScope prof(0, st);
		kern<<<grid, threads, 0, st>>>(P, cols, units, bounds);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

static int launch_bounds_kernel(const DevPlan& P, const uint32_t* pos, uint64_t ndata, uint64_t* bounds, cudaStream_t st) {
	const uint64_t nb = P.unit_hi - P.unit_lo + 1;
	const uint64_t want = (nb + 255) / 256;
	const uint64_t cap = (uint64_t)num_sms() * 8;
	k_bp_bounds<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, pos, ndata, bounds);
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// scan mode (L.hgw): `units` is scanned in place into SUF, `pre` receives PRE; see k_block_scan
template <class Stat>
static int launch_windows(const DevPlan& P, typename Stat::Acc* units, uint64_t units_base, const uint32_t* pos, const pgt_windows& out,
                          bool hgw, typename Stat::Acc* pre, uint64_t blk_lo, uint64_t blk_hi, cudaStream_t st) {
	const uint64_t nwin = P.win_hi - P.win_lo;
	if (nwin == 0) return PGT_OK;
	const uint64_t cap = (uint64_t)num_sms() * 8;
	ProfScope prof(1, st);
	if (hgw) {
		const uint64_t nblk = blk_hi - blk_lo;
		const uint64_t bcap = (uint64_t)num_sms() * 8;
		k_block_scan<Stat><<<(unsigned)(nblk < bcap ? nblk : bcap), 256, 0, st>>>(P, units, pre, units_base, blk_lo, blk_hi);
		g_launches++;
		const uint64_t want = (nwin + 255) / 256;
		k_windows_hgw<Stat><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, units, pre, units_base, pos, out);
	} else if (P.g.wunits <= 32 && g_tune_level2 != 1) {
		// fine windows: a thread per window
		void (*kern)(DevPlan, const typename Stat::Acc*, uint64_t, const uint32_t*, pgt_windows);
		const uint32_t wu = P.g.wunits;
		if (wu <= 1) kern = k_windows_small<Stat, 1>;
		else if (wu <= 2) kern = k_windows_small<Stat, 2>;
		else if (wu <= 4) kern = k_windows_small<Stat, 4>;
		else if (wu <= 8) kern = k_windows_small<Stat, 8>;
		else if (wu <= 16) kern = k_windows_small<Stat, 16>;
		else kern = k_windows_small<Stat, 32>;
		const uint64_t want = (nwin + 255) / 256;
		kern<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, units, units_base, pos, out);
	} else {
		const uint64_t want = (nwin + 7) / 8;  // few windows: a warp each; many: runs per block, stored 32 rows at a time
		k_windows<Stat><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, units, units_base, pos, out);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

template <class Stat>
static int launch_global(const typename Stat::Acc*, uint64_t, double*, double*, cudaStream_t) {
	return PGT_OK;
}
template <>
int launch_global<DxyStat>(const DxyStat::Acc* units, uint64_t n, double* scratch, double* g3, cudaStream_t st) {
	k_global_partial<DxyStat><<<kGlobalBlocks, 1024, 0, st>>>(units, n, scratch);
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, kGlobalBlocks, g3);
	g_launches += 2;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}
template <>
int launch_global<FusedStat>(const FusedStat::Acc* units, uint64_t n, double* scratch, double* g3, cudaStream_t st) {
	k_global_partial<FusedStat><<<kGlobalBlocks, 1024, 0, st>>>(units, n, scratch);
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, kGlobalBlocks, g3);
	g_launches += 2;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// Sliding-tile scan of the windows [P.win_lo, P.win_hi): the columns' element 0 is site P.site_origin and
// they hold `valid_elems` elements.  `pos` may be NULL (host mode gathers positions on the host).
template <class Stat>
static int launch_slide(const pgt_plan* plan, pgt_stat stat, const DevPlan& P, const Cols& cols, uint64_t valid_elems, const uint32_t* pos,
                        const pgt_windows& out, double* gpart, double* g3, cudaStream_t st) {
	const uint64_t nwin = P.win_hi - P.win_lo;
	if (nwin == 0) return PGT_OK;
	SlideShape sh;
	if (!slide_shape(plan->g, plan->mode, stat, &sh)) return pgt_set_error(PGT_ERR_ARGS, "internal: sliding-tile scan not applicable");
	const StatCols scol = tile_columns<Stat>(cols);
	TileCfg tc;
	memset(&tc, 0, sizeof(tc));
	tc.ncol = scol.n + (pos ? 1u : 0u);  // positions ride the ring too: read from HBM once, gathered from shared memory
	tc.nstages = sh.nstages;
	tc.stage_bytes = sh.stage_bytes;
	tc.minind = cols.minind;
	tc.valid_elems = valid_elems;
	for (uint32_t c = 0; c < tc.ncol; ++c) {
		tc.gcol[c] = c < scol.n ? (const char*)scol.ptr[c] : (const char*)pos;
		tc.elem[c] = c < scol.n ? scol.elem[c] : 4u;
		tc.col_off[c] = sh.col_off[c];
		tc.col_cap[c] = sh.col_cap[c];
	}
	SlideCfg sc;
	memset(&sc, 0, sizeof(sc));
	sc.E = sh.E;
	sc.wpb = sh.wpb;
	sc.G = sh.G;
	sc.wp = sh.wp;
	sc.sf_off = sh.sf_off;
	sc.pr_off = sh.pr_off;
	sc.pos_off = sh.pos_off;
	sc.pos_col = pos ? scol.n : 0xffffffffu;
	sc.stage_off = sh.stage_off;
	sc.gpart = gpart;
	auto finish_global = [&](unsigned grid) -> int {  // one partial triple per CTA -> the line
		if (!gpart) return PGT_OK;
		k_global_final<<<1, kGlobalBlocks, 0, st>>>(gpart, grid, g3);
		g_launches++;
		PGT_CUDA(cudaGetLastError());
		return PGT_OK;
	};
	// chunks: ~4 per resident CTA for balance, but long enough (>= 32 steps of windows) that the W - S sites
	// shared with the next chunk stay a few percent of what a chunk reads
	const uint64_t wps = ((uint64_t)sh.G * plan->g.W + plan->g.S - 1) / plan->g.S;  // windows starting in one step
	auto set_chunks = [&](uint64_t slots) {
		sc.chunk_windows = std::max<uint64_t>(32 * wps, (nwin + slots * 4 - 1) / (slots * 4));
		sc.nchunks = (nwin + sc.chunk_windows - 1) / sc.chunk_windows;
	};
	if constexpr (std::is_same<Stat, FusedStat>::value) {
		if (sh.G == 1 && g_tune_fused2 != 1) {
			// two crews (fst + het, dxy) share the staged block; their SUF / PRE arrays partition the fused ones
			SlideCfg2 s2;
			const uint32_t part[2] = {24u, 16u};  // accumulator bytes per site: FstHetStat, DxyStat
			uint32_t so = sh.sf_off, po = sh.pr_off;
			for (int c = 0; c < 2; ++c) {
				s2.wt_off[c] = kTileCtlBytes + 256u * c;
				s2.sf_off[c] = so;
				s2.pr_off[c] = po;
				so += 2u * sh.wp * part[c];
				po += sh.wp * part[c];
			}
			void (*k2)(DevPlan, TileCfg, SlideCfg, SlideCfg2, pgt_windows) =
			    sh.E <= 4 ? k_slide_fused2<4> : (sh.E == 5 ? k_slide_fused2<5> : k_slide_fused2<6>);
			PGT_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
			const uint64_t slots = (uint64_t)num_sms();
			set_chunks(slots);
			const unsigned grid2 = (unsigned)std::min<uint64_t>(sc.nchunks, slots);
			{
				ProfScope prof(0, st);
				k2<<<grid2, kSlide2Threads, sh.smem, st>>>(P, tc, sc, s2, out);
			}
			g_launches++;
			PGT_CUDA(cudaGetLastError());
			return finish_global(grid2);
		}
	}
	void (*kern)(DevPlan, TileCfg, SlideCfg, pgt_windows);
	if (sh.G > 1) {
		if (sh.E <= 2) kern = k_slide<Stat, 2, true>;
		else if (sh.E <= 4) kern = k_slide<Stat, 4, true>;
		else kern = k_slide<Stat, 6, true>;
	} else {
		if (sh.E <= 4) kern = k_slide<Stat, 4, false>;
		else if (sh.E == 5) kern = k_slide<Stat, 5, false>;  // W = 1000
		else kern = k_slide<Stat, 6, false>;
	}
	PGT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
	int per_sm = 0;
	PGT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSlideThreads, sh.smem));
	if (per_sm < 1) per_sm = 1;
	const uint64_t slots = (uint64_t)num_sms() * per_sm;
	set_chunks(slots);
	const unsigned grid = (unsigned)std::min<uint64_t>(sc.nchunks, slots);
	{
		ProfScope prof(0, st);
		kern<<<grid, kSlideThreads, sh.smem, st>>>(P, tc, sc, out);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return finish_global(grid);
}

// Do the sites the sliding-tile kernel counts as owned by its runs -- per segment from the first window's start to the
// start of the first window after the range, or to the end of the segment's last window -- make up exactly the range
// [gsite_lo, gsite_hi) of dxyWindow's global line?  (No when a segment inside the range has sites after its last window
// or no window at all: then the line is taken by a pass over the columns.)
static bool slide_runs_cover_owned_sites(const pgt_plan* plan, const Layout& L) {
	if (L.w_hi <= L.w_lo) return false;
	const uint64_t W = plan->g.W, S = plan->g.S;
	uint32_t si = pgt_plan_seg_of_window(plan, L.w_lo);
	uint64_t w = L.w_lo, cursor = L.gsite_lo;
	while (w < L.w_hi) {
		while (si < plan->segs.size() && w - plan->segs[si].win_base >= plan->segs[si].nwin) ++si;
		if (si >= plan->segs.size()) return false;
		const pgt_seg& sg = plan->segs[si];
		const uint64_t ka = w - sg.win_base, kb = std::min<uint64_t>(L.w_hi - sg.win_base, sg.nwin);
		if (sg.site_base + ka * S != cursor) return false;
		const uint64_t hi = kb == sg.nwin ? std::min<uint64_t>((kb - 1) * S + W, sg.nsites) : kb * S;
		cursor = sg.site_base + hi;
		w = sg.win_base + kb;
	}
	return cursor == L.gsite_hi;
}

// W = S = 1: the vectorised kernel when its preconditions hold (one segment; every used column and output pointer
// 16-byte aligned at the scan's first window -- 4-byte aligned for the genotype column), else the scalar one.
template <class Stat>
static int launch_persite(const pgt_plan* plan, pgt_stat stat, const DevPlan& P, const Cols& C, const pgt_windows& out, double* gpart,
                          double* g3, cudaStream_t st) {
	const uint64_t nwin = P.win_hi - P.win_lo;
	if (nwin == 0) return PGT_OK;
	bool vec = g_tune_persite != 1 && plan->segs.size() == 1 && nwin >= 4;
	uint64_t site0 = 0;
	if (vec) {
		site0 = plan->segs[0].site_base + (P.win_lo - plan->segs[0].win_base) - P.site_origin;
		auto ok = [&](const void* p, size_t elem, size_t align) { return !p || (((uintptr_t)p + site0 * elem) & (align - 1)) == 0; };
		ColDesc d[8];
		pgt_columns pc;
		memset(&pc, 0, sizeof(pc));
		pc.a = C.a;
		pc.b = C.b;
		pc.geno = C.g;
		pc.f1 = C.f1;
		pc.f2 = C.f2;
		pc.n1 = C.n1;
		pc.n2 = C.n2;
		const int nc = stat_columns(stat, PGT_MODE_SITES, &pc, d);
		for (int i = 0; i < nc; ++i) vec = vec && d[i].ptr && ok(d[i].ptr, d[i].elem, d[i].elem == 1 ? 4 : 16);
		vec = vec && ok(C.pos, 4, 16);
		const void* outs[14] = {out.label, out.start_pos, out.end_pos, out.mid_pos, out.nsites, out.sum_a, out.sum_b, out.fst,
		                        out.nhet, out.nonmissing, out.het, out.dxy, out.neffective, out.nskip};
		for (const void* p : outs) vec = vec && (((uintptr_t)p) & 15) == 0;
	}
	ProfScope prof(1, st);
	unsigned grid;
	if (vec) {
		grid = std::min<unsigned>(one_wave_grid(k_windows_persite4<Stat>, 256, (nwin + 1023) / 1024), kGlobalMaxPartials);
		k_windows_persite4<Stat><<<grid, 256, 0, st>>>(P, C, out, site0, gpart);
	} else {
		grid = std::min<unsigned>(one_wave_grid(k_windows_persite<Stat>, 256, (nwin + 1023) / 1024), kGlobalMaxPartials);  // four windows per thread and turn
		k_windows_persite<Stat><<<grid, 256, 0, st>>>(P, C, out, gpart);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	if (gpart) {  // the global line over exactly these windows' sites: one partial triple per block
		k_global_final<<<1, kGlobalBlocks, 0, st>>>(gpart, grid, g3);
		g_launches++;
		PGT_CUDA(cudaGetLastError());
	}
	return PGT_OK;
}

// dxyWindow's global line straight from the columns (scans without a unit array)
static int launch_global_sites(const Cols& cols, uint64_t i0, uint64_t i1, double* scratch, double* g3, cudaStream_t st) {
	k_global_sites<<<kGlobalBlocks, 1024, 0, st>>>(cols, i0, i1, scratch);
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, kGlobalBlocks, g3);
	g_launches += 2;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// host-side lower bound of bp entry e on the site axis (bp mode, PGT_MEM_HOST)
static uint64_t host_site_lower_bound(const pgt_plan* plan, const uint64_t* site_off, const uint32_t* pos, uint64_t origin,
                                      uint64_t ndata, uint64_t e) {
	const uint32_t nc = (uint32_t)plan->off.size() - 1;
	uint64_t res;
	if (e >= plan->off[nc]) {
		res = site_off[nc];
	} else {
		const uint32_t c = pgt_plan_contig_of(plan, e);
		const uint32_t p = (uint32_t)(e - plan->off[c]) + 1u;
		uint64_t lo = std::max(site_off[c], origin), hi = std::min(site_off[c + 1], origin + ndata);
		if (hi < lo) hi = lo;
		const uint32_t* b = pos + (lo - origin);
		res = lo + (uint64_t)(std::lower_bound(b, pos + (hi - origin), p) - b);
	}
	return std::min(std::max(res, origin), origin + ndata);
}

struct HostStreams {
	cudaStream_t copy = nullptr;
	cudaEvent_t ready[2] = {nullptr, nullptr};
	cudaEvent_t freed[2] = {nullptr, nullptr};
	~HostStreams() {
		for (int i = 0; i < 2; ++i) {
			if (ready[i]) cudaEventDestroy(ready[i]);
			if (freed[i]) cudaEventDestroy(freed[i]);
		}
		if (copy) cudaStreamDestroy(copy);
	}
};

template <class Stat>
static int run_scan(const pgt_plan* plan, const pgt_range* range, pgt_stat stat, const pgt_columns* cols, int minind,
                    const uint64_t* site_offsets, const pgt_windows* out, void* workspace, size_t workspace_bytes, pgt_mem mem,
                    cudaStream_t st) {
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return pgt_set_error(PGT_ERR_CUDA, "no usable CUDA device (libpgtscan has no CPU fallback)");
	Layout L;
	PGT_TRY(make_layout(plan, range, stat, mem, &L));
	if (!workspace) return pgt_set_error(PGT_ERR_ARGS, "workspace is NULL");
	if (workspace_bytes < L.total) return pgt_set_error(PGT_ERR_NOMEM, "workspace too small, see pgt_scan_workspace_bytes");
	const bool bp = plan->mode == PGT_MODE_BP;
	if (bp && stat != PGT_STAT_DXY) return pgt_set_error(PGT_ERR_ARGS, "PGT_MODE_BP plans are for PGT_STAT_DXY only");
	if (bp && (!site_offsets || !cols->pos)) return pgt_set_error(PGT_ERR_ARGS, "bp mode needs site_offsets and pos");
	const uint64_t nunits = L.u_hi - L.u_lo;
	const uint64_t nwin = L.w_hi - L.w_lo;
	if (nunits == 0 && nwin == 0) return PGT_OK;  // empty input / empty shard: nothing to read or write
	ColDesc cd[8];
	const int ncol = stat_columns(stat, plan->mode, cols, cd);
	for (int i = 0; i < ncol; ++i)
		if (!cd[i].ptr) return pgt_set_error(PGT_ERR_ARGS, "a column required by this statistic is NULL");
	if (!bp && nunits > 0 && pgt_plan_unit_start(plan, L.u_lo) < L.origin)
		return pgt_set_error(PGT_ERR_ARGS, "site_origin lies after the first site this scan must read");

	char* ws = (char*)workspace;
	const size_t seg_bytes = plan->segs.size() * sizeof(pgt_seg);
	const size_t off_bytes = plan->off.size() * sizeof(uint64_t);
	const bool resident = plan->d_tables != nullptr;  // pgt_plan_bind_device: nothing to upload (and graph-capturable)
	if (!resident) {
		if (seg_bytes) PGT_CUDA(cudaMemcpyAsync(ws + L.segs_off, plan->segs.data(), seg_bytes, cudaMemcpyHostToDevice, st));
		PGT_CUDA(cudaMemcpyAsync(ws + L.off_off, plan->off.data(), off_bytes, cudaMemcpyHostToDevice, st));
	}
	if (bp) PGT_CUDA(cudaMemcpyAsync(ws + L.siteoff_off, site_offsets, off_bytes, cudaMemcpyHostToDevice, st));

	DevPlan P;
	P.g = plan->g;
	P.segs = resident ? (const pgt_seg*)plan->d_tables : (const pgt_seg*)(ws + L.segs_off);
	P.off = resident ? (const uint64_t*)(plan->d_tables + align_up(seg_bytes, 256)) : (const uint64_t*)(ws + L.off_off);
	P.site_off = bp ? (const uint64_t*)(ws + L.siteoff_off) : nullptr;
	P.nseg = (uint32_t)plan->segs.size();
	P.ncontig = (uint32_t)plan->off.size() - 1;
	P.unit_lo = L.u_lo;
	P.unit_hi = L.u_hi;
	P.win_lo = L.w_lo;
	P.win_hi = L.w_hi;
	P.site_origin = L.origin;
	P.nunits_total = plan->nunits;
	P.col_elems = ~0ull;
	P.mode = (int)plan->mode;

	typename Stat::Acc* units = (typename Stat::Acc*)(ws + L.units_off);
	uint64_t* bounds = (bp || L.unit_table) ? (uint64_t*)(ws + L.bounds_off) : nullptr;
	uint64_t ndata = 0;  // bp mode: sites the columns hold
	if (bp) {
		if (site_offsets[P.ncontig] < L.origin) return pgt_set_error(PGT_ERR_ARGS, "site_origin beyond the last site");
		ndata = (range && range->site_count) ? range->site_count : site_offsets[P.ncontig] - L.origin;
	}
	const bool want_global = out->dxy_global && (stat == PGT_STAT_DXY || stat == PGT_STAT_FUSED);
	const uint64_t nglobal = L.g_hi > L.u_lo ? L.g_hi - L.u_lo : 0;

	if (mem == PGT_MEM_DEVICE) {
		Cols C;
		memset(&C, 0, sizeof(C));
		C.pos = cols->pos;
		C.a = cols->a;
		C.b = cols->b;
		C.g = cols->geno;
		C.f1 = cols->f1;
		C.f2 = cols->f2;
		C.n1 = cols->n1;
		C.n2 = cols->n2;
		C.minind = minind;
		if (L.persite || L.slide) {
			// windows straight from the sites: W = S = 1 (default arguments of fstWindow / hetWindow: an
			// elementwise map) or the sliding tile for fine steps; no unit array, no level 2
			if (nwin) {
				uint64_t last_site;
				pgt_plan_window(plan, L.w_hi - 1, nullptr, &last_site, nullptr);
				P.col_elems = last_site + 1 - L.origin;
			}
			bool global_done = false;
			if (nwin && L.persite) {
				// the windows ARE the sites: when the owned range of the global line is exactly their span (always, unless the
				// scan owns sites outside any window) the kernel adds its sites up on the way and the dxy columns are read once
				uint64_t first_site, last_site;
				pgt_plan_window(plan, L.w_lo, &first_site, nullptr, nullptr);
				pgt_plan_window(plan, L.w_hi - 1, nullptr, &last_site, nullptr);
				global_done = want_global && GlobalTerm<Stat>::has && L.gsite_lo == first_site && L.gsite_hi == last_site + 1;
				PGT_TRY(launch_persite<Stat>(plan, stat, P, C, *out, global_done ? (double*)(ws + L.gpart_off) : nullptr, out->dxy_global, st));
			} else if (nwin) {
				uint64_t last;
				pgt_plan_window(plan, L.w_hi - 1, nullptr, &last, nullptr);
				global_done = want_global && GlobalTerm<Stat>::has && g_tune_slideglobal != 1 && slide_runs_cover_owned_sites(plan, L);
				PGT_TRY(launch_slide<Stat>(plan, stat, P, C, last + 1 - L.origin, cols->pos, *out,
				                           global_done ? (double*)(ws + L.gpart_off) : nullptr, out->dxy_global, st));
			}
			if (want_global && !global_done) {
				if (L.gsite_lo < L.origin) return pgt_set_error(PGT_ERR_ARGS, "site_origin lies after the first site this scan must read");
				PGT_TRY(launch_global_sites(C, L.gsite_lo - L.origin, L.gsite_hi - L.origin, (double*)(ws + L.gpart_off), out->dxy_global, st));
			}
			return PGT_OK;
		}
		if (bp && nunits) PGT_TRY(launch_bounds_kernel(P, cols->pos, ndata, bounds, st));
		if (L.unit_table) {  // very many segments: unit starts once, then the INDIRECT level-1 kernels
			const uint64_t want = (nunits + 256) / 256, cap = (uint64_t)num_sms() * 8;
			k_unit_starts<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, bounds);
			g_launches++;
			PGT_CUDA(cudaGetLastError());
		}
		// elements the caller's columns are known to hold: up to the end of the last unit read
		const uint64_t valid = bp ? ndata : pgt_plan_unit_start(plan, L.u_hi) - L.origin;
		P.col_elems = valid;
		PGT_TRY(launch_units<Stat>(P, C, units, bounds, valid, L.tileseg_cap ? (uint32_t*)(ws + L.tileseg_off) : nullptr, L.tileseg_cap, st));
		// the global line reads the raw unit partials: before level 2, which may scan them in place
		if (want_global) PGT_TRY(launch_global<Stat>(units, nglobal, (double*)(ws + L.gpart_off), out->dxy_global, st));
		PGT_TRY(launch_windows<Stat>(P, units, L.u_lo, cols->pos, *out, L.hgw, (typename Stat::Acc*)(ws + L.pre_off), L.blk_lo, L.blk_hi, st));
		return PGT_OK;
	}

	// ---- PGT_MEM_HOST: slabs of the axis are staged H2D on a copy stream (double-buffered) while
	// the previous slab is reduced; only unit partials stay on the device.  Positions are not
	// copied in site mode: the two edge positions per window are gathered on the host.
	HostStreams hs;
	PGT_CUDA(cudaStreamCreateWithFlags(&hs.copy, cudaStreamNonBlocking));
	for (int i = 0; i < 2; ++i) {
		PGT_CUDA(cudaEventCreateWithFlags(&hs.ready[i], cudaEventDisableTiming));
		PGT_CUDA(cudaEventCreateWithFlags(&hs.freed[i], cudaEventDisableTiming));
	}
	// the copy stream must not start before the caller's stream reaches this point
	PGT_CUDA(cudaEventRecord(hs.freed[0], st));
	PGT_CUDA(cudaStreamWaitEvent(hs.copy, hs.freed[0], 0));
	// Window bookkeeping on the host -- label, nsites, the two edge positions per window -- runs on
	// worker threads from here on, next to the copy pipeline.  (It cannot simply follow the slab
	// loop: enqueueing ~10 operations per slab blocks once the driver's queue is full, so the host
	// would only get to it when most of the transfer is over; 3e7 windows take ~1 s on one thread.)
	struct Bookkeeper {
		std::vector<std::thread> th;
		void join() {
			for (auto& t : th)
				if (t.joinable()) t.join();
		}
		~Bookkeeper() { join(); }
	} book;
	if (nwin && (out->label || out->nsites || out->start_pos || out->end_pos || out->mid_pos)) {
		unsigned nt = nwin < (1u << 16) ? 1u : std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
		const uint64_t per = (nwin + nt - 1) / nt;
		const uint64_t origin = L.origin;
		for (unsigned t = 0; t < nt; ++t) {
			const uint64_t wa = L.w_lo + per * t, wb = std::min<uint64_t>(L.w_hi, wa + per);
			if (wa >= wb) break;
			book.th.emplace_back([plan, out, cols, bp, wa, wb, origin, w_lo = L.w_lo]() {
				uint32_t si = pgt_plan_seg_of_window(plan, wa);
				for (uint64_t w = wa; w < wb; ++w) {
					while (w >= plan->segs[si].win_base + plan->segs[si].nwin) ++si;
					const pgt_seg& sg = plan->segs[si];
					uint64_t fs;
					const uint32_t n = pgt_window_sites(plan->g, sg, w - sg.win_base, &fs);
					const uint64_t first = sg.site_base + fs, last = first + n - 1;
					const uint32_t label = pgt_plan_contig_of(plan, last);
					const uint64_t o = w - w_lo;
					if (out->label) out->label[o] = label;
					if (out->nsites) out->nsites[o] = n;
					uint32_t sp = 0, ep = 0;
					bool have = false;
					if (bp) {
						sp = (uint32_t)(first - plan->off[pgt_plan_contig_of(plan, first)]) + 1u;
						ep = (uint32_t)(last - plan->off[label]) + 1u;
						have = true;
					} else if (cols->pos) {
						sp = cols->pos[first - origin];
						ep = cols->pos[last - origin];
						have = true;
					}
					if (have) {
						if (out->start_pos) out->start_pos[o] = sp;
						if (out->end_pos) out->end_pos[o] = ep;
						if (out->mid_pos) out->mid_pos[o] = (sp + ep) / 2u;
					}
				}
			});
		}
	}
	char* stage[2][8];
	{
		size_t o = L.stage_off;
		for (int s = 0; s < 2; ++s)
			for (int i = 0; i < ncol; ++i) {
				stage[s][i] = ws + o;
				o += L.stage_col_bytes[i];
			}
	}
	// per-window values are produced into device staging, then copied back: one table for the whole scan, or --
	// when the windows come straight from the sites, slab by slab -- two slab-sized tables used in turn
	const bool from_units = !(L.slide || L.persite);
	const uint64_t out_rows = from_units ? nwin : L.out_slab_rows;
	const size_t ob = align_up((size_t)out_rows * 8 + 8, 256);
	auto device_table = [&](int which) -> pgt_windows {
		pgt_windows d;
		memset(&d, 0, sizeof(d));
		char* ob0 = ws + L.outs_off + (size_t)which * 12 * ob;
		int k = 0;
		auto dptr = [&](const void* want) -> void* {
			void* r = want ? (void*)(ob0 + ob * k) : nullptr;
			++k;
			return r;
		};
		d.sum_a = (double*)dptr(out->sum_a);
		d.sum_b = (double*)dptr(out->sum_b);
		d.fst = (double*)dptr(out->fst);
		d.nhet = (uint32_t*)dptr(out->nhet);
		d.nonmissing = (uint32_t*)dptr(out->nonmissing);
		d.het = (double*)dptr(out->het);
		d.dxy = (double*)dptr(out->dxy);
		d.neffective = (uint32_t*)dptr(out->neffective);
		d.nskip = (uint32_t*)dptr(out->nskip);
		d.dxy_global = want_global ? (double*)(ob0 + ob * 11) : nullptr;
		return d;
	};
	// rows [row0, row0 + n) of the caller's arrays <- rows [0, n) of a device table, on `st`
	auto copy_back = [&](const pgt_windows& d, uint64_t row0, uint64_t n) -> int {
		auto back = [&](void* h, const void* dp, size_t elem) -> cudaError_t {
			return (h && dp && n) ? cudaMemcpyAsync((char*)h + row0 * elem, dp, n * elem, cudaMemcpyDeviceToHost, st) : cudaSuccess;
		};
		PGT_CUDA(back(out->sum_a, d.sum_a, 8));
		PGT_CUDA(back(out->sum_b, d.sum_b, 8));
		PGT_CUDA(back(out->fst, d.fst, 8));
		PGT_CUDA(back(out->nhet, d.nhet, 4));
		PGT_CUDA(back(out->nonmissing, d.nonmissing, 4));
		PGT_CUDA(back(out->het, d.het, 8));
		PGT_CUDA(back(out->dxy, d.dxy, 8));
		PGT_CUDA(back(out->neffective, d.neffective, 4));
		PGT_CUDA(back(out->nskip, d.nskip, 4));
		return PGT_OK;
	};
	const pgt_windows dev = device_table(0);
	const uint64_t slab_target = L.slab_sites - kSlabSlack;
	const uint64_t axis_end = plan->off[0] + plan->nsites;
	int slot = 0;
	bool slot_used[2] = {false, false};
	// one slab: H2D of sites [s0, s0 + ns) of every column into `slot`, ordered after the kernels that last read it
	auto stage_slab = [&](uint64_t s0, uint64_t ns, Cols* C) -> int {
		if (slot_used[slot]) PGT_CUDA(cudaStreamWaitEvent(hs.copy, hs.freed[slot], 0));
		memset(C, 0, sizeof(*C));
		C->minind = minind;
		for (int i = 0; i < ncol; ++i) {
			const char* src = (const char*)cd[i].ptr + (s0 - L.origin) * cd[i].elem;
			if (ns) PGT_CUDA(cudaMemcpyAsync(stage[slot][i], src, ns * cd[i].elem, cudaMemcpyHostToDevice, hs.copy));
			const void* p = stage[slot][i];
			memcpy((char*)C + cd[i].offset_in_cols, &p, sizeof(p));
		}
		PGT_CUDA(cudaEventRecord(hs.ready[slot], hs.copy));
		PGT_CUDA(cudaStreamWaitEvent(st, hs.ready[slot], 0));
		return PGT_OK;
	};
	auto release_slab = [&]() -> int {
		PGT_CUDA(cudaEventRecord(hs.freed[slot], st));
		slot_used[slot] = true;
		slot ^= 1;
		return PGT_OK;
	};
	if (L.slide || L.persite) {
		// ---- windows straight from the sites (sliding tile / W = S = 1): slabs of consecutive WINDOWS; a
		// slab holds the sites of its windows, neighbours share the W - S sites of the overlap.  The global
		// line is taken from the same staged sites: every slab folds the part of the owned range
		// [gsite_lo, gsite_hi) that no earlier slab has folded, one partial triple per slab, added in slab
		// order on the host.
		double* gslab = (double*)(ws + L.gslab_off);
		uint64_t ngslab = 0, gcursor = L.gsite_lo;
		auto fold_global = [&](const Cols& C, uint64_t s0, uint64_t s1) -> int {  // owned sites inside the staged [s0, s1)
			const uint64_t a = std::max(gcursor, s0), b = std::min(L.gsite_hi, s1);
			if (!want_global || b <= a) return PGT_OK;
			if (ngslab >= L.gslab_cap) return pgt_set_error(PGT_ERR_NOMEM, "internal: global-line slab table too small");
			PGT_TRY(launch_global_sites(C, a - s0, b - s0, (double*)(ws + L.gpart_off), gslab + 3 * ngslab, st));
			++ngslab;
			gcursor = b;
			return PGT_OK;
		};
		uint64_t w = L.w_lo;
		while (w < L.w_hi) {
			uint64_t fs, last;
			pgt_plan_window(plan, w, &fs, nullptr, nullptr);
			const uint64_t wb = pgt_plan_windows_within(plan, w, std::min<uint64_t>(L.w_hi, w + L.out_slab_rows), fs + slab_target);
			pgt_plan_window(plan, wb - 1, nullptr, &last, nullptr);
			// (in site mode no owned site precedes the scan's first window: only an EOF segment can lack windows)
			const uint64_t s0 = fs;
			const uint64_t ns = last + 1 - s0;
			if (ns > L.slab_sites) return pgt_set_error(PGT_ERR_ARGS, "internal: window slab larger than the staging buffers");
			Cols C;
			PGT_TRY(stage_slab(s0, ns, &C));
			DevPlan Ps = P;
			Ps.win_lo = w;
			Ps.win_hi = wb;
			Ps.site_origin = s0;
			Ps.col_elems = ns;
			const pgt_windows o = device_table(slot);  // rewritten two slabs later, after this slab's copy-back (order of `st`)
			if (L.persite) {
				PGT_TRY(launch_persite<Stat>(plan, stat, Ps, C, o, nullptr, nullptr, st));  // (global line: fold_global below)
			} else {
				PGT_TRY(launch_slide<Stat>(plan, stat, Ps, C, ns, nullptr, o, nullptr, nullptr, st));  // (global line: fold_global below)
			}
			PGT_TRY(fold_global(C, s0, s0 + ns));
			PGT_TRY(copy_back(o, w - L.w_lo, wb - w));
			PGT_TRY(release_slab());
			w = wb;
		}
		// owned sites after the last window (dropped EOF partial) or of a scan without windows
		while (want_global && gcursor < L.gsite_hi) {
			const uint64_t ns = std::min<uint64_t>(L.gsite_hi - gcursor, slab_target);
			const uint64_t s0 = gcursor;
			Cols C;
			PGT_TRY(stage_slab(s0, ns, &C));
			PGT_TRY(fold_global(C, s0, s0 + ns));
			PGT_TRY(release_slab());
		}
		if (want_global) {
			std::vector<double> part(3 * ngslab + 3, 0.0);
			if (ngslab) PGT_CUDA(cudaMemcpyAsync(part.data(), gslab, 3 * ngslab * sizeof(double), cudaMemcpyDeviceToHost, st));
			PGT_CUDA(cudaStreamSynchronize(st));
			double g3[3] = {0.0, 0.0, 0.0};
			for (uint64_t i = 0; i < ngslab; ++i)
				for (int q = 0; q < 3; ++q) g3[q] += part[3 * i + q];
			memcpy(out->dxy_global, g3, sizeof(g3));
		}
	}
	uint64_t ua = (L.slide || L.persite) ? L.u_hi : L.u_lo;
	while (ua < L.u_hi) {
		const uint64_t e0 = pgt_plan_unit_start(plan, ua);
		uint64_t ub;
		if (e0 + slab_target >= axis_end) ub = L.u_hi;
		else ub = std::min<uint64_t>(L.u_hi, pgt_plan_unit_containing(plan, e0 + slab_target));
		if (ub <= ua) ub = ua + 1;
		const uint64_t e1 = pgt_plan_unit_start(plan, ub);
		uint64_t s0 = e0, s1 = e1;  // data-site range of this slab
		if (bp) {
			s0 = host_site_lower_bound(plan, site_offsets, cols->pos, L.origin, ndata, e0);
			s1 = host_site_lower_bound(plan, site_offsets, cols->pos, L.origin, ndata, e1);
		}
		const uint64_t ns = s1 - s0;
		if (ns > L.slab_sites) return pgt_set_error(PGT_ERR_INPUT, "bp mode: more sites than bp in a slab (positions not strictly increasing?)");
		Cols C;
		PGT_TRY(stage_slab(s0, ns, &C));
		DevPlan Ps = P;
		Ps.unit_lo = ua;
		Ps.unit_hi = ub;
		Ps.site_origin = s0;
		Ps.col_elems = ns;
		uint64_t* sb = bounds ? bounds + (ua - L.u_lo) : nullptr;
		// bp: bounds are relative to the staged slice.  Slab i writes entries [ua, ub]; the shared
		// entry ub is rewritten by slab i+1 relative to ITS slice, after slab i's unit kernel
		// (stream order on `st`).
		if (bp) PGT_TRY(launch_bounds_kernel(Ps, C.pos, ns, sb, st));
		PGT_TRY(launch_units<Stat>(Ps, C, units + (ua - L.u_lo), sb, ns, L.tileseg_cap ? (uint32_t*)(ws + L.tileseg_off) : nullptr,
		                           L.tileseg_cap, st));
		PGT_TRY(release_slab());
		ua = ub;
	}

	if (from_units) {
		if (want_global) PGT_TRY(launch_global<Stat>(units, nglobal, (double*)(ws + L.gpart_off), dev.dxy_global, st));
		PGT_TRY(launch_windows<Stat>(P, units, L.u_lo, nullptr, dev, L.hgw, (typename Stat::Acc*)(ws + L.pre_off), L.blk_lo, L.blk_hi, st));
		PGT_TRY(copy_back(dev, 0, nwin));
		if (want_global) PGT_CUDA(cudaMemcpyAsync(out->dxy_global, dev.dxy_global, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
	}

	book.join();
	PGT_CUDA(cudaStreamSynchronize(st));
	PGT_CUDA(cudaStreamSynchronize(hs.copy));
	return PGT_OK;
}

extern "C" int pgt_scan(const pgt_plan* plan, const pgt_range* range, pgt_stat stat, const pgt_columns* cols, int minind,
                        const uint64_t* site_offsets, const pgt_windows* out, void* workspace, size_t workspace_bytes, pgt_mem mem,
                        void* stream) {
	if (!plan || !cols || !out) return pgt_set_error(PGT_ERR_ARGS, "pgt_scan: plan, cols and out must not be NULL");
	if (mem != PGT_MEM_DEVICE && mem != PGT_MEM_HOST) return pgt_set_error(PGT_ERR_ARGS, "pgt_scan: unknown memory space");
	cudaStream_t st = (cudaStream_t)stream;
	switch (stat) {
		case PGT_STAT_FST: return run_scan<FstStat>(plan, range, stat, cols, minind, site_offsets, out, workspace, workspace_bytes, mem, st);
		case PGT_STAT_HET: return run_scan<HetStat>(plan, range, stat, cols, minind, site_offsets, out, workspace, workspace_bytes, mem, st);
		case PGT_STAT_DXY:
			if (minind < 1) return pgt_set_error(PGT_ERR_ARGS, "-minind must be at least 1");
			return run_scan<DxyStat>(plan, range, stat, cols, minind, site_offsets, out, workspace, workspace_bytes, mem, st);
		case PGT_STAT_FUSED:
			if (minind < 1) return pgt_set_error(PGT_ERR_ARGS, "-minind must be at least 1");
			return run_scan<FusedStat>(plan, range, stat, cols, minind, site_offsets, out, workspace, workspace_bytes, mem, st);
	}
	return pgt_set_error(PGT_ERR_ARGS, "pgt_scan: unknown statistic");
}

extern "C" int pgt_scan_fst(const pgt_plan* plan, const pgt_range* range, const pgt_columns* cols, const pgt_windows* out,
                            void* workspace, size_t workspace_bytes, pgt_mem mem, void* stream) {
	return pgt_scan(plan, range, PGT_STAT_FST, cols, 1, nullptr, out, workspace, workspace_bytes, mem, stream);
}
extern "C" int pgt_scan_het(const pgt_plan* plan, const pgt_range* range, const pgt_columns* cols, const pgt_windows* out,
                            void* workspace, size_t workspace_bytes, pgt_mem mem, void* stream) {
	return pgt_scan(plan, range, PGT_STAT_HET, cols, 1, nullptr, out, workspace, workspace_bytes, mem, stream);
}
extern "C" int pgt_scan_dxy(const pgt_plan* plan, const pgt_range* range, const pgt_columns* cols, int minind,
                            const uint64_t* site_offsets, const pgt_windows* out, void* workspace, size_t workspace_bytes, pgt_mem mem,
                            void* stream) {
	return pgt_scan(plan, range, PGT_STAT_DXY, cols, minind, site_offsets, out, workspace, workspace_bytes, mem, stream);
}
extern "C" int pgt_scan_fused(const pgt_plan* plan, const pgt_range* range, const pgt_columns* cols, int minind,
                              const pgt_windows* out, void* workspace, size_t workspace_bytes, pgt_mem mem, void* stream) {
	return pgt_scan(plan, range, PGT_STAT_FUSED, cols, minind, nullptr, out, workspace, workspace_bytes, mem, stream);
}
