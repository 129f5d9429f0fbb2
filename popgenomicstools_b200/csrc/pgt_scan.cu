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
//                              of 16-byte partials from L2, never of sites), plus the epilogue
//                              (ratio, position gather at the two window edges, label lookup).
//
// There is no CPU fallback: every entry point fails with PGT_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pgt_synth.h"
#include "pgt_internal.h"

// ----------------------------------------------------------------------------- utilities

static std::atomic<uint64_t> g_launches{0};
extern "C" uint64_t pgt_kernel_launch_count(void) { return g_launches.load(); }
void pgt_count_launch() { g_launches++; }

static int cuda_fail(cudaError_t e, const char* what) {
	return pgt_set_error(PGT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define PGT_CUDA(call)                                   \
	do {                                                 \
		cudaError_t e__ = (call);                        \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
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
	PGT_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
	return PGT_OK;
}
extern "C" int pgt_host_free(void* p) {
	PGT_CUDA(cudaFreeHost(p));
	return PGT_OK;
}

static int g_num_sms = 0;
static int num_sms() {
	if (g_num_sms == 0) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess) return 0;
		cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
	}
	return g_num_sms;
}

// ----------------------------------------------------------------------------- device plan

struct DevPlan {
	pgt_geom g;
	const pgt_seg* segs;
	const uint64_t* off;  // contig offsets
	uint32_t nseg;
	uint32_t ncontig;
	uint64_t unit_lo, unit_hi;  // global unit range of this scan
	uint64_t win_lo, win_hi;    // global window range of this scan
	uint64_t site_origin;       // global index of element 0 of the columns
	int mode;
};

// last segment with key <= x, key = unit_base (BY_UNIT) or win_base
template <bool BY_UNIT>
__device__ __forceinline__ uint32_t find_seg(const DevPlan& P, uint64_t x) {
	uint32_t lo = 0, hi = P.nseg;
	while (hi - lo > 1) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		uint64_t key = BY_UNIT ? P.segs[mid].unit_base : P.segs[mid].win_base;
		if (key <= x) lo = mid;
		else hi = mid;
	}
	return lo;
}

// contig c in [c0, c0+nc) with off[c] <= x < off[c+1]
__device__ __forceinline__ uint32_t find_contig(const uint64_t* off, uint32_t c0, uint32_t nc, uint64_t x) {
	uint32_t lo = c0, hi = c0 + nc;
	while (hi - lo > 1) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		if (off[mid] <= x) lo = mid;
		else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ----------------------------------------------------------------------------- statistics
//
// A Stat describes: the input columns (In), the unit partial (Acc, stored as-is in the unit
// array), how one site is folded into a lane's partial, and the butterfly combine.

struct FstStat {
	struct In {
		const double* a;
		const double* b;
	};
	struct Acc {
		double a, b;
	};
	struct Site {
		double a, b;
	};
	static __device__ __forceinline__ Acc zero() { return Acc{0.0, 0.0}; }
	static __device__ __forceinline__ Site load(const In& in, uint64_t i) { return Site{__ldg(in.a + i), __ldg(in.b + i)}; }
	static __device__ __forceinline__ Site none() { return Site{0.0, 0.0}; }
	// fstWindow.cpp:80-83: asum += a; bsum += b  (plain adds; order is the unit tree)
	static __device__ __forceinline__ void fold(Acc& acc, const Site& s) {
		acc.a = __dadd_rn(acc.a, s.a);
		acc.b = __dadd_rn(acc.b, s.b);
	}
	static __device__ __forceinline__ void add(Acc& acc, const Acc& o) {
		acc.a = __dadd_rn(acc.a, o.a);
		acc.b = __dadd_rn(acc.b, o.b);
	}
	static __device__ __forceinline__ Acc shfl_xor(const Acc& v, int m) { return Acc{shfl_xor_f64(v.a, m), shfl_xor_f64(v.b, m)}; }
};

template <class Stat>
__device__ __forceinline__ typename Stat::Acc warp_butterfly(typename Stat::Acc acc) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) Stat::add(acc, Stat::shfl_xor(acc, m));
	return acc;
}

// ----------------------------------------------------------------------------- level 1

// One warp per unit, persistent grid-stride over the scan's unit range.  Lane l folds sites
// l, l+32, l+64, ... of the unit in that order (all loads of a unit are issued before the
// first add: UPL independent 8-byte loads per column per lane in flight), then the butterfly.
template <class Stat, int UPL>
__global__ void __launch_bounds__(256) k_units(DevPlan P, typename Stat::In in, typename Stat::Acc* __restrict__ units) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.unit_base = 0;
	sg.nunits = 0;
	for (uint64_t j = P.unit_lo + warp; j < P.unit_hi; j += nwarp) {
		if (si == 0xffffffffu || j - sg.unit_base >= sg.nunits) {
			si = find_seg<true>(P, j);
			sg = P.segs[si];
		}
		uint64_t st;
		const uint32_t len = pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
		const uint64_t i0 = sg.site_base + st - P.site_origin + lane;
		typename Stat::Acc acc = Stat::zero();
		if (UPL > 0) {
			typename Stat::Site v[UPL > 0 ? UPL : 1];
#pragma unroll
			for (int t = 0; t < UPL; ++t) v[t] = (lane + 32u * t < len) ? Stat::load(in, i0 + 32u * t) : Stat::none();
#pragma unroll
			for (int t = 0; t < UPL; ++t) Stat::fold(acc, v[t]);
		} else {
			for (uint32_t x = lane; x < len; x += 32u) Stat::fold(acc, Stat::load(in, i0 + (x - lane)));
		}
		acc = warp_butterfly<Stat>(acc);
		if (lane == 0) units[j - P.unit_lo] = acc;
	}
}

// ----------------------------------------------------------------------------- level 2

struct WinInfo {
	uint64_t first, last;  // global site (entry) indices, inclusive
	uint32_t nsites;
	uint32_t label;
};

// One warp per window: lane l adds unit partials l, l+32, ... (from +0.0, so a window of
// -0.0 values sums to +0.0 exactly as the reference's `double asum = 0`), then the butterfly.
template <class Stat>
__device__ __forceinline__ typename Stat::Acc window_reduce(const DevPlan& P, const typename Stat::Acc* __restrict__ units, uint64_t w,
                                                            uint32_t lane, WinInfo* wi) {
	const uint32_t si = find_seg<false>(P, w);
	const pgt_seg sg = P.segs[si];
	const uint64_t k = w - sg.win_base;
	uint64_t fu;
	const uint64_t cnt = pgt_window_units(P.g, sg, k, &fu);
	const typename Stat::Acc* up = units + (sg.unit_base + fu - P.unit_lo);
	typename Stat::Acc acc = Stat::zero();
	for (uint64_t x = lane; x < cnt; x += 32u) Stat::add(acc, up[x]);
	acc = warp_butterfly<Stat>(acc);
	uint64_t fs;
	wi->nsites = pgt_window_sites(P.g, sg, k, &fs);
	wi->first = sg.site_base + fs;
	wi->last = wi->first + wi->nsites - 1;
	wi->label = find_contig(P.off, sg.first_contig, sg.ncontig, wi->last);
	return acc;
}

__global__ void __launch_bounds__(256) k_windows_fst(DevPlan P, const FstStat::Acc* __restrict__ units, const uint32_t* __restrict__ pos,
                                                      pgt_fst_out out) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t w = P.win_lo + warp; w < P.win_hi; w += nwarp) {
		WinInfo wi;
		FstStat::Acc acc = window_reduce<FstStat>(P, units, w, lane, &wi);
		if (lane == 0) {
			const uint64_t o = w - P.win_lo;
			if (out.label) out.label[o] = wi.label;
			if (pos) {
				// fstWindow.cpp:71-73 (uint32 arithmetic for the midpoint)
				const uint32_t sp = pos[wi.first - P.site_origin], ep = pos[wi.last - P.site_origin];
				if (out.start_pos) out.start_pos[o] = sp;
				if (out.end_pos) out.end_pos[o] = ep;
				if (out.mid_pos) out.mid_pos[o] = (sp + ep) / 2u;
			}
			if (out.sum_a) out.sum_a[o] = acc.a;
			if (out.sum_b) out.sum_b[o] = acc.b;
			if (out.fst) out.fst[o] = acc.b != 0.0 ? __ddiv_rn(acc.a, acc.b) : 0.0;  // fstWindow.cpp:85
			if (out.nsites) out.nsites[o] = wi.nsites;
		}
	}
}

// ----------------------------------------------------------------------------- host side of a scan

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ScanCtx {
	DevPlan P;
	char* ws;          // device workspace
	size_t units_off;  // byte offset of the unit array
	uint64_t nunits, nwin;
};

static const size_t kMaxAccBytes = 48;  // largest Stat::Acc (fused)

static int resolve_range(const pgt_plan* plan, const pgt_range* range, uint64_t* w_lo, uint64_t* w_hi, uint64_t* u_lo,
                         uint64_t* u_hi, uint64_t* origin) {
	if (!plan) return pgt_set_error(PGT_ERR_ARGS, "plan is NULL");
	uint64_t lo = 0, hi = plan->nwin, org = plan->off[0];
	if (range) {
		if (range->w_hi != 0 || range->w_lo != 0) {
			lo = range->w_lo;
			hi = range->w_hi;
		}
		org = range->site_origin;
	}
	if (lo > hi || hi > plan->nwin) return pgt_set_error(PGT_ERR_ARGS, "window range out of bounds");
	*w_lo = lo;
	*w_hi = hi;
	*origin = org;
	*u_lo = *u_hi = 0;
	if (hi > lo) {
		uint64_t f, c;
		pgt_plan_window_units(plan, lo, &f, &c);
		*u_lo = f;
		pgt_plan_window_units(plan, hi - 1, &f, &c);
		*u_hi = f + c;
		uint64_t fs;
		pgt_plan_window(plan, lo, &fs, nullptr, nullptr);
		if (fs < org) return pgt_set_error(PGT_ERR_ARGS, "site_origin lies after the first site of the window range");
	}
	return PGT_OK;
}

extern "C" size_t pgt_scan_workspace_bytes(const pgt_plan* plan, const pgt_range* range) {
	uint64_t wl, wh, ul, uh, org;
	if (resolve_range(plan, range, &wl, &wh, &ul, &uh, &org) != PGT_OK) return 0;
	size_t b = 0;
	b += align_up(plan->segs.size() * sizeof(pgt_seg), 256);
	b += align_up(plan->off.size() * sizeof(uint64_t), 256);
	b += align_up((size_t)(uh - ul) * kMaxAccBytes, 256);
	return b + 256;
}

// Uploads the plan tables to the head of the workspace and fills ctx.
static int begin_scan(const pgt_plan* plan, const pgt_range* range, void* workspace, size_t workspace_bytes, cudaStream_t st,
                      ScanCtx* ctx) {
	uint64_t wl, wh, ul, uh, org;
	int rc = resolve_range(plan, range, &wl, &wh, &ul, &uh, &org);
	if (rc != PGT_OK) return rc;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) return pgt_set_error(PGT_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
	if (!workspace) return pgt_set_error(PGT_ERR_ARGS, "workspace is NULL");
	if (workspace_bytes < pgt_scan_workspace_bytes(plan, range)) return pgt_set_error(PGT_ERR_NOMEM, "workspace too small, see pgt_scan_workspace_bytes");
	char* ws = (char*)workspace;
	size_t o = 0;
	const size_t seg_bytes = plan->segs.size() * sizeof(pgt_seg);
	const size_t off_bytes = plan->off.size() * sizeof(uint64_t);
	if (seg_bytes) PGT_CUDA(cudaMemcpyAsync(ws + o, plan->segs.data(), seg_bytes, cudaMemcpyHostToDevice, st));
	ctx->P.segs = (const pgt_seg*)(ws + o);
	o += align_up(seg_bytes, 256);
	PGT_CUDA(cudaMemcpyAsync(ws + o, plan->off.data(), off_bytes, cudaMemcpyHostToDevice, st));
	ctx->P.off = (const uint64_t*)(ws + o);
	o += align_up(off_bytes, 256);
	ctx->P.g = plan->g;
	ctx->P.nseg = (uint32_t)plan->segs.size();
	ctx->P.ncontig = (uint32_t)plan->off.size() - 1;
	ctx->P.unit_lo = ul;
	ctx->P.unit_hi = uh;
	ctx->P.win_lo = wl;
	ctx->P.win_hi = wh;
	ctx->P.site_origin = org;
	ctx->P.mode = (int)plan->mode;
	ctx->ws = ws;
	ctx->units_off = o;
	ctx->nunits = uh - ul;
	ctx->nwin = wh - wl;
	return PGT_OK;
}

template <class Stat>
static int launch_units(const ScanCtx& c, typename Stat::In in, cudaStream_t st) {
	if (c.nunits == 0) return PGT_OK;
	typename Stat::Acc* units = (typename Stat::Acc*)(c.ws + c.units_off);
	const int threads = 256;
	const uint64_t want = (c.nunits + 7) / 8;  // one warp per unit, 8 warps per block
	int per_sm = 0;
	const uint32_t upl = c.P.g.u / 32u;
	void (*kern)(DevPlan, typename Stat::In, typename Stat::Acc*) = upl == 8 ? k_units<Stat, 8> : (upl == 4 ? k_units<Stat, 4> : k_units<Stat, 0>);
	PGT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
	uint64_t cap = (uint64_t)num_sms() * (per_sm > 0 ? per_sm : 1);
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	kern<<<grid, threads, 0, st>>>(c.P, in, units);
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

static unsigned window_grid(const ScanCtx& c, int per_sm_hint) {
	const uint64_t want = (c.nwin + 7) / 8;
	const uint64_t cap = (uint64_t)num_sms() * per_sm_hint;
	return (unsigned)(want < cap ? want : cap);
}

extern "C" int pgt_scan_fst(const pgt_plan* plan, const pgt_range* range, const uint32_t* pos, const double* a, const double* b,
                            const pgt_fst_out* out, void* workspace, size_t workspace_bytes, pgt_mem mem, void* stream) {
	if (!a || !b || !out) return pgt_set_error(PGT_ERR_ARGS, "pgt_scan_fst: a, b and out must not be NULL");
	if (mem != PGT_MEM_DEVICE) return pgt_set_error(PGT_ERR_ARGS, "pgt_scan_fst: PGT_MEM_HOST not implemented yet");
	if (plan && plan->mode != PGT_MODE_SITES) return pgt_set_error(PGT_ERR_ARGS, "pgt_scan_fst: plan must be PGT_MODE_SITES");
	cudaStream_t st = (cudaStream_t)stream;
	ScanCtx c{};
	int rc = begin_scan(plan, range, workspace, workspace_bytes, st, &c);
	if (rc != PGT_OK) return rc;
	if (c.nwin == 0) return PGT_OK;
	FstStat::In in{a, b};
	rc = launch_units<FstStat>(c, in, st);
	if (rc != PGT_OK) return rc;
	k_windows_fst<<<window_grid(c, 8), 256, 0, st>>>(c.P, (const FstStat::Acc*)(c.ws + c.units_off), pos, *out);
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}
