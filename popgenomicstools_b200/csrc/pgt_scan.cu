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
//   slide: 0 auto, 1 never use the sliding-tile kernel (k_slide), 2 use it whenever it fits shared memory (any W <= 1048)
static int g_tune_slide = 0;

static int num_sms() {
	int dev = 0, n = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 0;
	cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
	return n;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// -DPGT_BOUNDS (make -C csrc bounds -> libpgtscan_bounds.so): every index a kernel forms from the plan is checked
// against the extent it must stay in -- column slices against the elements the caller's columns hold, positions
// inside a shared-memory stage against the staged slice, unit / window / block indices against this launch's ranges
// -- and a violation traps (the launch fails with an error instead of reading or writing out of bounds).  The GPU
// test-suite is run against this build with PGT_LIB=.../libpgtscan_bounds.so (tools/run_bounds_gpu.sh).
#ifdef PGT_BOUNDS
#define PGT_CHECK(cond)                                                                             \
	do {                                                                                            \
		if (!(cond)) {                                                                              \
			printf("PGT_BOUNDS violated at %s:%d: %s\n", __FILE__, __LINE__, #cond);                \
			__trap();                                                                               \
		}                                                                                           \
	} while (0)
#else
#define PGT_CHECK(cond) ((void)0)
#endif

// ----------------------------------------------------------------------------- device plan

struct Cols {
	const uint32_t* pos;
	const double* a;
	const double* b;
	const int8_t* g;
	const double* f1;
	const double* f2;
	const int32_t* n1;
	const int32_t* n2;
	int minind;
};

struct DevPlan {
	pgt_geom g;
	const pgt_seg* segs;
	const uint64_t* off;       // contig offsets on the plan axis (sites, or bp entries)
	const uint64_t* site_off;  // bp mode: cumulative site counts per chromosome (global site indices)
	uint32_t nseg;
	uint32_t ncontig;
	uint64_t unit_lo, unit_hi;  // global unit range reduced by this launch
	uint64_t win_lo, win_hi;    // global window range of this scan
	uint64_t site_origin;       // global index of element 0 of the columns
	uint64_t nunits_total;
	uint64_t col_elems;         // elements the columns hold from element 0 (PGT_BOUNDS checks; ~0 = unknown)
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
// last segment with blk_base <= x AND at least one scan block
__device__ __forceinline__ uint32_t find_seg_by_block(const DevPlan& P, uint64_t x) {
	uint32_t lo = 0, hi = P.nseg;
	while (hi - lo > 1) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		if (P.segs[mid].blk_base <= x) lo = mid;
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
__device__ __forceinline__ uint32_t shfl_xor_u32(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ----------------------------------------------------------------------------- statistics
//
// A Stat describes: the per-site record loaded from the columns (Site), the unit partial (Acc,
// stored as-is in the unit array), how one site is folded into a lane's partial (the per-site
// statistic), the combine used by the butterflies, and the window epilogue.

// --- fstWindow: asum += a; bsum += b (fstWindow.cpp:80-83); fst = bsum != 0 ? asum/bsum : 0 (:85)
struct FstStat {
	struct Acc {
		double a, b;
	};
	struct Site {
		double a, b;
	};
	static __device__ __forceinline__ Acc zero() { return Acc{0.0, 0.0}; }
	static __device__ __forceinline__ Site load(const Cols& c, uint64_t i) { return Site{__ldg(c.a + i), __ldg(c.b + i)}; }
	// tile columns in staging order (a, b); generic loads: the tile lives in shared memory
	static __device__ __forceinline__ Site load_tile(const char* const* cp, uint32_t i) {
		return Site{((const double*)cp[0])[i], ((const double*)cp[1])[i]};
	}
	static __device__ __forceinline__ void fold(Acc& acc, const Site& s, int) {
		acc.a = __dadd_rn(acc.a, s.a);
		acc.b = __dadd_rn(acc.b, s.b);
	}
	static __device__ __forceinline__ void add(Acc& acc, const Acc& o) {
		acc.a = __dadd_rn(acc.a, o.a);
		acc.b = __dadd_rn(acc.b, o.b);
	}
	static __device__ __forceinline__ Acc shfl_xor(const Acc& v, int m) { return Acc{shfl_xor_f64(v.a, m), shfl_xor_f64(v.b, m)}; }
	static __device__ __forceinline__ void emit(const pgt_windows& out, uint64_t o, const Acc& acc) {
		if (out.sum_a) out.sum_a[o] = acc.a;
		if (out.sum_b) out.sum_b[o] = acc.b;
		if (out.fst) out.fst[o] = acc.b != 0.0 ? __ddiv_rn(acc.a, acc.b) : 0.0;
	}
};

// --- hetWindow: nonmissing += (g >= 0); nhet += (g == 1) (hetWindow.cpp:77-82); h = nhet/nonmissing (:84)
struct HetStat {
	struct Acc {
		uint32_t nonmissing, nhet;
	};
	struct Site {
		int g;
	};
	static __device__ __forceinline__ Acc zero() { return Acc{0u, 0u}; }
	static __device__ __forceinline__ Site load(const Cols& c, uint64_t i) { return Site{(int)__ldg(c.g + i)}; }
	static __device__ __forceinline__ Site load_tile(const char* const* cp, uint32_t i) { return Site{(int)((const int8_t*)cp[0])[i]}; }
	static __device__ __forceinline__ void fold(Acc& acc, const Site& s, int) {
		acc.nonmissing += (s.g >= 0);
		acc.nhet += (s.g == 1);
	}
	static __device__ __forceinline__ void add(Acc& acc, const Acc& o) {
		acc.nonmissing += o.nonmissing;
		acc.nhet += o.nhet;
	}
	static __device__ __forceinline__ Acc shfl_xor(const Acc& v, int m) { return Acc{shfl_xor_u32(v.nonmissing, m), shfl_xor_u32(v.nhet, m)}; }
	static __device__ __forceinline__ void emit(const pgt_windows& out, uint64_t o, const Acc& acc) {
		if (out.nhet) out.nhet[o] = acc.nhet;
		if (out.nonmissing) out.nonmissing[o] = acc.nonmissing;
		if (out.het) out.het[o] = acc.nonmissing != 0 ? __ddiv_rn((double)acc.nhet, (double)acc.nonmissing) : 0.0;
	}
};

// --- dxyWindow: per-site dxy (dxyWindow.cpp:381) with explicit non-fused IEEE ops so the value is
// bit-identical to the x86-64 reference (no FMA there); window fold dxyWindow.cpp:179-186:
// v >= 0 -> dxy += v, ++neffective; v == -9 -> ++nskip.
__device__ __forceinline__ double dxy_site_value(double f1, double f2, int n1, int n2, int minind) {
	return (n1 >= minind && n2 >= minind) ? __dadd_rn(__dmul_rn(f1, __dsub_rn(1.0, f2)), __dmul_rn(f2, __dsub_rn(1.0, f1))) : -9.0;
}
struct DxyStat {
	struct Acc {
		double dxy;
		uint32_t neff, nskip;
	};
	struct Site {
		double f1, f2;
		int n1, n2;
	};
	static __device__ __forceinline__ Acc zero() { return Acc{0.0, 0u, 0u}; }
	static __device__ __forceinline__ Site load(const Cols& c, uint64_t i) {
		return Site{__ldg(c.f1 + i), __ldg(c.f2 + i), __ldg(c.n1 + i), __ldg(c.n2 + i)};
	}
	static __device__ __forceinline__ Site load_tile(const char* const* cp, uint32_t i) {
		return Site{((const double*)cp[0])[i], ((const double*)cp[1])[i], ((const int32_t*)cp[2])[i], ((const int32_t*)cp[3])[i]};
	}
	static __device__ __forceinline__ void fold(Acc& acc, const Site& s, int minind) {
		const double v = dxy_site_value(s.f1, s.f2, s.n1, s.n2, minind);
		if (v >= 0.0) {
			acc.dxy = __dadd_rn(acc.dxy, v);
			++acc.neff;
		} else if (v == -9.0) {
			++acc.nskip;
		}
	}
	static __device__ __forceinline__ void add(Acc& acc, const Acc& o) {
		acc.dxy = __dadd_rn(acc.dxy, o.dxy);
		acc.neff += o.neff;
		acc.nskip += o.nskip;
	}
	static __device__ __forceinline__ Acc shfl_xor(const Acc& v, int m) {
		return Acc{shfl_xor_f64(v.dxy, m), shfl_xor_u32(v.neff, m), shfl_xor_u32(v.nskip, m)};
	}
	static __device__ __forceinline__ void emit(const pgt_windows& out, uint64_t o, const Acc& acc) {
		if (out.dxy) out.dxy[o] = acc.dxy;
		if (out.neffective) out.neffective[o] = acc.neff;
		if (out.nskip) out.nskip[o] = acc.nskip;
	}
};

// --- fused fst + dxy + het over one site axis (BASELINE config 5): one pass, 41 B/site
struct FusedStat {
	struct Acc {
		FstStat::Acc fst;
		DxyStat::Acc dxy;
		HetStat::Acc het;
	};
	struct Site {
		FstStat::Site fst;
		DxyStat::Site dxy;
		HetStat::Site het;
	};
	static __device__ __forceinline__ Acc zero() { return Acc{FstStat::zero(), DxyStat::zero(), HetStat::zero()}; }
	static __device__ __forceinline__ Site load(const Cols& c, uint64_t i) { return Site{FstStat::load(c, i), DxyStat::load(c, i), HetStat::load(c, i)}; }
	// staging order: a, b, f1, f2, n1, n2, geno
	static __device__ __forceinline__ Site load_tile(const char* const* cp, uint32_t i) {
		return Site{FstStat::load_tile(cp, i), DxyStat::load_tile(cp + 2, i), HetStat::load_tile(cp + 6, i)};
	}
	static __device__ __forceinline__ void fold(Acc& acc, const Site& s, int minind) {
		FstStat::fold(acc.fst, s.fst, minind);
		DxyStat::fold(acc.dxy, s.dxy, minind);
		HetStat::fold(acc.het, s.het, minind);
	}
	static __device__ __forceinline__ void add(Acc& acc, const Acc& o) {
		FstStat::add(acc.fst, o.fst);
		DxyStat::add(acc.dxy, o.dxy);
		HetStat::add(acc.het, o.het);
	}
	static __device__ __forceinline__ Acc shfl_xor(const Acc& v, int m) {
		return Acc{FstStat::shfl_xor(v.fst, m), DxyStat::shfl_xor(v.dxy, m), HetStat::shfl_xor(v.het, m)};
	}
	static __device__ __forceinline__ void emit(const pgt_windows& out, uint64_t o, const Acc& acc) {
		FstStat::emit(out, o, acc.fst);
		DxyStat::emit(out, o, acc.dxy);
		HetStat::emit(out, o, acc.het);
	}
};

template <class Stat>
__device__ __forceinline__ typename Stat::Acc warp_butterfly(typename Stat::Acc acc) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) Stat::add(acc, Stat::shfl_xor(acc, m));
	return acc;
}
// butterfly inside aligned groups of G lanes
template <class Stat, int G>
__device__ __forceinline__ typename Stat::Acc group_butterfly(typename Stat::Acc acc) {
#pragma unroll
	for (int m = G / 2; m >= 1; m >>= 1) Stat::add(acc, Stat::shfl_xor(acc, m));
	return acc;
}

// ----------------------------------------------------------------------------- level 1

// One warp per unit, persistent grid-stride over the launch's unit range.  Lane l folds sites
// l, l+32, l+64, ... of the unit in that order (all loads of a unit are issued before the first
// fold: UPL independent loads per column per lane in flight), then the butterfly.
// INDIRECT (bp mode): the unit's site range comes from `bounds` instead of the closed form.
template <class Stat, int UPL, bool INDIRECT>
__global__ void __launch_bounds__(256) k_units(DevPlan P, Cols cols, typename Stat::Acc* __restrict__ units, const uint64_t* __restrict__ bounds) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.unit_base = 0;
	sg.nunits = 0;
	for (uint64_t j = P.unit_lo + warp; j < P.unit_hi; j += nwarp) {
		uint64_t i0;
		uint32_t len;
		if (INDIRECT) {
			const uint64_t b0 = bounds[j - P.unit_lo], b1 = bounds[j - P.unit_lo + 1];
			i0 = b0 + lane;
			len = (uint32_t)(b1 - b0);
		} else {
			if (si == 0xffffffffu || j - sg.unit_base >= sg.nunits) {
				si = find_seg<true>(P, j);
				sg = P.segs[si];
			}
			uint64_t st;
			len = pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
			i0 = sg.site_base + st - P.site_origin + lane;
		}
		PGT_CHECK(len == 0 || (i0 - lane) + len <= P.col_elems);
		typename Stat::Acc acc = Stat::zero();
		if (UPL > 0 && len <= 32u * UPL) {
			typename Stat::Site v[UPL > 0 ? UPL : 1];
#pragma unroll
			for (int t = 0; t < UPL; ++t)
				if (lane + 32u * t < len) v[t] = Stat::load(cols, i0 + 32u * t);
#pragma unroll
			for (int t = 0; t < UPL; ++t)
				if (lane + 32u * t < len) Stat::fold(acc, v[t], cols.minind);
		} else {
			for (uint32_t x = lane; x < len; x += 32u) Stat::fold(acc, Stat::load(cols, i0 + (x - lane)), cols.minind);
		}
		acc = warp_butterfly<Stat>(acc);
		if (lane == 0) units[j - P.unit_lo] = acc;
	}
}

// ----------------------------------------------------------------------------- level 1, tiled
//
// Persistent CTAs (one per SM) walk tiles of `m` consecutive units.  Warp 0 is the producer: it
// stages the tile's slice of every column in shared memory with 1-D bulk async copies
// (cp.async.bulk -> UBLKCP, completion on an mbarrier), two stages deep, so the bytes in flight
// per SM are one whole tile (64-96 KB) and cost no registers.  Warps 1..15 are consumers: groups
// of G lanes reduce one unit each straight from shared memory (lane g of a group folds sites
// g, g+G, g+2G, ... in that order, then a log2(G)-level butterfly), so the summation order is the
// same function of (W, S, u) as in k_units and does not depend on tiles, CTAs or shards.  G is
// small when units are short (pgt_geom.gw), which keeps all lanes busy for fine windows.
// Only the 16-byte-aligned interior of a slice is bulk-copied; the <16-byte head and tail are
// copied by the producer's lanes, so nothing outside [column, column + n) is ever read.

static constexpr int kTileThreads = 512;
static constexpr int kTileMaxStages = 4;
static constexpr int kTileConsumerWarps = kTileThreads / 32 - 1;
static constexpr int kMaxTileCols = 7;  // of a statistic; the sliding tile stages `pos` as one more (kMaxStageCols)
static constexpr int kMaxStageCols = kMaxTileCols + 1;
static constexpr uint32_t kTileCtlBytes = 384;  // >= sizeof(TileCtl) = 368


// With more segments than this the tiled kernel gets a precomputed tile -> segment table: a CTA's
// consecutive tiles lie gridDim * m units apart, i.e. in different segments once contigs are shorter
// than ~1e6 sites, and the producer then paid two binary searches over the segment table per tile
// (measured: 1e3 contigs 5.5 TB/s, 1e5 contigs 2.6 TB/s, against 7.0 TB/s for 24 contigs).
static constexpr size_t kTileSegTableMin = 32;

struct TileCfg {
	const uint32_t* tile_seg;  // [ntiles] segment of each tile's first unit, or NULL (few segments / bp mode)
	const char* gcol[kMaxStageCols];  // global column pointers (element 0 = site_origin), staging order
	uint32_t elem[kMaxStageCols];     // bytes per site
	uint32_t col_off[kMaxStageCols];  // byte offset of the column's region inside a stage
	uint32_t col_cap[kMaxStageCols];  // capacity of that region in bytes
	uint32_t ncol;
	uint32_t m;            // units per tile
	uint32_t stage_bytes;
	uint32_t nstages;      // 2..kTileMaxStages
	int minind;
	uint64_t valid_elems;  // elements every column holds from element 0 (bounds the aligned superset copies)
};

struct TileCtl {
	uint64_t full[kTileMaxStages];
	uint64_t empty[kTileMaxStages];
	uint64_t s0[kTileMaxStages];                   // column element index of the tile's first site
	uint32_t ns[kTileMaxStages];                   // elements staged (PGT_BOUNDS checks)
	const char* cp[kTileMaxStages][kMaxStageCols];  // where site s0 of each column lives (shared, or global if unstaged)
};
static_assert(sizeof(TileCtl) <= kTileCtlBytes, "control block");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "LAB_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra DONE;\n"
	    "bra LAB_WAIT;\n"
	    "DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
	             "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// Producer side of a stage (all 32 lanes of the producer warp; the caller has waited for the stage to
// drain): lane c stages elements [s0, s1) of column c.  The bulk copy covers the 16-byte-aligned SUPERSET
// of the slice whenever that stays inside the column (always, except at the first/last elements of a
// column that is not 16-byte aligned/padded); only then are head/tail bytes copied by hand, so nothing
// outside [column, column + valid_elems) is ever read.
__device__ __forceinline__ void producer_fill_stage(const TileCfg& tc, TileCtl* ctl, unsigned char* stages, uint32_t stg, uint64_t s0,
                                                    uint64_t s1, uint32_t lane) {
	PGT_CHECK(s0 <= s1 && s1 <= tc.valid_elems && stg < tc.nstages);
	// generic-proxy reads of this stage are done; order them before the async-proxy writes
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	unsigned char* stage = stages + (size_t)stg * tc.stage_bytes;
	const uint32_t c = lane < tc.ncol ? lane : 0u;
	const char* A = tc.gcol[c] + s0 * tc.elem[c];
	const uint64_t nbytes = (s1 - s0) * tc.elem[c];
	const uint32_t pad = (uint32_t)((uintptr_t)A & 15u);
	const bool staged = lane < tc.ncol && pad + nbytes + 16u <= tc.col_cap[c];
	unsigned char* region = stage + tc.col_off[c];
	const char* col_lo = tc.gcol[c];
	const char* col_hi = tc.gcol[c] + tc.valid_elems * tc.elem[c];
	const char* B0 = A - pad;  // aligned superset [B0, B1)
	const char* B1 = (const char*)(((uintptr_t)(A + nbytes) + 15u) & ~(uintptr_t)15u);
	uint32_t nh = 0, ntl = 0;
	if (B0 < col_lo) {  // cannot read before the column: copy the head by hand
		B0 += 16;
		nh = 16u - pad;
		if (nh > nbytes) nh = (uint32_t)nbytes;
	}
	if (B1 > col_hi) {  // cannot read past the column: copy the tail by hand
		B1 -= 16;
		ntl = (uint32_t)((A + nbytes) - B1);
		if (B1 < A + nh) ntl = (uint32_t)(nbytes - nh);
	}
	uint32_t tx = (staged && B1 > B0) ? (uint32_t)(B1 - B0) : 0u;
	if (lane < tc.ncol) ctl->cp[stg][c] = staged ? (const char*)(region + pad) : A;
	uint32_t txsum = tx;
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) txsum += __shfl_xor_sync(0xffffffffu, txsum, m);
	PGT_CHECK(!tx || (B0 >= col_lo && B1 <= col_hi && (uint32_t)(pad + (B0 - A)) + tx <= tc.col_cap[c] && tc.col_off[c] + tc.col_cap[c] <= tc.stage_bytes));
	if (lane == 0) {
		ctl->s0[stg] = s0;
		ctl->ns[stg] = (uint32_t)(s1 - s0);
		mbar_arrive_expect_tx(&ctl->full[stg], txsum);
	}
	__syncwarp();
	if (tx) bulk_g2s(region + pad + (B0 - A), B0, tx, &ctl->full[stg]);
	// rare: hand-copied head / tail bytes
	const uint32_t any = __ballot_sync(0xffffffffu, staged && (nh | ntl));
	for (uint32_t cc = 0; cc < tc.ncol; ++cc) {
		if (!((any >> cc) & 1u)) continue;
		const uint32_t nh_c = __shfl_sync(0xffffffffu, nh, cc), nt_c = __shfl_sync(0xffffffffu, ntl, cc);
		const uint32_t pad_c = __shfl_sync(0xffffffffu, pad, cc);
		const unsigned long long A_c = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)A, cc);
		const unsigned long long nb_c = __shfl_sync(0xffffffffu, (unsigned long long)nbytes, cc);
		unsigned char* reg_c = stage + tc.col_off[cc];
		const unsigned char* Ac = (const unsigned char*)(uintptr_t)A_c;
		if (lane < nh_c) reg_c[pad_c + lane] = __ldg(Ac + lane);
		if (lane < nt_c) reg_c[pad_c + (uint32_t)(nb_c - nt_c) + lane] = __ldg(Ac + (nb_c - nt_c) + lane);
	}
	__syncwarp();
	if (lane == 0) mbar_arrive(&ctl->full[stg]);  // control words (and any head/tail bytes) are in place
}

// global site (entry) index where global unit j starts / ends
__device__ __forceinline__ uint64_t unit_bounds_global(const DevPlan& P, uint64_t j, uint64_t* end) {
	const pgt_seg sg = P.segs[find_seg<true>(P, j)];
	uint64_t st;
	const uint32_t len = pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
	*end = sg.site_base + st + len;
	return sg.site_base + st;
}

__global__ void __launch_bounds__(256) k_tile_segs(DevPlan P, uint32_t m, uint64_t ntiles, uint32_t* __restrict__ tile_seg) {
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < ntiles) tile_seg[t] = find_seg<true>(P, P.unit_lo + t * m);
}

// hetWindow's 1-byte genotype column inside the tiled kernel: a warp reduces one unit from the staged tile
// with 16-byte loads + byte-SIMD compare / popc (the generic per-site loop would need one load per byte).
// The staged copy keeps the column's alignment modulo 16, so the aligned chunks are the same in shared
// memory and -- for a slice that was too long to stage -- in global memory (generic loads serve both).
// Integer counts: independent of the order, identical to the per-site fold.
// Four genotypes per word.  Both tests end in a word with 0x80 in every byte that counts, and an unsigned
// dp4a against 0x01010101 adds 128 per such byte: the accumulators hold 128 x count (shifted down once per
// unit; a lane sees < 2^25 bytes of a unit).  ~7 integer instructions per word and no POPC, which runs at a
// quarter of the integer rate (the POPC / __vcmpeq4 version needed 13 and was issue-bound at 4.5 TB/s).
__device__ __forceinline__ void het_count_word(uint32_t w, uint32_t& nonmissing128, uint32_t& nhet128) {
	nonmissing128 = __dp4a(~w & 0x80808080u, 0x01010101u, nonmissing128);  // g >= 0  (hetWindow.cpp:78): sign bit clear
	const uint32_t x = w ^ 0x01010101u;                                     // g == 1  (hetWindow.cpp:80): byte of x is zero
	const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;                     // bit 7 of t | x is set iff the byte of x is not zero
	nhet128 = __dp4a(~(t | x) & 0x80808080u, 0x01010101u, nhet128);
}
__device__ __forceinline__ void het_unit_from_tile(HetStat::Acc& acc, const char* col, uint32_t rel, uint32_t len, uint32_t lane) {
	const int8_t* A = (const int8_t*)col + rel;
	const int8_t* E = A + len;
	const int8_t* A0 = (const int8_t*)(((uintptr_t)A + 15u) & ~(uintptr_t)15u);  // first aligned chunk inside
	const int8_t* A1 = (const int8_t*)((uintptr_t)E & ~(uintptr_t)15u);          // end of the last aligned chunk
	uint32_t nonmissing = 0, nhet = 0;
	if (A1 > A0) {
		const uint32_t nch = (uint32_t)(A1 - A0) >> 4;
		uint32_t nm128 = 0, h128 = 0;
		for (uint32_t c = lane; c < nch; c += 32u) {
			const uint4 v = *(reinterpret_cast<const uint4*>(A0) + c);
			het_count_word(v.x, nm128, h128);
			het_count_word(v.y, nm128, h128);
			het_count_word(v.z, nm128, h128);
			het_count_word(v.w, nm128, h128);
		}
		nonmissing = nm128 >> 7;
		nhet = h128 >> 7;
		const uint32_t nh = (uint32_t)(A0 - A), nt = (uint32_t)(E - A1);  // < 16 each
		if (lane < nh) {
			const int g = A[lane];
			nonmissing += (g >= 0);
			nhet += (g == 1);
		}
		if (lane >= 16u && lane - 16u < nt) {
			const int g = A1[lane - 16u];
			nonmissing += (g >= 0);
			nhet += (g == 1);
		}
	} else {
		for (uint32_t x = lane; x < len; x += 32u) {  // < 32 bytes without an aligned chunk
			const int g = A[x];
			nonmissing += (g >= 0);
			nhet += (g == 1);
		}
	}
	acc.nonmissing += nonmissing;
	acc.nhet += nhet;
}

template <class Stat, int G, bool INDIRECT>
__global__ void __launch_bounds__(kTileThreads, 1)
    k_units_tiled(DevPlan P, TileCfg tc, typename Stat::Acc* __restrict__ units, const uint64_t* __restrict__ bounds) {
	extern __shared__ __align__(128) unsigned char smem[];
	TileCtl* ctl = reinterpret_cast<TileCtl*>(smem);
	unsigned char* stages = smem + kTileCtlBytes;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint64_t nunits = P.unit_hi - P.unit_lo;
	const uint64_t ntiles = (nunits + tc.m - 1) / tc.m;

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < tc.nstages; ++s) {
			mbar_init(&ctl->full[s], 2);
			mbar_init(&ctl->empty[s], kTileConsumerWarps);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == 0) {
		// ------------------------------------------------------------------ producer
		// lane c stages column c: the bulk copy covers the 16-byte-aligned SUPERSET of the slice
		// whenever that stays inside the column (always, except at the first/last elements of a
		// column that is not 16-byte aligned/padded); only then are head/tail bytes copied by hand.
		uint32_t it = 0;
		uint32_t psi = 0xffffffffu;  // cached segment of the producer
		pgt_seg psg;
		psg.unit_base = 0;
		psg.nunits = 0;
		auto unit_span = [&](uint64_t j, uint64_t* end) -> uint64_t {  // global [start, end) of unit j
			if (psi == 0xffffffffu || j - psg.unit_base >= psg.nunits) {
				if (tc.tile_seg && psi != 0xffffffffu && j >= psg.unit_base) {
					do {  // forward from the tile's first segment (set from the table below)
						++psi;
						psg = P.segs[psi];
					} while (j - psg.unit_base >= psg.nunits);
				} else {
					psi = find_seg<true>(P, j);
					psg = P.segs[psi];
				}
			}
			uint64_t st;
			const uint32_t len = pgt_unit_range(P.g, psg.nsites, j - psg.unit_base, &st);
			*end = psg.site_base + st + len;
			return psg.site_base + st;
		};
		for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
			const uint32_t stg = it % tc.nstages;
			// the tile's element range is computed BEFORE waiting for the stage to drain
			const uint64_t j0 = P.unit_lo + t * tc.m;
			const uint64_t j1 = (P.unit_hi - j0 < tc.m) ? P.unit_hi : j0 + tc.m;
			uint64_t s0, s1;  // element range of the tile in the columns
			if (INDIRECT) {
				s0 = bounds[j0 - P.unit_lo];
				s1 = bounds[j1 - P.unit_lo];
			} else {
				uint64_t e;
				if (tc.tile_seg) {
					const uint32_t ts = tc.tile_seg[t];
					if (ts != psi) {
						psi = ts;
						psg = P.segs[psi];
					}
				}
				s0 = unit_span(j0, &e) - P.site_origin;
				unit_span(j1 - 1, &e);
				s1 = e - P.site_origin;
			}
			if (it >= tc.nstages) mbar_wait(&ctl->empty[stg], ((it / tc.nstages) - 1u) & 1u);
			producer_fill_stage(tc, ctl, stages, stg, s0, s1, lane);
		}
	} else {
		// ------------------------------------------------------------------ consumers
		constexpr uint32_t GPW = 32u / G;  // groups per warp
		const uint32_t gl = lane % G;      // lane inside its group
		const uint32_t wgroup0 = (warp - 1u) * GPW;
		constexpr uint32_t NGROUPS = kTileConsumerWarps * GPW;
		uint32_t si = 0xffffffffu;
		pgt_seg sg;
		sg.unit_base = 0;
		sg.nunits = 0;
		uint32_t it = 0;
		for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
			const uint32_t stg = it % tc.nstages;
			mbar_wait(&ctl->full[stg], (it / tc.nstages) & 1u);
			const uint64_t s0 = ctl->s0[stg];
			const char* cp[kMaxTileCols];
#pragma unroll
			for (int c = 0; c < kMaxTileCols; ++c) cp[c] = ctl->cp[stg][c];
			const uint64_t j0 = P.unit_lo + t * tc.m;
			const uint32_t cnt = (uint32_t)((P.unit_hi - j0 < tc.m) ? (P.unit_hi - j0) : tc.m);
			// Units are dealt to the consumer groups round-robin by GLOBAL unit index, so the group that
			// gets the extra unit rotates from tile to tile and warps that finish early run ahead
			// into the next stage: no warp is systematically idle.
			const uint32_t rot = (uint32_t)((t * tc.m) % NGROUPS);
			const uint32_t myg = wgroup0 + lane / G;
			const uint32_t first = myg >= rot ? myg - rot : myg + NGROUPS - rot;
			for (uint32_t base = 0; base < cnt; base += NGROUPS) {  // warp-uniform trip count
				const uint32_t ul = base + first;
				const bool active = ul < cnt;
				const uint64_t j = j0 + (active ? ul : 0u);
				uint32_t rel, len;
				if (INDIRECT) {
					const uint64_t b0 = bounds[j - P.unit_lo], b1 = bounds[j - P.unit_lo + 1];
					rel = (uint32_t)(b0 - s0);
					len = (uint32_t)(b1 - b0);
				} else {
					if (si == 0xffffffffu || j - sg.unit_base >= sg.nunits) {
						if (tc.tile_seg) {  // forward from the tile's first segment
							if (si == 0xffffffffu || j < sg.unit_base || j - sg.unit_base >= sg.nunits + (uint64_t)tc.m) {
								si = tc.tile_seg[t];
								sg = P.segs[si];
							}
							while (j - sg.unit_base >= sg.nunits) {
								++si;
								sg = P.segs[si];
							}
						} else {
							si = find_seg<true>(P, j);
							sg = P.segs[si];
						}
					}
					uint64_t st;
					len = pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
					rel = (uint32_t)(sg.site_base + st - P.site_origin - s0);
				}
				if (!active) len = 0;
				PGT_CHECK(len == 0 || (rel + len <= ctl->ns[stg] && j >= P.unit_lo && j < P.unit_hi));
				typename Stat::Acc acc = Stat::zero();
				if constexpr (std::is_same<Stat, HetStat>::value && G == 32) {
					het_unit_from_tile(acc, cp[0], rel, len, lane);
				} else {
					for (uint32_t x = gl; x < len; x += G) Stat::fold(acc, Stat::load_tile(cp, rel + x), tc.minind);
				}
				acc = group_butterfly<Stat, G>(acc);
				if (active && gl == 0) units[j - P.unit_lo] = acc;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&ctl->empty[stg]);
		}
	}
}

// hetWindow's 1-byte genotype column: a warp per unit, lanes take 16-byte aligned chunks
// (LDG.128) and count with byte-SIMD + popc; the partial chunks at the two ends of the unit are
// read byte by byte, so nothing outside [unit start, unit end) is touched.  Integer counts: the
// result is independent of the order, identical to the generic kernels.
// volatile asm: the eight loads of a round stay back to back (the compiler otherwise interleaves
// them with the counting and keeps only ~3 in flight)
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
template <bool INDIRECT>
__global__ void __launch_bounds__(256, 4) k_units_het_vec(DevPlan P, const int8_t* __restrict__ geno, HetStat::Acc* __restrict__ units,
                                                        const uint64_t* __restrict__ bounds) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.unit_base = 0;
	sg.nunits = 0;
	for (uint64_t j = P.unit_lo + warp; j < P.unit_hi; j += nwarp) {
		uint64_t i0;
		uint32_t len;
		if (INDIRECT) {
			const uint64_t b0 = bounds[j - P.unit_lo], b1 = bounds[j - P.unit_lo + 1];
			i0 = b0;
			len = (uint32_t)(b1 - b0);
		} else {
			if (si == 0xffffffffu || j - sg.unit_base >= sg.nunits) {
				si = find_seg<true>(P, j);
				sg = P.segs[si];
			}
			uint64_t st;
			len = pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
			i0 = sg.site_base + st - P.site_origin;
		}
		const int8_t* A = geno + i0;
		const int8_t* E = A + len;
		PGT_CHECK(len == 0 || i0 + len <= P.col_elems);
		const int8_t* A0 = (const int8_t*)(((uintptr_t)A + 15u) & ~(uintptr_t)15u);  // first aligned chunk inside
		const int8_t* A1 = (const int8_t*)((uintptr_t)E & ~(uintptr_t)15u);          // end of the last aligned chunk
		uint32_t nonmissing = 0, nhet = 0;
		if (A1 > A0) {
			uint32_t nm128 = 0, h128 = 0;
			const uint32_t nch = (uint32_t)(A1 - A0) >> 4;
			// 8 x 16 bytes in flight per lane (one 4096-site unit = one round); chunks past the end read
			// as 0x80 bytes = missing genotypes, which count for nothing
			const uint4 kMissing = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
			for (uint32_t c0 = lane; c0 < nch; c0 += 256u) {
				uint4 v[8];
#pragma unroll
				for (int q = 0; q < 8; ++q) {  // unconditional loads (index clamped) so that all 8 are issued back to back
					const uint32_t c = c0 + 32u * q;
					v[q] = ldg_stream_u4(reinterpret_cast<const uint4*>(A0) + (c < nch ? c : nch - 1u));
				}
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					if (c0 + 32u * q >= nch) v[q] = kMissing;
					het_count_word(v[q].x, nm128, h128);
					het_count_word(v[q].y, nm128, h128);
					het_count_word(v[q].z, nm128, h128);
					het_count_word(v[q].w, nm128, h128);
				}
			}
			nonmissing = nm128 >> 7;
			nhet = h128 >> 7;
			const uint32_t nh = (uint32_t)(A0 - A), nt = (uint32_t)(E - A1);  // < 16 each
			if (lane < nh) {
				const int g = __ldg(A + lane);
				nonmissing += (g >= 0);
				nhet += (g == 1);
			}
			if (lane >= 16u && lane - 16u < nt) {
				const int g = __ldg(A1 + (lane - 16u));
				nonmissing += (g >= 0);
				nhet += (g == 1);
			}
		} else {
			for (uint32_t x = lane; x < len; x += 32u) {  // < 32 bytes without an aligned chunk
				const int g = __ldg(A + x);
				nonmissing += (g >= 0);
				nhet += (g == 1);
			}
		}
		HetStat::Acc acc{nonmissing, nhet};
		acc = warp_butterfly<HetStat>(acc);
		if (lane == 0) units[j - P.unit_lo] = acc;
	}
}

// bp mode: bounds[t] = index (relative to the columns' element 0) of the first site at or after
// the first bp of unit unit_lo+t, t in [0, unit_hi-unit_lo]; the unit's bp -> (chromosome, pos) is
// closed form, the site is a lower_bound in that chromosome's slice of `pos`.
__global__ void __launch_bounds__(256) k_bp_bounds(DevPlan P, const uint32_t* __restrict__ pos, uint64_t ndata, uint64_t* __restrict__ bounds) {
	const uint64_t nb = P.unit_hi - P.unit_lo + 1;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nb; t += stride) {
		const uint64_t j = P.unit_lo + t;
		uint64_t res;
		if (j >= P.nunits_total) {
			res = P.site_off[P.ncontig];
		} else {
			const pgt_seg sg = P.segs[find_seg<true>(P, j)];
			uint64_t st;
			pgt_unit_range(P.g, sg.nsites, j - sg.unit_base, &st);
			const uint64_t e = sg.site_base + st;  // global bp entry
			const uint32_t c = find_contig(P.off, sg.first_contig, sg.ncontig, e);
			const uint32_t p = (uint32_t)(e - P.off[c]) + 1u;  // 1-based position on chromosome c
			uint64_t lo = P.site_off[c], hi = P.site_off[c + 1];
			// clip to the sites this call holds
			if (lo < P.site_origin) lo = P.site_origin;
			if (hi > P.site_origin + ndata) hi = P.site_origin + ndata;
			if (hi < lo) hi = lo;
			while (lo < hi) {
				const uint64_t mid = lo + ((hi - lo) >> 1);
				if (__ldg(pos + (mid - P.site_origin)) < p) lo = mid + 1;
				else hi = mid;
			}
			res = lo;
		}
		if (res < P.site_origin) res = P.site_origin;
		if (res > P.site_origin + ndata) res = P.site_origin + ndata;
		bounds[t] = res - P.site_origin;
	}
}

// ----------------------------------------------------------------------------- level 2

template <class Stat>
__device__ __forceinline__ void emit_window(const DevPlan& P, const pgt_seg& sg, uint64_t w, uint64_t k, const typename Stat::Acc& acc,
                                            const uint32_t* __restrict__ pos, const pgt_windows& out, bool have_edges = false,
                                            uint32_t edge_start = 0, uint32_t edge_end = 0);

// One warp per window: lane l adds unit partials l, l+32, ... (from +0.0, so a window of
// -0.0 values sums to +0.0 exactly as the reference's `double asum = 0`), then the butterfly.
// units_base = global index of units[0].
template <class Stat>
__global__ void __launch_bounds__(256) k_windows(DevPlan P, const typename Stat::Acc* __restrict__ units, uint64_t units_base,
                                                  const uint32_t* __restrict__ pos, pgt_windows out) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	// The per-window work is a chain of dependent memory round trips; keep it short: the segment is
	// cached (w only grows, so it changes a few dozen times per warp instead of costing a binary
	// search in global memory per window), and lane 0 fetches the two edge positions BEFORE the
	// partials are summed, so that gather overlaps the partial loads instead of following them.
	// Every warp takes a contiguous run of windows, so the segment only ever moves forward by a step or
	// two: one binary search per warp, then a walk (genomes of 1e5 contigs with W larger than the contigs
	// paid a 17-probe search in L2 per window: 0.86 ms for 1e5 windows).
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.win_base = 0;
	sg.nwin = 0;
	const uint64_t nwin_all = P.win_hi - P.win_lo;
	const uint64_t per = (nwin_all + nwarp - 1) / nwarp;
	const uint64_t w_begin = P.win_lo + warp * per;
	const uint64_t w_end = w_begin + per < P.win_hi ? w_begin + per : P.win_hi;
	for (uint64_t w = w_begin; w < w_end; ++w) {
		if (si == 0xffffffffu) {
			si = find_seg<false>(P, w);
			sg = P.segs[si];
		}
		while (w - sg.win_base >= sg.nwin) {  // also skips segments without windows
			++si;
			sg = P.segs[si];
		}
		const uint64_t k = w - sg.win_base;
		uint64_t fu;
		const uint64_t cnt = pgt_window_units(P.g, sg, k, &fu);
		uint32_t sp = 0, ep = 0;
		if (lane == 0 && pos && P.mode != PGT_MODE_BP) {
			uint64_t fs;
			const uint32_t nsites = pgt_window_sites(P.g, sg, k, &fs);
			sp = __ldg(pos + (sg.site_base + fs - P.site_origin));
			ep = __ldg(pos + (sg.site_base + fs + nsites - 1 - P.site_origin));
		}
		const typename Stat::Acc* up = units + (sg.unit_base + fu - units_base);
		PGT_CHECK(sg.unit_base + fu >= P.unit_lo && sg.unit_base + fu + cnt <= P.unit_hi && w >= P.win_lo && w < P.win_hi);
		typename Stat::Acc acc = Stat::zero();
		for (uint64_t x = lane; x < cnt; x += 128u) {  // four partials per lane in flight; added in index order
			typename Stat::Acc v[4];
#pragma unroll
			for (int q = 0; q < 4; ++q)
				if (x + 32u * q < cnt) v[q] = up[x + 32u * q];
#pragma unroll
			for (int q = 0; q < 4; ++q)
				if (x + 32u * q < cnt) Stat::add(acc, v[q]);
		}
		acc = warp_butterfly<Stat>(acc);
		if (lane == 0) emit_window<Stat>(P, sg, w, k, acc, pos, out, true, sp, ep);
	}
}

// Fine windows (<= 32 units each): one THREAD per window, outputs written coalesced.  The value is
// bit-identical to k_windows: leaf i is (+0.0 + unit i) for i < cnt and +0.0 beyond, combined in
// the butterfly's order V(i, s) = V(i, 2s) + V(i + s, 2s), evaluated depth-first so only log2(P2)
// partials are live.
template <class Stat, int P2, int S>
struct SmallTree {
	static __device__ __forceinline__ typename Stat::Acc eval(const typename Stat::Acc* __restrict__ up, uint32_t i, uint32_t cnt) {
		typename Stat::Acc a = SmallTree<Stat, P2, S * 2>::eval(up, i, cnt);
		if (i + S < cnt) Stat::add(a, SmallTree<Stat, P2, S * 2>::eval(up, i + S, cnt));  // adding the +0.0 subtree is the identity
		return a;
	}
};
template <class Stat, int P2>
struct SmallTree<Stat, P2, P2> {
	static __device__ __forceinline__ typename Stat::Acc eval(const typename Stat::Acc* __restrict__ up, uint32_t i, uint32_t cnt) {
		typename Stat::Acc a = Stat::zero();
		if (i < cnt) Stat::add(a, up[i]);
		return a;
	}
};

template <class Stat>
__device__ __forceinline__ void emit_window(const DevPlan& P, const pgt_seg& sg, uint64_t w, uint64_t k, const typename Stat::Acc& acc,
                                            const uint32_t* __restrict__ pos, const pgt_windows& out, bool have_edges, uint32_t edge_start,
                                            uint32_t edge_end) {
	const uint64_t o = w - P.win_lo;
	uint64_t fs;
	const uint32_t nsites = pgt_window_sites(P.g, sg, k, &fs);
	const uint64_t first = sg.site_base + fs, last = first + nsites - 1;
	const uint32_t label = find_contig(P.off, sg.first_contig, sg.ncontig, last);
	if (out.label) out.label[o] = label;
	if (out.nsites) out.nsites[o] = nsites;
	uint32_t sp = 0, ep = 0;
	bool have = false;
	if (P.mode == PGT_MODE_BP) {
		// dxyWindow.cpp:190 prints the bp position of the first / last buffer entry
		const uint32_t cf = find_contig(P.off, sg.first_contig, sg.ncontig, first);
		sp = (uint32_t)(first - P.off[cf]) + 1u;
		ep = (uint32_t)(last - P.off[label]) + 1u;
		have = true;
	} else if (pos) {
		PGT_CHECK(first >= P.site_origin && last - P.site_origin < P.col_elems);
		sp = have_edges ? edge_start : __ldg(pos + (first - P.site_origin));  // the caller may have fetched them early
		ep = have_edges ? edge_end : __ldg(pos + (last - P.site_origin));
		have = true;
	}
	if (have) {
		if (out.start_pos) out.start_pos[o] = sp;
		if (out.end_pos) out.end_pos[o] = ep;
		if (out.mid_pos) out.mid_pos[o] = (sp + ep) / 2u;  // uint32 arithmetic, fstWindow.cpp:73
	}
	Stat::emit(out, o, acc);
}

template <class Stat, int P2>
__global__ void __launch_bounds__(256) k_windows_small(DevPlan P, const typename Stat::Acc* __restrict__ units, uint64_t units_base,
                                                        const uint32_t* __restrict__ pos, pgt_windows out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.win_base = 0;
	sg.nwin = 0;
	for (uint64_t w = P.win_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < P.win_hi; w += stride) {
		if (si == 0xffffffffu || w - sg.win_base >= sg.nwin) {
			si = find_seg<false>(P, w);
			sg = P.segs[si];
		}
		const uint64_t k = w - sg.win_base;
		uint64_t fu;
		const uint32_t cnt = (uint32_t)pgt_window_units(P.g, sg, k, &fu);
		PGT_CHECK(cnt <= (uint32_t)P2 && sg.unit_base + fu >= P.unit_lo && sg.unit_base + fu + cnt <= P.unit_hi);
		const typename Stat::Acc acc = SmallTree<Stat, P2, 1>::eval(units + (sg.unit_base + fu - units_base), 0u, cnt);
		emit_window<Stat>(P, sg, w, k, acc, pos, out);
	}
}

// W = S = 1 (the tools' default arguments): every window is one site, so the window table is an
// elementwise map of the columns; level 1 is skipped and the per-site statistic is evaluated here.
// Output-bound (36-76 bytes of rows per 1-41 bytes of site): a thread takes four windows per turn and issues
// their column loads together before the first row is stored; the label comes from a cached contig range
// instead of a binary search per window.
template <class Stat>
__global__ void __launch_bounds__(256) k_windows_persite(DevPlan P, Cols cols, pgt_windows out) {
	constexpr int U = 4;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.win_base = 0;
	sg.nwin = 0;
	uint32_t lc = 0;
	uint64_t lc_lo = 1, lc_hi = 0;  // sites of contig lc: [lc_lo, lc_hi); empty = nothing cached
	for (uint64_t w0 = P.win_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < P.win_hi; w0 += U * stride) {
		uint64_t site[U];
		typename Stat::Site v[U];
		uint32_t ps[U];
#pragma unroll
		for (int q = 0; q < U; ++q) {
			const uint64_t w = w0 + q * stride;
			site[q] = ~0ull;
			if (w >= P.win_hi) continue;
			if (si == 0xffffffffu || w - sg.win_base >= sg.nwin) {
				si = find_seg<false>(P, w);
				sg = P.segs[si];
			}
			site[q] = sg.site_base + (w - sg.win_base);  // window k of a segment is its site k
		}
#pragma unroll
		for (int q = 0; q < U; ++q) {
			if (site[q] == ~0ull) continue;
			PGT_CHECK(site[q] >= P.site_origin && site[q] - P.site_origin < P.col_elems);
			v[q] = Stat::load(cols, site[q] - P.site_origin);
			ps[q] = cols.pos ? __ldg(cols.pos + (site[q] - P.site_origin)) : 0u;
		}
#pragma unroll
		for (int q = 0; q < U; ++q) {
			if (site[q] == ~0ull) continue;
			const uint64_t o = w0 + q * stride - P.win_lo;
			if (site[q] < lc_lo || site[q] >= lc_hi) {
				lc = find_contig(P.off, 0, P.ncontig, site[q]);
				lc_lo = P.off[lc];
				lc_hi = P.off[lc + 1];
			}
			typename Stat::Acc acc = Stat::zero();
			Stat::fold(acc, v[q], cols.minind);
			if (out.label) out.label[o] = lc;
			if (out.nsites) out.nsites[o] = 1u;
			if (cols.pos) {
				if (out.start_pos) out.start_pos[o] = ps[q];
				if (out.end_pos) out.end_pos[o] = ps[q];
				if (out.mid_pos) out.mid_pos[o] = (ps[q] + ps[q]) / 2u;  // uint32 arithmetic, fstWindow.cpp:73
			}
			Stat::emit(out, o, acc);
		}
	}
}

// ----------------------------------------------------------------------------- level 2, scan mode
//
// Fine steps with long windows (e.g. W = 1000, S = 1): summing W/S unit partials per window
// would cost O(n * W / S^2).  Instead the unit array is cut into blocks of B = wunits units
// (aligned to the segment's first unit), and an inclusive prefix scan PRE and suffix scan SUF are
// taken inside every block.  A window covers at most two adjacent blocks, so
//     window = SUF[first unit] + PRE[last unit]          (van Herk / Gil-Werman)
// -- two reads per window, only additions of true partial sums (no subtraction, no cancellation).
// Order inside a block: chunks of 256 units; warp shuffle scan, warp totals, running carry.

template <class Acc>
__device__ __forceinline__ Acc shfl_up_acc(const Acc& v, unsigned delta) {
	static_assert(sizeof(Acc) % 4 == 0, "Acc is made of 32-bit words");
	uint32_t w[sizeof(Acc) / 4];
	memcpy(w, &v, sizeof(Acc));
#pragma unroll
	for (unsigned i = 0; i < sizeof(Acc) / 4; ++i) w[i] = __shfl_up_sync(0xffffffffu, w[i], delta);
	Acc r;
	memcpy(&r, w, sizeof(Acc));
	return r;
}

// one direction of the block scan: items x0 + i (forward) or x1 - 1 - i (backward), i = 0..n-1
template <class Stat, bool BACKWARD>
__device__ __forceinline__ void cta_scan_dir(const typename Stat::Acc* __restrict__ in, typename Stat::Acc* __restrict__ outp, uint64_t x0, uint64_t x1,
                                             typename Stat::Acc* s_wtot, typename Stat::Acc* s_carry) {
	typedef typename Stat::Acc Acc;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint64_t n = x1 - x0;
	if (threadIdx.x == 0) *s_carry = Stat::zero();
	__syncthreads();
	for (uint64_t base = 0; base < n; base += blockDim.x) {
		const uint64_t i = base + threadIdx.x;
		const bool ok = i < n;
		const uint64_t idx = BACKWARD ? (x1 - 1 - i) : (x0 + i);
		Acc v = Stat::zero();
		if (ok) Stat::add(v, in[idx]);
		// inclusive scan inside the warp
#pragma unroll
		for (unsigned d = 1; d < 32; d <<= 1) {
			Acc o = shfl_up_acc(v, d);
			if (lane >= d) {
				Acc t = o;          // earlier items first: t = earlier + v
				Stat::add(t, v);
				v = t;
			}
		}
		if (lane == 31) s_wtot[warp] = v;
		__syncthreads();
		Acc pre = *s_carry;  // everything before this chunk
		for (uint32_t w = 0; w < warp; ++w) Stat::add(pre, s_wtot[w]);
		Stat::add(pre, v);
		if (ok) outp[idx] = pre;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) *s_carry = pre;  // inclusive total through this chunk
		__syncthreads();
	}
}

template <class Stat>
__global__ void __launch_bounds__(256) k_block_scan(DevPlan P, typename Stat::Acc* __restrict__ units, typename Stat::Acc* __restrict__ pre,
                                                     uint64_t units_base, uint64_t blk_lo, uint64_t blk_hi) {
	__shared__ typename Stat::Acc s_wtot[8];
	__shared__ typename Stat::Acc s_carry;
	const uint64_t B = P.g.wunits;
	for (uint64_t gb = blk_lo + blockIdx.x; gb < blk_hi; gb += gridDim.x) {
		const pgt_seg sg = P.segs[find_seg_by_block(P, gb)];
		const uint64_t lb = gb - sg.blk_base;
		uint64_t u0 = sg.unit_base + lb * B;
		uint64_t u1 = sg.unit_base + ((lb + 1) * B < sg.nunits ? (lb + 1) * B : sg.nunits);
		// clip to the units this scan computed (shards): see DESIGN.md, the clipped values are never used
		if (u0 < P.unit_lo) u0 = P.unit_lo;
		if (u1 > P.unit_hi) u1 = P.unit_hi;
		if (u1 <= u0) continue;
		PGT_CHECK(u0 >= units_base && u0 >= P.unit_lo && u1 <= P.unit_hi);
		cta_scan_dir<Stat, false>(units, pre, u0 - units_base, u1 - units_base, s_wtot, &s_carry);  // PRE (reads raw units)
		cta_scan_dir<Stat, true>(units, units, u0 - units_base, u1 - units_base, s_wtot, &s_carry);  // SUF in place
	}
}

template <class Stat>
__global__ void __launch_bounds__(256) k_windows_hgw(DevPlan P, const typename Stat::Acc* __restrict__ suf, const typename Stat::Acc* __restrict__ pre,
                                                      uint64_t units_base, const uint32_t* __restrict__ pos, pgt_windows out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t B = P.g.wunits;
	uint32_t si = 0xffffffffu;
	pgt_seg sg;
	sg.win_base = 0;
	sg.nwin = 0;
	for (uint64_t w = P.win_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < P.win_hi; w += stride) {
		if (si == 0xffffffffu || w - sg.win_base >= sg.nwin) {
			si = find_seg<false>(P, w);
			sg = P.segs[si];
		}
		const uint64_t k = w - sg.win_base;
		uint64_t fu;
		const uint64_t cnt = pgt_window_units(P.g, sg, k, &fu);
		const uint64_t lu = fu + cnt - 1;  // segment-local first / last unit
		const uint64_t gf = sg.unit_base + fu - units_base, gl = sg.unit_base + lu - units_base;
		PGT_CHECK(sg.unit_base + fu >= P.unit_lo && sg.unit_base + lu < P.unit_hi);
		typename Stat::Acc acc = Stat::zero();
		if (fu / B == lu / B) {
			// inside one block: block-aligned start (PRE up to the last unit), or it runs to the block /
			// segment end (SUF from the first unit)
			if (fu % B == 0) Stat::add(acc, pre[gl]);
			else Stat::add(acc, suf[gf]);
		} else {
			Stat::add(acc, suf[gf]);
			Stat::add(acc, pre[gl]);
		}
		emit_window<Stat>(P, sg, w, k, acc, pos, out);
	}
}


// ----------------------------------------------------------------------------- fine steps: sliding tile
//
// W = 1000, S = 1 and its kin: steps so short that a "unit" would be a handful of sites (pgt_geom.h:
// units never span a step), so the two-level scheme degenerates into copying every site into the unit
// array and scanning that array twice in global memory (~3x the compulsory traffic).  Here the windows
// are formed straight from the sites, in shared memory, and no unit array exists:
//
//   * the site axis of a segment is cut into BLOCKS of W sites from the segment origin; window k
//     (first site f = k*S, last site l) touches block m = f div W and at most block m + 1, so
//         window = SUF_m[f mod W] (+ PRE_{m+1}[l mod W] when l is in block m + 1)
//     with PRE / SUF the inclusive prefix / suffix sums inside a block (van Herk / Gil-Werman): only
//     additions of true partial sums, no differences, no cancellation;
//   * persistent CTAs take CHUNKS of consecutive windows; a chunk walks its blocks in order, G blocks per
//     step: the producer warp stages the step's slice of every column (positions included) with bulk
//     async copies (the ring of k_units_tiled); every block is scanned by its own team of WPB warps
//     (thread q of a team owns elements [q*E, (q+1)*E) of its block: forward total, two warp shuffle scans,
//     the team's warp totals through shared memory, then PRE running forward from the thread's base and
//     SUF running backward from it); the windows that start in the blocks of the step -- all but its last
//     block, plus the last block of the step before -- are emitted from SUF and PRE in shared memory with
//     coalesced row stores.  SUF and the positions are double-buffered by step parity, so a step has two
//     CTA-wide barriers.  Every site is read once per chunk; neighbouring chunks share W - S sites.
//     (E, WPB, G) are chosen per W so that a step covers as many sites as the 7 consumer warps can own:
//     W = 1000: E = 5, one block of 7 warps; W = 256: E = 4, three blocks of 2 warps.
//
// Summation order: a pure function of (W, S) and the position of a site inside its block -- not of
// chunks, CTAs, stages or shards (a chunk that starts or ends inside a block simply has the sites
// outside its windows absent, and those only ever enter prefix / suffix values no window of the chunk
// reads), so results are bit-identical for any GPU count.  Reference semantics: fstWindow.cpp:80-99.

static constexpr int kSlideWarps = 7;                          // consumer warps (+ the producer warp = 256 threads: 128 registers each at two CTAs per SM)
static constexpr int kSlideConsumers = kSlideWarps * 32;
static constexpr int kSlideThreads = kSlideConsumers + 32;     // + the producer warp (the last one)
static constexpr uint32_t kSlideMaxE = 6;                      // W <= 1344 (and the fused step must fit shared memory: W <= 1048)
static constexpr uint32_t kSlideWtBytes = 512;                 // warp totals: 7 x Acc (<= 40 B)

// The team shape for a window size: E elements per thread, WPB warps per block, G = 7 / WPB blocks per step --
// the (E, WPB) with the most sites per step, the smaller E on ties.  Part of the summation order.
struct SlideTeam {
	uint32_t E, wpb, G;
};
// (a step holds at most 1024 sites -- G * (W rounded up to 32) -- so that the fused statistic's step fits shared memory)
__host__ __device__ inline SlideTeam slide_team(uint32_t W) {
	SlideTeam best{1, kSlideWarps, 0};
	const uint32_t wp = (W + 31u) / 32u * 32u;
	for (uint32_t E = 1; E <= kSlideMaxE; ++E) {
		const uint32_t tpb = (W + E - 1) / E;
		const uint32_t wpb = (tpb + 31) / 32;
		if (wpb > (uint32_t)kSlideWarps) continue;
		uint32_t G = (uint32_t)kSlideWarps / wpb;
		if (G > 1024u / wp) G = 1024u / wp;
		if (G < 1) G = 1;
		if (G > best.G) best = SlideTeam{E, wpb, G};
	}
	return best;
}

struct SlideCfg {
	uint32_t E, wpb, G;      // slide_team(W)
	uint32_t wp;             // stride of one block's arrays in shared memory (W rounded up to 32)
	uint32_t sf_off, pr_off, pos_off, stage_off;  // byte offsets inside the dynamic shared memory
	uint32_t pos_col;        // staging index of the position column, 0xffffffff = positions not wanted
	uint64_t chunk_windows;  // windows per chunk
	uint64_t nchunks;
};

template <class Acc>
__device__ __forceinline__ Acc shfl_down_acc(const Acc& v, unsigned delta) {
	static_assert(sizeof(Acc) % 4 == 0, "Acc is made of 32-bit words");
	uint32_t w[sizeof(Acc) / 4];
	memcpy(w, &v, sizeof(Acc));
#pragma unroll
	for (unsigned i = 0; i < sizeof(Acc) / 4; ++i) w[i] = __shfl_down_sync(0xffffffffu, w[i], delta);
	Acc r;
	memcpy(&r, w, sizeof(Acc));
	return r;
}

__device__ __forceinline__ void slide_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kSlideConsumers) : "memory"); }

// the part of a chunk that lies in one segment: windows [ka, kb) of segment `sg` (segment-local), reading
// the segment-local sites [lo, hi)
struct SlideRun {
	pgt_seg sg;
	uint32_t si;
	uint64_t ka, kb, lo, hi;
};
// first run of the chunk [wa, wb) / the run after `r`; false when the chunk is exhausted
__device__ __forceinline__ bool slide_run_at(const DevPlan& P, uint64_t w, uint64_t wb, SlideRun& r, bool first) {
	if (w >= wb) return false;
	if (first) {
		r.si = find_seg<false>(P, w);
		r.sg = P.segs[r.si];
	}
	while (w - r.sg.win_base >= r.sg.nwin) {  // also skips segments without windows
		++r.si;
		r.sg = P.segs[r.si];
	}
	r.ka = w - r.sg.win_base;
	r.kb = wb - r.sg.win_base < r.sg.nwin ? wb - r.sg.win_base : r.sg.nwin;
	r.lo = r.ka * P.g.S;
	const uint64_t e = (r.kb - 1) * P.g.S + P.g.W;
	r.hi = e < r.sg.nsites ? e : r.sg.nsites;
	return true;
}

// (register budget: two CTAs per SM for the narrow statistics -- 128 registers -- and one for the fused
// scan, whose step fills the shared memory of an SM anyway)
template <class Stat>
struct SlideMinBlocks {
	static constexpr int value = 2;
};
template <>
struct SlideMinBlocks<FusedStat> {
	static constexpr int value = 1;
};
// MULTI = several blocks per step (G > 1, W <= 512); with one block per step the team arithmetic folds away
template <class Stat, int EMAX, bool MULTI>
__global__ void __launch_bounds__(kSlideThreads, SlideMinBlocks<Stat>::value) k_slide(DevPlan P, TileCfg tc, SlideCfg sc, pgt_windows out) {
	typedef typename Stat::Acc Acc;
	extern __shared__ __align__(128) unsigned char smem[];
	TileCtl* ctl = reinterpret_cast<TileCtl*>(smem);
	Acc* wt = reinterpret_cast<Acc*>(smem + kTileCtlBytes);
	Acc* SfBase = reinterpret_cast<Acc*>(smem + sc.sf_off);           // SUF of the step's blocks: [2][G * wp], by step parity
	Acc* Pr = reinterpret_cast<Acc*>(smem + sc.pr_off);               // PRE of the step's blocks: [G * wp]
	uint32_t* PosBase = reinterpret_cast<uint32_t*>(smem + sc.pos_off);  // positions of the step's blocks: [2][G * wp]
	unsigned char* stages = smem + sc.stage_off;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint64_t W = P.g.W, S = P.g.S;
	const uint32_t Wu = P.g.W, Su = P.g.S, G = MULTI ? sc.G : 1u, gw = G * sc.wp;
	const uint32_t wpb = MULTI ? sc.wpb : (uint32_t)kSlideWarps;
	const bool has_pos = sc.pos_col != 0xffffffffu;

	if (threadIdx.x == 0) {
		for (uint32_t s = 0; s < tc.nstages; ++s) {
			mbar_init(&ctl->full[s], 2);
			mbar_init(&ctl->empty[s], kSlideWarps);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == kSlideWarps) {
		// ------------------------------------------------------------------ producer: one stage per step (G blocks)
		uint32_t it = 0;
		for (uint64_t c = blockIdx.x; c < sc.nchunks; c += gridDim.x) {
			const uint64_t wa = P.win_lo + c * sc.chunk_windows;
			const uint64_t wb = P.win_hi - wa < sc.chunk_windows ? P.win_hi : wa + sc.chunk_windows;
			SlideRun r;
			for (bool ok = slide_run_at(P, wa, wb, r, true); ok; ok = slide_run_at(P, r.sg.win_base + r.kb, wb, r, false)) {
				for (uint64_t b = r.lo / W; b * W < r.hi; b += G, ++it) {
					const uint64_t x0 = b * W > r.lo ? b * W : r.lo;
					const uint64_t x1 = (b + G) * W < r.hi ? (b + G) * W : r.hi;
					const uint32_t stg = it % tc.nstages;
					if (it >= tc.nstages) mbar_wait(&ctl->empty[stg], ((it / tc.nstages) - 1u) & 1u);
					producer_fill_stage(tc, ctl, stages, stg, r.sg.site_base + x0 - P.site_origin, r.sg.site_base + x1 - P.site_origin, lane);
				}
			}
		}
		return;
	}

	// ---------------------------------------------------------------------- consumers
	const uint32_t t = threadIdx.x;
	const uint32_t team = MULTI ? warp / wpb : 0u;  // which block of a step this thread works on (>= G: none)
	const uint32_t tw = warp - team * wpb;          // warp inside its team
	const uint32_t e0 = (tw * 32u + lane) * sc.E;   // first element of the block this thread owns
	const bool member = MULTI ? team < G : true;
	uint32_t it = 0, step = 0;  // `step` selects the halves of the double buffers
	for (uint64_t c = blockIdx.x; c < sc.nchunks; c += gridDim.x) {
		const uint64_t wa = P.win_lo + c * sc.chunk_windows;
		const uint64_t wb = P.win_hi - wa < sc.chunk_windows ? P.win_hi : wa + sc.chunk_windows;
		SlideRun r;
		for (bool ok = slide_run_at(P, wa, wb, r, true); ok; ok = slide_run_at(P, r.sg.win_base + r.kb, wb, r, false)) {
			const uint64_t m_first = r.lo / W, m_last = (r.hi - 1) / W;
			// label = contig of the window's last site: the windows a thread emits go up the axis, so a cursor
			// that only walks forward replaces a binary search per window
			uint32_t lc = r.sg.first_contig;
			uint64_t lc_end = r.sg.ncontig > 1 ? P.off[lc + 1] : ~0ull;  // global site where contig lc ends
			// kc = ceil(b*W / S): first window that starts at or after block b; kept incrementally
			uint64_t kq = m_first * W / S;
			uint32_t krem = (uint32_t)(m_first * W - kq * S);
			uint64_t kc_from = kq + (krem != 0u);  // first window not yet emitted (by block start)
			for (uint64_t b = m_first; b <= m_last; b += G, ++step) {
				// blocks [b, b + gi) arrive; the windows starting in blocks [b - 1, b + gi - 2] leave -- and those of
				// block b + gi - 1 too when it is the run's last block (they end inside it)
				const uint32_t gi = (uint32_t)(m_last + 1 - b < G ? m_last + 1 - b : G);
				Acc* Sf = SfBase + (step & 1u) * gw;
				const Acc* SfPrev = SfBase + ((step & 1u) ^ 1u) * gw + (G - 1u) * sc.wp;  // SUF of block b - 1
				uint32_t* Pos = PosBase + (step & 1u) * gw;
				const uint32_t* PosPrev = PosBase + ((step & 1u) ^ 1u) * gw + (G - 1u) * sc.wp;
				const uint32_t stg = it % tc.nstages;
				mbar_wait(&ctl->full[stg], (it / tc.nstages) & 1u);
				++it;
				Acc leaf[EMAX];
				uint32_t pv[EMAX];
				Acc up = Stat::zero(), dn = Stat::zero();
				{
					const char* cp[kMaxTileCols];
#pragma unroll
					for (int cc = 0; cc < kMaxTileCols; ++cc) cp[cc] = ctl->cp[stg][cc];
					const uint32_t* pstage = has_pos ? reinterpret_cast<const uint32_t*>(ctl->cp[stg][sc.pos_col]) : nullptr;
					const uint64_t x0 = b * W > r.lo ? b * W : r.lo;            // first staged site
					const uint64_t x1 = (b + gi) * W < r.hi ? (b + gi) * W : r.hi;
					// sites of a block outside [x0, x1) are absent (other chunk / beyond the segment): +0 leaves
					const uint32_t s_lo = (uint32_t)(x0 - b * W), s_hi = (uint32_t)(x1 - b * W);  // relative to block b
					const uint32_t base = team * Wu + e0;                                          // this thread's first element, same frame
					Acc tot = Stat::zero();
#pragma unroll
					for (int e = 0; e < EMAX; ++e) {
						leaf[e] = Stat::zero();
						pv[e] = 0u;
						const uint32_t i = base + (uint32_t)e;
						if (member && (uint32_t)e < sc.E && e0 + (uint32_t)e < Wu && i >= s_lo && i < s_hi) {
							PGT_CHECK(i - s_lo < ctl->ns[stg]);
							Stat::fold(leaf[e], Stat::load_tile(cp, i - s_lo), tc.minind);
							if (has_pos) pv[e] = pstage[i - s_lo];
						}
						Stat::add(tot, leaf[e]);
					}
					__syncwarp();
					if (lane == 0) mbar_arrive(&ctl->empty[stg]);  // the step's sites now live in registers
					up = tot;
					dn = tot;
#pragma unroll
					for (unsigned d = 1; d < 32; d <<= 1) {
						Acc o = shfl_up_acc(up, d);
						if (lane >= d) {  // earlier threads first
							Stat::add(o, up);
							up = o;
						}
						Acc q = shfl_down_acc(dn, d);
						if (lane + d < 32u) {  // later threads first
							Stat::add(q, dn);
							dn = q;
						}
					}
					if (lane == 31u) wt[warp] = up;  // the warp's total, forward order
				}
				slide_bar();  // warp totals visible; every thread has left the emit phase of the step before
				{
					Acc bpre = Stat::zero(), bsuf = Stat::zero();
#pragma unroll
					for (int w2 = 0; w2 < kSlideWarps - 1; ++w2)  // the team's warps before this one, in order
						if ((uint32_t)w2 < tw && member) Stat::add(bpre, wt[team * wpb + w2]);
#pragma unroll
					for (int w2 = kSlideWarps - 1; w2 > 0; --w2)  // the team's warps after it, last first
						if ((uint32_t)w2 > tw && (uint32_t)w2 < wpb && member) Stat::add(bsuf, wt[team * wpb + w2]);
					Acc xu = shfl_up_acc(up, 1), xd = shfl_down_acc(dn, 1);
					if (lane == 0) xu = Stat::zero();
					if (lane == 31u) xd = Stat::zero();
					Stat::add(bpre, xu);
					Stat::add(bsuf, xd);
					const uint32_t o0 = team * sc.wp + e0;
#pragma unroll
					for (int e = 0; e < EMAX; ++e) {  // PRE: running forward from everything before this thread
						if (member && (uint32_t)e < sc.E && e0 + (uint32_t)e < Wu) {
							Stat::add(bpre, leaf[e]);
							Pr[o0 + e] = bpre;
							if (has_pos) Pos[o0 + e] = pv[e];
						}
					}
#pragma unroll
					for (int e = EMAX - 1; e >= 0; --e) {  // SUF: running backward from everything after it
						if (member && (uint32_t)e < sc.E && e0 + (uint32_t)e < Wu) {
							Stat::add(bsuf, leaf[e]);
							Sf[o0 + e] = bsuf;
						}
					}
				}
				slide_bar();  // PRE, SUF and the positions of the step's blocks complete
				{
					// windows that start in blocks [eb_lo, eb_hi]; everything below is relative to the start of block b - 1
					const uint64_t eb_lo = b > m_first ? b - 1 : b;
					const uint64_t eb_hi = (b + gi - 1 == m_last) ? m_last : b + gi - 2;  // may be eb_lo - 1: nothing to emit yet
					// ceil((eb_hi + 1) * W / S): the quotient / remainder of (blocks emitted so far) * W / S stepped forward
					uint64_t kc_to = kc_from;
					if (eb_hi + 1 > eb_lo) {
						const uint32_t nb = (uint32_t)(eb_hi + 1 - eb_lo);
						kq += (uint64_t)nb * P.g.q;
						krem += nb * P.g.r;
						while (krem >= Su) {
							krem -= Su;
							++kq;
						}
						kc_to = kq + (krem != 0u);
					}
					const uint64_t k_lo = kc_from > r.ka ? kc_from : r.ka, k_hi = kc_to < r.kb ? kc_to : r.kb;
					if (eb_hi + 1 > eb_lo && k_hi > k_lo) {
						const uint32_t cnt = (uint32_t)(k_hi - k_lo);
						const uint32_t rel0 = (uint32_t)(k_lo * S + W - b * W);  // first window's start, relative to block b - 1
						const uint64_t left = r.sg.nsites + W - b * W;            // sites of the segment from block b - 1 on
						const uint32_t last_rel = (uint32_t)(left < (uint64_t)(G + 2u) * Wu ? left : (uint64_t)(G + 2u) * Wu) - 1u;
						const uint64_t obase = r.sg.win_base + k_lo - P.win_lo;
						const uint64_t gbase = r.sg.site_base + b * W - W;  // global site of relative position 0
						for (uint32_t i = t; i < cnt; i += kSlideConsumers) {
							const uint32_t rel = rel0 + i * Su;
							uint32_t blk = 0, brel = 0;  // blk = rel / W; brel = blk * W
							if (MULTI) {
								while (rel >= brel + Wu) {  // <= G <= 7 steps
									brel += Wu;
									++blk;
								}
							} else if (rel >= Wu) {
								brel = Wu;
								blk = 1u;
							}
							const uint32_t j = rel - brel;  // blk = 0: block b - 1; blk = 1 + g: block g of this step
							uint32_t lr = rel + Wu - 1u;
							if (lr > last_rel) lr = last_rel;
							const bool two = lr >= brel + Wu;  // the window ends in the next block
							const uint32_t jl = two ? lr - brel - Wu : lr - brel;
							PGT_CHECK(j < Wu && jl < Wu && blk <= gi && (!two || blk < gi) && obase + i < P.win_hi - P.win_lo);
							const uint32_t boff = (blk - 1u) * sc.wp;  // (unused when blk == 0)
							Acc acc = blk == 0u ? SfPrev[j] : Sf[boff + j];
							if (two) Stat::add(acc, Pr[blk * sc.wp + jl]);
							const uint64_t o = obase + i;
							const uint64_t glast = gbase + lr;
							while (glast >= lc_end) {
								++lc;
								lc_end = P.off[lc + 1];
							}
							if (out.label) out.label[o] = lc;
							if (out.nsites) out.nsites[o] = lr - rel + 1u;
							if (has_pos) {
								const uint32_t* ps = blk == 0u ? PosPrev : Pos + boff;  // positions of the window's first block
								const uint32_t sp = ps[j];
								const uint32_t ep = two ? Pos[blk * sc.wp + jl] : ps[jl];
								if (out.start_pos) out.start_pos[o] = sp;
								if (out.end_pos) out.end_pos[o] = ep;
								if (out.mid_pos) out.mid_pos[o] = (sp + ep) / 2u;  // uint32 arithmetic, fstWindow.cpp:73
							}
							Stat::emit(out, o, acc);
						}
					}
					if (eb_hi + 1 > eb_lo) kc_from = kc_to;
				}
			}
		}
	}
}

// dxyWindow's global line (dxyWindow.cpp:382-385,429-433) over the unit partials [0, n): one
// block; thread t adds partials t, t+1024, ...; warp butterflies; counts in 64 bit.
template <class Stat>
struct GlobalOf;
template <>
struct GlobalOf<DxyStat> {
	static __device__ __forceinline__ const DxyStat::Acc& get(const DxyStat::Acc& a) { return a; }
};
template <>
struct GlobalOf<FusedStat> {
	static __device__ __forceinline__ const DxyStat::Acc& get(const FusedStat::Acc& a) { return a.dxy; }
};

static constexpr int kGlobalBlocks = 256;  // fixed (part of the summation order of the global line)

__device__ __forceinline__ void block_reduce_global(double d, unsigned long long ne, unsigned long long nk, double* __restrict__ out3) {
	__shared__ double s_d[32];
	__shared__ unsigned long long s_e[32], s_k[32];
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) {
		d = __dadd_rn(d, shfl_xor_f64(d, m));
		ne += __shfl_xor_sync(0xffffffffu, ne, m);
		nk += __shfl_xor_sync(0xffffffffu, nk, m);
	}
	if ((threadIdx.x & 31u) == 0) {
		s_d[threadIdx.x >> 5] = d;
		s_e[threadIdx.x >> 5] = ne;
		s_k[threadIdx.x >> 5] = nk;
	}
	__syncthreads();
	if (threadIdx.x < 32) {
		const uint32_t nw = blockDim.x >> 5;
		d = threadIdx.x < nw ? s_d[threadIdx.x] : 0.0;
		ne = threadIdx.x < nw ? s_e[threadIdx.x] : 0ull;
		nk = threadIdx.x < nw ? s_k[threadIdx.x] : 0ull;
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) {
			d = __dadd_rn(d, shfl_xor_f64(d, m));
			ne += __shfl_xor_sync(0xffffffffu, ne, m);
			nk += __shfl_xor_sync(0xffffffffu, nk, m);
		}
		if (threadIdx.x == 0) {
			out3[0] = d;
			out3[1] = (double)ne;
			out3[2] = (double)nk;
		}
	}
}

// stage 1: block b adds partials b*1024 + t + k*(256*1024) per thread t, then reduces the block
template <class Stat>
__global__ void __launch_bounds__(1024) k_global_partial(const typename Stat::Acc* __restrict__ units, uint64_t n, double* __restrict__ partial3) {
	double d = 0.0;
	unsigned long long ne = 0, nk = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * 1024u + threadIdx.x; i < n; i += (uint64_t)kGlobalBlocks * 1024u) {
		const DxyStat::Acc& a = GlobalOf<Stat>::get(units[i]);
		d = __dadd_rn(d, a.dxy);
		ne += a.neff;
		nk += a.nskip;
	}
	block_reduce_global(d, ne, nk, partial3 + 3 * blockIdx.x);
}
// stage 2: one block over the 256 block partials
__global__ void __launch_bounds__(kGlobalBlocks) k_global_final(const double* __restrict__ partial3, double* __restrict__ global3) {
	block_reduce_global(partial3[3 * threadIdx.x], (unsigned long long)partial3[3 * threadIdx.x + 1],
	                    (unsigned long long)partial3[3 * threadIdx.x + 2], global3);
}

// dxyWindow's global line when no unit array exists (sliding-tile and per-site scans): thread t of
// block b folds sites i0 + b*1024 + t + k*(256*1024) of the owned range [i0, i1) (column element
// indices) in that order, then the block reductions of k_global_partial.  Classification as
// DxyStat::fold (dxyWindow.cpp:179-186).
__global__ void __launch_bounds__(1024) k_global_sites(Cols cols, uint64_t i0, uint64_t i1, double* __restrict__ partial3) {
	double d = 0.0;
	unsigned long long ne = 0, nk = 0;
	constexpr int U = 4;  // four sites' column loads in flight per thread; folded in index order
	const uint64_t stride = (uint64_t)kGlobalBlocks * 1024u;
	for (uint64_t i = i0 + (uint64_t)blockIdx.x * 1024u + threadIdx.x; i < i1; i += U * stride) {
		DxyStat::Site v[U];
#pragma unroll
		for (int q = 0; q < U; ++q)
			if (i + q * stride < i1) v[q] = DxyStat::load(cols, i + q * stride);
#pragma unroll
		for (int q = 0; q < U; ++q) {
			if (i + q * stride >= i1) break;
			DxyStat::Acc a = DxyStat::zero();
			DxyStat::fold(a, v[q], cols.minind);
			d = __dadd_rn(d, a.dxy);
			ne += a.neff;
			nk += a.nskip;
		}
	}
	block_reduce_global(d, ne, nk, partial3 + 3 * blockIdx.x);
}

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
	o += align_up((size_t)kGlobalBlocks * 3 * sizeof(double), 256);
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
	if (plan->mode == PGT_MODE_BP) o += align_up((size_t)(nunits + 1) * sizeof(uint64_t), 256);
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
	else if (key && strcmp(key, "stages") == 0 && value >= 2 && value <= kTileMaxStages) g_tune_stages = value;
	else if (key && strcmp(key, "stage_kb") == 0 && value >= 8 && value <= 110) g_tune_stage_kb = value;
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
		const uint64_t want = (nwin + 7) / 8;
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
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, g3);
	g_launches += 2;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}
template <>
int launch_global<FusedStat>(const FusedStat::Acc* units, uint64_t n, double* scratch, double* g3, cudaStream_t st) {
	k_global_partial<FusedStat><<<kGlobalBlocks, 1024, 0, st>>>(units, n, scratch);
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, g3);
	g_launches += 2;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// Sliding-tile scan of the windows [P.win_lo, P.win_hi): the columns' element 0 is site P.site_origin and
// they hold `valid_elems` elements.  `pos` may be NULL (host mode gathers positions on the host).
template <class Stat>
static int launch_slide(const pgt_plan* plan, pgt_stat stat, const DevPlan& P, const Cols& cols, uint64_t valid_elems, const uint32_t* pos,
                        const pgt_windows& out, cudaStream_t st) {
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
	// chunks: ~4 per resident CTA for balance, but long enough (>= 32 steps of windows) that the W - S sites
	// shared with the next chunk stay a few percent of what a chunk reads
	const uint64_t wpb = ((uint64_t)sh.G * plan->g.W + plan->g.S - 1) / plan->g.S;  // windows starting in one step
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
	sc.chunk_windows = std::max<uint64_t>(32 * wpb, (nwin + slots * 4 - 1) / (slots * 4));
	sc.nchunks = (nwin + sc.chunk_windows - 1) / sc.chunk_windows;
	const unsigned grid = (unsigned)std::min<uint64_t>(sc.nchunks, slots);
	{
		ProfScope prof(0, st);
		kern<<<grid, kSlideThreads, sh.smem, st>>>(P, tc, sc, out);
	}
	g_launches++;
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

// dxyWindow's global line straight from the columns (scans without a unit array)
static int launch_global_sites(const Cols& cols, uint64_t i0, uint64_t i1, double* scratch, double* g3, cudaStream_t st) {
	k_global_sites<<<kGlobalBlocks, 1024, 0, st>>>(cols, i0, i1, scratch);
	k_global_final<<<1, kGlobalBlocks, 0, st>>>(scratch, g3);
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
	uint64_t* bounds = bp ? (uint64_t*)(ws + L.bounds_off) : nullptr;
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
			if (nwin && L.persite) {
				const uint64_t want = (nwin + 255) / 256, cap = (uint64_t)num_sms() * 8;
				ProfScope prof(1, st);
				k_windows_persite<Stat><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(P, C, *out);
				g_launches++;
				PGT_CUDA(cudaGetLastError());
			} else if (nwin) {
				uint64_t last;
				pgt_plan_window(plan, L.w_hi - 1, nullptr, &last, nullptr);
				PGT_TRY(launch_slide<Stat>(plan, stat, P, C, last + 1 - L.origin, cols->pos, *out, st));
			}
			if (want_global) {
				if (L.gsite_lo < L.origin) return pgt_set_error(PGT_ERR_ARGS, "site_origin lies after the first site this scan must read");
				PGT_TRY(launch_global_sites(C, L.gsite_lo - L.origin, L.gsite_hi - L.origin, (double*)(ws + L.gpart_off), out->dxy_global, st));
			}
			return PGT_OK;
		}
		if (bp && nunits) PGT_TRY(launch_bounds_kernel(P, cols->pos, ndata, bounds, st));
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
				const uint64_t want = (wb - w + 255) / 256, cap = (uint64_t)num_sms() * 8;
				ProfScope prof(1, st);
				k_windows_persite<Stat><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(Ps, C, o);
				g_launches++;
				PGT_CUDA(cudaGetLastError());
			} else {
				PGT_TRY(launch_slide<Stat>(plan, stat, Ps, C, ns, nullptr, o, st));
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
