// pgt_synth.cu -- on-device counter-based synthetic site generator.
// Same functions as the CPU twin (include/pgt_synth.h), so device columns, oracle arrays and
// the 6-decimal text fed to the reference binaries hold bit-identical values.  Column formats
// follow /root/reference/fstWindow.cpp:17-21, hetWindow.cpp:18 and dxyWindow.cpp:24-32.
#include <cuda_runtime.h>

#include <string>

#include "../../include/pgt_synth.h"
#include "pgt_internal.h"


#define PGT_CUDA(call)                                                                                   \
	do {                                                                                                 \
		cudaError_t e__ = (call);                                                                        \
		if (e__ != cudaSuccess)                                                                          \
			return pgt_set_error(PGT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
	} while (0)

static unsigned synth_grid(uint64_t n, int threads) {
	uint64_t want = (n + threads - 1) / threads;
	uint64_t cap = 148ull * 16ull;
	return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

__global__ void k_synth_fst(uint64_t seed, uint64_t site0, uint64_t n, double* __restrict__ a, double* __restrict__ b) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		a[i] = pgt_synth_fst_a(seed, site0 + i);
		b[i] = pgt_synth_fst_b(seed, site0 + i);
	}
}

__global__ void k_synth_het(uint64_t seed, uint64_t site0, uint64_t n, int8_t* __restrict__ g) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] = (int8_t)pgt_synth_het_g(seed, site0 + i);
}

__global__ void k_synth_dxy(uint64_t seed, uint64_t site0, uint64_t n, double* __restrict__ f1, double* __restrict__ f2,
                            int32_t* __restrict__ n1, int32_t* __restrict__ n2) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		f1[i] = pgt_synth_dxy_f1(seed, site0 + i);
		f2[i] = pgt_synth_dxy_f2(seed, site0 + i);
		n1[i] = pgt_synth_dxy_n1(seed, site0 + i);
		n2[i] = pgt_synth_dxy_n2(seed, site0 + i);
	}
}

__global__ void k_synth_pos(uint64_t seed, uint64_t site0, uint64_t n, const uint64_t* __restrict__ off, uint32_t ncontig,
                            uint32_t density, uint32_t* __restrict__ pos) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t s = site0 + i;
		uint32_t lo = 0, hi = ncontig;  // contig with off[c] <= s < off[c+1]
		while (hi - lo > 1) {
			uint32_t mid = lo + ((hi - lo) >> 1);
			if (off[mid] <= s) lo = mid;
			else hi = mid;
		}
		pos[i] = pgt_synth_pos(seed, s, s - off[lo], density);
	}
}

extern "C" int pgt_synth_fst(uint64_t seed, uint64_t site0, uint64_t n, double* a, double* b, void* stream) {
	if (n == 0) return PGT_OK;
	if (!a || !b) return pgt_set_error(PGT_ERR_ARGS, "pgt_synth_fst: NULL column");
	k_synth_fst<<<synth_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, site0, n, a, b);
	pgt_count_launch();
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

extern "C" int pgt_synth_het(uint64_t seed, uint64_t site0, uint64_t n, int8_t* g, void* stream) {
	if (n == 0) return PGT_OK;
	if (!g) return pgt_set_error(PGT_ERR_ARGS, "pgt_synth_het: NULL column");
	k_synth_het<<<synth_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, site0, n, g);
	pgt_count_launch();
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

extern "C" int pgt_synth_dxy(uint64_t seed, uint64_t site0, uint64_t n, double* f1, double* f2, int32_t* n1, int32_t* n2, void* stream) {
	if (n == 0) return PGT_OK;
	if (!f1 || !f2 || !n1 || !n2) return pgt_set_error(PGT_ERR_ARGS, "pgt_synth_dxy: NULL column");
	k_synth_dxy<<<synth_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, site0, n, f1, f2, n1, n2);
	pgt_count_launch();
	PGT_CUDA(cudaGetLastError());
	return PGT_OK;
}

extern "C" int pgt_synth_pos(uint64_t seed, uint64_t site0, uint64_t n, const uint64_t* contig_offsets, uint32_t ncontig,
                             uint32_t density, uint32_t* pos, void* stream) {
	if (n == 0) return PGT_OK;
	if (!pos || !contig_offsets || ncontig == 0) return pgt_set_error(PGT_ERR_ARGS, "pgt_synth_pos: NULL argument");
	cudaStream_t st = (cudaStream_t)stream;
	uint64_t* d_off = nullptr;
	const size_t bytes = (size_t)(ncontig + 1) * sizeof(uint64_t);
	PGT_CUDA(cudaMalloc(&d_off, bytes));
	cudaError_t e = cudaMemcpyAsync(d_off, contig_offsets, bytes, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess) {
		k_synth_pos<<<synth_grid(n, 256), 256, 0, st>>>(seed, site0, n, d_off, ncontig, density, pos);
		pgt_count_launch();
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaFree(d_off);
	if (e != cudaSuccess) return pgt_set_error(PGT_ERR_CUDA, std::string("pgt_synth_pos: ") + cudaGetErrorString(e));
	return PGT_OK;
}
