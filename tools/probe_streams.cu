// How fast can one kernel WRITE K interleaved output columns?  (The per-site scan writes 8-16 columns of rows; a fill of
// one array reaches 7.5 TB/s on this GPU.)  Each thread stores 16 bytes (four u32 rows) to each of K arrays per turn, the
// shape of k_windows_persite4's row stores; variant B stores 32 bytes per array and turn.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/probe_streams.cu -o tools/probe_streams
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K, int WIDE>
__global__ void __launch_bounds__(256) k_write(uint32_t* base, size_t n_per, size_t pitch) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	const size_t per = WIDE ? 8 : 4;
	for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g * per + per <= n_per; g += stride) {
#pragma unroll
		for (int k = 0; k < K; ++k) {
			uint32_t* p = base + (size_t)k * pitch + g * per;
			if (WIDE)
				asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"((uint32_t)g) : "memory");
			else
				*reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)g, k, 2, 3);
		}
	}
}

template <int K, int WIDE>
static void run(uint32_t* buf, size_t total_elems, int sms) {
	const size_t n_per = total_elems / K / 1024 * 1024, pitch = n_per;
	int per_sm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_write<K, WIDE>, 256, 0);
	const int grid = sms * per_sm;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int i = 0; i < 3; ++i) k_write<K, WIDE><<<grid, 256>>>(buf, n_per, pitch);
	cudaEventRecord(e0);
	for (int i = 0; i < 10; ++i) k_write<K, WIDE><<<grid, 256>>>(buf, n_per, pitch);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	ms /= 10;
	printf("K=%2d %s: %.3f ms for %.2f GB = %.0f GB/s (grid %d)\n", K, WIDE ? "32 B/thread/array" : "16 B/thread/array", ms,
	       n_per * K * 4 / 1e9, n_per * K * 4 / ms / 1e6, grid);
}

int main() {
	const size_t total = (size_t)1 << 30;  // 4 GiB of u32
	uint32_t* buf;
	if (cudaMalloc(&buf, total * 4) != cudaSuccess) return 1;
	cudaDeviceProp pr;
	cudaGetDeviceProperties(&pr, 0);
	const int sms = pr.multiProcessorCount;
	run<1, 0>(buf, total, sms);
	run<2, 0>(buf, total, sms);
	run<4, 0>(buf, total, sms);
	run<8, 0>(buf, total, sms);
	run<9, 0>(buf, total, sms);
	run<16, 0>(buf, total, sms);
	run<1, 1>(buf, total, sms);
	run<8, 1>(buf, total, sms);
	run<16, 1>(buf, total, sms);
	return cudaDeviceSynchronize() != cudaSuccess;
}
