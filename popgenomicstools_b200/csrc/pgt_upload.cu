// pgt_upload.cu -- streaming host -> device upload through a persistent ring of pinned slots.
//
// The reference streams its input (read a line, maybe flush a window, append: /root/reference/fstWindow.cpp:123-146).
// The drop-in tools parse on all cores into ordinary (pageable) arrays, and a pageable cudaMemcpy runs at ~11 GB/s
// on these hosts against ~54 GB/s from page-locked memory (tools/probe_pageable.py); page-locking whole columns is
// no way out either (pinning runs at 1-3 GB/s).  The uploader owns a small ring of pinned slots that lives as long
// as the caller wants (a tool keeps it for the whole run; a service would keep it for the process): a pool of copy
// threads moves queued ranges -- host memory, or a byte range of a FILE read with pread, so a `.pgtc` cache never
// goes through a page-cache mapping -- into free slots and sends them on with cudaMemcpyAsync; the memcpy / pread of
// one slot overlaps the DMA of the others.  put() only enqueues, so a tool uploads the rows its parser threads have
// finished while they are still parsing the rest; the scan then runs on device-resident columns (PGT_MEM_DEVICE).
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pgt_internal.h"

struct pgt_uploader {
	struct Piece {
		char* dst;          // device
		const char* src;    // host memory, or nullptr for a file range
		int fd;
		uint64_t file_off;
		size_t bytes;
	};
	int device = 0;
	char* pinned = nullptr;
	bool own_pinned = false;
	size_t slot_bytes = 0;
	uint32_t nslots = 0;
	std::vector<cudaEvent_t> slot_done;
	std::vector<cudaStream_t> streams;  // one per worker
	std::vector<std::thread> workers;
	std::mutex mu;
	std::condition_variable cv_work, cv_idle;
	std::deque<Piece> queue;
	uint64_t queued = 0, finished = 0;  // pieces
	bool stop = false;
	int rc = PGT_OK;
	std::string err;
	std::atomic<uint64_t> bytes_sent{0};

	void fail(int code, const std::string& msg) {
		std::lock_guard<std::mutex> lk(mu);
		if (rc == PGT_OK) {
			rc = code;
			err = msg;
		}
	}

	void work(uint32_t wi, uint32_t nworkers) {
		if (cudaSetDevice(device) != cudaSuccess) fail(PGT_ERR_CUDA, "cudaSetDevice failed in an upload thread");
		uint32_t turn = 0;
		std::vector<bool> used(nslots, false);
		for (;;) {
			Piece p;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv_work.wait(lk, [&] { return stop || !queue.empty(); });
				if (queue.empty()) return;  // stop requested and nothing left
				p = queue.front();
				queue.pop_front();
			}
			// this worker's slots: wi, wi + nworkers, ...
			const uint32_t mine = (nslots - wi + nworkers - 1) / nworkers;
			const uint32_t s = wi + (turn++ % mine) * nworkers;
			bool ok = rc == PGT_OK;
			if (ok && used[s]) ok = cudaEventSynchronize(slot_done[s]) == cudaSuccess;  // the slot's previous DMA has finished
			char* slot = pinned + (size_t)s * slot_bytes;
			if (ok) {
				if (p.src) {
					memcpy(slot, p.src, p.bytes);
				} else {
					size_t got = 0;
					while (got < p.bytes) {
						const ssize_t k = pread(p.fd, slot + got, p.bytes - got, (off_t)(p.file_off + got));
						if (k <= 0) {
							fail(PGT_ERR_INPUT, "pgt_uploader: short read from the input file");
							ok = false;
							break;
						}
						got += (size_t)k;
					}
				}
			}
			if (ok) {
				cudaError_t e = cudaMemcpyAsync(p.dst, slot, p.bytes, cudaMemcpyHostToDevice, streams[wi]);
				if (e == cudaSuccess) e = cudaEventRecord(slot_done[s], streams[wi]);
				if (e != cudaSuccess) fail(PGT_ERR_CUDA, std::string("pgt_uploader: ") + cudaGetErrorString(e));
				else {
					used[s] = true;
					bytes_sent += p.bytes;
				}
			}
			{
				std::lock_guard<std::mutex> lk(mu);
				++finished;
			}
			cv_idle.notify_all();
		}
	}
};

extern "C" size_t pgt_uploader_pinned_bytes(uint32_t nslots, size_t slot_bytes) { return (size_t)nslots * slot_bytes; }

extern "C" int pgt_uploader_create(pgt_uploader** out, void* pinned, size_t pinned_bytes, uint32_t nslots, size_t slot_bytes, uint32_t nthreads) {
	if (!out) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_create: NULL");
	*out = nullptr;
	if (nslots == 0 || slot_bytes < 4096 || nthreads == 0) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_create: need >= 1 slot of >= 4096 bytes and >= 1 thread");
	if (nthreads > nslots) nthreads = nslots;
	if (pinned && pinned_bytes < (size_t)nslots * slot_bytes) return pgt_set_error(PGT_ERR_NOMEM, "pgt_uploader_create: pinned buffer too small, see pgt_uploader_pinned_bytes");
	pgt_uploader* u = new (std::nothrow) pgt_uploader();
	if (!u) return pgt_set_error(PGT_ERR_NOMEM, "pgt_uploader_create: out of memory");
	cudaError_t e = cudaGetDevice(&u->device);
	u->nslots = nslots;
	u->slot_bytes = slot_bytes;
	u->pinned = (char*)pinned;
	if (e == cudaSuccess && !pinned) {
		e = cudaHostAlloc((void**)&u->pinned, (size_t)nslots * slot_bytes, cudaHostAllocPortable);
		u->own_pinned = e == cudaSuccess;
	}
	u->slot_done.assign(nslots, nullptr);
	for (uint32_t s = 0; e == cudaSuccess && s < nslots; ++s) e = cudaEventCreateWithFlags(&u->slot_done[s], cudaEventDisableTiming);
	u->streams.assign(nthreads, nullptr);
	for (uint32_t t = 0; e == cudaSuccess && t < nthreads; ++t) e = cudaStreamCreateWithFlags(&u->streams[t], cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		const std::string msg = std::string("pgt_uploader_create: ") + cudaGetErrorString(e);
		pgt_uploader_destroy(u);
		return pgt_set_error(PGT_ERR_CUDA, msg);
	}
	for (uint32_t t = 0; t < nthreads; ++t) u->workers.emplace_back([u, t, nthreads] { u->work(t, nthreads); });
	*out = u;
	return PGT_OK;
}

static int enqueue(pgt_uploader* u, void* dev_dst, const void* host_src, int fd, uint64_t file_off, size_t bytes) {
	if (!u || !dev_dst) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_put: NULL");
	{
		std::lock_guard<std::mutex> lk(u->mu);
		if (u->rc != PGT_OK) return pgt_set_error(u->rc, u->err);
		for (size_t o = 0; o < bytes; o += u->slot_bytes) {
			const size_t k = std::min(u->slot_bytes, bytes - o);
			u->queue.push_back(pgt_uploader::Piece{(char*)dev_dst + o, host_src ? (const char*)host_src + o : nullptr, fd, file_off + o, k});
			++u->queued;
		}
	}
	u->cv_work.notify_all();
	return PGT_OK;
}

extern "C" int pgt_uploader_put(pgt_uploader* u, void* dev_dst, const void* host_src, size_t bytes) {
	if (!host_src && bytes) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_put: host_src is NULL");
	return enqueue(u, dev_dst, host_src, -1, 0, bytes);
}

extern "C" int pgt_uploader_put_file(pgt_uploader* u, void* dev_dst, int fd, uint64_t file_offset, size_t bytes) {
	if (fd < 0) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_put_file: bad file descriptor");
	return enqueue(u, dev_dst, nullptr, fd, file_offset, bytes);
}

extern "C" int pgt_uploader_drain(pgt_uploader* u, uint64_t* bytes_sent) {
	if (!u) return pgt_set_error(PGT_ERR_ARGS, "pgt_uploader_drain: NULL");
	{
		std::unique_lock<std::mutex> lk(u->mu);
		u->cv_idle.wait(lk, [&] { return u->finished == u->queued; });
	}
	int prev = 0;
	cudaGetDevice(&prev);
	cudaSetDevice(u->device);
	cudaError_t e = cudaSuccess;
	for (cudaStream_t st : u->streams)
		if (st && e == cudaSuccess) e = cudaStreamSynchronize(st);
	cudaSetDevice(prev);
	if (bytes_sent) *bytes_sent = u->bytes_sent.load();
	if (e != cudaSuccess) return pgt_set_error(PGT_ERR_CUDA, std::string("pgt_uploader_drain: ") + cudaGetErrorString(e));
	std::lock_guard<std::mutex> lk(u->mu);
	if (u->rc != PGT_OK) return pgt_set_error(u->rc, u->err);
	return PGT_OK;
}

extern "C" void pgt_uploader_destroy(pgt_uploader* u) {
	if (!u) return;
	{
		std::lock_guard<std::mutex> lk(u->mu);
		u->stop = true;
	}
	u->cv_work.notify_all();
	for (auto& t : u->workers)
		if (t.joinable()) t.join();
	int prev = 0;
	cudaGetDevice(&prev);
	cudaSetDevice(u->device);
	for (cudaStream_t st : u->streams)
		if (st) {
			cudaStreamSynchronize(st);
			cudaStreamDestroy(st);
		}
	for (cudaEvent_t ev : u->slot_done)
		if (ev) cudaEventDestroy(ev);
	if (u->own_pinned && u->pinned) cudaFreeHost(u->pinned);
	cudaSetDevice(prev);
	delete u;
}

extern "C" int pgt_memcpy_to_host(void* host, const void* dev, size_t bytes) {
	if (bytes == 0) return PGT_OK;
	if (!host || !dev) return pgt_set_error(PGT_ERR_ARGS, "pgt_memcpy_to_host: NULL");
	const cudaError_t e = cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
	if (e != cudaSuccess) return pgt_set_error(PGT_ERR_CUDA, std::string("pgt_memcpy_to_host: ") + cudaGetErrorString(e));
	return PGT_OK;
}

extern "C" int pgt_device_mem_info(size_t* free_bytes, size_t* total_bytes) {
	size_t f = 0, t = 0;
	const cudaError_t e = cudaMemGetInfo(&f, &t);
	if (e != cudaSuccess) return pgt_set_error(PGT_ERR_CUDA, std::string("cudaMemGetInfo: ") + cudaGetErrorString(e));
	if (free_bytes) *free_bytes = f;
	if (total_bytes) *total_bytes = t;
	return PGT_OK;
}
