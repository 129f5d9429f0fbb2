// pgt_hostcopy.h -- a memcpy spread over a few persistent host threads (used by the experimental pinned staging
// ring of PGT_MEM_HOST, pgt_scan.cu PinnedRing).  Header-only so that tests/integration/copypool_check.cpp can
// exercise it on the CPU (also under -fsanitize=thread).
#ifndef PGT_HOSTCOPY_H
#define PGT_HOSTCOPY_H

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// copy(): the calling thread and `nworkers` workers each copy one 4096-byte-aligned part; returns when all bytes
// are in place.  One copy at a time per pool.
struct CopyPool {
	std::vector<std::thread> th;
	std::mutex m;
	std::condition_variable cv_go, cv_done;
	uint64_t gen = 0;
	unsigned left = 0;
	bool stop = false;
	char* dst = nullptr;
	const char* src = nullptr;
	size_t bytes = 0;
	static size_t cut(size_t n, unsigned p, unsigned parts) { return p >= parts ? n : (size_t)((unsigned __int128)n * p / parts) & ~(size_t)4095; }
	void start(unsigned nworkers) {
		for (unsigned i = 0; i < nworkers; ++i)
			th.emplace_back([this, i, nworkers]() {
				uint64_t seen = 0;
				std::unique_lock<std::mutex> lk(m);
				for (;;) {
					cv_go.wait(lk, [&] { return stop || gen != seen; });
					if (stop) return;
					seen = gen;
					char* d = dst;
					const char* sp = src;
					const size_t n = bytes;
					lk.unlock();
					const size_t lo = cut(n, i + 1, nworkers + 1), hi = cut(n, i + 2, nworkers + 1);
					if (hi > lo) memcpy(d + lo, sp + lo, hi - lo);
					lk.lock();
					if (--left == 0) cv_done.notify_one();
				}
			});
	}
	void copy(char* d, const char* sp, size_t n) {  // returns when all n bytes are in place
		if (th.empty() || n < (1u << 20)) {
			memcpy(d, sp, n);
			return;
		}
		const unsigned parts = (unsigned)th.size() + 1;
		{
			std::lock_guard<std::mutex> lk(m);
			dst = d;
			src = sp;
			bytes = n;
			left = (unsigned)th.size();
			++gen;
		}
		cv_go.notify_all();
		const size_t hi = cut(n, 1, parts);
		if (hi) memcpy(d, sp, hi);  // the calling thread takes the first part
		std::unique_lock<std::mutex> lk(m);
		cv_done.wait(lk, [&] { return left == 0; });
	}
	~CopyPool() {
		{
			std::lock_guard<std::mutex> lk(m);
			stop = true;
		}
		cv_go.notify_all();
		for (auto& t : th) t.join();
	}
};

#endif
