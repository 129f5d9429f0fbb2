// copypool_check.cpp -- TEST: CopyPool (popgenomicstools_b200/csrc/pgt_hostcopy.h), the threaded memcpy of the
// experimental pinned staging ring: every size class (serial path, aligned cuts, ragged tails), unaligned sources,
// 0..7 workers, back-to-back jobs.  Built by tests/test_hostcopy_cpu.py, also with -fsanitize=thread.
#include <cstdio>
#include <cstdlib>

#include "pgt_hostcopy.h"

int main() {
	std::vector<char> src((40u << 20) + 64), dst((40u << 20) + 64);
	for (size_t i = 0; i < src.size(); ++i) src[i] = (char)((i * 2654435761u) >> 13);
	const size_t sizes[] = {0, 1, 4095, 4096, 4097, (1u << 20) - 1, 1u << 20, (1u << 20) + 1, 5000000, 16u << 20, (40u << 20) - 7};
	unsigned long long jobs = 0;
	for (unsigned nw : {0u, 1u, 2u, 3u, 7u}) {
		CopyPool pool;
		pool.start(nw);
		for (int rep = 0; rep < 6; ++rep)
			for (size_t n : sizes) {
				const size_t so = (size_t)(rep % 3), dof = (size_t)(rep % 2) * 5;
				memset(dst.data(), 0x5a, n + dof + 8);
				pool.copy(dst.data() + dof, src.data() + so, n);
				++jobs;
				if (memcmp(dst.data() + dof, src.data() + so, n) != 0 || dst[dof + n] != 0x5a || (dof && dst[dof - 1] != 0x5a)) {
					printf("FAIL workers=%u n=%zu rep=%d\n", nw, n, rep);
					return 1;
				}
			}
	}
	printf("%llu copies ok\n", jobs);
	return 0;
}
