// Stress test of flatland-marl_b200/csrc/expand_pool.h (host threads of the compact wire format): many short jobs back to back,
// every index of every job must run exactly once with that job's own function.  Built and run by tests/test_expand_pool.py.
#include "expand_pool.h"

#include <chrono>
#include <cstdio>
#include <vector>

int main(int argc, char **argv) {
    const int jobs = argc > 1 ? atoi(argv[1]) : 20000, threads = argc > 2 ? atoi(argv[2]) : 8;
    ExpandPool pool;
    pool.set_threads(threads);
    std::vector<std::atomic<int>> hits(4096);
    long long total = 0;
    for (int j = 0; j < jobs; j++) {
        const int n = 1 + (j * 7919) % 200;
        for (int k = 0; k < n; k++) hits[k].store(0);
        const int tag = j;                                   // lives on this stack frame only while the job runs
        std::atomic<int> wrong{0};
        const std::function<void(int)> fn = [&](int k) {
            if (k < 0 || k >= n || tag != j) wrong.fetch_add(1);
            hits[k].fetch_add(1);
        };
        pool.parallel_for(n, fn);
        for (int k = 0; k < n; k++)
            if (hits[k].load() != 1) { std::printf("job %d: index %d ran %d times\n", j, k, hits[k].load()); return 1; }
        if (wrong.load()) { std::printf("job %d: %d calls with foreign arguments\n", j, wrong.load()); return 1; }
        total += n;
        if (j % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(j % 50));   // let workers fall asleep sometimes
    }
    std::printf("ok %d jobs %lld items %d threads\n", jobs, total, pool.threads());
    return 0;
}
