// Host-side helper: contiguous, ordered chunks of an index range on a few std::threads.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

namespace spand {

constexpr int kMaxHostThreads = 8;

// fn(chunk, begin, end) over [0, n); returns the number of chunks used (<= kMaxHostThreads). Chunk t covers
// [n t / nth, n (t + 1) / nth): concatenating per-chunk outputs in chunk order reproduces the serial order.
template <class F>
int parallel_chunks(size_t n, F fn, size_t min_parallel = 65536) {
    // SPAND_HOST_THREADS caps the thread count (1: serial; the results do not depend on it)
    const char* cap = std::getenv("SPAND_HOST_THREADS");
    const unsigned hw = cap ? (unsigned)std::max(1, std::atoi(cap)) : std::thread::hardware_concurrency();
    const int nth = n < min_parallel ? 1 : (int)std::min<unsigned>(kMaxHostThreads, std::max(1u, hw));
    if (nth == 1) {
        fn(0, (size_t)0, n);
        return 1;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < nth; t++) pool.emplace_back(fn, t, n * t / nth, n * (t + 1) / nth);
    fn(0, (size_t)0, n / nth);
    for (auto& th : pool) th.join();
    return nth;
}

}  // namespace spand
