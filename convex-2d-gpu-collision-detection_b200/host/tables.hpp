// tables.hpp -- the pose / variance tables of generate_dataset (generate_dataset.cu:275-336), generated in parallel.
//
// Upstream fills both tables from ONE default-constructed std::default_random_engine, variances first (5 draws per
// row, also for dimensions whose min == max), then poses (3 draws per row), each draw through a
// std::uniform_real_distribution<float>(min_k, max_k) (generate_dataset.cu:279-300,319-330).  At the defaults that is
// 1.3e8 sequential draws -- 5 s on one core, ten times the GPU work of a 10-batch run.  The engine is libstdc++'s
// minstd_rand0, a Lehmer generator x <- 16807 x mod (2^31 - 1): it can be advanced by k steps with one modular
// exponentiation, so every thread jumps to its first draw and the tables come out bit-identical to the sequential
// loop (host_selftest `tables` compares both on this toolchain; the distribution's arithmetic is libstdc++'s
// generate_canonical<float, 24>: float(x - 1) / 2^31 clamped below 1, then v * (max - min) + min).
#pragma once
#include <cmath>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

namespace tables {

constexpr uint64_t kM = 2147483647ull, kA = 16807ull;

inline uint64_t pow_mod(uint64_t a, uint64_t k) {
    uint64_t r = 1;
    a %= kM;
    while (k) {
        if (k & 1) r = r * a % kM;
        a = a * a % kM;
        k >>= 1;
    }
    return r;
}

// state of minstd_rand0 after `draws` calls from the default seed (1)
inline uint64_t state_after(uint64_t draws) { return pow_mod(kA, draws); }

inline float canonical(uint64_t x) {                      // generate_canonical<float, 24>(minstd_rand0), one call per value
    float r = float(x - 1) / 2147483648.0f;
    if (r >= 1.0f) r = std::nextafter(1.0f, 0.0f);
    return r;
}

inline unsigned default_threads() {
    unsigned t = std::thread::hardware_concurrency();
    return t ? t : 1;
}

// rows[i][k] = uniform(min[k], max[k]) with draw number first_draw + i * dim + k of the default engine
inline void fill_uniform_rows(float* rows, size_t n_rows, int dim, const float* mn, const float* mx, uint64_t first_draw,
                              unsigned threads = 0) {
    if (threads == 0) threads = default_threads();
    if (n_rows < 4096) threads = 1;
    auto work = [=](size_t lo, size_t hi) {
        uint64_t x = state_after(first_draw + (uint64_t)lo * (uint64_t)dim);
        for (size_t i = lo; i < hi; i++)
            for (int k = 0; k < dim; k++) {
                x = x * kA % kM;
                rows[i * dim + k] = canonical(x) * (mx[k] - mn[k]) + mn[k];
            }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; t++) pool.emplace_back(work, n_rows * t / threads, n_rows * (t + 1) / threads);
    work(0, n_rows / threads);
    for (std::thread& t : pool) t.join();
}

// the sequential upstream loop, verbatim in behaviour (used by the self-test as the ground truth)
inline void fill_uniform_rows_std(float* rows, size_t n_rows, int dim, const float* mn, const float* mx, std::default_random_engine& gen) {
    std::vector<std::uniform_real_distribution<float>> u;
    for (int k = 0; k < dim; k++) u.emplace_back(mn[k], mx[k]);
    for (size_t i = 0; i < n_rows; i++)
        for (int k = 0; k < dim; k++) rows[i * dim + k] = u[k](gen);
}

}  // namespace tables
