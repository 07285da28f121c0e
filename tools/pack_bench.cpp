// Host packer throughput vs threads (build: g++ -O2 -std=c++17 -pthread tools/pack_bench.cpp -o /tmp/pack_bench)
#include "../telomeric_identifier_b200/csrc/host_pack.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
int main() {
    const size_t n = 512u << 20;
    std::vector<uint8_t> s(n);
    const char *al = "ACGTacgt";
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; s[i] = al[x & 7]; }
    const size_t nw = n / 32;
    // 64-byte aligned outputs (like the pinned staging buffers): the packer then uses non-temporal stores;
    // PACK_BENCH_MISALIGN=1 offsets them by 4 bytes to measure the plain-store path
    const size_t mis = getenv("PACK_BENCH_MISALIGN") ? 1 : 0;
    uint32_t *lo_b = (uint32_t *)aligned_alloc(64, (nw + 16) * 4), *hi_b = (uint32_t *)aligned_alloc(64, (nw + 16) * 4),
             *va_b = (uint32_t *)aligned_alloc(64, (nw + 16) * 4);
    memset(lo_b, 0, (nw + 16) * 4); memset(hi_b, 0, (nw + 16) * 4); memset(va_b, 0, (nw + 16) * 4);
    struct V { uint32_t *p; uint32_t *data() { return p; } } lo{lo_b + mis}, hi{hi_b + mis}, va{va_b + mis};
    printf("avx2=%d avx512bw=%d hw_threads=%u\n", __builtin_cpu_supports("avx2"), __builtin_cpu_supports("avx512bw"),
           std::thread::hardware_concurrency());
    const bool three = getenv("PACK_BENCH_PLANES") && atoi(getenv("PACK_BENCH_PLANES")) == 3;
    for (int T : {1, 2, 4, 8, 16, 32}) {
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    uint32_t f = 0;
                    size_t w0 = nw * t / T, w1 = nw * (t + 1) / T;
                    // PACK_BENCH_PLANES=3: the three-plane form; default: two planes + validity exceptions (what the upload
                    // uses; TIDK_PACK_ISA=avx2 keeps the AVX2 form on an AVX-512 machine)
                    if (three) tidk::pack_words(s.data() + w0 * 32, (w1 - w0) * 32, w1 - w0, lo.data() + w0, hi.data() + w0, va.data() + w0, &f);
                    else {
                        std::vector<uint64_t> exc;
                        tidk::pack_words2(s.data() + w0 * 32, (w1 - w0) * 32, w1 - w0, lo.data() + w0, hi.data() + w0, w0, exc, &f);
                    }
                });
            for (auto &q : th) q.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
        }
        printf("T=%2d  %.2f ms  %.1f GB/s\n", T, best * 1e3, n / best / 1e9);
    }
}
