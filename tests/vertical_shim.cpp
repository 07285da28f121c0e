// Host build of the vertical explore arithmetic (csrc/explore_vertical_core.hpp) for tests/test_vertical_core.py:
// events of one slice computed tile by tile exactly as a lane of explore_vertical_kernel does, and the same events
// straight from the definition (explore.rs:178-233 restated in explore_vertical_core.hpp).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../telomeric_identifier_b200/csrc/explore_vertical_core.hpp"

using namespace tidk;

template <int K>
static void one_k(const uint32_t (&lo)[32], const uint32_t (&hi)[32], uint32_t h0, uint64_t rel0, uint64_t m, bool first, bool last,
                  std::vector<uint64_t> &out) {
    uint32_t E[32];
    const uint32_t phase = (uint32_t)(rel0 % K);
    vt_compute<K>(lo, hi, [&](int i) { return vt_mask_word(K, phase + (uint32_t)i); }, [&](int i, uint32_t e) { E[i] = e; });
    if (first && vt_head_flip(h0, K)) E[2 * K - 1] ^= 1u;
    vt_emit(K, rel0, m, first, last, [&](uint32_t i) { return E[i]; }, [&](uint64_t j) { out.push_back(((uint64_t)K << 56) | j); });
}

extern "C" {

// events (k << 56 | chunk) of slice [beg, beg + m) of seq[0..n); bytes outside seq read as 0.  Returns the count.
uint64_t vt_events(const uint8_t *seq, uint64_t n, uint64_t beg, uint64_t m, uint32_t kmin, uint32_t kmax, uint64_t *out, uint64_t cap) {
    std::vector<uint64_t> ev;
    const uint64_t nt = vt_count(m);
    for (uint64_t t = 0; t < nt; t++) {
        const uint64_t rel0 = (uint64_t)kVStep * t, a0 = beg + rel0;
        uint32_t lo[32], hi[32];
        for (int b = 0; b < 32; b++) {
            lo[b] = hi[b] = 0;
            for (int i = 0; i < 32; i++) {
                const uint64_t p = a0 + 32u * b + i;
                const uint8_t c = p < n ? seq[p] : 0;
                lo[b] |= (uint32_t)((c >> 1) & 1) << i;
                hi[b] |= (uint32_t)((c >> 2) & 1) << i;
            }
        }
        const uint32_t h0 = lo[0] | hi[0];
        vt_transpose32(lo);
        vt_transpose32(hi);
        const bool first = t == 0, last = t + 1 == nt;
#define ONE(K) if (kmin <= K && K <= kmax) one_k<K>(lo, hi, h0, rel0, m, first, last, ev);
        ONE(1) ONE(2) ONE(3) ONE(4) ONE(5) ONE(6) ONE(7) ONE(8) ONE(9) ONE(10) ONE(11) ONE(12) ONE(13) ONE(14) ONE(15) ONE(16)
#undef ONE
    }
    for (uint64_t i = 0; i < ev.size() && i < cap; i++) out[i] = ev[i];
    return ev.size();
}

// the definition: G[q] over the chunk ends of the slice, an event wherever it changes
uint64_t vt_events_ref(const uint8_t *seq, uint64_t n, uint64_t beg, uint64_t m, uint32_t kmin, uint32_t kmax, uint64_t *out, uint64_t cap) {
    (void)n;
    std::vector<uint64_t> ev;
    const uint8_t *t = seq + beg;
    for (uint32_t k = kmin; k <= kmax; k++) {
        bool prev = false;
        for (uint64_t j = 0;; j++) { // chunk end q = (j + 2) k - 1, up to the first one at or behind m
            const uint64_t q = (j + 2) * k - 1;
            bool g = false;
            if (q < m) g = memcmp(t + q - k + 1, t + q - 2 * k + 1, k) == 0;
            if (g != prev) ev.push_back(((uint64_t)k << 56) | j);
            prev = g;
            if (q >= m) break;
        }
    }
    for (uint64_t i = 0; i < ev.size() && i < cap; i++) out[i] = ev[i];
    return ev.size();
}

// transpose check: y[i] bit b == x[b] bit i
int vt_transpose_ok(const uint32_t *x) {
    uint32_t y[32];
    for (int i = 0; i < 32; i++) y[i] = x[i];
    vt_transpose32(y);
    for (int i = 0; i < 32; i++)
        for (int b = 0; b < 32; b++)
            if (((y[i] >> b) & 1u) != ((x[b] >> i) & 1u)) return 0;
    return 1;
}
}
