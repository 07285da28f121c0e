// host_pack.hpp -- host-side packing of ASCII bases into the three bit-planes the matchers use
// (lo = ASCII bit 1, hi = ASCII bit 2, va = byte is one of ACGTacgt), 32 bases per word, LSB first.
//
// Used by the host-pointer entry point tidk_cuda_window_counts: moving the bases (PCIe about 55 GB/s,
// host memory bandwidth) is the end-to-end bound, so 2 bits per base plus the rare validity exceptions
// cross the bus instead of 8 (pack_words2; pack_words is the three-plane form).  The exact same planes
// are what planes.cuh builds in registers from ASCII on the device path.
#pragma once
#include <immintrin.h>

#ifndef TIDK_PACK_NT
#define TIDK_PACK_NT 1 // non-temporal stores into the pinned staging buffers
#endif

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace tidk {

inline void pack_words_scalar(const uint8_t *src, size_t n_bytes, size_t n_words, uint32_t *lo, uint32_t *hi,
                              uint32_t *va, uint32_t *nonascii) {
    uint32_t na = 0;
    for (size_t w = 0; w < n_words; w++) {
        uint32_t l = 0, h = 0, v = 0;
        const size_t base = w * 32;
        const size_t m = base >= n_bytes ? 0 : (n_bytes - base < 32 ? n_bytes - base : 32);
        for (size_t b = 0; b < m; b++) {
            const uint8_t c = src[base + b];
            const uint8_t u = c | 0x20;
            l |= (uint32_t)((c >> 1) & 1u) << b;
            h |= (uint32_t)((c >> 2) & 1u) << b;
            v |= (uint32_t)(u == 'a' || u == 'c' || u == 'g' || u == 't') << b;
            na |= c & 0x80u;
        }
        lo[w] = l; hi[w] = h; va[w] = v;
    }
    if (na) *nonascii = 1;
}

__attribute__((target("avx2"))) inline void pack_words_avx2(const uint8_t *src, size_t n_bytes, size_t n_words,
                                                            uint32_t *lo, uint32_t *hi, uint32_t *va,
                                                            uint32_t *nonascii) {
    const size_t full = n_bytes / 32 < n_words ? n_bytes / 32 : n_words;
    const __m256i c20 = _mm256_set1_epi8(0x20), ca = _mm256_set1_epi8('a'), cc = _mm256_set1_epi8('c'),
                  cg = _mm256_set1_epi8('g'), ct = _mm256_set1_epi8('t');
    uint32_t na = 0;
    // movemask takes bit 7 of every byte: shift the wanted bit there (16-bit lanes: bit 1 / 2 of a byte
    // lands on bit 7 of the same byte)
#define TIDK_PACK_ONE(W, L, H, V)                                                                        \
    do {                                                                                                 \
        const __m256i v = _mm256_loadu_si256((const __m256i *)(src + (W) * 32));                         \
        (L) = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));                                   \
        (H) = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));                                   \
        const __m256i u = _mm256_or_si256(v, c20);                                                       \
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, ca), _mm256_cmpeq_epi8(u, cc)), \
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, cg), _mm256_cmpeq_epi8(u, ct))); \
        (V) = (uint32_t)_mm256_movemask_epi8(ok);                                                        \
        na |= (uint32_t)_mm256_movemask_epi8(v);                                                         \
    } while (0)
    size_t w = 0;
    // eight words at a time with non-temporal stores when the destinations are 32-byte aligned (they
    // are for the pinned staging buffers): the planes are written once and read only by the DMA engine,
    // so they should not be pulled into the cache first (no read-for-ownership traffic)
    if ((((uintptr_t)lo | (uintptr_t)hi | (uintptr_t)va) & 31) == 0) {
        alignas(32) uint32_t tl[8], th[8], tv[8];
        for (; w + 8 <= full; w += 8) {
            for (int i = 0; i < 8; i++) TIDK_PACK_ONE(w + i, tl[i], th[i], tv[i]);
            _mm256_stream_si256((__m256i *)(lo + w), _mm256_load_si256((const __m256i *)tl));
            _mm256_stream_si256((__m256i *)(hi + w), _mm256_load_si256((const __m256i *)th));
            _mm256_stream_si256((__m256i *)(va + w), _mm256_load_si256((const __m256i *)tv));
        }
        _mm_sfence();
    }
    for (; w < full; w++) TIDK_PACK_ONE(w, lo[w], hi[w], va[w]);
#undef TIDK_PACK_ONE
    if (na) *nonascii = 1;
    if (full < n_words)
        pack_words_scalar(src + full * 32, n_bytes - full * 32, n_words - full, lo + full, hi + full, va + full, nonascii);
}

// n_words words starting at src; only the first n_bytes bytes exist (the rest packs as invalid)
inline void pack_words(const uint8_t *src, size_t n_bytes, size_t n_words, uint32_t *lo, uint32_t *hi, uint32_t *va,
                       uint32_t *nonascii) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) pack_words_avx2(src, n_bytes, n_words, lo, hi, va, nonascii);
    else pack_words_scalar(src, n_bytes, n_words, lo, hi, va, nonascii);
}

// Two-plane variant: lo and hi are written, the validity plane is not -- it is all ones almost everywhere
// (every base one of ACGTacgt), so only the words where it is not are reported, as (global word index << 32 |
// va) records appended to `exc`.  The device fills its va plane with ones and scatters the exceptions.
// 2 bits per base cross the bus instead of 3, and the host writes a third less.
// STRICT (the explore upload): a base is valid only if it is one of A C G T in UPPER case, so that the planes of an
// exception-free word determine its bytes exactly (explore compares raw bytes, explore.rs:208).
template <bool STRICT = false, class Vec>
inline void pack_words2_scalar(const uint8_t *src, size_t n_bytes, size_t n_words, uint32_t *lo, uint32_t *hi,
                               uint64_t word0, Vec &exc, uint32_t *nonascii) {
    uint32_t na = 0;
    for (size_t w = 0; w < n_words; w++) {
        uint32_t l = 0, h = 0, v = 0;
        const size_t base = w * 32;
        const size_t m = base >= n_bytes ? 0 : (n_bytes - base < 32 ? n_bytes - base : 32);
        for (size_t b = 0; b < m; b++) {
            const uint8_t c = src[base + b];
            const uint8_t u = STRICT ? (uint8_t)(c ^ 0x20) : (uint8_t)(c | 0x20); // STRICT: only 'A' maps to 'a', ...
            l |= (uint32_t)((c >> 1) & 1u) << b;
            h |= (uint32_t)((c >> 2) & 1u) << b;
            v |= (uint32_t)(u == 'a' || u == 'c' || u == 'g' || u == 't') << b;
            na |= c & 0x80u;
        }
        lo[w] = l; hi[w] = h;
        if (v != 0xFFFFFFFFu) exc.push_back(((word0 + w) << 32) | v);
    }
    if (na) *nonascii = 1;
}

template <bool STRICT = false, class Vec>
__attribute__((target("avx2"))) inline void pack_words2_avx2(const uint8_t *src, size_t n_bytes, size_t n_words,
                                                             uint32_t *lo, uint32_t *hi, uint64_t word0, Vec &exc,
                                                             uint32_t *nonascii) {
    const size_t full = n_bytes / 32 < n_words ? n_bytes / 32 : n_words;
    const __m256i c20 = _mm256_set1_epi8(0x20), ca = _mm256_set1_epi8('a'), cc = _mm256_set1_epi8('c'),
                  cg = _mm256_set1_epi8('g'), ct = _mm256_set1_epi8('t');
    uint32_t na = 0;
#define TIDK_PACK_ONE(W, L, H, V)                                                                        \
    do {                                                                                                 \
        const __m256i v = _mm256_loadu_si256((const __m256i *)(src + (W) * 32));                         \
        (L) = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));                                   \
        (H) = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));                                   \
        const __m256i u = STRICT ? _mm256_xor_si256(v, c20) : _mm256_or_si256(v, c20);                                                       \
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, ca), _mm256_cmpeq_epi8(u, cc)), \
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, cg), _mm256_cmpeq_epi8(u, ct))); \
        (V) = (uint32_t)_mm256_movemask_epi8(ok);                                                        \
        na |= (uint32_t)_mm256_movemask_epi8(v);                                                         \
    } while (0)
    size_t w = 0;
    if ((((uintptr_t)lo | (uintptr_t)hi) & 31) == 0) {
        alignas(32) uint32_t tl[8], th[8], tv[8];
        for (; w + 8 <= full; w += 8) {
            uint32_t all = 0xFFFFFFFFu;
            for (int i = 0; i < 8; i++) { TIDK_PACK_ONE(w + i, tl[i], th[i], tv[i]); all &= tv[i]; }
#if TIDK_PACK_NT
            _mm256_stream_si256((__m256i *)(lo + w), _mm256_load_si256((const __m256i *)tl));
            _mm256_stream_si256((__m256i *)(hi + w), _mm256_load_si256((const __m256i *)th));
#else
            _mm256_store_si256((__m256i *)(lo + w), _mm256_load_si256((const __m256i *)tl));
            _mm256_store_si256((__m256i *)(hi + w), _mm256_load_si256((const __m256i *)th));
#endif
            if (all != 0xFFFFFFFFu)
                for (int i = 0; i < 8; i++)
                    if (tv[i] != 0xFFFFFFFFu) exc.push_back(((word0 + w + i) << 32) | tv[i]);
        }
        _mm_sfence();
    }
    for (; w < full; w++) {
        uint32_t va1;
        TIDK_PACK_ONE(w, lo[w], hi[w], va1);
        if (va1 != 0xFFFFFFFFu) exc.push_back(((word0 + w) << 32) | va1);
    }
#undef TIDK_PACK_ONE
    if (na) *nonascii = 1;
    if (full < n_words)
        pack_words2_scalar<STRICT>(src + full * 32, n_bytes - full * 32, n_words - full, lo + full, hi + full, word0 + full, exc, nonascii);
}

// AVX-512 (BW + VBMI) form: 64 bases per step.  A plane is ONE instruction -- vptestmb against the bit's mask
// yields the 64 plane bits as a mask register -- and validity is a 64-entry table lookup on the low six bits of
// the lower-cased byte (vpermb) plus one compare: about 8 uops per 64 bases against about 28 for the AVX2 form,
// so a packer thread is bound by memory instead of the movemask port.
template <bool STRICT = false, class Vec>
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) inline void pack_words2_avx512(const uint8_t *src, size_t n_bytes,
                                                                                     size_t n_words, uint32_t *lo, uint32_t *hi,
                                                                                     uint64_t word0, Vec &exc, uint32_t *nonascii) {
    const size_t full = n_bytes / 32 < n_words ? n_bytes / 32 : n_words;
    alignas(64) uint8_t tab[64];
    memset(tab, 0, sizeof tab); // 0 never equals a byte with bit 5 set
    tab[0] = 0xFF;              // STRICT: u = v ^ 0x20 is 0 for a space; 0xFF equals none of 0x00 / 0x40 / 0x80 / 0xC0
    tab['a' & 63] = 'a'; tab['c' & 63] = 'c'; tab['g' & 63] = 'g'; tab['t' & 63] = 't';
    const __m512i table = _mm512_load_si512(tab);
    const __m512i c02 = _mm512_set1_epi8(0x02), c04 = _mm512_set1_epi8(0x04), c20 = _mm512_set1_epi8(0x20);
    __m512i acc = _mm512_setzero_si512(); // OR of all bytes: bit 7 set somewhere <=> a non-ASCII byte
    size_t w = 0;
    // sixteen words (512 bases) per round: the eight 64-bit plane pieces of a round are gathered in a register and
    // leave as ONE full-line non-temporal store per plane (8-byte streaming stores were tried: partially filled
    // write-combining buffers get evicted and the memory controller serialises every thread at 6 GB/s)
    if ((((uintptr_t)lo | (uintptr_t)hi) & 63) == 0) {
        for (; w + 16 <= full; w += 16) {
            __m512i gl = _mm512_setzero_si512(), gh = _mm512_setzero_si512();
            uint64_t okall = ~0ull;
            uint64_t oks[8];
#define TIDK_PACK_STEP(I)                                                                                  \
    do {                                                                                                   \
        const __m512i v = _mm512_loadu_si512(src + (w + 2 * (I)) * 32);                                    \
        const uint64_t l = _mm512_test_epi8_mask(v, c02), h = _mm512_test_epi8_mask(v, c04);               \
        const __m512i u = STRICT ? _mm512_xor_si512(v, c20) : _mm512_or_si512(v, c20);                     \
        oks[I] = _mm512_cmpeq_epi8_mask(u, _mm512_permutexvar_epi8(u, table));                             \
        okall &= oks[I];                                                                                   \
        acc = _mm512_or_si512(acc, v);                                                                     \
        gl = _mm512_mask_set1_epi64(gl, (__mmask8)(1u << (I)), (long long)l);                              \
        gh = _mm512_mask_set1_epi64(gh, (__mmask8)(1u << (I)), (long long)h);                              \
    } while (0)
            TIDK_PACK_STEP(0); TIDK_PACK_STEP(1); TIDK_PACK_STEP(2); TIDK_PACK_STEP(3);
            TIDK_PACK_STEP(4); TIDK_PACK_STEP(5); TIDK_PACK_STEP(6); TIDK_PACK_STEP(7);
#undef TIDK_PACK_STEP
#if TIDK_PACK_NT
            _mm512_stream_si512((__m512i *)(lo + w), gl);
            _mm512_stream_si512((__m512i *)(hi + w), gh);
#else
            _mm512_store_si512((__m512i *)(lo + w), gl);
            _mm512_store_si512((__m512i *)(hi + w), gh);
#endif
            if (okall != ~0ull)
                for (int i = 0; i < 8; i++) {
                    const uint64_t ok = oks[i];
                    if ((uint32_t)ok != 0xFFFFFFFFu) exc.push_back(((word0 + w + 2 * i) << 32) | (uint32_t)ok);
                    if ((uint32_t)(ok >> 32) != 0xFFFFFFFFu) exc.push_back(((word0 + w + 2 * i + 1) << 32) | (uint32_t)(ok >> 32));
                }
        }
        _mm_sfence();
    }
    if (_mm512_movepi8_mask(acc)) *nonascii = 1;
    if (w < n_words)
        pack_words2_scalar<STRICT>(src + w * 32, n_bytes - w * 32, n_words - w, lo + w, hi + w, word0 + w, exc, nonascii);
}

template <bool STRICT = false, class Vec>
inline void pack_words2(const uint8_t *src, size_t n_bytes, size_t n_words, uint32_t *lo, uint32_t *hi, uint64_t word0,
                        Vec &exc, uint32_t *nonascii) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    // TIDK_PACK_ISA=avx2 keeps the AVX2 form on an AVX-512 machine (measurements)
    static const bool avx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                               !(getenv("TIDK_PACK_ISA") && strcmp(getenv("TIDK_PACK_ISA"), "avx2") == 0);
    if (avx512) pack_words2_avx512<STRICT>(src, n_bytes, n_words, lo, hi, word0, exc, nonascii);
    else if (avx2) pack_words2_avx2<STRICT>(src, n_bytes, n_words, lo, hi, word0, exc, nonascii);
    else pack_words2_scalar<STRICT>(src, n_bytes, n_words, lo, hi, word0, exc, nonascii);
}

} // namespace tidk
