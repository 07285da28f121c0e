// explore_vertical_core.hpp -- the arithmetic of the vertical explore scan (explore_vertical.cuh), written so that
// the same code compiles for the device and, for tests/vertical_shim.cpp, for the host.
//
// What is computed (explore.rs:178-233 in the look-behind form of explore_planes.cuh): for a slice t[0..m), a chunk
// length k and a chunk end q = (j + 2) k - 1,
//     D_k[q] = t[q] == t[q-k],     G_k[q] = D_k[q-k+1] & ... & D_k[q]     (chunk j + 1 equals chunk j, raw bytes)
// and an EVENT at chunk j wherever G_k[q] != G_k[q-k] (G is 0 before the first full chunk pair and beyond the slice).
//
// How.  A lane owns a TILE of 1024 consecutive bases of ONE slice as 2 x 32 plane words (lo = ASCII bit 1,
// hi = ASCII bit 2: the planes of planes_build.cuh, which ARE the bytes wherever a word is "plain", i.e. 32 upper-case
// A C G T) and transposes each 32 x 32 bit matrix once: vertical word i, bit b  <->  tile position 32 b + i.  In that
// form "k bases back" is "word i - k" -- a register, not a shift -- for all 32 bit columns at once, and the word
// before word 0 is word 32 - j shifted up by one column (position 32 b - j = 32 (b - 1) + (32 - j)):
//     M[i]  = ~((lo[i] ^ lo[i-k]) | (hi[i] ^ hi[i-k]))                       2 LOP3        D_k for 32 positions
//     A[i]  = M[i] & M[i-1] & ... & M[i-k+1]                                 2-3 LOP3      (3-input ANDs of partial windows)
//     E[i]  = (A[i] ^ A[i-k]) & chunk_end_mask[(tile phase + i) mod k]       1 LOP3 + 1 LDS
// about 6 logic instructions per k and 32 bases where the horizontal form (explore_planes.cuh: funnel shifts, a 64-bit
// add per word, a shuffle for the neighbour's word) needs 19.  The column shifted in at the bottom is 0: positions
// before the tile compare against 'A' and count as "no match" in A, so the first 3k - 1 positions of a tile are not
// usable; tiles of a slice therefore advance by 960 bases and a tile owns the chunk ends from its position 64 on
// (the first tile of a slice starts AT the slice start and owns everything from 2k - 1 on, see vt_emit).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TIDK_HD __host__ __device__ __forceinline__
#else
#define TIDK_HD inline
#endif

namespace tidk {

constexpr uint32_t kVTile = 1024;              // bases per lane tile (32 words x 32 bits)
constexpr uint32_t kVHalo = 64;                // look-behind of a later tile: >= 3 * 16 - 1
constexpr uint32_t kVStep = kVTile - kVHalo;   // 960: tile t of a slice starts at slice position 960 t
constexpr int kVMaxK = 16;

struct VTile {
    uint32_t slice; // 2 * record + (0 left | 1 right)
    uint32_t t;     // tile of the slice
};

// tiles of a slice of m bases: tile 0 covers [0, 1024), tile t > 0 owns [960 t + 64, 960 t + 1024)
TIDK_HD uint64_t vt_count(uint64_t m) { return m == 0 ? 0 : (m <= kVTile ? 1 : 1 + (m - kVTile + kVStep - 1) / kVStep); }

TIDK_HD uint32_t vt_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// One LOP3: bit t of LUT is the result for (a, b, c) = (t >> 2 & 1, t >> 1 & 1, t & 1), i.e. LUT = f(0xF0, 0xCC, 0xAA).
// Written as inline PTX on the device: left to the compiler, the window ANDs are re-associated across levels and
// scheduled as one large expression graph, which costs a few hundred spilled registers per tile.
template <int LUT>
TIDK_HD uint32_t vt_lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
#else
    uint32_t r = 0;
    for (int t = 0; t < 8; t++)
        if ((LUT >> t) & 1) r |= ((t & 4) ? a : ~a) & ((t & 2) ? b : ~b) & ((t & 1) ? c : ~c);
    return r;
#endif
}

// (a & m) | (b & ~m)
TIDK_HD uint32_t vt_sel(uint32_t m, uint32_t a, uint32_t b) { return vt_lop3<0xCA>(m, a, b); }

// In-place transpose of a 32 x 32 bit matrix, LSB-first: on return bit b of x[i] is what bit i of x[b] was.
// Block-recursive: swap the off-diagonal 16 x 16 blocks (two PRMT per word pair), then 8 x 8 (PRMT), then 4, 2, 1
// (two shifts + two LOP3 per pair): 256 instructions per matrix.
template <int J>
TIDK_HD void vt_transpose_stage(uint32_t (&x)[32]) {
    constexpr uint32_t m = J == 4 ? 0x0F0F0F0Fu : J == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        if (k & J) continue;
        const uint32_t a = x[k], b = x[k + J];
        if (J == 16) {
            x[k] = vt_prmt(a, b, 0x5410);
            x[k + J] = vt_prmt(a, b, 0x7632);
        } else if (J == 8) {
            x[k] = vt_prmt(a, b, 0x6240);
            x[k + J] = vt_prmt(a, b, 0x7351);
        } else {
            x[k] = vt_sel(m, a, b << J);
            x[k + J] = vt_sel(~m, b, a >> J);
        }
    }
}
TIDK_HD void vt_transpose32(uint32_t (&x)[32]) {
    vt_transpose_stage<16>(x);
    vt_transpose_stage<8>(x);
    vt_transpose_stage<4>(x);
    vt_transpose_stage<2>(x);
    vt_transpose_stage<1>(x);
}

// word j of a vertical array, j in [-32, 32): the word before word 0 is word 32 - |j| one bit column up
TIDK_HD uint32_t vt_ext(const uint32_t (&a)[32], int j) { return j >= 0 ? a[j & 31] : (a[(j + 32) & 31] << 1); }

// Chunk-end masks.  Row k holds, for x = 0 .. k + 31, the word whose bit b says "tile position 32 b + i is a chunk
// end" when phase + i = x, phase = (slice position of the tile's position 0) mod k: bit b set iff
// (x + 32 b + 1) mod k == 0.  One table per CTA in shared memory on the device; row k starts vt_mask_row(k) words in.
TIDK_HD constexpr uint32_t vt_mask_row(int k) { return (uint32_t)(32 * (k - 1) + k * (k - 1) / 2); } // sum over j < k of (j + 32)
constexpr uint32_t kVMaskWords = 32 * kVMaxK + (kVMaxK + 1) * kVMaxK / 2; // = vt_mask_row(kVMaxK + 1)
TIDK_HD uint32_t vt_mask_word(int k, uint32_t x) {
    uint32_t w = 0;
    for (uint32_t b = 0; b < 32; b++)
        if ((x + 32u * b + 1u) % (uint32_t)k == 0) w |= 1u << b;
    return w;
}

// First tile of a slice: the column shifted in below position 0 is lo = hi = 0, i.e. 'A', so A[k-1] comes out as
// "the first k bases are all A" instead of 0 and E at the first chunk end that counts, 2k - 1, is off by exactly
// that bit (nothing else reads it).  h0 = lo | hi of the tile's first HORIZONTAL word.  True: flip bit 0 of E[2k - 1].
TIDK_HD bool vt_head_flip(uint32_t h0, uint32_t k) { return (h0 & ((1u << k) - 1u)) == 0u; }

// E[i] for one k and one tile.  lo / hi: the tile's vertical planes; mask(i): chunk-end mask word of vertical word i
// (table row of k at phase + i); out(i, e): receives E[i], once per i, in the order given by vt_out_key.  Returns the
// OR of all E.
//
// The windows are ANDed in levels: A3 = three consecutive M; for k <= 9 three A3 windows (overlapping unless k = 9)
// make A; for k >= 10 three A3 make A9 and two A9 (overlapping unless k = 18) make A.  Every level reaches back across
// word 0 into the top words (shifted up one column), which therefore have to exist first.  The code below is written
// as ONE pass over a virtual index u = 33 - 2k .. 63 - k in which each step produces one word of every level: a first
// lap over the top words 33 - 2k .. 31 (nothing there depends on the wrap), then the lap proper over the words
// 0 .. 31 - k, with word u - 32 at virtual index u and references to virtual indices below 32 shifted.  A[i] for the top k
// words comes out of the first lap; E of those words is formed as soon as the second lap delivers A[i - k].  In this
// order only a sliding window of each level (and the k top words of A) is live, where computing level after level over
// the whole tile keeps two whole 32-word arrays alive next to the planes.
template <int K>
TIDK_HD constexpr int vt_out_key(int i) { return i < 32 - K ? 2 * (32 + i) : 2 * (32 + i - K) + 1; } // production order of E[i]
template <int K>
TIDK_HD constexpr bool vt_group_done(int i) { // E[i] is the last word of its group of four to be produced
    const int g = i & ~3;
    int best = 0;
    for (int j = g; j < g + 4; j++) best = vt_out_key<K>(j) > best ? vt_out_key<K>(j) : best;
    return vt_out_key<K>(i) == best;
}

template <int K, class MaskFn, class OutFn>
TIDK_HD uint32_t vt_compute(const uint32_t (&lo)[32], const uint32_t (&hi)[32], MaskFn mask, OutFn out) {
    static_assert(K >= 1 && K <= 16, "the first lap starts at word 33 - 2k");
    constexpr int PM = 33 - 2 * K, END = 64 - K;
    uint32_t Mv[64], A3v[64], A9v[64], Av[64];
    uint32_t any = 0;
#pragma unroll
    for (int u = PM; u < END; u++) {
        const int i = u & 31;
        const bool lap2 = u >= 32;
        // a word `o` back, seen from virtual index u: the second lap sees the first lap's (top) words one column up
#define TIDK_VREF(X, o) ((lap2 && u - (o) < 32) ? (X[u - (o)] << 1) : X[u - (o)])
        Mv[u] = vt_lop3<0x09>(vt_lop3<0x3C>(lo[i], vt_ext(lo, i - K), 0u), hi[i], vt_ext(hi, i - K)); // ~(a | (b ^ c))
        bool have_a = false;
        if (K == 1) {
            Av[u] = Mv[u];
            have_a = true;
        } else if (K == 2) {
            if (lap2 || u >= 32 - K) { Av[u] = vt_lop3<0xC0>(Mv[u], TIDK_VREF(Mv, 1), 0u); have_a = true; }
        } else if (K == 3) {
            if (lap2 || u >= 32 - K) { Av[u] = vt_lop3<0x80>(Mv[u], TIDK_VREF(Mv, 1), TIDK_VREF(Mv, 2)); have_a = true; }
        } else {
            if (lap2 || u >= PM + 2) A3v[u] = vt_lop3<0x80>(Mv[u], TIDK_VREF(Mv, 1), TIDK_VREF(Mv, 2));
            if (K <= 9) {
                constexpr int a = K - 3 < 3 ? K - 3 : 3, b = K - 3;
                if (lap2 || u >= 32 - K) { Av[u] = vt_lop3<0x80>(A3v[u], TIDK_VREF(A3v, a), TIDK_VREF(A3v, b)); have_a = true; }
            } else {
                if (lap2 || u >= PM + 8) A9v[u] = vt_lop3<0x80>(A3v[u], TIDK_VREF(A3v, 3), TIDK_VREF(A3v, 6));
                if (lap2 || u >= 32 - K) { Av[u] = vt_lop3<0xC0>(A9v[u], TIDK_VREF(A9v, K - 9), 0u); have_a = true; }
            }
        }
        if (lap2 && have_a) {
            const uint32_t e = vt_lop3<0x28>(Av[u], TIDK_VREF(Av, K), mask(i)); // (a ^ b) & c
            out(i, e);
            if (i >= 32 - 2 * K) { // a top word: its A is a first-lap value, the word k below it is this one
                const uint32_t e2 = vt_lop3<0x28>(Av[i + K], Av[u], mask(i + K));
                out(i + K, e2);
                any = vt_lop3<0xFE>(any, e, e2);
            } else {
                any |= e;
            }
        }
#undef TIDK_VREF
    }
    return any;
}

// What a tile contributes for one k, given its E words (e(i), already corrected by vt_head_flip): an event at chunk
// j = (q + 1) / k - 2 for every E bit at a slice position q the tile owns, and -- last tile of the slice only -- the
// event that closes a run which touches the end of the slice: the E bits of a tile telescope (E[q] = A[q] ^ A[q-k],
// A = 0 before the tile), so G at the last chunk end inside the slice is the parity of the bits at positions
// 2k - 1 <= q < m, owned or not.
//   rel0   slice position of the tile's position 0 (960 t),  m  slice length,  last  the tile owns position m - 1
template <class EFn, class EmitFn>
TIDK_HD void vt_emit(uint32_t k, uint64_t rel0, uint64_t m, bool first, bool last, EFn e, EmitFn emit) {
    const uint64_t own_lo = first ? 2ull * k - 1 : rel0 + kVHalo;
    const uint64_t own_hi = last ? m : rel0 + kVTile; // exclusive
    uint32_t parity = 0;
    for (uint32_t i = 0; i < 32; i++) {
        uint32_t w = e(i);
        while (w) {
            uint32_t b = 0;
            while (!((w >> b) & 1u)) b++;
            w &= w - 1;
            const uint64_t q = rel0 + 32u * b + i;
            if (q >= 2ull * k - 1 && q < m) parity ^= 1u;
            if (q >= own_lo && q < own_hi) emit((q + 1) / k - 2);
        }
    }
    if (last && parity) emit(m / k - 1); // the run's last entry is the last full chunk
}

} // namespace tidk
