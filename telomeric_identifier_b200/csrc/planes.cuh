// planes.cuh -- ASCII -> bit-plane transposition in registers (sm_100a).
//
// A lane owns 32 consecutive sequence bytes (one 256-bit load) and turns them
// into 32-bit planes, bit b <-> base b (LSB = lowest address):
//   lo = ASCII bit 1, hi = ASCII bit 2   (A=00 C=01 G=11 T=10 as (hi,lo); case-blind;
//                                         complement == flip hi)
//   b0 = ASCII bit 0                      (1 for A/C/G, 0 for T)
//   ok = all ones on the fast path; on the exact path the positions whose other
//        ASCII bits are consistent with a letter of {A,C,G,T,a,c,g,t}
// A base is one of A C G T a c g t  <=>  ok & (b0 ^ (hi & ~lo)).  Everything else
// (N, IUPAC codes, gaps, ...) never matches an A/C/G/T query, which is the
// reference's rule: the window is upper-cased and compared bytewise
// (search.rs:107-110, finder.rs:146-149).
//
// How a plane is formed.  Two words wa (bases 0-3) and wb (bases 4-7) are merged
// into nibble words  X = low nibbles, Y = high nibbles  (one shift + one LOP3
// each); bit o of the eight nibbles of a word is then gathered into one byte by a
// single dot product (dp4a with weights 1,2,4,8) -- or, in the alternative build,
// by a single multiply whose 32 partial products land on distinct bits -- and the
// four bytes are strung together with multiply-adds.  IDP/IMAD run on the fma
// pipe, LOP3/SHF/PRMT on the alu pipe, so the two halves of the work overlap.
//
// Validity needs five more ASCII bits (0,3,4,6,7).  b7=0, b6=1, b3=0 are constant
// over the valid set and b4 == !b0 on it, so those four are checked in aggregate
// on the nibble words (AND/OR accumulators, no extraction); only b0 is
// extracted.  If the aggregate test fails for any lane of the warp (N runs,
// IUPAC codes, anything else) the exact `ok` plane is built from four more planes.
#pragma once
#include <cstdint>

namespace tidk {

__device__ __forceinline__ void ld256_stream(const void *p, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),
                   "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

// (a & m) | (b & ~m) in one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t m, uint32_t a, uint32_t b) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(a), "r"(b));
    return r;
}

__host__ __device__ constexpr uint32_t gather_mul(int o) {
    return (1u << (24 - o)) | (1u << (17 - o)) | (1u << (10 - o)) | (1u << (3 - o));
}

// bit `O` of each of the 32 nibbles held in N[0..3] (8 nibbles per word) -> 32-bit plane.
//
// TIDK_GATHER 2 (default): dp4a gather.  Byte b of (N & M) is l*2^O + h*2^(O+4) (l / h = the wanted
// bit of the low / high nibble); dotting the four bytes with the weights (1,2,4,8) gives the eight
// gathered bits, shifted left by O.  Three multiply-adds string the four groups together and one
// funnel shift removes the O.  IDP.4A and IMAD issue on the fma pipe, which this kernel leaves
// mostly idle, so only the four masks (and two shifts for O > 0) land on the busy alu pipe.
// TIDK_GATHER 0: multiply gather ((N & M) * C_O puts the eight bits into the top byte; the 32 partial
// products land on distinct bit positions, so there are no carries) + a three-PRMT tree.
#ifndef TIDK_GATHER
#define TIDK_GATHER 2
#endif
template <int O>
__device__ __forceinline__ uint32_t nibble_plane(const uint32_t (&N)[4]) {
    constexpr uint32_t M = 0x11111111u << O;
#if TIDK_GATHER == 2
    const uint32_t d0 = __dp4a(N[0] & M, 0x08040201u, 0u);
    const uint32_t d1 = __dp4a(N[1] & M, 0x08040201u, 0u);
    const uint32_t d2 = __dp4a(N[2] & M, 0x08040201u, 0u);
    const uint32_t d3 = __dp4a(N[3] & M, 0x08040201u, 0u);
    const uint32_t r = d3 * 0x1000000u + (d2 * 0x10000u + (d1 * 0x100u + d0)); // (plane << O) mod 2^32
    if (O == 0) return r;
    return __funnelshift_r(r, d3 >> 8, O); // the O bits that fell off the top are d3's bits 8..8+O
#else
    constexpr uint32_t C = gather_mul(O);
    uint32_t p0 = (N[0] & M) * C;
    uint32_t p1 = (N[1] & M) * C;
    uint32_t p2 = (N[2] & M) * C;
    uint32_t p3 = (N[3] & M) * C;
    return __byte_perm(__byte_perm(p0, p1, 0x7373), __byte_perm(p2, p3, 0x7373), 0x5410);
#endif
}

struct Nibbles {
    uint32_t X[4], Y[4]; // low / high nibbles of 32 bases, 8 per word
};

__device__ __forceinline__ Nibbles to_nibbles(const uint32_t (&w)[8]) {
    Nibbles n;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t wa = w[2 * g], wb = w[2 * g + 1];
        n.X[g] = bitsel(0x0F0F0F0Fu, wa, wb << 4);
        n.Y[g] = bitsel(0x0F0F0F0Fu, wa >> 4, wb);
    }
    return n;
}

struct Planes {
    uint32_t lo, hi, b0, ok;
    uint32_t isn; // exact path only: byte is 'N' or 'n' (0 on the fast path, where there is none)
    bool exact;   // warp-uniform: ok is not all ones
};

// The planes every path needs, plus the per-lane verdict of the aggregate test; no warp vote yet.
struct PlaneCore {
    uint32_t lo, hi, b0;
    uint32_t bad; // non-zero: some byte of this lane is not of the form 010x0xxx with b4 == !b0
};

__device__ __forceinline__ PlaneCore extract_core(const Nibbles &n) {
    uint32_t accA = 0xFFFFFFFFu, accO = 0u, accB = 0xFFFFFFFFu;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        accA &= n.X[g] ^ n.Y[g]; // nibble bit 0: b0 ^ b4 must be 1
        accO |= n.X[g] | n.Y[g]; // nibble bit 3: b3 | b7 must be 0
        accB &= n.Y[g];          // nibble bit 2: b6 must be 1
    }
    PlaneCore c;
    c.bad = (~accA & 0x11111111u) | (accO & 0x88888888u) | (~accB & 0x44444444u);
    c.lo = nibble_plane<1>(n.X);
    c.hi = nibble_plane<2>(n.X);
    c.b0 = nibble_plane<0>(n.X);
    return c;
}

// exact = (warp-uniform) some lane failed the aggregate test: build the exact `ok` and N planes.
// nonascii is OR-ed with a non-zero value if any byte >= 0x80 (such a byte always fails the aggregate test).
__device__ __forceinline__ Planes finish_planes(const Nibbles &n, const PlaneCore &c, bool exact, uint32_t &nonascii) {
    Planes P;
    P.lo = c.lo; P.hi = c.hi; P.b0 = c.b0;
    P.ok = 0xFFFFFFFFu;
    P.isn = 0u;
    P.exact = exact;
    if (P.exact) {
        uint32_t b3 = nibble_plane<3>(n.X);
        uint32_t b4 = nibble_plane<0>(n.Y);
        uint32_t b6 = nibble_plane<2>(n.Y);
        uint32_t b7 = nibble_plane<3>(n.Y);
        nonascii |= b7;
        P.ok = ~b7 & b6 & ~b3 & (b4 ^ P.b0);
        P.isn = ~b7 & b6 & b3 & ~b4 & P.hi & P.lo & ~P.b0; // 0x4E / 0x6E (explore treats N as a fifth letter)
    }
    return P;
}

__device__ __forceinline__ Planes extract_planes(const Nibbles &n, uint32_t &nonascii) {
    const PlaneCore c = extract_core(n);
    return finish_planes(n, c, __any_sync(0xFFFFFFFFu, c.bad != 0), nonascii);
}

// prev[lane] = lane > 0 ? cur[lane-1] : (this lane 31's value from the previous block)
__device__ __forceinline__ uint32_t lane_lookbehind(uint32_t cur, uint32_t old, int lane) {
    uint32_t x = (lane == 31) ? old : cur;
    return __shfl_sync(0xFFFFFFFFu, x, (lane + 31) & 31);
}

// bits b in [0,32) with lo_rel <= b < hi_rel (branch-free: shl.b32 clamps shift counts above 31 to "all out")
__device__ __forceinline__ uint32_t range_mask32(int lo_rel, int hi_rel) {
    uint32_t from_lo, from_hi; // PTX shl.b32 clamps the count: 32 or more shifts everything out
    asm("shl.b32 %0, %1, %2;" : "=r"(from_lo) : "r"(0xFFFFFFFFu), "r"((uint32_t)max(lo_rel, 0)));
    asm("shl.b32 %0, %1, %2;" : "=r"(from_hi) : "r"(0xFFFFFFFFu), "r"((uint32_t)max(hi_rel, 0)));
    return from_lo & ~from_hi;
}

} // namespace tidk
