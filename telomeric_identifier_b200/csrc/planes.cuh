// planes.cuh -- ASCII -> bit-plane transposition in registers (sm_100a).
//
// A lane owns 32 consecutive sequence bytes (one 256-bit load) and turns them
// into 32-bit planes, bit b <-> base b (LSB = lowest address):
//   lo = ASCII bit 1, hi = ASCII bit 2   (A=00 C=01 G=11 T=10 as (hi,lo); case-blind;
//                                         complement == flip hi)
//   va = byte is one of A C G T a c g t  (everything else never matches an A/C/G/T query,
//                                         which is the reference's rule: the window is
//                                         upper-cased and compared bytewise, search.rs:107-110)
//
// How a plane is formed.  Two words wa (bases 0-3) and wb (bases 4-7) are merged
// into nibble words  X = low nibbles, Y = high nibbles  (one shift + one LOP3
// each); bit o of the eight nibbles is then gathered into one byte by a single
// multiply:  ((X & 0x11111111<<o) * C_o) >> 24, C_o = 2^(24-o)+2^(17-o)+2^(10-o)+2^(3-o)
// (all 32 partial products land on distinct bit positions, so there are no
// carries).  Four such bytes are combined with three PRMT.  IMAD runs on the fma
// pipe, LOP3/SHF/PRMT on the alu pipe, so the two halves of the work overlap.
//
// Validity costs five more ASCII bits (0,3,4,6,7).  Three of them are constant
// over the valid set and b4 == !b0 on it, so they are checked in aggregate on
// the nibble words (AND/OR accumulators, no extraction); only b0 is extracted,
// and  va = b0 ^ (hi & ~lo).  If the aggregate test fails for any lane of the
// warp (N runs, IUPAC codes, anything else) the exact planes are built instead.
#pragma once
#include <cstdint>

namespace tidk {

__device__ __forceinline__ void ld256_stream(const void *p, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),
                   "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

__host__ __device__ constexpr uint32_t gather_mul(int o) {
    return (1u << (24 - o)) | (1u << (17 - o)) | (1u << (10 - o)) | (1u << (3 - o));
}

// bit `o` of each of the 32 nibbles held in N[0..3] (8 nibbles per word) -> 32-bit plane
template <int O>
__device__ __forceinline__ uint32_t nibble_plane(const uint32_t (&N)[4]) {
    constexpr uint32_t M = 0x11111111u << O;
    constexpr uint32_t C = gather_mul(O);
    uint32_t p0 = (N[0] & M) * C;
    uint32_t p1 = (N[1] & M) * C;
    uint32_t p2 = (N[2] & M) * C;
    uint32_t p3 = (N[3] & M) * C;
    return __byte_perm(__byte_perm(p0, p1, 0x7373), __byte_perm(p2, p3, 0x7373), 0x5410);
}

struct Planes {
    uint32_t lo, hi, va;
};

// w: 32 sequence bytes.  nonascii is OR-ed with a non-zero value if any byte >= 0x80.
__device__ __forceinline__ Planes extract_planes(const uint32_t (&w)[8], uint32_t &nonascii) {
    uint32_t X[4], Y[4];
    uint32_t accA = 0xFFFFFFFFu, accO = 0u, accB = 0xFFFFFFFFu;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t wa = w[2 * g], wb = w[2 * g + 1];
        X[g] = (wa & 0x0F0F0F0Fu) | ((wb << 4) & 0xF0F0F0F0u);
        Y[g] = ((wa >> 4) & 0x0F0F0F0Fu) | (wb & 0xF0F0F0F0u);
        accA &= X[g] ^ Y[g]; // nibble bit 0: b0 ^ b4 must be 1
        accO |= X[g] | Y[g]; // nibble bit 3: b3 | b7 must be 0
        accB &= Y[g];        // nibble bit 2: b6 must be 1
    }
    uint32_t bad = (~accA & 0x11111111u) | (accO & 0x88888888u) | (~accB & 0x44444444u);
    Planes P;
    P.lo = nibble_plane<1>(X);
    P.hi = nibble_plane<2>(X);
    uint32_t b0 = nibble_plane<0>(X);
    uint32_t t = P.hi & ~P.lo; // looks like T in (b2,b1)
    if (__any_sync(0xFFFFFFFFu, bad != 0)) {
        // exact validity: b7=0, b6=1, b3=0, b4=!b0, b0=!(b2&!b1)
        uint32_t b3 = nibble_plane<3>(X);
        uint32_t b4 = nibble_plane<0>(Y);
        uint32_t b6 = nibble_plane<2>(Y);
        uint32_t b7 = nibble_plane<3>(Y);
        nonascii |= b7;
        P.va = ~b7 & b6 & ~b3 & (b4 ^ b0) & (b0 ^ t);
    } else {
        P.va = b0 ^ t;
    }
    return P;
}

// prev[lane] = lane > 0 ? cur[lane-1] : carry_from_previous_block[31]
__device__ __forceinline__ uint32_t lane_lookbehind(uint32_t cur, uint32_t old, int lane) {
    uint32_t x = (lane == 31) ? old : cur;
    return __shfl_sync(0xFFFFFFFFu, x, (lane + 31) & 31);
}

} // namespace tidk
