// explore_planes.cuh -- tandem k-mer run detection on bit-planes (k <= 16).
//
// Same contract as explore_events_kernel (explore.cuh): emit one RunEvent per raw-run
// boundary of every (slice, k).  Restated in look-behind form, with q a slice-relative
// position and t the slice:
//     D_k[q] = t[q] == t[q-k]                       (both inside the slice)
//     G_k[q] = D_k[q-k+1] & ... & D_k[q]            (the k-mer ending at q equals the one before it)
// At a chunk end q = (j+2)k - 1, G_k[q] is the reference's raw_j = (c_j == c_{j+1})
// (explore.rs:208) and G_k[q-k] is raw_{j-1}; where they differ chunk j starts a raw run or is
// the last entry of one.  Positions up to m + k - 1 are evaluated so that the run that touches
// the end of the slice is closed (G is 0 beyond the slice: a partial chunk equals nothing).
//
// A lane owns 32 bases per 1 KiB block (one 256-bit load) as planes lo / hi / va
// (va = base is A/C/G/T/N in any case AND inside the slice).  The mismatch plane nD_k = ~D_k costs two
// funnel shifts and two LOP3 for the lane's word (one more of each with the validity plane); the
// previous lane's nD_k arrives by SHFL.  G_k at the chunk ends is an all-ones-field test on D over the
// (previous | current) word pair -- one 64-bit add, explore_k() -- with the chunk-end mask of the lane's
// phase (q mod k) read from a small shared-memory table.  Events are rare (4^-k per chunk on random
// sequence), so the append path is off the hot path.
//
// Exactness.  Plane equality on lo / hi is byte equality only for bases in {A,C,G,T} of uniform case.
// Blocks are sorted (warp-uniformly) into: plain (one ballot: interior, all upper-case A/C/G/T) ->
// lo / hi only; clean at a slice or tile edge -> + validity plane; mixed case, case change or N present
// -> + case and N planes (then plane equality is byte equality for A/C/G/T/N in either case); any other
// byte (IUPAC codes, gaps) in this or the previous block -> the block's owned chunk ends are evaluated
// byte-wise while the plane path only keeps its carries consistent.
#pragma once
#include <cstddef>

#include "explore.cuh"
#include "planes.cuh"
#include "planes_build.cuh"

#ifndef TIDK_EX_LDS_PRMT
#define TIDK_EX_LDS_PRMT 1 // chunk-end masks: table address and phase folded by one PRMT (see phase_addr)
#endif

namespace tidk {

constexpr int kPlaneMaxK = 16;
constexpr uint32_t kExTile = 16384;      // chunk-end positions owned by one work item ...
constexpr uint32_t kExTileSmall = 4096;  // ... or this when there are too few items to fill the GPU

struct __align__(64) ExItem {
    uint64_t scan0;     // absolute offset of block 0 (multiple of 32)
    int64_t rel0;       // scan0 - slice_beg (slice-relative position of block 0, lane 0, bit 0)
    uint64_t slice_len; // m
    uint32_t slice;     // 2 * record + (0 left | 1 right)
    uint16_t nblk;
    uint16_t own_lo, own_hi; // owned chunk-end positions, relative to scan0
    uint16_t v_lo, v_hi;     // bases inside the slice, relative to scan0
    uint16_t pad;
    uint8_t pad2[8];
    uint8_t r0[kPlaneMaxK];  // r0[k-1] = rel0 mod k (non-negative); 16-byte aligned (offset 48)
};
static_assert(offsetof(ExItem, r0) % 16 == 0, "r0 is read as one uint4");
static_assert(sizeof(ExItem) == 64, "ExItem is two 256-bit loads");

struct ExPrepJob {
    const SliceDesc *slices;
    const uint64_t *tile_base; // [n_slices + 1]
    ExItem *items;
    uint64_t n_items;
    uint32_t n_slices, kmax, tile;
};

// The work item that owns the chunk-end positions [own_beg, own_end) (slice-relative) of slice `slice`.  A range that
// does not start at the slice start is scanned from one block earlier: real look-behind planes for its first block.
__device__ __forceinline__ ExItem make_ex_item(const SliceDesc sd, uint32_t slice, uint64_t own_beg, uint64_t own_end) {
    ExItem d;
    memset(&d, 0, sizeof d);
    uint64_t scan0 = (sd.beg + own_beg) & ~31ull;
    if (own_beg > 0) scan0 -= 1024;
    d.scan0 = scan0;
    d.rel0 = (int64_t)scan0 - (int64_t)sd.beg;
    d.slice_len = sd.len;
    d.slice = slice;
    const uint64_t end_abs = sd.beg + own_end;
    d.nblk = (uint16_t)((end_abs - scan0 + 1023) >> 10);
    d.own_lo = (uint16_t)(sd.beg + own_beg - scan0);
    d.own_hi = (uint16_t)(end_abs - scan0);
    d.v_lo = (uint16_t)(d.rel0 < 0 ? -d.rel0 : 0);
    const int64_t vhi = (int64_t)sd.len - d.rel0;
    d.v_hi = (uint16_t)(vhi < 0 ? 0 : (vhi > 65535 ? 65535 : vhi));
    for (int k = 1; k <= kPlaneMaxK; k++) {
        int64_t r = d.rel0 % k;
        if (r < 0) r += k;
        d.r0[k - 1] = (uint8_t)r;
    }
    return d;
}

__global__ void __launch_bounds__(256) explore_items_kernel(const ExPrepJob job) {
    const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= job.n_items) return;
    uint32_t lo = 0, hi = job.n_slices;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (job.tile_base[mid] <= item) lo = mid; else hi = mid;
    }
    const SliceDesc sd = job.slices[lo];
    const uint64_t tile_off = (item - job.tile_base[lo]) * (uint64_t)job.tile;
    const uint64_t q_end = sd.len + job.kmax; // chunk ends up to m + k - 1 close the last run
    const uint64_t own_end = tile_off + job.tile < q_end ? tile_off + job.tile : q_end;
    job.items[item] = make_ex_item(sd, lo, tile_off, own_end);
}

struct ExJob {
    const uint8_t *seq;
    const SliceDesc *slices;
    const ExItem *items;
    uint64_t n_items;
    uint32_t kmin, kmax;
    RunEvent *events;
    unsigned long long *n_events;  // slots handed out (events + the sentinels that fill abandoned reservations)
    unsigned long long *n_unused;  // sentinels among them
    uint64_t cap_events;
    unsigned long long *work_counter; // kQueues ticket counters, one per 128-byte line (window_counts.cuh)
    uint32_t zero;                 // 0: keeps ptxas from proving an atomic's address warp-uniform (see reserve_slots)
    uint32_t *err_flags;
    uint32_t packed; // events are 64-bit keys (pack_event) instead of RunEvent structs
    KeyFmt fmt;
    // resident bit-planes of the batch (planes_build.cuh), PLANES variant of the scan: lo / hi words in the chunked
    // layout and one "all 32 bytes are upper-case A C G T" bit per word
    const uint32_t *planes;
    const uint32_t *plainbits;
    // when set, the number of items is read from here (the list the vertical scan leaves behind, explore_vertical.cuh)
    const unsigned long long *n_items_dev;
};

// Events are appended with one atomicAdd each on a global counter (the compiler aggregates the lanes
// of a warp that emit together).  A per-warp shared-memory staging queue with one atomic per work
// item was tried and measured slower on read-like inputs (1.07 vs 0.99 ms): the events are too sparse
// (about one per KiB) for the same-address atomics to matter, and the staging code costs registers.
__device__ __forceinline__ void emit_event(const ExJob &job, uint32_t slice, uint32_t k, uint64_t chunk,
                                           bool start, bool joins) {
    const uint32_t type = start ? 0u : 1u, jn = joins ? 1u : 0u;
    // a plain atom: one or two lanes of a warp get here together, the compiler's warp-aggregated atomicAdd (vote,
    // leader election, popc, shuffle) costs more than the second atomic it saves
    unsigned long long idx;
    asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(idx) : "l"(job.n_events) : "memory");
    if (idx < job.cap_events) {
        if (job.packed) ((uint64_t *)job.events)[idx] = pack_event(job.fmt, slice, k, chunk, type, jn);
        else {
            RunEvent ev;
            ev.slice = slice; ev.k = (uint8_t)k; ev.type = (uint8_t)type; ev.joins = (uint8_t)jn; ev.pad = 0;
            ev.chunk = chunk;
            job.events[idx] = ev;
        }
    }
}

// Event slots are handed to a warp in ranges of kExReserve, requested ahead of need: a returning atomic per event made
// the emitting lane -- and with it the warp -- wait out an L2 round trip for every run boundary (20 % of the stall
// samples of the scan).  The part of a range that is never used is filled with sentinel keys (all ones: above every
// real key, the k field of a real key is never all ones) that the sort sends to the end, and counted in n_unused.
constexpr uint32_t kExReserve = 32;

__device__ __forceinline__ void store_sentinel(const ExJob &job, unsigned long long idx) {
    if (idx >= job.cap_events) return;
    if (job.packed) ((uint64_t *)job.events)[idx] = ~0ull;
    else {
        RunEvent ev;
        ev.slice = 0xFFFFFFFFu; ev.k = 0xFF; ev.type = 0xFF; ev.joins = 0; ev.pad = 0; ev.chunk = ~0ull;
        job.events[idx] = ev;
    }
}

// lane 0 asks for the next range; nothing waits for the answer here.  The address is "lane-dependent" (job.zero is 0):
// for a provably uniform address ptxas emits its warp-aggregated form, whose closing SHFL consumes the result at once.
__device__ __forceinline__ void reserve_slots(const ExJob &job, int lane, unsigned long long &raw) {
    if (lane == 0) {
        const unsigned long long *a = job.n_events + (size_t)threadIdx.x * job.zero; // (lane is known to be 0 here, threadIdx.x is not)
        asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(raw) : "l"(a), "l"((unsigned long long)kExReserve) : "memory");
    }
}

struct SlotState {
    unsigned long long base; // next free slot of the current range
    unsigned long long raw;  // lane 0: answer to the outstanding request
    uint32_t left;           // free slots in the current range
    uint32_t unused;         // sentinels written so far
    bool pending;            // a request is outstanding
};

// make room for n (<= 32) events: abandon what is left of the current range, move on to the requested one
__device__ __forceinline__ void next_range(const ExJob &job, int lane, SlotState &st) {
    if ((uint32_t)lane < st.left) store_sentinel(job, st.base + (uint32_t)lane);
    st.unused += st.left;
    if (!st.pending) reserve_slots(job, lane, st.raw); // (first events of a warp, or a burst: wait for it)
    st.base = __shfl_sync(0xFFFFFFFFu, st.raw, 0);
    st.left = kExReserve;
    st.pending = false;
}

// byte-wise G_k[q] (exact for any bytes)
__device__ __forceinline__ bool g_bytes(const uint8_t *t, int64_t m, int k, int64_t q) {
    if (q < 2 * (int64_t)k - 1 || q >= m) return false;
    const uint8_t *a = t + (q - k + 1);
    for (int j = 0; j < k; j++)
        if (a[j] != a[j - k]) return false;
    return true;
}

__device__ __forceinline__ bool chunks_equal_upper(const uint8_t *t, int k, uint64_t j) {
    const uint8_t *a = t + (j - 1) * (uint64_t)k;
    for (int i = 0; i < k; i++)
        if (dev_upper(a[i]) != dev_upper(a[i + k])) return false;
    return true;
}

// per-CTA stash of the run boundaries found in the current block: [k - KBASE][thread].  The hot loop only stores
// the boundary word of a k and sets the k's bit in a per-lane mask; whether a boundary starts or ends a run is
// read back from the sequence on the (rare) emission path.
template <int ROWS> // one row per k of the kernel's range (row = k - KBASE)
struct EvStash {
    uint32_t ev[ROWS][256];
};

struct Planes3 {
    uint32_t lo, hi;
    uint32_t va; // base is one of A C G T N (either case) and inside the slice
    uint32_t cs; // letter case (ASCII bit 5); extracted only when the block needs it
    uint32_t nn; // base is N / n (shares lo/hi with G, so it needs its own plane)
};

// Chunk-end masks.  htab[K-1][s] = ~H for the (previous | current) 64-bit pair when the first chunk end
// of the previous word sits at bit s: H has a bit every K positions from s on (the chunk ends).  Filled
// once per CTA; one LDS.64 per k and block replaces two shifts of a 64-bit constant.
struct HTable {
    uint2 nh[kPlaneMaxK][16];
};

__device__ __forceinline__ void fill_htable(HTable &tab) {
    for (int e = threadIdx.x; e < kPlaneMaxK * 16; e += blockDim.x) {
        const int k = e / 16 + 1, sft = e % 16;
        uint64_t per = 0;
        for (int b = 0; b < 64; b += k) per |= 1ull << b;
        const uint64_t nh = ~(per << sft);
        tab.nh[k - 1][sft] = make_uint2((uint32_t)nh, (uint32_t)(nh >> 32));
    }
}

// Per-lane chunk phases, four k per register, one byte each: 8 * s_k, where s_k (0..k-1) is the bit of the
// first chunk end in the PREVIOUS word.  A block later the phase is (s_k - 1024) mod k; the update runs on
// all four bytes at once (no byte overflows: 16k - 16 + 128 - 8k < 256 for k <= 16).
template <int KBASE, int REG>
struct PhaseConst {
    __host__ __device__ static constexpr uint32_t byte_of(int k, int what) { // what 0: 8 * ((k - 1024 % k) % k), 1: 0x80 - 8k, 2: 8k,
        return k > kPlaneMaxK ? 0u                                                  //      3: 0x80 - k, 4: k, 5: k - 1
               : what == 0 ? 8u * (uint32_t)((k - 1024 % k) % k)
               : what == 1 ? 0x80u - 8u * (uint32_t)k
               : what == 2 ? 8u * (uint32_t)k
               : what == 3 ? 0x80u - (uint32_t)k
               : what == 4 ? (uint32_t)k
                           : (uint32_t)(k - 1);
    }
    __host__ __device__ static constexpr uint32_t word(int what) {
        return byte_of(KBASE + 4 * REG, what) | byte_of(KBASE + 4 * REG + 1, what) << 8 |
               byte_of(KBASE + 4 * REG + 2, what) << 16 | byte_of(KBASE + 4 * REG + 3, what) << 24;
    }
};

template <int KBASE, int REG>
__device__ __forceinline__ uint32_t phase_step(uint32_t p) {
    constexpr uint32_t INC = PhaseConst<KBASE, REG>::word(0), CMP = PhaseConst<KBASE, REG>::word(1),
                       MOD = PhaseConst<KBASE, REG>::word(2);
    p += INC;
    const uint32_t ge = ((p + CMP) >> 7) & 0x01010101u; // byte >= 8k
    return p - ((ge * 0xFFu) & MOD);
}

// Phases of an item's block 0 for this lane, four k at once.  r = the item's (rel0 mod k) bytes, c = this lane's
// (32 * lane + 32 * (k - 1)) mod k bytes (computed once per thread): x = (r + c) mod k is the offset of the lane's
// bit 0 inside its chunk (of the PREVIOUS word), s = k - 1 - x the bit of the first chunk end there; returns 8 * s.
// Byte-parallel like phase_step: r + c < 2k <= 32, so no byte overflows.
template <int KBASE, int REG>
__device__ __forceinline__ uint32_t phase_init(uint32_t r, uint32_t c) {
    constexpr uint32_t CMP = PhaseConst<KBASE, REG>::word(3), MOD = PhaseConst<KBASE, REG>::word(4),
                       KM1 = PhaseConst<KBASE, REG>::word(5);
    uint32_t x = r + c;
    const uint32_t ge = ((x + CMP) >> 7) & 0x01010101u; // byte >= k
    x -= (ge * 0xFFu) & MOD;
    return (KM1 - x) << 3;
}

// One k for one 32-base word.  WITH_VA: honour the validity plane (edge blocks); interior blocks of
// a clean stretch have va == all ones for both words and skip it.
//
// "All K positions of an aligned chunk match K behind" is an all-ones-field test on D over the
// (previous | current) 64-bit pair, fields = chunks, H = top bit of each field (the chunk ends).  With
// nD = ~D (the MISMATCH plane, which is what the XORs produce) and nH = ~H:
//     Z = ~((nD & nH) + nH) & D & H  =  ~(sum | nD | nH)
// Inside a field the addend nH is all ones below the top bit, so the sum carries into the top bit exactly
// when some low position mismatches; both addends are 0 at the top bit, so no carry leaves a field.  Two
// LOP3, one 64-bit add, two LOP3 -- regardless of K.
//
// s8 = 8 * s, s = bit position (0..K-1) of the first chunk end in the PREVIOUS word (see PhaseConst).
// FULL: also require equal letter case and equal N-ness (mixed-case blocks, soft-mask edges, N runs):
// with the case and N planes in the comparison, plane equality is byte equality for A/C/G/T/N in either
// case, so only blocks with other bytes (IUPAC codes, gaps) need the byte-wise evaluation.
template <int K, int ROW, bool WITH_VA, bool FULL, bool EMIT, class Stash>
__device__ __forceinline__ void explore_k(Stash &stash, const HTable &htab, uint32_t &any_ev,
                                          const Planes3 &cur, const Planes3 &prv, uint32_t &carry_nd, uint32_t m31,
                                          int src_lane, uint32_t s8, uint32_t own) {
    const uint32_t lo_k = __funnelshift_l(prv.lo, cur.lo, K);
    const uint32_t hi_k = __funnelshift_l(prv.hi, cur.hi, K);
    uint32_t nd_hi = (cur.lo ^ lo_k) | (cur.hi ^ hi_k); // mismatch plane of this word
    if (WITH_VA) nd_hi |= ~(cur.va & __funnelshift_l(prv.va, cur.va, K));
    if (FULL) nd_hi |= (cur.cs ^ __funnelshift_l(prv.cs, cur.cs, K)) | (cur.nn ^ __funnelshift_l(prv.nn, cur.nn, K));
    // previous lane's mismatch plane (lane 0: lane 31's of the previous block)
    const uint32_t nd_lo = __shfl_sync(0xFFFFFFFFu, bitsel(m31, carry_nd, nd_hi), src_lane);
    carry_nd = nd_hi;
    if (!EMIT) return; // special block: only the carries are kept consistent
#if TIDK_EX_LDS_PRMT
    // s8 already IS the shared-window address of the entry's row start + 8 * s (phase_addr): one LDS with an immediate
    uint32_t nh_lo, nh_hi;
    asm("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(nh_lo), "=r"(nh_hi) : "r"(s8), "n"((K - 1) * 128));
    const unsigned long long nh = ((unsigned long long)nh_hi << 32) | nh_lo;
    (void)htab;
#else
    const unsigned long long nh = *reinterpret_cast<const unsigned long long *>(reinterpret_cast<const char *>(&htab.nh[K - 1][0]) + s8);
    const uint32_t nh_lo = (uint32_t)nh, nh_hi = (uint32_t)(nh >> 32);
#endif
    // one 64-bit add: IADD3 + IMAD.X (issuing the low word as mad.wide to reach the fma pipe was tried: ptxas
    // splits it into IMAD.WIDE + IADD3 + IADD3.X, one instruction more)
    const unsigned long long sum = (((unsigned long long)(nd_hi & nh_hi) << 32) | (nd_lo & nh_lo)) + nh;
    const uint32_t sum_hi = (uint32_t)(sum >> 32);
    // x = ~Z; a funnel shift commutes with the complement, so the boundary test needs no NOT
    const uint32_t x_lo = (uint32_t)sum | nd_lo | nh_lo;
    const uint32_t x_hi = sum_hi | nd_hi | nh_hi;
    const uint32_t x_prev = __funnelshift_l(x_lo, x_hi, K); // the chunk before
    // run boundaries are rare; they are only stashed here (one shared-memory slot per k and lane) and
    // turned into events after all k of the block, in ONE copy of the emission code: inlining it
    // after every k made the hot loop three times larger than the instruction cache
    // x is 1 off the chunk ends in both words, so only chunk ends survive the XOR; interior blocks
    // (WITH_VA false) own every position of the word
    // ev = (x_hi ^ x_prev) [& own] and "ev != 0" out of ONE LOP3 (lop3 with a predicate result), then a
    // predicated OR of the k's bit: which k to look at afterwards
    uint32_t ev;
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, 0, 0;\n\t"
        "lop3.or.b32 %0|p, %2, %3, %4, %5, q;\n\t"
        "@p or.b32 %1, %1, %6;\n\t}"
        : "=r"(ev), "+r"(any_ev)
        : "r"(x_hi), "r"(x_prev), "r"(WITH_VA ? own : 0xFFFFFFFFu), "n"(0x28), "n"(1u << ROW)); // 0x28 = (a ^ b) & c
    stash.ev[ROW][threadIdx.x] = ev;
}

// byte-wise evaluation of this lane's owned chunk ends (special blocks).  x0 = (rel0 mod k) + offset
// of bit 0, so the first chunk end of the word is at bit k - 1 - (x0 mod k) and the rest follow
// every k bits.
__device__ __forceinline__ void explore_k_bytes(const ExJob &job, int k, uint32_t x0, uint32_t own,
                                                uint32_t slice, const uint8_t *t, int64_t m, int64_t p) {
    for (int b = k - 1 - (int)(x0 % (uint32_t)k); b < 32; b += k) {
        if (!((own >> b) & 1u)) continue;
        const int64_t q = p + b;
        const bool cur = g_bytes(t, m, k, q), prv = g_bytes(t, m, k, q - k);
        if (cur == prv) continue;
        const uint64_t j = (uint64_t)((q + 1) / k) - 2;
        const bool joins = cur && j > 0 && chunks_equal_upper(t, k, j);
        emit_event(job, slice, (uint32_t)k, j, cur, joins);
    }
}

__device__ __forceinline__ uint32_t r0_of(const uint4 &v, int i) { // byte i of a 16-byte vector
    const uint32_t wsel = i < 4 ? v.x : i < 8 ? v.y : i < 12 ? v.z : v.w;
    return (wsel >> (8 * (i & 3))) & 0xFFu;
}

template <int KBASE, int NREG, int G = 0>
__device__ __forceinline__ void phase_step_all(uint32_t (&ph)[NREG]) {
    if constexpr (G < NREG) {
        ph[G] = phase_step<KBASE, G>(ph[G]);
        phase_step_all<KBASE, NREG, G + 1>(ph);
    }
}

// TIDK_EX_LDS_PRMT: the chunk-end table sits 256-byte aligned in shared memory, so "table address + 8 * s_K" is ONE
// PRMT (byte 0 from the phase register, bytes 1-3 from the table's shared-window address) and the row of K is the
// load's immediate -- no address add per k, and no generic-to-shared conversion inside the block loop.
template <int KBASE, int NREG, int K>
__device__ __forceinline__ uint32_t phase_addr(const uint32_t (&ph)[NREG], uint32_t htab_s) {
    if constexpr (K >= KBASE && K < KBASE + 4 * NREG)
        return __byte_perm(ph[(K - KBASE) / 4], htab_s, 0x7650u + (uint32_t)((K - KBASE) % 4));
    else
        return htab_s;
}

template <int KBASE, int NREG, int G = 0>
__device__ __forceinline__ void phase_init_all(uint32_t (&ph)[NREG], const uint32_t (&rw)[4], const uint32_t (&lane_c)[NREG]) {
    if constexpr (G < NREG) {
        ph[G] = phase_init<KBASE, G>(rw[(KBASE - 1) / 4 + G], lane_c[G]);
        phase_init_all<KBASE, NREG, G + 1>(ph, rw, lane_c);
    }
}

// 8 * s_K out of the packed phase registers (one PRMT); K outside the registers' range never runs
template <int KBASE, int NREG, int K>
__device__ __forceinline__ uint32_t phase_byte(const uint32_t (&ph)[NREG]) {
    if constexpr (K >= KBASE && K < KBASE + 4 * NREG)
        return __byte_perm(ph[(K - KBASE) / 4], 0u, 0x4440u + (uint32_t)((K - KBASE) % 4));
    else
        return 0u;
}

// q / k for a q that IS a multiple of k (chunk end + 1), k <= 16, quotient < 2^32: shift out k's twos, multiply by the
// inverse of its odd part modulo 2^32.  Two constant-bank reads, a shift and a multiply.
__constant__ uint32_t kExactInv[17] = {0u, 1u, 1u, 0xAAAAAAABu, 1u, 0xCCCCCCCDu, 0xAAAAAAABu, 0xB6DB6DB7u, 1u,
                                       0x38E38E39u, 0xCCCCCCCDu, 0xBA2E8BA3u, 0xAAAAAAABu, 0xC4EC4EC5u, 0xB6DB6DB7u, 0xEEEEEEEFu, 1u};
__constant__ uint32_t kExactShift[17] = {0u, 0u, 1u, 0u, 2u, 0u, 1u, 0u, 3u, 0u, 1u, 0u, 2u, 0u, 1u, 0u, 4u};
__device__ __forceinline__ uint32_t div_exact(uint32_t q, uint32_t k) { return (q >> kExactShift[k]) * kExactInv[k]; }

// KMIN > 0: the k range is a compile-time constant (default 5..12); KMIN == 0: run-time range.
#ifndef TIDK_EX_MINB
#define TIDK_EX_MINB 3
#endif
// PLANES: the batch's resident bit-planes are the source.  A block whose 32 words are all plain (one bit per word,
// built with the planes) IS its lo / hi words: no ASCII is read and nothing is transposed; any other block -- lower
// case, N, IUPAC codes, the tail behind the last base -- is taken from the ASCII copy by the general path below,
// exactly as in the ASCII variant.
template <int KMIN, int KMAX, bool PLANES = false>
__global__ void __launch_bounds__(256, TIDK_EX_MINB) explore_planes_kernel(const ExJob job) {
    __shared__ EvStash<(KMIN > 0 ? KMAX - KMIN + 1 : kPlaneMaxK)> stash;
    __shared__ __align__(256) HTable htab;
    constexpr int KBASE = KMIN > 0 ? KMIN : 1;                         // first k of the packed phase registers
    constexpr int NREG = KMIN > 0 ? (KMAX - KMIN) / 4 + 1 : kPlaneMaxK / 4;
    fill_htable(htab);
    __syncthreads();
    uint32_t htab_s; // shared-window address of the table, low byte 0 (256-byte aligned); volatile: computed once, not per block
    asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(htab_s) : "l"(&htab));
    const int lane = threadIdx.x & 31;
    const uint32_t m31 = lane == 31 ? 0xFFFFFFFFu : 0u; // look-behind: lane 31 sends its previous-block value
    const int src_lane = (lane + 31) & 31;
    uint32_t lane_c[NREG]; // byte b of lane_c[g]: (32 * lane + 32 * (k - 1)) mod k for k = KBASE + 4g + b (see phase_init)
#pragma unroll
    for (int g = 0; g < NREG; g++) {
        lane_c[g] = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = KBASE + 4 * g + b;
            if (k <= kPlaneMaxK) lane_c[g] |= ((32u * (uint32_t)lane + 32u * (uint32_t)(k - 1)) % (uint32_t)k) << (8 * b);
        }
    }
    uint32_t nonascii = 0;
    // Work queue, pipelined like the window counter's: the ticket of the item after next is drawn while
    // an item is scanned, the next item's block 0 address is fetched at the start of the current item,
    // and its first 1 KiB is requested before the current item's last block is evaluated -- a warp never
    // sits through ticket -> descriptor -> data round trips between items.
    // The grid is the resident set: the first two items of every warp are assigned statically, the rest come from
    // kQueues ticket counters (queue q hands out the items q, q + kQueues, ... behind the static ones), drawn with
    // atom.inc two items ahead and not looked at before the end of the item (see window_counts.cuh).
    const unsigned long long n_warps = (unsigned long long)gridDim.x * 8;
    const unsigned long long gw = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned long long item = gw, item_next = gw + n_warps;
    const unsigned long long ticket0 = 2 * n_warps + (gw & (kQueues - 1));
    uint32_t *const queue = reinterpret_cast<uint32_t *>(job.work_counter + (gw & (kQueues - 1)) * kQueueStride);
    SlotState slots = {0ull, 0ull, 0u, 0u, false};
    uint32_t w[8];
    uint32_t n_lo = 0, n_hi = 0, n_f0 = 0, n_f1 = 0; // PLANES: the next block's words and plain bits, requested a block ahead
    auto request = [&](uint64_t byte_off) {       // first lane word of a block -> this lane's data
        if constexpr (PLANES) {
            const uint32_t *pw = plane_word_ptr(job.planes, (byte_off >> 5) + (uint32_t)lane);
            n_lo = __ldg(pw); n_hi = __ldg(pw + kPackWords);
            n_f0 = __ldg(job.plainbits + (byte_off >> 10)); n_f1 = __ldg(job.plainbits + (byte_off >> 10) + 1);
        } else {
            ld256_stream(job.seq + byte_off + (uint32_t)lane * 32u, w);
        }
    };
    const unsigned long long n_items = job.n_items_dev ? *job.n_items_dev : job.n_items;
    if (item < n_items) request(job.items[item].scan0);
    while (item < n_items) {
        const ExItem *ip = job.items + item;
        const bool have_next = item_next < n_items;
        uint64_t scan0_next = 0;
        if (have_next) scan0_next = job.items[item_next].scan0;
        uint32_t t2 = 0;
        if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0xffffffff;" : "=r"(t2) : "l"(queue) : "memory");
        const uint64_t scan0 = ip->scan0;
        const int64_t rel0 = ip->rel0;
        const int64_t m = (int64_t)ip->slice_len;
        const uint32_t slice = ip->slice;
        const uint32_t nblk = ip->nblk;
        const int own_lo = ip->own_lo, own_hi = ip->own_hi, v_lo = ip->v_lo, v_hi = ip->v_hi;
        const uint4 rq0 = *reinterpret_cast<const uint4 *>(ip->r0);
        const uint8_t *t = job.seq + ((int64_t)scan0 - rel0); // slice start

        Planes3 own_prev = {0u, 0u, 0u, 0u, 0u};
        bool had_n_prev = false;
        uint32_t carry_nd[kPlaneMaxK]; // mismatch planes of lane 31's word of the previous block
#pragma unroll
        for (int i = 0; i < kPlaneMaxK; i++) carry_nd[i] = 0xFFFFFFFFu;
        // chunk phases of block 0 for this lane: s_k = (k-1) - (rel0 + 32 * lane - 32) mod k, kept as 8 * s_k
        uint32_t ph[NREG];
        static_assert((KBASE - 1) % 4 == 0, "the r0 bytes of a phase register are one word of the descriptor's r0 vector");
        {
            const uint32_t rw[4] = {rq0.x, rq0.y, rq0.z, rq0.w}; // r0[k - 1], k = 1..16
            phase_init_all<KBASE, NREG>(ph, rw, lane_c);
        }
        bool dirty_prev = false;
        int case_prev = -1; // -1 unknown (first block), 0 all upper, 1 all lower, 2 mixed
        bool prev_last_plain = false; // lane 31's word of the previous block: in the slice, A/C/G/T, upper case
        const uint32_t nlead = own_lo >= 1024 ? 2u : 1u; // a later tile starts one block before its owned range
        // the slice may end up to kmax + 31 positions before the owned range does: then it ends in the block
        // before the last one
        const uint32_t ntail = v_hi <= (int)((nblk - 1u) << 10) ? 2u : 1u;

        const uint32_t fsh = (uint32_t)(scan0 >> 5) & 31u; // PLANES: bit of the block's first word in its plain-bits word
        if (nblk == 0 && have_next) request(scan0_next); // (never: every tile owns a position)
        for (uint32_t blk = 0; blk < nblk; blk++) {
            Nibbles nib;
            PlaneCore core;
            uint32_t plain_lanes;
            if constexpr (PLANES) {
                core.lo = n_lo; core.hi = n_hi; core.b0 = 0u; core.bad = 0u;
                plain_lanes = __funnelshift_r(n_f0, n_f1, fsh);
            } else {
                nib = to_nibbles(w); // w is dead from here: refill it
            }
            if (blk + 1 < nblk) request(scan0 + ((uint64_t)(blk + 1) << 10));
            else if (have_next) request(scan0_next);
            if constexpr (!PLANES) core = extract_core(nib);
            const int prel = (int)(blk << 10) + lane * 32;
            // slice edges (and the item's owned range) can only fall into the first block (the first two
            // of a later tile, which starts one block early) and the last block (or two) of an item
            const bool edge = blk < nlead || blk + ntail >= nblk;
            const int64_t p = rel0 + prel;
            uint32_t any_ev = 0;
            uint32_t own = 0xFFFFFFFFu;
            bool need_cs = false, special = false;
            Planes3 cur, prv;

            // The common block -- interior, every base one of A C G T in upper case, and the block before it
            // of the same kind -- is recognised with ONE warp vote and needs nothing but the lo / hi planes.
            uint32_t any_l = 0u;
            if constexpr (!PLANES) {
                any_l = nib.Y[0] | nib.Y[1] | nib.Y[2] | nib.Y[3]; // ASCII bit 5 = bit 1 of the high nibbles
                const bool lane_plain = core.bad == 0u && (core.b0 ^ (core.hi & ~core.lo)) == 0xFFFFFFFFu &&
                                        (any_l & 0x22222222u) == 0u;
                plain_lanes = __ballot_sync(0xFFFFFFFFu, lane_plain);
            }
            // (after a block whose in-slice bases were all upper-case A/C/G/T the case / N / dirty state is
            // the initial one, whichever path evaluated it; the fast path leaves it untouched)
            const bool all_plain = plain_lanes == 0xFFFFFFFFu, calm = !had_n_prev && !dirty_prev;
            const bool fast = !edge && all_plain && prev_last_plain && case_prev == 0 && calm;
            // An edge block whose 32 x 32 bytes are all plain (the bytes beyond the slice edge belong to the
            // neighbouring slice or read and usually are) only needs the slice masks on top: the word behind
            // lane 0 is either not in the slice at all (first block of an item: validity 0) or plain.
            const bool edge_plain = edge && all_plain && calm && (blk == 0 || (prev_last_plain && case_prev == 0));
            uint32_t inslice = 0xFFFFFFFFu;
            cur.lo = core.lo; cur.hi = core.hi;
            const uint32_t xoff = (uint32_t)prel; // < 2^15
#define TIDK_EXK(K, VA, CS, EM)                                                                   \
    if ((KMIN > 0) ? (K >= KMIN && K <= KMAX) : (job.kmin <= K && K <= job.kmax)) {              \
        explore_k<K, (K >= KBASE ? K - KBASE : 0), VA, CS, EM>(stash, htab, any_ev, cur, prv, carry_nd[K - 1], m31, src_lane,  \
                                 TIDK_EX_LDS_PRMT ? phase_addr<KBASE, NREG, K>(ph, htab_s) : phase_byte<KBASE, NREG, K>(ph), own); \
    }
#define TIDK_EXALL(VA, CS, EM)                                                                    \
    TIDK_EXK(1, VA, CS, EM) TIDK_EXK(2, VA, CS, EM) TIDK_EXK(3, VA, CS, EM) TIDK_EXK(4, VA, CS, EM)     \
    TIDK_EXK(5, VA, CS, EM) TIDK_EXK(6, VA, CS, EM) TIDK_EXK(7, VA, CS, EM) TIDK_EXK(8, VA, CS, EM)     \
    TIDK_EXK(9, VA, CS, EM) TIDK_EXK(10, VA, CS, EM) TIDK_EXK(11, VA, CS, EM) TIDK_EXK(12, VA, CS, EM)  \
    TIDK_EXK(13, VA, CS, EM) TIDK_EXK(14, VA, CS, EM) TIDK_EXK(15, VA, CS, EM) TIDK_EXK(16, VA, CS, EM)
            if (fast) {
                // self-contained: nothing of the case / N / validity state is touched.  Only lane 31's own_prev is ever
                // read (lane_lookbehind), and its va / cs / nn already are ~0 / 0 / 0 here: prev_last_plain says the
                // word was in the slice and A/C/G/T, case_prev == 0 that it was upper case; both stay as they are.
                cur.va = 0xFFFFFFFFu; cur.cs = 0u; cur.nn = 0u;
                prv.lo = lane_lookbehind(cur.lo, own_prev.lo, lane);
                prv.hi = lane_lookbehind(cur.hi, own_prev.hi, lane);
                prv.va = 0xFFFFFFFFu; prv.cs = 0u; prv.nn = 0u;
                own_prev.lo = cur.lo; own_prev.hi = cur.hi;
                TIDK_EXALL(false, false, true)
            } else if (edge_plain) {
                // self-contained like the plain path, plus the slice masks.  Lane 31's case / N planes of the block
                // before are 0 already (block 0: the item's initial state; later: prev_last_plain && case_prev == 0)
                // and stay 0, the N / dirty state stays calm.
                inslice = range_mask32(v_lo - prel, v_hi - prel);
                own = range_mask32(own_lo - prel, own_hi - prel);
                cur.va = inslice; cur.cs = 0u; cur.nn = 0u;
                prv.lo = lane_lookbehind(cur.lo, own_prev.lo, lane);
                prv.hi = lane_lookbehind(cur.hi, own_prev.hi, lane);
                prv.va = lane_lookbehind(cur.va, own_prev.va, lane);
                prv.cs = 0u; prv.nn = 0u;
                case_prev = 0; // all upper case
                own_prev.lo = cur.lo; own_prev.hi = cur.hi; own_prev.va = cur.va;
                prev_last_plain = __shfl_sync(0xFFFFFFFFu, inslice, 31) == 0xFFFFFFFFu; // (all 32 lanes are plain here)
                TIDK_EXALL(true, false, true)
            } else {
            {
                if constexpr (PLANES) { // not a plain block, or the state behind one: the bytes themselves
                    uint32_t wa[8];
                    ld256_stream(job.seq + scan0 + ((uint64_t)blk << 10) + (uint32_t)lane * 32u, wa);
                    nib = to_nibbles(wa);
                    core = extract_core(nib);
                    any_l = nib.Y[0] | nib.Y[1] | nib.Y[2] | nib.Y[3];
                }
                uint32_t na = 0;
                const Planes P = finish_planes(nib, core, __any_sync(0xFFFFFFFFu, core.bad != 0u), na);
                if (edge) {
                    inslice = range_mask32(v_lo - prel, v_hi - prel);
                    own = range_mask32(own_lo - prel, own_hi - prel);
                }
                nonascii |= na & inslice;
                cur.nn = P.isn & inslice;
                cur.va = ((P.b0 ^ (P.hi & ~P.lo)) & P.ok & inslice) | cur.nn;

                // Letter case.  Without the case plane, plane equality is byte equality only under uniform
                // case, so a block whose case is mixed, or differs from the previous block's, compares the
                // case plane as well (FULL variant).  A lane that straddles the slice edge is judged on all
                // its 32 bytes, which can only make the block "mixed" more often (still exact).
                const uint32_t all_l = nib.Y[0] & nib.Y[1] & nib.Y[2] & nib.Y[3];
                const bool lane_in = inslice != 0;
                const bool lane_any_lower = (any_l & 0x22222222u) != 0;
                const bool lane_all_lower = (all_l & 0x22222222u) == 0x22222222u;
                const bool blk_upper = __all_sync(0xFFFFFFFFu, !lane_in || !lane_any_lower);
                const bool blk_lower = __all_sync(0xFFFFFFFFu, !lane_in || lane_all_lower);
                const int case_now = blk_upper ? 0 : (blk_lower ? 1 : 2);
                const bool had_n = P.exact && __any_sync(0xFFFFFFFFu, cur.nn != 0);
                need_cs = case_now == 2 || case_now != case_prev || had_n || had_n_prev; // FULL variant
                case_prev = case_now;
                had_n_prev = had_n;
                // bytes other than A/C/G/T/N (IUPAC codes, gaps, ...): their identity is not in the planes,
                // so this block and the next one are evaluated byte-wise
                const bool dirty_now = __any_sync(0xFFFFFFFFu, (~cur.va & inslice) != 0);
                special = dirty_now || dirty_prev;
                dirty_prev = dirty_now;
                // the case plane: extracted when needed, otherwise the constant the block's state implies
                cur.cs = need_cs ? nibble_plane<1>(nib.Y) : (case_now == 1 ? 0xFFFFFFFFu : 0u);

                prv.lo = lane_lookbehind(cur.lo, own_prev.lo, lane);
                prv.hi = lane_lookbehind(cur.hi, own_prev.hi, lane);
                prv.va = lane_lookbehind(cur.va, own_prev.va, lane);
                prv.cs = need_cs ? lane_lookbehind(cur.cs, own_prev.cs, lane) : 0u;
                prv.nn = need_cs ? lane_lookbehind(cur.nn, own_prev.nn, lane) : 0u;
            }
            own_prev = cur;
            prev_last_plain = (plain_lanes >> 31) != 0u && (!edge || __shfl_sync(0xFFFFFFFFu, inslice, 31) == 0xFFFFFFFFu);
            // four straight-line variants, chosen per block (warp-uniform):
            //   interior of a clean stretch (every base of both words valid, uniform case) -> lo/hi only
            //   clean block at a slice edge                                                -> + va plane
            //   mixed letter case / case change / N present                                -> + va + case + N planes
            //   special block (bytes other than A C G T N)                                 -> carries only + byte-wise
            bool lohi_only = false;
            if (!special && !need_cs && !edge) // (edge blocks own only part of their positions: VA variant)
                lohi_only = __all_sync(0xFFFFFFFFu, (cur.va & prv.va) == 0xFFFFFFFFu);
            if (lohi_only) { TIDK_EXALL(false, false, true) }
            else if (special) { TIDK_EXALL(true, false, false) }
            else if (need_cs) { TIDK_EXALL(true, true, true) }
            else { TIDK_EXALL(true, false, true) }
            } // !fast
#undef TIDK_EXALL
#undef TIDK_EXK
            if (__any_sync(0xFFFFFFFFu, any_ev != 0u)) { // some lane found run boundaries in this block: emitted by the warp together
                uint32_t pend = any_ev, cur_ev = 0u, cur_k = 0u;
                for (;;) {
                    if (cur_ev == 0u && pend != 0u) { // next k of this lane that has boundaries
                        const uint32_t row = (uint32_t)(__ffs(pend) - 1);
                        pend &= pend - 1;
                        cur_ev = stash.ev[row][threadIdx.x];
                        cur_k = (uint32_t)KBASE + row;
                    }
                    const bool has = cur_ev != 0u;
                    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, has);
                    if (mask == 0u) break;
                    const uint32_t n = (uint32_t)__popc(mask);
                    if (slots.left < n) next_range(job, lane, slots);
                    if (has) {
                        const int b = __ffs(cur_ev) - 1;
                        cur_ev &= cur_ev - 1;
                        const uint32_t k = cur_k;
                        const uint64_t q1 = (uint64_t)(p + b) + 1; // chunk end + 1, a multiple of k
                        const uint64_t j = (q1 >> 32 ? q1 / k : (uint64_t)div_exact((uint32_t)q1, k)) - 2;
                        // a run start in a uniform-case block never joins the previous chunk (they differ
                        // as bytes and share the case, so they differ after upper-casing too); with mixed
                        // case the two chunks are compared after upper-casing (explore.rs:212 vs :324)
                        // The boundary says G[q] != G[q - k].  Which of the two holds is not needed here: the
                        // boundaries of a (slice, k) alternate start / end from a start on (G is 0 before the first
                        // full chunk pair and beyond the slice), so after the sort an event's type is the parity of
                        // its rank and the run builder pairs (2i, 2i + 1).  Only `joins` needs it, in mixed-case
                        // blocks, where it is read back from the slice (the planes' verdict and the bytes agree
                        // wherever the plane path emits); elsewhere the type bit is left 0.
                        const bool start = !need_cs || g_bytes(t, m, (int)k, p + b);
                        const bool joins = need_cs && start && j > 0 && chunks_equal_upper(t, (int)k, j);
                        const unsigned long long idx = slots.base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                        if (idx < job.cap_events) {
                            const uint32_t type = start ? 0u : 1u, jn = joins ? 1u : 0u;
                            if (job.packed) ((uint64_t *)job.events)[idx] = pack_event(job.fmt, slice, k, j, type, jn);
                            else {
                                RunEvent ev;
                                ev.slice = slice; ev.k = (uint8_t)k; ev.type = (uint8_t)type; ev.joins = (uint8_t)jn; ev.pad = 0;
                                ev.chunk = j;
                                job.events[idx] = ev;
                            }
                        }
                    }
                    slots.base += n;
                    slots.left -= n;
                }
                if (!slots.pending && slots.left < kExReserve / 2) { // ask for the next range while this one lasts
                    reserve_slots(job, lane, slots.raw);
                    slots.pending = true;
                }
            }
            if (special) { // exact byte-wise evaluation of this block's owned chunk ends
                const uint32_t k0 = KMIN > 0 ? KMIN : job.kmin, k1 = KMIN > 0 ? KMAX : job.kmax;
                for (uint32_t k = k0; k <= k1; k++)
                    explore_k_bytes(job, (int)k, r0_of(rq0, (int)k - 1) + xoff, own, slice, t, m, p);
            }
            phase_step_all<KBASE, NREG>(ph); // the next block lies 1024 bases further on
        }
        item = item_next;
        item_next = ticket0 + (unsigned long long)__shfl_sync(0xFFFFFFFFu, t2, 0) * kQueues;
    }
    // slots this warp was handed and did not use
    if ((uint32_t)lane < slots.left) store_sentinel(job, slots.base + (uint32_t)lane);
    slots.unused += slots.left;
    if (slots.pending) {
        const unsigned long long b2 = __shfl_sync(0xFFFFFFFFu, slots.raw, 0);
        store_sentinel(job, b2 + (uint32_t)lane);
        slots.unused += kExReserve;
    }
    if (lane == 0 && slots.unused) atomicAdd(job.n_unused, (unsigned long long)slots.unused);
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) atomicOr(job.err_flags, 1u);
}

} // namespace tidk
