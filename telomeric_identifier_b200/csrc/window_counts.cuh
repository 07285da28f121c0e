// window_counts.cuh -- per-window forward / reverse-complement repeat counting.
//
// Replaces search::write_window_counts (search.rs:83-151) and
// finder::write_window_counts (finder.rs:111-178): for every window of every
// record, the number of (overlapping) start positions at which upper(text)
// equals the query, resp. its reverse complement, with matching confined to the
// window (an occurrence that straddles a window boundary is lost).
//
// Work item = (window, 32 KiB segment of it); windows up to 32 KiB are one item,
// so with the default W = 10000 a warp owns a whole window, needs no halo from
// anywhere, and writes its two counts with plain stores.  A small prep kernel
// turns (record offsets, W) into one 32-byte descriptor per item; the counting
// kernel pulls items from an atomic queue.  A warp walks its item in 1 KiB
// blocks (32 lanes x one 256-bit load), keeps the previous block's planes for
// the (L-1)-base look-behind, and counts matches by their END position:
//     match ending at e  <=>  text[e-L+1 .. e] == query.
// Window confinement is applied to the data, not the result: bases outside
// [beg, end) are marked invalid, so no match can reach across the boundary.
// The load of the next block -- or of the next item's first block -- is issued
// before the current block is matched, the next item's descriptor one item
// ahead, and the queue ticket two items ahead, so the memory pipe never drains
// at a window boundary.
#pragma once
#include <type_traits>

#include "planes.cuh"

#ifndef TIDK_L2_PREFETCH
#define TIDK_L2_PREFETCH 2 // blocks ahead (0 = off); the register load runs one block ahead
#endif

namespace tidk {

constexpr int kMaxFastLen = 32;        // plane kernels: one look-behind word
constexpr int kMaxFastQueries = 16;    // queries per fused multi-query launch
constexpr uint32_t kSegBytes = 32768;  // work-item span for windows larger than this
constexpr int kWarpsPerBlock = 8;
// Work queues per launch.  One shared ticket counter is an L2 atomic unit working on ONE address: at C4 size (309 000
// items in 0.52 ms, a ticket per item from 4 736 warps) the kernel ran at either of two rates, 0.52 or 0.94 ms per
// pass, the slower one bound by that unit (ncu: the wait for the ticket was the top stall).  Eight counters on eight
// lines, a warp bound to one of them, queue q handing out the items q, q + 8, q + 16, ... behind the statically
// assigned ones: an eighth of the traffic per address, and runs of slow items (N gaps) still spread over all queues.
constexpr int kQueues = 8;
constexpr int kQueueStride = 16;       // in 64-bit words: 128 bytes between counters

struct __align__(32) ItemDesc {
    uint64_t scan0;       // absolute offset of block 0 (multiple of 32)
    uint64_t out;         // output element index for query 0 of the launch's table
    uint32_t nwin_r;      // windows in this record = output stride between queries
    uint16_t vbeg, vend;  // bases inside the window: [vbeg, vend) relative to scan0
    uint16_t cbeg, cend;  // counted END positions:   [cbeg, cend) relative to scan0
    uint16_t nblk;        // 1 KiB blocks to scan (0: nothing to do)
    uint16_t flags;       // bit 0: window has several items -> accumulate with atomics
};
static_assert(sizeof(ItemDesc) == 32, "ItemDesc is one 256-bit load");

struct PrepJob {
    const uint64_t *rec_off;  // [n_records + 1]
    const uint64_t *win_base; // [n_records + 1] windows before record r
    ItemDesc *items;
    uint64_t window, n_items;
    uint32_t n_records, segs_per_window, nq_total;
};

__global__ void __launch_bounds__(256) window_items_kernel(const PrepJob job) {
    const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= job.n_items) return;
    const uint64_t win = item / job.segs_per_window;
    const uint32_t seg = (uint32_t)(item - win * job.segs_per_window);
    uint32_t r_lo = 0, r_hi = job.n_records;
    while (r_hi - r_lo > 1) { // last record with win_base[r] <= win
        uint32_t mid = (r_lo + r_hi) >> 1;
        if (job.win_base[mid] <= win) r_lo = mid; else r_hi = mid;
    }
    const uint64_t rec_beg = job.rec_off[r_lo], rec_end = job.rec_off[r_lo + 1];
    const uint64_t wi = win - job.win_base[r_lo];
    const uint64_t beg = rec_beg + wi * job.window;
    const uint64_t end = (beg + job.window < rec_end) ? beg + job.window : rec_end;
    const uint64_t sbeg = beg + (uint64_t)seg * kSegBytes;
    ItemDesc d;
    d.scan0 = 0; d.nblk = 0; d.vbeg = d.vend = d.cbeg = d.cend = 0;
    d.out = job.win_base[r_lo] * job.nq_total + wi;
    d.nwin_r = (uint32_t)(job.win_base[r_lo + 1] - job.win_base[r_lo]);
    d.flags = job.segs_per_window > 1 ? 1 : 0;
    if (sbeg < end) {
        const uint64_t send = (sbeg + kSegBytes < end) ? sbeg + kSegBytes : end;
        // Later segments start one lane word (32 bases) early: the look-behind is at most 31 bases, the
        // early word lies before the counted range, and lane 1 takes its look-behind from lane 0.
        uint64_t scan0 = sbeg & ~31ull;
        if (seg > 0) scan0 -= 32;
        d.scan0 = scan0;
        d.nblk = (uint16_t)((send - scan0 + 1023) >> 10);
        d.vbeg = (uint16_t)(beg > scan0 ? beg - scan0 : 0);
        d.vend = (uint16_t)(end - scan0 < 65535 ? end - scan0 : 65535);
        d.cbeg = (uint16_t)(sbeg - scan0);
        d.cend = (uint16_t)(send - scan0);
    }
    job.items[item] = d;
}

constexpr int kPackShift = 17;                    // 2^17 words = 4 MiB of sequence per packed chunk
constexpr uint32_t kPackWords = 1u << kPackShift; // chunk layout: [lo words][hi words][va words]

struct WindowJob {
    const uint8_t *seq;      // device, 256-B aligned, zero padding behind the data (ASCII source)
    const uint32_t *packed;  // device bit-planes packed by the host (packed source), else nullptr
    const ItemDesc *items;
    uint64_t *out_fwd;       // [record][query][window]
    uint64_t *out_rev;
    unsigned long long *work_counter; // kQueues ticket counters, one per 128-byte line, zero at launch
    unsigned long long *next_counter; // the other set: zeroed by this launch for the next one (no memset per call)
    uint32_t *err_flags;     // mapped host word; set to 1 when a non-ASCII byte is seen
    uint64_t n_items;
    uint32_t q0;             // first query index of this launch
    uint32_t nq;             // queries in this launch
};

// ------------------------------------------------------------------ matchers
// A matcher turns the planes of one block (+ the look-behind carry) into per-lane
// match counts.  Base codes in static patterns: A=0 C=1 G=2 T=3, complement = 3 - c.

// Pattern known at compile time: equality planes, one funnel shift + one AND per
// pattern position and strand, LOP3 truth tables folded into immediates.
template <int L, uint64_t F>
struct StaticMatcher {
    static constexpr int NQ = 1;
    struct Params { uint32_t unused; };
    struct Carry { uint32_t e[4]; };
    __host__ __device__ static constexpr int fwd_at(int j) { return (int)((F >> (2 * j)) & 3); }
    __device__ __forceinline__ static void reset(Carry &c) { c.e[0] = c.e[1] = c.e[2] = c.e[3] = 0; }

    // The forward chain is ANDed from the motif's FIRST base (largest shift) down, the reverse-complement
    // chain from its LAST base (shift 0) up: motifs of one clade that share a prefix (AACCC..., six of
    // length 11 in Hymenoptera) then share the leading sub-chains of both strands, which the compiler
    // merges when several motifs are fused into one kernel.
    template <int S>
    __device__ __forceinline__ static void steps(const uint32_t (&e)[4], const uint32_t (&p)[4],
                                                 uint32_t &mf, uint32_t &mr) {
        if constexpr (S < L) {
            constexpr int SF = L - 1 - S;         // forward strand: shift of the motif's base S
            constexpr int fb = fwd_at(S);         // forward letter SF bases behind the end
            constexpr int rb = 3 - fwd_at(S);     // revcomp letter S bases behind the end
            mf &= SF == 0 ? e[fb] : __funnelshift_l(p[fb], e[fb], SF);
            mr &= S == 0 ? e[rb] : __funnelshift_l(p[rb], e[rb], S);
            steps<S + 1>(e, p, mf, mr);
        }
    }

    __device__ __forceinline__ static void step(const Params &, const Planes &P, bool special,
                                                uint32_t cmask, int lane, Carry &carry,
                                                uint32_t *cf, uint32_t *cr) {
        uint32_t e[4], p[4];
        e[0] = ~P.lo & ~P.hi & P.b0; // A
        e[1] = P.lo & ~P.hi & P.b0;  // C
        e[2] = P.lo & P.hi & P.b0;   // G
        e[3] = ~P.lo & P.hi & ~P.b0; // T
        if (special) {
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] &= P.ok;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = lane_lookbehind(e[i], carry.e[i], lane);
        uint32_t mf = 0xFFFFFFFFu, mr = 0xFFFFFFFFu;
        steps<0>(e, p, mf, mr);
#pragma unroll
        for (int i = 0; i < 4; i++) carry.e[i] = e[i];
        cf[0] += __popc(mf & cmask);
        cr[0] += __popc(mr & cmask);
    }
};

// Several compile-time motifs in ONE pass over the sequence (a multi-repeat clade of `tidk find`,
// e.g. Hymenoptera's 15 repeats): the planes are extracted once, the shifted equality planes are
// shared between all motifs and both strands (common sub-expressions), and every motif costs only
// its ANDs and two popcounts.  The reference makes one full pass per repeat (finder.rs:121-176).
template <int L_, uint64_t F_>
struct Motif {
    static constexpr int L = L_;
    static constexpr uint64_t F = F_;
};

template <class... Ms>
struct StaticMultiMatcher {
    static constexpr int NQ = (int)sizeof...(Ms);
    struct Params { uint32_t unused; };
    struct Carry { uint32_t e[4]; };
    __device__ __forceinline__ static void reset(Carry &c) { c.e[0] = c.e[1] = c.e[2] = c.e[3] = 0; }

    template <class M>
    __device__ __forceinline__ static void one(const uint32_t (&e)[4], const uint32_t (&p)[4], uint32_t cmask,
                                               uint32_t &cf, uint32_t &cr) {
        uint32_t mf = 0xFFFFFFFFu, mr = 0xFFFFFFFFu;
        StaticMatcher<M::L, M::F>::template steps<0>(e, p, mf, mr);
        cf += __popc(mf & cmask);
        cr += __popc(mr & cmask);
    }

    __device__ __forceinline__ static void step(const Params &, const Planes &P, bool special,
                                                uint32_t cmask, int lane, Carry &carry,
                                                uint32_t *cf, uint32_t *cr) {
        uint32_t e[4], p[4];
        e[0] = ~P.lo & ~P.hi & P.b0; // A
        e[1] = P.lo & ~P.hi & P.b0;  // C
        e[2] = P.lo & P.hi & P.b0;   // G
        e[3] = ~P.lo & P.hi & ~P.b0; // T
        if (special) {
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] &= P.ok;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = lane_lookbehind(e[i], carry.e[i], lane);
#pragma unroll
        for (int i = 0; i < 4; i++) carry.e[i] = e[i];
        int q = 0;
        ((one<Ms>(e, p, cmask, cf[q], cr[q]), q++), ...);
    }
};

// One strand pattern for the runtime matchers: bit s of plo / phi is the expected
// lo / hi plane bit s bases behind the match end (s = 0 is the LAST base).
struct StrandPattern {
    uint32_t plo, phi;
};

struct RuntimeQueries {
    StrandPattern fwd[kMaxFastQueries];
    StrandPattern rev[kMaxFastQueries];
    uint8_t len[kMaxFastQueries];
    uint32_t nq;
};

// Pattern known only at run time: mismatch accumulation against broadcast pattern bits.
// LT > 0: one query of length LT (unrolled).  LT == 0: nq queries, run-time lengths.
template <int LT>
struct RuntimeMatcher {
    static constexpr int NQ = LT > 0 ? 1 : kMaxFastQueries;
    using Params = RuntimeQueries;
    struct Carry { uint32_t lo, hi, va; };
    __device__ __forceinline__ static void reset(Carry &c) { c.lo = c.hi = c.va = 0; }

    __device__ __forceinline__ static void one(const StrandPattern &f, const StrandPattern &r, int s,
                                               uint32_t lo_s, uint32_t hi_s, uint32_t &mf, uint32_t &mr) {
        uint32_t eflo = 0u - ((f.plo >> s) & 1u), efhi = 0u - ((f.phi >> s) & 1u);
        uint32_t erlo = 0u - ((r.plo >> s) & 1u), erhi = 0u - ((r.phi >> s) & 1u);
        mf |= (lo_s ^ eflo) | (hi_s ^ efhi);
        mr |= (lo_s ^ erlo) | (hi_s ^ erhi);
    }

    __device__ __forceinline__ static void step(const Params &q, const Planes &P, bool special,
                                                uint32_t cmask, int lane, Carry &carry,
                                                uint32_t *cf, uint32_t *cr) {
        uint32_t va = P.b0 ^ (P.hi & ~P.lo);
        if (special) va &= P.ok;
        const uint32_t plo = lane_lookbehind(P.lo, carry.lo, lane);
        const uint32_t phi = lane_lookbehind(P.hi, carry.hi, lane);
        const uint32_t pva = lane_lookbehind(va, carry.va, lane);
        carry.lo = P.lo; carry.hi = P.hi; carry.va = va;
        if constexpr (LT > 0) {
            uint32_t mf = 0, mr = 0, inv = 0;
#pragma unroll
            for (int s = 0; s < LT; s++) {
                uint32_t lo_s = s == 0 ? P.lo : __funnelshift_l(plo, P.lo, s);
                uint32_t hi_s = s == 0 ? P.hi : __funnelshift_l(phi, P.hi, s);
                uint32_t va_s = s == 0 ? va : __funnelshift_l(pva, va, s);
                inv |= ~va_s;
                one(q.fwd[0], q.rev[0], s, lo_s, hi_s, mf, mr);
            }
            cf[0] += __popc(~(mf | inv) & cmask);
            cr[0] += __popc(~(mr | inv) & cmask);
        } else {
#pragma unroll 1
            for (uint32_t i = 0; i < q.nq; i++) {
                uint32_t mf = 0, mr = 0, inv = 0;
                const int Lq = q.len[i];
                for (int s = 0; s < Lq; s++) {
                    uint32_t lo_s = __funnelshift_l(plo, P.lo, s);
                    uint32_t hi_s = __funnelshift_l(phi, P.hi, s);
                    uint32_t va_s = __funnelshift_l(pva, va, s);
                    inv |= ~va_s;
                    one(q.fwd[i], q.rev[i], s, lo_s, hi_s, mf, mr);
                }
                const uint32_t a = __popc(~(mf | inv) & cmask), b = __popc(~(mr | inv) & cmask);
#pragma unroll
                for (int t = 0; t < kMaxFastQueries; t++) // keep the counters in registers
                    if (t == (int)i) { cf[t] += a; cr[t] += b; }
            }
        }
    }
};

// ------------------------------------------------------------------ sources
// Where a lane's 32 bases come from.  AsciiSource: the sequence as uploaded (1 byte per base), turned
// into planes in registers.  PackedSource: planes built on the host by the upload path of
// tidk_cuda_window_counts (lo / hi cross PCIe, the validity plane is rebuilt on the device from its
// exceptions), read as three coalesced words.
struct AsciiSource {
    struct Raw { uint32_t w[8]; };
    using Cooked = Nibbles;
    __device__ __forceinline__ static void load(const WindowJob &job, uint64_t block_off, int lane, Raw &r) {
        ld256_stream(job.seq + block_off + (uint32_t)lane * 32u, r.w);
    }
    __device__ __forceinline__ static Cooked cook(const Raw &r) { return to_nibbles(r.w); }
    __device__ __forceinline__ static Planes planes(const Cooked &c, uint32_t &nonascii) { return extract_planes(c, nonascii); }
    // pull a block that will be loaded one step later into L2: more bytes in flight per warp without
    // spending registers on them
    __device__ __forceinline__ static void prefetch(const WindowJob &job, uint64_t block_off, int lane) {
#if TIDK_L2_PREFETCH
        asm volatile("prefetch.global.L2 [%0];" ::"l"(job.seq + block_off + (uint32_t)lane * 32u));
#endif
    }
};

struct PackedSource {
    struct Raw { uint32_t lo, hi, va; };
    using Cooked = Raw;
    __device__ __forceinline__ static void load(const WindowJob &job, uint64_t block_off, int lane, Raw &r) {
        const uint64_t word = (block_off >> 5) + (uint32_t)lane;
        const uint32_t *p = job.packed + (word >> kPackShift) * (3ull * kPackWords) + (word & (kPackWords - 1));
        r.lo = __ldg(p); r.hi = __ldg(p + kPackWords); r.va = __ldg(p + 2 * kPackWords);
    }
    __device__ __forceinline__ static Cooked cook(const Raw &r) { return r; }
    __device__ __forceinline__ static void prefetch(const WindowJob &, uint64_t, int) {}
    __device__ __forceinline__ static Planes planes(const Cooked &c, uint32_t &) {
        Planes P;
        P.lo = c.lo; P.hi = c.hi;
        P.b0 = c.va ^ (c.hi & ~c.lo); // the matchers' validity term b0 ^ (hi & ~lo) then equals va
        P.ok = 0xFFFFFFFFu; P.isn = 0u; P.exact = false;
        return P;
    }
};

// ------------------------------------------------------------------ the kernel
template <class M, int MINB, class Src = AsciiSource>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
window_counts_kernel(const WindowJob job, const typename M::Params mp) {
    const int lane = threadIdx.x & 31;
    uint32_t nonascii = 0;

    // Work queue.  The grid is exactly the resident set (persistent CTAs), so the first two items
    // of every warp are assigned statically -- thousands of warps hitting one atomic counter in the
    // first microsecond serialise in L2 -- and the shared counter hands out the rest.
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long gw = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    unsigned long long item = gw, item_next = gw + n_warps;
    const unsigned long long ticket0 = 2 * n_warps + (gw & (kQueues - 1)); // first item of this warp's queue
    uint32_t *const queue = reinterpret_cast<uint32_t *>(job.work_counter + (gw & (kQueues - 1)) * kQueueStride);
    if (blockIdx.x == 0 && threadIdx.x < kQueues) job.next_counter[threadIdx.x * kQueueStride] = 0ull; // stream order: the next launch starts after this one ends
    if (item >= job.n_items) return;

    const uint4 *ip = reinterpret_cast<const uint4 *>(job.items);
    uint4 da = __ldg(ip + 2 * item), db = __ldg(ip + 2 * item + 1);
    typename Src::Raw w;
    Src::load(job, ((uint64_t)da.y << 32) | da.x, lane, w);

    for (;;) {
        // decode the current descriptor
        const uint64_t scan0 = ((uint64_t)da.y << 32) | da.x;
        const uint64_t out0 = ((uint64_t)da.w << 32) | da.z;
        const uint32_t nwin_r = db.x;
        const int vbeg = (int)(db.y & 0xFFFFu), vend = (int)(db.y >> 16);
        const int cbeg = (int)(db.z & 0xFFFFu), cend = (int)(db.z >> 16);
        const uint32_t nblk = db.w & 0xFFFFu;
        const bool multi = (db.w >> 16) & 1u;

        // look ahead: descriptor of the next item, ticket of the one after
        const bool have_next = item_next < job.n_items;
        uint4 na = make_uint4(0, 0, 0, 0), nb = na;
        if (have_next) { na = __ldg(ip + 2 * item_next); nb = __ldg(ip + 2 * item_next + 1); }
        // atom.inc, not atom.add: ptxas wraps a uniform-address add in its warp aggregation (vote, leader, add, SHFL of
        // the result), which waits for the atomic on the spot.  The ticket is not touched before the end of the item.
        uint32_t t2 = 0;
        if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0xffffffff;" : "=r"(t2) : "l"(queue) : "memory");

        uint32_t cf[M::NQ], cr[M::NQ];
#pragma unroll
        for (int q = 0; q < M::NQ; q++) { cf[q] = 0; cr[q] = 0; }
        typename M::Carry carry;
        M::reset(carry);

        uint64_t off = scan0;
        // One 1 KiB block.  EDGE blocks (the first and the last one)
        // confine matching to the window / segment with masks; the blocks in between carry no mask
        // state at all, so their loop is peeled into its own lean copy.
        auto block = [&](uint32_t blk, auto edge_tag) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            const typename Src::Cooked ck = Src::cook(w); // w is dead from here: refill it
            if (blk + 1 < nblk) {
                off += 1024;
                Src::load(job, off, lane, w);
            } else if (have_next) {
                Src::load(job, ((uint64_t)na.y << 32) | na.x, lane, w);
            }
            // and the block TIDK_L2_PREFETCH steps ahead (possibly of the next item) is pulled into L2
            constexpr uint32_t PF = TIDK_L2_PREFETCH;
            if (PF > 0) {
                if (blk + PF < nblk) Src::prefetch(job, off + (PF - 1) * 1024, lane);
                else if (have_next) Src::prefetch(job, (((uint64_t)na.y << 32) | na.x) + ((uint64_t)(blk + PF - nblk) << 10), lane);
            }
            Planes P = Src::planes(ck, nonascii);
            uint32_t cmask = 0xFFFFFFFFu;
            bool special = P.exact;
            if (EDGE) {
                const int prel = (int)(blk << 10) + lane * 32;
                P.ok &= range_mask32(vbeg - prel, vend - prel);
                special = true;
                if (multi) cmask = range_mask32(cbeg - prel, cend - prel);
            }
            M::step(mp, P, special, cmask, lane, carry, cf, cr);
        };
        uint32_t blk = 0;
        if (nblk > 0) { block(0, std::true_type{}); blk = 1; }
        for (; blk + 1 < nblk; blk++) block(blk, std::false_type{});
        if (blk < nblk) block(blk, std::true_type{});

        if (nblk == 0 && have_next) // empty item (window shorter than its segment grid): nobody
            Src::load(job, ((uint64_t)na.y << 32) | na.x, lane, w); // prefetched for the next one
        if (nblk > 0) {
#pragma unroll
            for (int q = 0; q < M::NQ; q++) {
                if (M::NQ == 1 || (uint32_t)q < job.nq) {
                    const uint32_t f = __reduce_add_sync(0xFFFFFFFFu, cf[q]);
                    const uint32_t r = __reduce_add_sync(0xFFFFFFFFu, cr[q]);
                    if (lane == 0) {
                        const uint64_t o = out0 + (uint64_t)(job.q0 + q) * nwin_r;
                        if (!multi) {
                            job.out_fwd[o] = f;
                            job.out_rev[o] = r;
                        } else {
                            atomicAdd((unsigned long long *)&job.out_fwd[o], (unsigned long long)f);
                            atomicAdd((unsigned long long *)&job.out_rev[o], (unsigned long long)r);
                        }
                    }
                }
            }
        }
        if (!have_next) break;
        item = item_next;
        da = na; db = nb;
        item_next = ticket0 + (unsigned long long)__shfl_sync(0xFFFFFFFFu, t2, 0) * kQueues;
    }
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) *(volatile uint32_t *)job.err_flags = 1u; // mapped host memory
}

// Exact byte-wise kernel for queries the plane kernels do not take: letters other
// than A/C/G/T (compared literally against the upper-cased text, e.g. 'N' or the
// 'N's produced by utils::reverse_complement for non-ACGT letters, utils.rs:49-58)
// or length > 32 (incl. the reference's BOM branch, utils.rs:25-28).  One warp per
// window; correctness path, not a roofline path.
struct SlowJob {
    const uint8_t *seq;
    const uint64_t *rec_off;
    const uint64_t *win_base;
    uint64_t *out_fwd, *out_rev;
    unsigned long long *work_counter, *next_counter;
    uint32_t *err_flags;
    uint64_t window, n_windows;
    uint32_t n_records, q, nq_total;
    const uint8_t *fwd, *rev; // device, length len
    uint32_t len;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) window_counts_bytes_kernel(const SlowJob job) {
    const int lane = threadIdx.x & 31;
    uint32_t nonascii = 0;
    if (blockIdx.x == 0 && threadIdx.x < kQueues) job.next_counter[threadIdx.x * kQueueStride] = 0ull;
    for (;;) {
        unsigned long long win = 0;
        if (lane == 0) win = atomicAdd(job.work_counter, 1ull);
        win = __shfl_sync(0xFFFFFFFFu, win, 0);
        if (win >= job.n_windows) break;
        uint32_t r_lo = 0, r_hi = job.n_records;
        while (r_hi - r_lo > 1) {
            uint32_t mid = (r_lo + r_hi) >> 1;
            if (job.win_base[mid] <= win) r_lo = mid; else r_hi = mid;
        }
        const uint32_t rec = r_lo;
        const uint64_t rec_beg = job.rec_off[rec], rec_end = job.rec_off[rec + 1];
        const uint64_t wi = win - job.win_base[rec];
        const uint64_t nwin_r = job.win_base[rec + 1] - job.win_base[rec];
        const uint64_t beg = rec_beg + wi * job.window;
        const uint64_t end = (beg + job.window < rec_end) ? beg + job.window : rec_end;
        unsigned long long f = 0, r = 0;
        for (uint64_t p = beg + lane; p < end; p += 32) {
            nonascii |= job.seq[p] & 0x80u;
            if (p + job.len > end) continue;
            bool mf = true, mr = true;
            for (uint32_t j = 0; j < job.len && (mf || mr); j++) {
                uint8_t c = job.seq[p + j];
                c = (c >= 'a' && c <= 'z') ? c - 32 : c;
                mf = mf && (c == job.fwd[j]);
                mr = mr && (c == job.rev[j]);
            }
            f += mf; r += mr;
        }
        for (int o = 16; o > 0; o >>= 1) {
            f += __shfl_down_sync(0xFFFFFFFFu, f, o);
            r += __shfl_down_sync(0xFFFFFFFFu, r, o);
        }
        if (lane == 0) {
            const uint64_t o = job.win_base[rec] * job.nq_total + (uint64_t)job.q * nwin_r + wi;
            job.out_fwd[o] = f;
            job.out_rev[o] = r;
        }
    }
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) *(volatile uint32_t *)job.err_flags = 1u; // mapped host memory
}

} // namespace tidk
