// window_counts.cuh -- per-window forward / reverse-complement repeat counting.
//
// Replaces search::write_window_counts (search.rs:83-151) and
// finder::write_window_counts (finder.rs:111-178): for every window of every
// record, the number of (overlapping) start positions at which upper(text)
// equals the query, resp. its reverse complement, with matching confined to the
// window (an occurrence that straddles a window boundary is lost).
//
// Mapping: one warp per work item = (window, 32 KiB segment of it); windows up
// to 32 KiB are one item, so with the default W = 10000 a warp owns a window,
// needs no halo from anywhere else, and writes its two counts with plain stores.
// The warp walks its span in 1 KiB blocks (32 lanes x one 256-bit load), keeps
// the previous block's planes for the (L-1)-base look-behind, and counts matches
// by their END position:  match ending at e  <=>  text[e-L+1 .. e] == query.
// Window confinement is applied to the data, not the result: bases outside
// [beg, end) are marked invalid, so no match can reach across the boundary.
#pragma once
#include "planes.cuh"

namespace tidk {

constexpr int kMaxFastLen = 32;         // plane kernels: one look-behind word
constexpr int kMaxFastQueries = 16;     // queries per fused plane-kernel launch
constexpr uint32_t kSegBytes = 32768;   // work-item span for windows larger than this
constexpr int kWarpsPerBlock = 8;

// One strand pattern for the plane kernels: for look-behind distance s (s = 0 is
// the LAST base of the pattern) bit s of plo / phi is the expected lo / hi plane bit.
struct StrandPattern {
    uint32_t plo, phi;
};

struct WindowJob {
    const uint8_t *seq;        // device, 256-B aligned, >= 2 KiB of zero padding behind the data
    const uint64_t *rec_off;   // [n_records + 1] byte offsets
    const uint64_t *win_base;  // [n_records + 1] windows before record r (per query)
    uint64_t *out_fwd;         // [record][query][window]
    uint64_t *out_rev;
    unsigned long long *work_counter; // zeroed before launch
    uint32_t *err_flags;       // bit 0: non-ASCII byte seen
    uint64_t window;
    uint64_t n_windows;        // sum over records
    uint32_t n_records;
    uint32_t segs_per_window;  // ceil(window / kSegBytes)
    uint32_t nq;
    uint32_t q0;               // first query index of this launch (output addressing)
    uint32_t nq_total;
};

struct FastQueries {
    StrandPattern fwd[kMaxFastQueries];
    StrandPattern rev[kMaxFastQueries];
    uint8_t len[kMaxFastQueries];
};

__device__ __forceinline__ uint32_t range_mask(int64_t lo_rel, int64_t hi_rel) {
    // bits b with lo_rel <= b < hi_rel, b in [0, 32)
    if (hi_rel <= 0 || lo_rel >= 32 || hi_rel <= lo_rel) return 0u;
    uint32_t m = 0xFFFFFFFFu;
    if (lo_rel > 0) m &= 0xFFFFFFFFu << (int)lo_rel;
    if (hi_rel < 32) m &= 0xFFFFFFFFu >> (32 - (int)hi_rel);
    return m;
}

// mismatch accumulation for one strand over L look-behind steps.
// mism bit e stays 0 iff text[e-s] == pattern[L-1-s] for all s < L.
template <int L, int S>
struct MatchStep {
    __device__ __forceinline__ static void run(const Planes &c, const Planes &p, uint32_t flo,
                                               uint32_t fhi, uint32_t rlo, uint32_t rhi,
                                               uint32_t &mf, uint32_t &mr, uint32_t &inv) {
        uint32_t lo_s = S == 0 ? c.lo : __funnelshift_l(p.lo, c.lo, S);
        uint32_t hi_s = S == 0 ? c.hi : __funnelshift_l(p.hi, c.hi, S);
        uint32_t va_s = S == 0 ? c.va : __funnelshift_l(p.va, c.va, S);
        // expected bit for step S broadcast to all 32 positions
        uint32_t eflo = 0u - ((flo >> S) & 1u), efhi = 0u - ((fhi >> S) & 1u);
        uint32_t erlo = 0u - ((rlo >> S) & 1u), erhi = 0u - ((rhi >> S) & 1u);
        inv |= ~va_s;
        mf |= (lo_s ^ eflo) | (hi_s ^ efhi);
        mr |= (lo_s ^ erlo) | (hi_s ^ erhi);
        MatchStep<L, S + 1>::run(c, p, flo, fhi, rlo, rhi, mf, mr, inv);
    }
};
template <int L>
struct MatchStep<L, L> {
    __device__ __forceinline__ static void run(const Planes &, const Planes &, uint32_t, uint32_t,
                                               uint32_t, uint32_t, uint32_t &, uint32_t &,
                                               uint32_t &) {}
};

// Runtime-length version (multi-query launches with mixed lengths).
__device__ __forceinline__ void match_runtime(const Planes &c, const Planes &p, int L, uint32_t flo,
                                              uint32_t fhi, uint32_t rlo, uint32_t rhi,
                                              uint32_t &mf, uint32_t &mr, uint32_t &inv) {
    for (int s = 0; s < L; s++) {
        uint32_t lo_s = __funnelshift_l(p.lo, c.lo, s);
        uint32_t hi_s = __funnelshift_l(p.hi, c.hi, s);
        uint32_t va_s = __funnelshift_l(p.va, c.va, s);
        uint32_t eflo = 0u - ((flo >> s) & 1u), efhi = 0u - ((fhi >> s) & 1u);
        uint32_t erlo = 0u - ((rlo >> s) & 1u), erhi = 0u - ((rhi >> s) & 1u);
        inv |= ~va_s;
        mf |= (lo_s ^ eflo) | (hi_s ^ efhi);
        mr |= (lo_s ^ erlo) | (hi_s ^ erhi);
    }
}

// LT > 0: single query of compile-time length LT.  LT == 0: nq queries, runtime lengths.
template <int LT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
window_counts_planes_kernel(const WindowJob job, const FastQueries fq) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_items = job.n_windows * job.segs_per_window;
    uint32_t nonascii = 0;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(job.work_counter, 1ull);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_items) break;
        const uint64_t win = item / job.segs_per_window;
        const uint32_t seg = (uint32_t)(item - win * job.segs_per_window);

        // record of this window: last r with win_base[r] <= win
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
        // segment span: END positions counted by this item
        const uint64_t sbeg = beg + (uint64_t)seg * kSegBytes;
        if (sbeg >= end) continue; // last window of a record can be short
        const uint64_t send = (sbeg + kSegBytes < end) ? sbeg + kSegBytes : end;
        // scan from one block (1 KiB) before the segment when it is not the window start,
        // so that the look-behind planes are real
        uint64_t scan0 = sbeg & ~31ull;
        if (seg > 0) scan0 -= 1024;
        const uint32_t nblk = (uint32_t)((send - scan0 + 1023) >> 10);

        uint32_t cf[LT > 0 ? 1 : kMaxFastQueries], cr[LT > 0 ? 1 : kMaxFastQueries];
#pragma unroll
        for (int q = 0; q < (LT > 0 ? 1 : kMaxFastQueries); q++) { cf[q] = 0; cr[q] = 0; }

        Planes prev_own = {0u, 0u, 0u}; // this lane's planes of the previous block
        uint32_t w[8], wn[8];
        const uint8_t *ptr = job.seq + scan0 + (uint64_t)lane * 32;
        ld256_stream(ptr, w);
        for (uint32_t blk = 0; blk < nblk; blk++) {
            if (blk + 1 < nblk) ld256_stream(ptr + 1024, wn); // prefetch next block
            Planes cur = extract_planes(w, nonascii);
            const uint64_t P = scan0 + ((uint64_t)blk << 10) + (uint64_t)lane * 32;
            uint32_t cmask = 0xFFFFFFFFu;
            if (blk <= 1 || blk + 1 == nblk) { // warp-uniform: only edge blocks pay for masks
                cur.va &= range_mask((int64_t)beg - (int64_t)P, (int64_t)end - (int64_t)P);
                cmask = range_mask((int64_t)sbeg - (int64_t)P, (int64_t)send - (int64_t)P);
            }
            Planes prv;
            prv.lo = lane_lookbehind(cur.lo, prev_own.lo, lane);
            prv.hi = lane_lookbehind(cur.hi, prev_own.hi, lane);
            prv.va = lane_lookbehind(cur.va, prev_own.va, lane);
            if (LT > 0) {
                uint32_t mf = 0, mr = 0, inv = 0;
                MatchStep<LT, 0>::run(cur, prv, fq.fwd[0].plo, fq.fwd[0].phi, fq.rev[0].plo,
                                      fq.rev[0].phi, mf, mr, inv);
                cf[0] += __popc(~(mf | inv) & cmask);
                cr[0] += __popc(~(mr | inv) & cmask);
            } else {
#pragma unroll 1
                for (uint32_t q = 0; q < job.nq; q++) {
                    uint32_t mf = 0, mr = 0, inv = 0;
                    match_runtime(cur, prv, fq.len[q], fq.fwd[q].plo, fq.fwd[q].phi,
                                  fq.rev[q].plo, fq.rev[q].phi, mf, mr, inv);
                    uint32_t a = __popc(~(mf | inv) & cmask), b = __popc(~(mr | inv) & cmask);
#pragma unroll
                    for (int t = 0; t < kMaxFastQueries; t++) // keep cf/cr in registers
                        if (t == (int)q) { cf[t] += a; cr[t] += b; }
                }
            }
            prev_own = cur;
            ptr += 1024;
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = wn[i];
        }

        // reduce over the warp and write
        const uint32_t nq_here = LT > 0 ? 1u : job.nq;
#pragma unroll
        for (int q = 0; q < (LT > 0 ? 1 : kMaxFastQueries); q++) {
            if ((uint32_t)q < nq_here) {
                uint32_t f = __reduce_add_sync(0xFFFFFFFFu, cf[q]);
                uint32_t r = __reduce_add_sync(0xFFFFFFFFu, cr[q]);
                if (lane == 0) {
                    const uint64_t o = job.win_base[rec] * job.nq_total +
                                       (uint64_t)(job.q0 + q) * nwin_r + wi;
                    if (job.segs_per_window == 1) {
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
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) atomicOr(job.err_flags, 1u);
}

// Exact byte-wise kernel for queries the plane kernels do not take: letters other
// than A/C/G/T (compared literally against the upper-cased text, e.g. 'N' or the
// 'N's produced by utils::reverse_complement for non-ACGT letters, utils.rs:49-58)
// or length > 32 (incl. the reference's BOM branch, utils.rs:25-28).
struct SlowQuery {
    const uint8_t *fwd; // device, length len
    const uint8_t *rev;
    uint32_t len;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
window_counts_bytes_kernel(const WindowJob job, const SlowQuery sq) {
    const int lane = threadIdx.x & 31;
    uint32_t nonascii = 0;
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
            if (p + sq.len > end) continue;
            bool mf = true, mr = true;
            for (uint32_t j = 0; j < sq.len && (mf || mr); j++) {
                uint8_t c = job.seq[p + j];
                c = (c >= 'a' && c <= 'z') ? c - 32 : c;
                mf = mf && (c == sq.fwd[j]);
                mr = mr && (c == sq.rev[j]);
            }
            f += mf; r += mr;
        }
        for (int o = 16; o > 0; o >>= 1) {
            f += __shfl_down_sync(0xFFFFFFFFu, f, o);
            r += __shfl_down_sync(0xFFFFFFFFu, r, o);
        }
        if (lane == 0) {
            const uint64_t o = job.win_base[rec] * job.nq_total + (uint64_t)job.q0 * nwin_r + wi;
            job.out_fwd[o] = f;
            job.out_rev[o] = r;
        }
    }
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) atomicOr(job.err_flags, 1u);
}

} // namespace tidk
