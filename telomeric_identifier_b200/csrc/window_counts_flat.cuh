// window_counts_flat.cuh -- the window counter on resident bit-planes in FLAT form: the work is cut by position, not
// by window.
//
// window_counts_kernel (window_counts.cuh) gives a warp one window (or a 32 KiB segment of one) at a time.  That is the
// right shape for the ASCII source, where a lane must transpose its 32 bytes anyway, but on the resident planes
// (planes_build.cuh: 0.375 B per base) it leaves three 128-byte requests in flight per warp, pays the per-item costs
// (descriptor, queue ticket, edge masks, two reductions, two stores) once per window -- every 1 000 bases when
// `--window 1000` -- and was measured at 53 % of the HBM peak on the bytes it actually moves (3.1 Gb, W = 10 000),
// 3.1 x slower at W = 1 000.  Here a warp walks a SPAN of 32 Ki consecutive bases of the batch in steps of 4 096: a lane
// owns 128 consecutive bases = one 128-bit load per plane (three 512-byte requests per warp and step, the next step's
// in flight while this one is matched), the look-behind of a lane's words 1..3 is its own previous word (one SHFL per
// equality plane and STEP instead of per word), and windows are an accounting matter:
//
//   * match END positions are computed for the whole step with no regard to windows (search.rs:111-128 counts the
//     occurrences that lie inside one window: a match is lost iff a window START lies inside it behind its first base);
//   * the window starts ("boundaries") that fall into the step are visited in order -- warp-uniform state: record,
//     window index, next boundary -- and each one (a) closes the window in front of it: the match ends below it that are
//     still unclaimed are counted (popc under a mask, one REDUX per strand) and added to that window's two counters with
//     `red.global.add`, (b) wipes the L - 1 match ends that would straddle it;
//   * what is left at the end of the step is added to per-lane counters that belong to the window still open.
//
// Counters are pre-zeroed and every window receives its count by `red.add` (a window may be shared by two spans), so no
// span waits for another.  Cost per boundary: about 70 warp instructions, i.e. W = 1 000 costs what a 12-base run-time
// pattern costs; W = 10 000 pays 3 per KiB.
//
// Matchers: StaticMatcher<L, F> (compile-time motif, window_counts.cuh) and FlatRuntime<L> (any A/C/G/T query up to 32
// bases: mismatch accumulation against broadcast pattern bits, both strands share the shifted planes; the validity
// chain is skipped in steps whose 128 words are all valid, which planes_build.cuh recorded one bit per word).
// Multi-query clade kernels keep the per-window form.
#pragma once
#include "planes_build.cuh"
#include "window_counts.cuh"

namespace tidk {

constexpr uint32_t kFlatStepBases = 4096;  // 128 words: four per lane
#ifndef TIDK_FLAT_SPAN_STEPS
#define TIDK_FLAT_SPAN_STEPS 8
#endif
constexpr uint32_t kFlatSpanSteps = TIDK_FLAT_SPAN_STEPS; // (1, 2, 4 or 8: a span's blocks are one validbits word per lane)
constexpr uint32_t kFlatSpanBases = kFlatStepBases * kFlatSpanSteps; // 32 Ki bases = 32 plane blocks = one validbits line

struct __align__(32) SpanDesc { // the window that holds the first base of a span
    uint64_t next_b;  // where the next window starts (the end of this one)
    uint64_t cur_beg; // where this one starts
    uint64_t out;     // its output index for the launch's query: win_base[rec] * nq_total + q * nwin_rec + wi
    uint32_t rec;     // its record
    uint32_t wi;      // its index in the record
};
static_assert(sizeof(SpanDesc) == 32, "SpanDesc is one 256-bit load");

struct FlatPrepJob {
    const uint64_t *rec_off;  // [n_records + 1]
    const uint64_t *win_base; // [n_records + 1]
    SpanDesc *spans;
    uint64_t window, n_spans, span0; // span0: first span of the list (hybrid uploads start behind the ASCII part)
    uint32_t n_records, nq_total;
};

// one thread per span: binary search for the record that holds the span's first base
__global__ void __launch_bounds__(256) flat_spans_kernel(const FlatPrepJob job) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= job.n_spans) return;
    const uint64_t P = (job.span0 + s) * kFlatSpanBases;
    uint32_t lo = 0, hi = job.n_records; // last record with rec_off[r] <= P (it is not empty: P < total)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (job.rec_off[mid] <= P) lo = mid; else hi = mid;
    }
    const uint64_t rb = job.rec_off[lo], re = job.rec_off[lo + 1];
    const uint64_t wi = (P - rb) / job.window;
    SpanDesc d;
    d.cur_beg = rb + wi * job.window;
    d.next_b = job.window < re - d.cur_beg ? d.cur_beg + job.window : re;
    d.out = job.win_base[lo] * job.nq_total + wi; // (+ q * nwin_rec: added by the kernel)
    d.rec = lo;
    d.wi = (uint32_t)wi;
    job.spans[s] = d;
}

struct FlatJob {
    const uint32_t *planes;    // chunked [lo][hi][va] (planes_build.cuh)
    const uint32_t *validbits; // one bit per word: va is all ones
    const uint64_t *rec_off, *win_base;
    const SpanDesc *spans;
    unsigned long long *out_fwd, *out_rev; // zero at launch
    unsigned long long *work_counter, *next_counter; // as in window_counts.cuh: kQueues ticket counters per set
    uint64_t window, total, n_spans, span0;
    uint64_t n_vblocks;        // validbits entries that exist (blocks behind them count as not valid)
    uint4 *zero_ptr;           // zero_n16 16-byte words to zero on the side (the output buffer of the NEXT call), or nullptr
    uint64_t zero_n16;
    uint32_t n_records, nq_total, q; // this launch's query index
};

__device__ __forceinline__ uint4 ldg128_stream(const uint32_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Run-time pattern, one query: bit s of plo / phi is the expected lo / hi plane bit s bases behind the match end.
template <int L>
struct FlatRuntime {
    static constexpr int LEN = L;
    // word s of flo / fhi (rlo / rhi): all ones iff the forward (reverse-complement) pattern expects lo / hi = 1 s bases
    // behind the match end.  Kernel parameters live in the constant bank, which LOP3 reads as an operand: the 4 L
    // masks cost no register.
    struct Params { uint32_t flo[32], fhi[32], rlo[32], rhi[32]; };
};
template <class M> struct flat_is_runtime : std::false_type {};
template <int L> struct flat_is_runtime<FlatRuntime<L>> : std::true_type {};
template <class M> struct flat_len;
template <int L, uint64_t F> struct flat_len<StaticMatcher<L, F>> { static constexpr int value = L; };
template <int L> struct flat_len<FlatRuntime<L>> { static constexpr int value = L; };

template <class M, int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
window_counts_flat_kernel(const FlatJob job, const typename std::conditional<flat_is_runtime<M>::value, typename M::Params, uint32_t>::type mp) {
    constexpr int L = flat_len<M>::value;
    constexpr bool RT = flat_is_runtime<M>::value;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * kWarpsPerBlock;
    const unsigned long long gw = (unsigned long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x < kQueues) job.next_counter[threadIdx.x * kQueueStride] = 0ull; // for the launch after this one
    if (job.zero_ptr) { // the counters the next call will add into: no memset node between two passes
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < job.zero_n16; i += (uint64_t)gridDim.x * blockDim.x) job.zero_ptr[i] = z;
    }
    unsigned long long span = gw;
    if (span >= job.n_spans) return;
    uint32_t *const queue = reinterpret_cast<uint32_t *>(job.work_counter + (gw & (kQueues - 1)) * kQueueStride);
    const unsigned long long ticket0 = n_warps + (gw & (kQueues - 1));

    auto plane_ptr = [&](uint64_t word) { return plane_word_ptr(job.planes, word) + (uint32_t)lane * 4u; };
    const uint4 *sp = reinterpret_cast<const uint4 *>(job.spans);

    // first step of the first span
    uint64_t word0 = (job.span0 + span) * (kFlatSpanBases / 32);
    const uint32_t *pp = plane_ptr(word0);
    uint4 nlo = ldg128_stream(pp), nhi = ldg128_stream(pp + kPackWords), nva = ldg128_stream(pp + 2 * kPackWords);
    auto valid_word = [&](uint64_t w0) { // this lane's block of the span that starts at word w0
        const uint64_t blk = (w0 >> 5) + (uint32_t)lane;
        return blk < job.n_vblocks ? __ldg(job.validbits + blk) : 0u;
    };
    uint32_t vb = valid_word(word0);
    uint4 da = __ldg(sp + 2 * span), db = __ldg(sp + 2 * span + 1);
    uint32_t n_c_lo = 0, n_c_hi = 0, n_c_va = 0;
    auto load_front = [&](uint64_t w0) { // lane 31: the word in front of the span that starts at word w0
        n_c_lo = n_c_hi = n_c_va = 0u;
        if (lane == 31 && w0 > 0) {
            const uint32_t *q = plane_word_ptr(job.planes, w0 - 1);
            n_c_lo = __ldg(q); n_c_hi = __ldg(q + kPackWords); n_c_va = __ldg(q + 2 * kPackWords);
        }
    };
    load_front(word0);

    for (;;) {
        // ticket for the span after this one: used at the end of the span
        uint32_t t2 = 0;
        if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0xffffffff;" : "=r"(t2) : "l"(queue) : "memory");

        // window state (warp-uniform)
        uint64_t next_b = ((uint64_t)da.y << 32) | da.x, last_b = ((uint64_t)da.w << 32) | da.z;
        uint32_t rec = db.z, wi = db.w;
        uint64_t rec_end = __ldg(job.rec_off + rec + 1);
        uint32_t nwin_r = (uint32_t)(__ldg(job.win_base + rec + 1) - __ldg(job.win_base + rec));
        uint64_t out = (((uint64_t)db.y << 32) | db.x) + (uint64_t)job.q * nwin_r;
        bool open = true; // a window is open (false behind the last record)

        const uint64_t span_pos = (job.span0 + span) * (uint64_t)kFlatSpanBases;
        // the two boundaries as span-relative 32-bit numbers (the per-step tests are then single compares): nb = next
        // boundary, "far" (beyond any span) when it lies outside; lb = last boundary, clamped well below any straddle zone
        constexpr uint32_t kFar = 0x40000000u;
        auto rel_next = [&](uint64_t b) { return b - span_pos < (uint64_t)kFar ? (uint32_t)(b - span_pos) : kFar; };
        auto rel_last = [&](uint64_t b) { return b >= span_pos ? (int)rel_next(b) : (span_pos - b < 64 ? -(int)(span_pos - b) : -64); };
        uint32_t nb = rel_next(next_b);
        int lb = rel_last(last_b);
        const uint32_t valid_blocks = __ballot_sync(0xFFFFFFFFu, vb == 0xFFFFFFFFu); // bit b: block b of the span is all valid
        // look-behind across the span start: lane 31 holds the word in front of the span (loaded with the span's first step)
        uint32_t c_lo = n_c_lo, c_hi = n_c_hi, c_va = n_c_va;
        uint32_t cf = 0, cr = 0;
        const uint64_t left = job.total - span_pos; // > 0
        const uint32_t n_steps = left >= kFlatSpanBases ? kFlatSpanSteps : (uint32_t)((left + kFlatStepBases - 1) / kFlatStepBases);
        unsigned long long span_next = 0;
        bool have_next = false;

#pragma unroll 1 // (unrolled by two -- no register rotation -- and with an L2 prefetch two steps ahead: no gain, the alu pipe is the limiter)
        for (uint32_t st = 0; st < n_steps; st++) {
            const uint4 lo4 = nlo, hi4 = nhi;
            uint4 va4 = nva;
            const bool step_valid = ((valid_blocks >> (4u * st)) & 0xFu) == 0xFu;
            // next step's words (or the first step of the next span)
            if (st + 1 < n_steps) {
                pp += kFlatStepBases / 32;
                nlo = ldg128_stream(pp); nhi = ldg128_stream(pp + kPackWords);
                if (((valid_blocks >> (4u * (st + 1))) & 0xFu) != 0xFu) nva = ldg128_stream(pp + 2 * kPackWords); // (all valid: va is not read, 0.25 B per base)
            } else {
                span_next = ticket0 + (unsigned long long)__shfl_sync(0xFFFFFFFFu, t2, 0) * kQueues;
                have_next = span_next < job.n_spans;
                if (have_next) {
                    word0 = (job.span0 + span_next) * (kFlatSpanBases / 32);
                    pp = plane_ptr(word0);
                    nlo = ldg128_stream(pp); nhi = ldg128_stream(pp + kPackWords); nva = ldg128_stream(pp + 2 * kPackWords);
                    vb = valid_word(word0);
                    da = __ldg(sp + 2 * span_next); db = __ldg(sp + 2 * span_next + 1);
                    load_front(word0);
                }
            }
            if (step_valid) va4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            const uint32_t lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hi[4] = {hi4.x, hi4.y, hi4.z, hi4.w}, va[4] = {va4.x, va4.y, va4.z, va4.w};
            // the word in front of this lane's four: lane - 1's last word (lane 0: lane 31's of the previous step)
            const uint32_t p_lo = lane_lookbehind(lo[3], c_lo, lane), p_hi = lane_lookbehind(hi[3], c_hi, lane), p_va = lane_lookbehind(va[3], c_va, lane);
            c_lo = lo[3]; c_hi = hi[3]; c_va = va[3];
            // (run-time matcher) nothing invalid in the step nor in the word in front of any lane: no validity chain
            const bool all_valid = RT && step_valid && __all_sync(0xFFFFFFFFu, p_va == 0xFFFFFFFFu);

            uint32_t mf[4], mr[4];
            if constexpr (!RT) {
                uint32_t e[4][4], p0[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    e[j][0] = ~lo[j] & ~hi[j] & va[j]; // A
                    e[j][1] = lo[j] & ~hi[j] & va[j];  // C
                    e[j][2] = lo[j] & hi[j] & va[j];   // G
                    e[j][3] = ~lo[j] & hi[j] & va[j];  // T
                }
                p0[0] = ~p_lo & ~p_hi & p_va; p0[1] = p_lo & ~p_hi & p_va; p0[2] = p_lo & p_hi & p_va; p0[3] = ~p_lo & p_hi & p_va;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    mf[j] = 0xFFFFFFFFu; mr[j] = 0xFFFFFFFFu;
                    if (j == 0) M::template steps<0>(e[0], p0, mf[0], mr[0]);
                    else M::template steps<0>(e[j], e[j - 1], mf[j], mr[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t q_lo = j == 0 ? p_lo : lo[j - 1], q_hi = j == 0 ? p_hi : hi[j - 1];
                    uint32_t xf = 0, xr = 0;
#pragma unroll
                    for (int s = 0; s < L; s++) {
                        const uint32_t lo_s = s == 0 ? lo[j] : __funnelshift_l(q_lo, lo[j], s);
                        const uint32_t hi_s = s == 0 ? hi[j] : __funnelshift_l(q_hi, hi[j], s);
                        xf |= (lo_s ^ mp.flo[s]) | (hi_s ^ mp.fhi[s]);
                        xr |= (lo_s ^ mp.rlo[s]) | (hi_s ^ mp.rhi[s]);
                    }
                    mf[j] = ~xf; mr[j] = ~xr;
                }
                if (!all_valid) { // (warp-uniform) some base of the step is not A/C/G/T: wipe the match ends that touch one
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t nv = ~va[j], q_nv = j == 0 ? ~p_va : ~va[j - 1];
                        uint32_t inv = nv;
#pragma unroll
                        for (int s = 1; s < L; s++) inv |= __funnelshift_l(q_nv, nv, s);
                        mf[j] &= ~inv; mr[j] &= ~inv;
                    }
                }
            }

            // windows: the boundaries inside this step, in order
            const int sb = (int)(st * kFlatStepBases); // span-relative position of the step
            const int lane_rel = sb + lane * 128;
            if (lb + (L - 1) > sb) { // the straddle zone of the last boundary reaches into this step
                const int r0 = lb - lane_rel; // (the boundary may lie below the step: negative)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t keep = ~range_mask32(r0 - 32 * j, r0 - 32 * j + (L - 1));
                    mf[j] &= keep; mr[j] &= keep;
                }
            }
            while (nb < (uint32_t)sb + kFlatStepBases) {
                const int r0 = (int)nb - lane_rel;
                uint32_t bf = 0, br = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int rel = r0 - 32 * j;
                    const uint32_t below = range_mask32(-1000000, rel) ; // bits < rel
                    bf += __popc(mf[j] & below); br += __popc(mr[j] & below);
                    const uint32_t keep = ~(below | range_mask32(rel, rel + (L - 1)));
                    mf[j] &= keep; mr[j] &= keep;
                }
                const uint32_t f = __reduce_add_sync(0xFFFFFFFFu, cf + bf), r = __reduce_add_sync(0xFFFFFFFFu, cr + br);
                cf = 0; cr = 0;
                if (lane == 0 && open) {
                    if (f) atomicAdd(job.out_fwd + out, (unsigned long long)f);
                    if (r) atomicAdd(job.out_rev + out, (unsigned long long)r);
                }
                // the window that starts at next_b
                last_b = next_b;
                lb = (int)nb;
                wi++;
                if (wi < nwin_r) {
                    out++;
                } else { // first window of the next record that has one
                    do { rec++; } while (rec < job.n_records && __ldg(job.rec_off + rec + 1) == __ldg(job.rec_off + rec));
                    if (rec >= job.n_records) { open = false; next_b = ~0ull; nb = kFar; break; }
                    rec_end = __ldg(job.rec_off + rec + 1);
                    const uint64_t wb0 = __ldg(job.win_base + rec);
                    nwin_r = (uint32_t)(__ldg(job.win_base + rec + 1) - wb0);
                    out = wb0 * job.nq_total + (uint64_t)job.q * nwin_r;
                    wi = 0;
                }
                next_b = job.window < rec_end - last_b ? last_b + job.window : rec_end;
                nb = rel_next(next_b);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) { cf += __popc(mf[j]); cr += __popc(mr[j]); }
        }
        // what the span leaves for the window still open
        {
            const uint32_t f = __reduce_add_sync(0xFFFFFFFFu, cf), r = __reduce_add_sync(0xFFFFFFFFu, cr);
            if (lane == 0 && open) {
                if (f) atomicAdd(job.out_fwd + out, (unsigned long long)f);
                if (r) atomicAdd(job.out_rev + out, (unsigned long long)r);
            }
        }
        if (!have_next) break;
        span = span_next;
    }
}

} // namespace tidk
