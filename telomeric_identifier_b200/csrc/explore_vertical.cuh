// explore_vertical.cuh -- the explore scan on resident bit-planes in VERTICAL (transposed) form.
//
// Same contract as explore_planes_kernel (explore_planes.cuh): one event per raw-run boundary of every (slice, k)
// (explore.rs:178-233; the look-behind restatement and the arithmetic are in explore_vertical_core.hpp).  A lane owns
// a tile of 1024 consecutive bases of one slice, a warp 32 such tiles; tiles of a slice start every 960 bases at the
// slice start (not at a word boundary: the lane funnel-shifts its 33 plane words into place), so the first tile needs
// no validity mask -- what lies before the slice is the zero column the transposed form shifts in -- and a later
// tile's first 64 positions are look-behind only.  Nothing of the k loop needs a shuffle, a carry or a phase-dependent
// shift: all 32 bit columns of a vertical word are independent positions 32 bases apart.
//
// Exactness.  The lo / hi planes are the bytes only where a 32-base word is "plain" (32 upper-case A C G T; one bit
// per word, built with the planes).  A tile whose in-slice words are all plain is scanned here; any other tile -- lower
// case, N, IUPAC codes, whatever else -- is appended as an ordinary work item to a list that explore_planes_kernel
// (ASCII source, every byte value, letter case, joins) works off right behind this kernel.  Both own the same chunk-end
// positions per tile and write into the same event buffer, so which of them scanned a tile does not show in the result.
// What lies behind the end of a slice is never masked: events found there are dropped when they are emitted, and the
// run that touches the slice end is closed by the parity rule of vt_emit.
#pragma once
#include "explore_planes.cuh"
#include "explore_vertical_core.hpp"

// A/B switch (kernel-tuning builds): the warps of a CTA enter every item behind a barrier
#ifndef TIDK_V_SYNC
#define TIDK_V_SYNC 0
#endif

namespace tidk {

constexpr int kVWarps = 4;          // warps per CTA
constexpr uint32_t kVEvBuf = 64;    // events a warp collects in shared memory per 32 tiles before it asks for global slots

struct VJob {
    const uint32_t *planes;    // chunked [lo][hi][va] (planes_build.cuh)
    const uint32_t *plainbits; // one bit per 32-base word
    const SliceDesc *slices;
    const VTile *vtiles;       // [n_vtiles] (slice, tile of the slice), slice by slice
    uint64_t n_vtiles;
    uint32_t kmin, kmax;
    uint64_t *events;              // packed keys (pack_event)
    unsigned long long *n_events;  // slots handed out
    uint64_t cap_events;
    uint32_t *work_counter;        // zero at launch
    KeyFmt fmt;
    ExItem *slow_items;            // tiles left to explore_planes_kernel
    unsigned long long *n_slow;
};

// The tile list (base = prefix sums of vt_count over the slices): a thread writes the tiles of its slice when they are
// few (reads), the warp writes them together when they are many (chromosome ends).  (A binary search per tile cost
// 33 us on two million tiles; the same search inside the scan, 18 dependent loads per item, cost more than that in
// exposed latency; one warp per slice 21 us.)
__global__ void __launch_bounds__(256) vt_items_kernel(const uint64_t *base /* [n_slices + 1] */, uint32_t n_slices, VTile *out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint64_t b = 0, n = 0;
    if (s < n_slices) {
        b = base[s];
        n = base[s + 1] - b;
    }
    const bool big = n > 32;
    if (!big)
        for (uint32_t t = 0; t < (uint32_t)n; t++) out[b + t] = VTile{s, t};
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, big);
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const uint64_t bb = __shfl_sync(0xFFFFFFFFu, b, src), nn = __shfl_sync(0xFFFFFFFFu, n, src);
        const uint32_t ss = __shfl_sync(0xFFFFFFFFu, s, src);
        for (uint64_t t = (uint32_t)lane; t < nn; t += 32) out[bb + t] = VTile{ss, (uint32_t)t};
    }
}

// Shared memory of a CTA (dynamic), per warp: a staging buffer for the horizontal plane words of the item's 32 tiles, a stash
// for the E words of the k in work, the event buffer (and, in the cp.async build, a second staging buffer: there the buffer
// of the item in work doubles as the stash); one chunk-end mask table per CTA.
#ifndef TIDK_V_TMA
#define TIDK_V_TMA 1 // staging by bulk copies (TMA), one per run of tiles of a slice; 0: four-byte cp.async, row by row
#endif
#if TIDK_V_TMA
// TMA variant: ONE staging buffer per warp, filled by bulk copies (cp.async.bulk), one per RUN of tiles of the same slice and
// plane: consecutive tiles of a slice start 30 words apart, so the tiles of one slice in a warp's item are one contiguous
// stretch of each plane.  The run that starts in lane h lands 40 h words into the plane's half of the buffer (a run of n
// tiles needs at most 30 n + 8 words); completion is counted on the warp's mbarrier.  A separate 4 KB stash for the E words.
constexpr uint32_t kVRunSlot = 40;                 // words per lane slot: 160 bytes, a multiple of 16
constexpr uint32_t kVPlaneWords = 32 * kVRunSlot;  // one plane's half of the staging buffer
constexpr uint32_t kVStageWords = 2 * kVPlaneWords;
struct __align__(16) VWarpShared {
    uint32_t stage[1][kVStageWords];
    uint4 stash[8 * 32];
    unsigned long long ev[kVEvBuf];
    unsigned long long mbar;
    unsigned long long pad;
};
#else
constexpr uint32_t kVStageWords = 2 * 32 * 33;
struct __align__(16) VWarpShared {
    uint32_t stage[2][kVStageWords];
    unsigned long long ev[kVEvBuf];
};
#endif
struct __align__(16) VShared {
    VWarpShared w[kVWarps];
    uint32_t mask[kVMaskWords];
};

// the chunk-end mask table (explore_vertical_core.hpp), built at compile time; every CTA copies it into shared memory
struct VtMaskTable {
    uint32_t w[kVMaskWords];
    constexpr VtMaskTable() : w{} {
        for (int k = 1; k <= kVMaxK; k++)
            for (int x = 0; x < k + 32; x++) {
                uint32_t m = 0;
                for (int b = 0; b < 32; b++)
                    if ((x + 32 * b + 1) % k == 0) m |= 1u << b;
                w[vt_mask_row(k) + x] = m;
            }
    }
};
__constant__ VtMaskTable c_vt_mask = VtMaskTable();

// chunk-end masks of a k that divides the tile step (960 = 2^6 * 3 * 5: k = 5, 6, 8, 10, 12, ...): every tile of a slice
// starts at chunk phase 0, the 32 mask words are compile-time constants and end up as LOP3 immediates
template <int K>
struct VtMaskConst {
    uint32_t w[32];
    constexpr VtMaskConst() : w{} {
        for (int i = 0; i < 32; i++) {
            uint32_t x = 0;
            for (int b = 0; b < 32; b++)
                if ((i + 32 * b + 1) % K == 0) x |= 1u << b;
            w[i] = x;
        }
    }
};

// one k of one tile: E words into the lane's stash column (uint4 [8][32] over the staging area), OR of them back
template <int K>
__device__ __forceinline__ uint32_t vt_run_k(const uint32_t (&lo)[32], const uint32_t (&hi)[32], const uint32_t *mask_tab, uint32_t t, uint4 *stash_lane) {
    uint32_t eb[32];
    auto out = [&](int i, uint32_t e) {
        eb[i] = e;
        if (vt_group_done<K>(i)) stash_lane[(i >> 2) * 32] = make_uint4(eb[i & ~3], eb[(i & ~3) + 1], eb[(i & ~3) + 2], eb[(i & ~3) + 3]);
    };
    if constexpr (kVStep % K == 0) {
        constexpr VtMaskConst<K> mc{};
        return vt_compute<K>(lo, hi, [&](int i) { return mc.w[i]; }, out);
    } else {
        const uint32_t *mrow = mask_tab + vt_mask_row(K) + ((t % K) * (kVStep % K)) % K; // phase = (960 t) mod K
        return vt_compute<K>(lo, hi, [&](int i) { return mrow[i]; }, out);
    }
}

// event number `pos` of the warp's current item: into the warp's buffer, or, in a burst (tiles full of short repeats),
// straight to the global buffer
__device__ __forceinline__ void vt_put(const VJob &job, unsigned long long *evbuf, uint32_t pos, uint64_t key) {
    if (pos < kVEvBuf) {
        evbuf[pos] = key;
    } else {
        const unsigned long long idx = atomicAdd(job.n_events, 1ull);
        if (idx < job.cap_events) job.events[idx] = key;
    }
}

struct VDesc { // one lane's tile
    VTile d;
    SliceDesc sd;
    bool valid;
};
__device__ __forceinline__ VDesc vt_desc(const VJob &job, uint64_t item, int lane) {
    VDesc r;
    const uint64_t v = item * 32 + (uint32_t)lane;
    r.valid = v < job.n_vtiles;
    r.d = VTile{0u, 0u};
    r.sd = SliceDesc{0ull, 0ull};
    if (r.valid) {
        r.d = job.vtiles[v];
        r.sd = job.slices[r.d.slice];
    }
    return r;
}

// the 32 x 33 words of each plane that the warp's tiles start in, row by row (a tile's words are consecutive in the
// plane: coalesced), straight into shared memory; wa = this lane's first word.  Every lane forms the address of ITS
// tile's first word once; the loop over the 32 tiles then broadcasts it (two SHFL) together with the number of words
// left in the tile's 4 MiB chunk of the plane layout (a row that runs over the end of its chunk continues 2 * kPackWords
// further on, behind the chunk's hi and va parts): 8 instructions per row where forming the address per row took 17.
__device__ __forceinline__ void vt_stage_async(const uint32_t *planes, uint32_t wa, uint32_t *stage, int lane) {
    const uint32_t s_lo = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)lane * 4u, s_hi = s_lo + 32 * 33 * 4;
    const unsigned long long row = (unsigned long long)(uintptr_t)plane_word_ptr(planes, (uint64_t)wa);
    const uint32_t left = kPackWords - (wa & (kPackWords - 1u)); // words of this chunk from wa on (>= 1)
    constexpr unsigned long long kSkip = 2ull * kPackWords * 4ull;
    if (!__any_sync(0xFFFFFFFFu, left < 33u)) { // (nearly always) no row of this item reaches the end of its chunk
#pragma unroll 8
        for (int L = 0; L < 32; L++) {
            const unsigned long long pp = __shfl_sync(0xFFFFFFFFu, row, L) + (uint32_t)lane * 4u;
            const uint32_t off = (uint32_t)(L * 33) * 4u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_lo + off), "l"(pp) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_hi + off), "l"(pp + (unsigned long long)kPackWords * 4ull) : "memory");
        }
        const uint32_t off = (uint32_t)(lane * 32 + 32) * 4u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_lo + off), "l"(row + 128ull) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_hi + off), "l"(row + 128ull + (unsigned long long)kPackWords * 4ull) : "memory");
        return;
    }
#pragma unroll 1
    for (int L = 0; L < 32; L++) {
        const unsigned long long rl = __shfl_sync(0xFFFFFFFFu, row, L);
        const uint32_t ll = __shfl_sync(0xFFFFFFFFu, left, L);
        const unsigned long long p = rl + (uint32_t)lane * 4u + ((uint32_t)lane >= ll ? kSkip : 0ull);
        const uint32_t off = (uint32_t)(L * 33) * 4u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_lo + off), "l"(p) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1+%2], 4;" ::"r"(s_hi + off), "l"(p), "n"(kPackWords * 4u) : "memory"); // the hi part of the chunk
    }
    const unsigned long long p = row + 32ull * 4ull + (32u >= left ? kSkip : 0ull);
    const uint32_t off = (uint32_t)(lane * 32 + 32) * 4u; // (+ lane * 4 from s_lo: word lane * 33 + 32)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_lo + off), "l"(p) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1+%2], 4;" ::"r"(s_hi + off), "l"(p), "n"(kPackWords * 4u) : "memory");
}

#if TIDK_V_TMA
// One bulk copy (TMA) of `bytes` (a multiple of 16) from global to shared memory, completion counted on `mbar`.
__device__ __forceinline__ void vt_bulk(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// Stages the plane words of the warp's 32 tiles run by run.  slice / wa / valid: this lane's tile (valid lanes are a prefix of
// the warp; the tiles of a slice sit in consecutive lanes, 30 words apart).  Returns the word offset of this lane's row in a
// plane's half of the buffer (its 33 words start there).
__device__ __forceinline__ uint32_t vt_stage_tma(const uint32_t *planes, uint32_t slice, uint32_t wa, bool valid, uint32_t *stage,
                                                 uint32_t mbar, int lane) {
    const uint32_t prev_slice = __shfl_up_sync(0xFFFFFFFFu, slice, 1);
    const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, valid && (lane == 0 || slice != prev_slice));
    const uint32_t n_valid = (uint32_t)__popc(vm);
    // my run: head lane h, and (for heads) its length
    const uint32_t below = heads & (0xFFFFFFFFu >> (31 - lane));          // heads at or below this lane
    const uint32_t h = below ? 31u - (uint32_t)__clz((int)below) : 0u;
    const uint32_t above = lane < 31 ? heads & (0xFFFFFFFEu << lane) : 0u; // heads above this lane
    const uint32_t run_end = above ? (uint32_t)__ffs((int)above) - 1u : n_valid;
    const uint32_t wa_h = __shfl_sync(0xFFFFFFFFu, wa, (int)h);
    const bool head = valid && h == (uint32_t)lane;
    uint32_t words = 0;
    if (head) words = ((wa & 3u) + 30u * (run_end - (uint32_t)lane - 1u) + 33u + 3u) & ~3u;
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, words);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total * 8u) : "memory"); // two planes
    __syncwarp();
    if (head) {
        const uint32_t a4 = wa & ~3u;
        const uint32_t *src = plane_word_ptr(planes, (uint64_t)a4);
        const uint32_t left = kPackWords - (a4 & (kPackWords - 1u)); // words of this chunk from a4 on: a multiple of 4, >= 4
        const uint32_t n1 = left < words ? left : words, n2 = words - n1;
        const uint32_t d_lo = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)lane * kVRunSlot * 4u, d_hi = d_lo + kVPlaneWords * 4u;
        vt_bulk(d_lo, src, n1 * 4u, mbar);
        vt_bulk(d_hi, src + kPackWords, n1 * 4u, mbar);
        if (n2) { // the run crosses the end of its 4 MiB chunk of the plane layout: the next chunk's lo part starts 2 * kPackWords on
            const uint32_t *src2 = src + left + 2u * kPackWords;
            vt_bulk(d_lo + n1 * 4u, src2, n2 * 4u, mbar);
            vt_bulk(d_hi + n1 * 4u, src2 + kPackWords, n2 * 4u, mbar);
        }
    }
    return kVRunSlot * h + (wa_h & 3u) + 30u * ((uint32_t)lane - h);
}
// wait for the phase with the given parity (bounded: a wrong byte count must not hang the device -- the results would be wrong
// and the parity tests would say so)
__device__ __forceinline__ void vt_mbar_wait(uint32_t mbar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return;
    }
}
#endif

// KMIN > 0: the k range is a compile-time constant; KMIN == 0: run-time range (job.kmin .. job.kmax, <= kVMaxK).
#ifndef TIDK_V_MINB
#define TIDK_V_MINB 3
#endif
template <int KMIN, int KMAX>
__global__ void __launch_bounds__(kVWarps * 32, TIDK_V_MINB) explore_vertical_kernel(const VJob job) {
    extern __shared__ __align__(16) unsigned char vt_smem[];
    VShared &sh = *reinterpret_cast<VShared *>(vt_smem);
    for (uint32_t e = threadIdx.x; e < kVMaskWords; e += blockDim.x) sh.mask[e] = c_vt_mask.w[e];
#if TIDK_V_TMA
    if ((threadIdx.x & 31) == 0) {
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&sh.w[threadIdx.x >> 5].mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the barrier is used by the async proxy (bulk copies)
    }
#endif
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    VWarpShared &ws = sh.w[wid];
#if TIDK_V_TMA
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&ws.mbar);
    uint32_t mb_parity = 0, row_off = 0, row_off_next = 0;
#endif
    unsigned long long *const evbuf = ws.ev;
    const uint32_t k0 = KMIN > 0 ? KMIN : job.kmin, k1 = KMIN > 0 ? KMAX : job.kmax;

    // Items (32 tiles each): the first two of a warp are assigned statically, the rest come from one ticket counter,
    // drawn two items ahead.  The words of item n + 1 are on their way into the other staging buffer while item n is
    // in work; its descriptors are fetched as soon as its ticket has arrived.  (One CTA of twelve warps per SM whose
    // warps enter the code of every k together, behind a barrier, was measured: 0.77 against 0.68 ms.)
    const uint64_t n_items = (job.n_vtiles + 31) >> 5;
    const uint64_t n_warps = (uint64_t)gridDim.x * kVWarps;
    uint64_t item = (uint64_t)blockIdx.x * kVWarps + wid, item_next = item + n_warps;
    VDesc cur = vt_desc(job, item, lane);
#if TIDK_V_TMA
    if (item < n_items)
        row_off = vt_stage_tma(job.planes, cur.d.slice, (uint32_t)((cur.sd.beg + (uint64_t)kVStep * cur.d.t) >> 5), cur.valid, ws.stage[0], mbar, lane);
#else
    if (item < n_items) vt_stage_async(job.planes, (uint32_t)((cur.sd.beg + (uint64_t)kVStep * cur.d.t) >> 5), ws.stage[0], lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    VDesc nxt = vt_desc(job, item_next < n_items ? item_next : n_items, lane);
    uint32_t buf = 0;
#if TIDK_V_SYNC
    for (;;) { // the warps of a CTA enter every item together: they run the same stretch of code at about the same time
        const bool active_ = item < n_items;
        if (!__syncthreads_or(active_)) break;
        if (!active_) continue;
#else
    while (item < n_items) {
#endif
        uint32_t ticket = 0;
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(job.work_counter) : "memory");
#if TIDK_V_TMA
        uint32_t *const stage = ws.stage[0];
        uint4 *const stash_lane = ws.stash + lane;
#else
        uint32_t *const stage = ws.stage[buf];
        uint4 *const stash_lane = reinterpret_cast<uint4 *>(stage) + lane;
        if (item_next < n_items) vt_stage_async(job.planes, (uint32_t)((nxt.sd.beg + (uint64_t)kVStep * nxt.d.t) >> 5), ws.stage[buf ^ 1], lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif

        const VTile d = cur.d;
        const SliceDesc sd = cur.sd;
        const bool valid = cur.valid;
        const uint64_t rel0 = (uint64_t)kVStep * d.t, a0 = sd.beg + rel0, m = sd.len;
        const bool first = d.t == 0, last = (uint64_t)d.t + 1 == vt_count(m);
        const uint32_t wa = (uint32_t)(a0 >> 5), sft = (uint32_t)a0 & 31u; // (the host checks that word indices fit 32 bits)
        // every word of the tile that holds a base of the slice must be plain
        bool plain;
        {
            const uint64_t end = a0 + kVTile < sd.beg + m ? a0 + kVTile : sd.beg + m; // exclusive, > a0 for a valid tile
            const uint32_t nw = valid ? (uint32_t)((end - 1) >> 5) - wa + 1u : 0u;    // 1 .. 33
            const uint32_t *pb = job.plainbits + (wa >> 5);
            const uint32_t p0 = __ldg(pb), p1 = __ldg(pb + 1), p2 = __ldg(pb + 2);
            const uint32_t x = __funnelshift_r(p0, p1, wa & 31u), y = __funnelshift_r(p1, p2, wa & 31u);
            const uint32_t need = nw >= 32u ? 0xFFFFFFFFu : (1u << nw) - 1u;
            plain = (x & need) == need && (nw < 33u || (y & 1u));
        }
        // this item's words have arrived (all but the group just committed are complete); each lane takes its own
        // 33 + 33, shifted to the tile's first base
#if TIDK_V_TMA
        vt_mbar_wait(mbar, mb_parity);
        mb_parity ^= 1u;
        uint32_t lo[32], hi[32];
        {
            const uint32_t *sl = stage + row_off, *sh2 = stage + kVPlaneWords + row_off; // (rows of a run are 30 words apart: two-way bank conflicts)
            uint32_t pl = sl[0], ph = sh2[0];
#pragma unroll
            for (int b = 0; b < 32; b++) {
                const uint32_t nl = sl[b + 1], nh = sh2[b + 1];
                lo[b] = __funnelshift_r(pl, nl, sft);
                hi[b] = __funnelshift_r(ph, nh, sft);
                pl = nl; ph = nh;
            }
        }
        __syncwarp(); // every lane has its rows in registers: the staging buffer takes the next item's
        if (item_next < n_items)
            row_off_next = vt_stage_tma(job.planes, nxt.d.slice, (uint32_t)((nxt.sd.beg + (uint64_t)kVStep * nxt.d.t) >> 5), nxt.valid, ws.stage[0], mbar, lane);
#else
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        uint32_t lo[32], hi[32];
        {
            const uint32_t *sl = stage + lane * 33, *sh2 = stage + 32 * 33 + lane * 33;
            uint32_t pl = sl[0], ph = sh2[0];
#pragma unroll
            for (int b = 0; b < 32; b++) {
                const uint32_t nl = sl[b + 1], nh = sh2[b + 1];
                lo[b] = __funnelshift_r(pl, nl, sft);
                hi[b] = __funnelshift_r(ph, nh, sft);
                pl = nl; ph = nh;
            }
        }
        __syncwarp(); // the staging buffer becomes the stash
#endif
        const uint32_t h0 = lo[0] | hi[0];
        vt_transpose32(lo);
        vt_transpose32(hi);
        // the item after next: its ticket has had time to arrive; fetch its descriptors now, use them at the end
        const uint64_t item_after = 2 * n_warps + __shfl_sync(0xFFFFFFFFu, ticket, 0);
        const VDesc aft = vt_desc(job, item_after < n_items ? item_after : n_items, lane);

        const bool scan = valid && plain;
        if (valid && !plain) { // left to the general kernel, as an ordinary item that owns the same chunk ends
            const unsigned long long idx = atomicAdd(job.n_slow, 1ull);
            job.slow_items[idx] = make_ex_item(sd, d.slice, first ? 0ull : rel0 + kVHalo, last ? m + k1 : rel0 + kVTile);
        }
        // tile-relative bounds: the tile owns the chunk ends in [own_lo, own_hi); the closing parity counts [par_lo, m_t)
        const uint32_t m_t = m - rel0 < (uint64_t)kVTile ? (uint32_t)(m - rel0) : kVTile;
        const uint32_t own_hi = last ? m_t : kVTile;
        uint32_t evcount = 0; // events of this item so far (warp-uniform)
        for (uint32_t k = k0; k <= k1; k++) {
            uint32_t any = 0;
            switch (k) {
#define TIDK_VK(K)                                                                                          \
    case K:                                                                                                 \
        if constexpr (KMIN == 0 || (K >= KMIN && K <= KMAX)) any = vt_run_k<K>(lo, hi, sh.mask, d.t, stash_lane); \
        break;
                TIDK_VK(1) TIDK_VK(2) TIDK_VK(3) TIDK_VK(4) TIDK_VK(5) TIDK_VK(6) TIDK_VK(7) TIDK_VK(8)
                TIDK_VK(9) TIDK_VK(10) TIDK_VK(11) TIDK_VK(12) TIDK_VK(13) TIDK_VK(14) TIDK_VK(15) TIDK_VK(16)
#undef TIDK_VK
            default: break;
            }
            // Run boundaries of this k (vt_emit, explore_vertical_core.hpp).  Few lanes have any, and those that do
            // have one or two: the lanes step through "my next non-zero E word" / "its next bit" TOGETHER (ballots
            // decide when to stop), so the warp pays for the busiest lane instead of for the sum over all lanes, and
            // event slots come from a ballot, not from an atomic per event.
            const uint32_t iflip = 2u * k - 1u; // (< 32: k <= 16)
            const bool flip = first && vt_head_flip(h0, k);
            const bool mine = scan && (any != 0u || flip);
            if (!__any_sync(0xFFFFFFFFu, mine)) continue;
            const uint32_t fg = iflip >> 2, fc = iflip & 3u;
            uint32_t gm = 0; // groups of four E words with a bit in them
            if (mine) {
#pragma unroll
                for (uint32_t g = 0; g < 8; g++) {
                    const uint4 e4 = stash_lane[g * 32];
                    if ((e4.x | e4.y | e4.z | e4.w) != 0u) gm |= 1u << g;
                }
                if (flip) gm |= 1u << fg;
            }
            const uint32_t par_lo = first ? iflip : 0u, own_lo = first ? iflip : kVHalo;
            uint32_t parity = 0, pend = 0, pi = 0, cw = 0, cg = 0;
            uint4 e4 = make_uint4(0u, 0u, 0u, 0u);
            for (;;) {
                while (pend == 0u && (cw | gm) != 0u) { // next non-zero word of this lane (a group can turn out empty: the head flip)
                    if (cw == 0u) {
                        cg = (uint32_t)__ffs((int)gm) - 1u;
                        gm &= gm - 1u;
                        e4 = stash_lane[cg * 32];
                        if (flip && cg == fg) {
                            if (fc == 0u) e4.x ^= 1u; else if (fc == 1u) e4.y ^= 1u; else if (fc == 2u) e4.z ^= 1u; else e4.w ^= 1u;
                        }
                        cw = (e4.x != 0u ? 1u : 0u) | (e4.y != 0u ? 2u : 0u) | (e4.z != 0u ? 4u : 0u) | (e4.w != 0u ? 8u : 0u);
                    }
                    if (cw != 0u) {
                        const uint32_t c = (uint32_t)__ffs((int)cw) - 1u;
                        cw &= cw - 1u;
                        pend = c == 0u ? e4.x : c == 1u ? e4.y : c == 2u ? e4.z : e4.w;
                        pi = 4u * cg + c;
                    }
                }
                if (!__any_sync(0xFFFFFFFFu, pend != 0u)) break;
                bool has_ev = false;
                uint64_t key = 0;
                if (pend != 0u) {
                    const uint32_t b = (uint32_t)__ffs((int)pend) - 1u;
                    pend &= pend - 1u;
                    const uint32_t q = 32u * b + pi; // tile position
                    if (q >= par_lo && q < m_t) parity ^= 1u;
                    if (q >= own_lo && q < own_hi) {
                        const uint64_t q1 = rel0 + q + 1; // slice position + 1: a multiple of k
                        const uint64_t j = (q1 >> 32 ? q1 / k : (uint64_t)div_exact((uint32_t)q1, k)) - 2;
                        key = pack_event(job.fmt, d.slice, k, j, 0u, 0u);
                        has_ev = true;
                    }
                }
                const uint32_t em = __ballot_sync(0xFFFFFFFFu, has_ev);
                if (has_ev) vt_put(job, evbuf, evcount + (uint32_t)__popc(em & ((1u << lane) - 1u)), key);
                evcount += (uint32_t)__popc(em);
            }
            { // the run that touches the end of the slice
                const bool closer = mine && last && parity != 0u;
                const uint32_t em = __ballot_sync(0xFFFFFFFFu, closer);
                if (closer) vt_put(job, evbuf, evcount + (uint32_t)__popc(em & ((1u << lane) - 1u)), pack_event(job.fmt, d.slice, k, m / k - 1, 0u, 0u));
                evcount += (uint32_t)__popc(em);
            }
            __syncwarp();
        }
        // the warp's events of these 32 tiles: one range of global slots
        __syncwarp();
        {
            const uint32_t n = min(evcount, kVEvBuf);
            if (n) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(job.n_events, (unsigned long long)n);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                for (uint32_t e = (uint32_t)lane; e < n; e += 32)
                    if (base + e < job.cap_events) job.events[base + e] = evbuf[e];
            }
        }
        __syncwarp();
#if TIDK_V_TMA
        row_off = row_off_next;
#endif
        item = item_next;
        item_next = item_after;
        cur = nxt;
        nxt = aft;
        buf ^= 1;
    }
#if !TIDK_V_TMA
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

} // namespace tidk
