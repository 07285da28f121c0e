// explore.cuh -- tandem k-mer run detection in the chromosome-end slices.
//
// Replaces split_seq_by_distance + chunk_fasta + calculate_indexes
// (explore.rs:151-160, 178-233, 289-347) for every record, slice and k.
//
// What the reference computes, restated per slice t[0..m) and chunk length k:
//   chunks   c_i = t[i*k .. min((i+1)*k, m))
//   raw_i    = (c_i == c_{i+1}) on RAW bytes, lengths included        (explore.rs:208)
//   a "raw run" is a maximal stretch raw_s .. raw_{e-1}; its entries are chunks s..e
//   two raw runs that abut (next s == e + 1) and whose chunks agree after
//   upper-casing are one run (labels are upper-cased before calculate_indexes
//   compares them, explore.rs:212 vs :324)
//   run -> start = s*k (0 for the first run of the slice, explore.rs:297), end = (e+1)*k
//
// Device side (this file): a streaming pass over the slices that emits one
// 16-byte EVENT per raw-run boundary (run starts carry the "agrees with the
// previous chunk after upper-casing" bit).  Events are sparse (~4^-k per chunk on
// random sequence), so the pass is read-dominated.  Host side: sort events into
// (k, record, slice, chunk) order -- the reference's FASTA-order arrival order --
// pair them into runs, merge, apply the first-run rule and the threshold.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/tidk_cuda.h"

namespace tidk {

constexpr int kMaxExploreK = 255;
constexpr uint32_t kExploreTile = 65536; // bytes of a slice per work item

struct SliceDesc {
    uint64_t beg; // absolute offset in the sequence buffer
    uint64_t len; // m
};

struct RunEvent {
    uint32_t slice;   // 2 * record + (0 left | 1 right)
    uint8_t k;
    uint8_t type;     // 0 = raw run starts at `chunk`, 1 = raw run's last entry is `chunk`
    uint8_t joins;    // start only: upper(c_{chunk-1}) == upper(c_chunk)
    uint8_t pad;
    uint64_t chunk;
};

// Packed form of an event: one sortable 64-bit key (k | slice | chunk | type | joins), used whenever
// the batch fits 64 bits; sorting the keys puts the events into the reference's arrival order
// (k, record, slice, position).  The field widths follow the batch (KeyFmt), so that the radix sort only
// runs over the bits that are in use: 4 passes for C3 (60 slices), 5 for two million reads, not 8.
struct KeyFmt {
    uint32_t cb, sb; // bits of the chunk and slice fields; k sits above them, type and joins below
    __host__ __device__ uint32_t total_bits(uint32_t kmax) const {
        uint32_t kb = 1;
        while ((1u << kb) <= kmax + 1) kb++; // a real k is never all ones: the all-ones key is the filler of unused event slots
        return kb + sb + cb + 2;
    }
};
__host__ inline uint32_t bits_for(uint64_t max_value) { // smallest b with max_value < 2^b (at least 1)
    uint32_t b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}
__host__ __device__ inline uint64_t pack_event(KeyFmt f, uint32_t slice, uint32_t k, uint64_t chunk, uint32_t type, uint32_t joins) {
    return ((uint64_t)k << (f.sb + f.cb + 2)) | ((uint64_t)slice << (f.cb + 2)) | (chunk << 2) | ((uint64_t)type << 1) | joins;
}
__host__ __device__ inline void unpack_event(KeyFmt f, uint64_t key, uint32_t &slice, uint32_t &k, uint64_t &chunk, uint32_t &type, uint32_t &joins) {
    k = (uint32_t)(key >> (f.sb + f.cb + 2));
    slice = (uint32_t)((key >> (f.cb + 2)) & ((1ull << f.sb) - 1));
    chunk = (key >> 2) & ((1ull << f.cb) - 1);
    type = (uint32_t)((key >> 1) & 1u);
    joins = (uint32_t)(key & 1u);
}

struct ExploreJob {
    const uint8_t *seq;
    const SliceDesc *slices;
    const uint64_t *tile_base; // [n_slices + 1] tiles before slice s
    uint64_t n_tiles;
    uint32_t n_slices;
    uint32_t kmin, kmax;
    RunEvent *events;
    unsigned long long *n_events; // counts every event, stored or not
    uint64_t cap_events;
    unsigned long long *work_counter;
    uint32_t *err_flags;
    uint32_t packed; // events are 64-bit keys (pack_event) instead of RunEvent structs
    KeyFmt fmt;
};

struct RunBlockPool; // pinned host blocks for run lists (explore_post.cuh); owned by the context, closed by tidk_cuda_destroy
struct ExploreScratch {
    RunBlockPool *run_pool = nullptr;
    void *slices = nullptr, *tile_base = nullptr, *tiles = nullptr, *events = nullptr, *misc = nullptr, *items = nullptr;
    void *keys2 = nullptr, *temp = nullptr, *runs = nullptr, *flags = nullptr, *runs_out = nullptr;
    void *vtiles = nullptr; // tile list of the vertical scan (explore_vertical.cuh)
    size_t vtiles_cap = 0;
    size_t slices_cap = 0, tile_cap = 0, tiles_cap = 0, events_cap = 0, items_cap = 0;
    size_t keys2_cap = 0, temp_cap = 0, runs_cap = 0, flags_cap = 0, runs_out_cap = 0;
};

inline void explore_scratch_release(ExploreScratch &s) {
    for (void *p : {s.slices, s.tile_base, s.tiles, s.events, s.misc, s.items, s.keys2, s.temp, s.runs, s.flags, s.runs_out, s.vtiles})
        if (p) cudaFree(p);
    s = ExploreScratch();
}

__device__ __forceinline__ uint8_t dev_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// raw_i for chunk pair (i, i+1) of a slice: both chunks full and bytewise equal
__device__ __forceinline__ bool chunk_pair_equal(const uint8_t *t, uint64_t m, uint32_t k, uint64_t i) {
    if ((i + 2) * k > m) return false;
    const uint8_t *a = t + i * k;
    for (uint32_t j = 0; j < k; j++)
        if (a[j] != a[j + k]) return false;
    return true;
}

__device__ __forceinline__ bool chunk_pair_equal_upper(const uint8_t *t, uint64_t m, uint32_t k, uint64_t i) {
    if ((i + 2) * k > m) return false;
    const uint8_t *a = t + i * k;
    for (uint32_t j = 0; j < k; j++)
        if (dev_upper(a[j]) != dev_upper(a[j + k])) return false;
    return true;
}

// v1: byte-wise compare, one lane per chunk, a warp per 64 KiB tile of a slice,
// all k in one pass so a tile is read from HBM once and re-read from L1/L2.
__global__ void __launch_bounds__(256) explore_events_kernel(const ExploreJob job) {
    const int lane = threadIdx.x & 31;
    uint32_t nonascii = 0;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(job.work_counter, 1ull);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= job.n_tiles) break;
        uint32_t lo = 0, hi = job.n_slices;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (job.tile_base[mid] <= item) lo = mid; else hi = mid;
        }
        const uint32_t sl = lo;
        const SliceDesc sd = job.slices[sl];
        const uint64_t tbeg = (item - job.tile_base[sl]) * (uint64_t)kExploreTile;
        const uint64_t tend = (tbeg + kExploreTile < sd.len) ? tbeg + kExploreTile : sd.len;
        const uint8_t *t = job.seq + sd.beg;
        // non-ASCII check over the tile (str::from_utf8(a).unwrap() panics, explore.rs:212)
        for (uint64_t p = tbeg + (uint64_t)lane * 4; p < tend; p += 128) {
            uint32_t acc = 0;
            for (int j = 0; j < 4 && p + j < tend; j++) acc |= t[p + j];
            nonascii |= acc & 0x80u;
        }
        for (uint32_t k = job.kmin; k <= job.kmax; k++) {
            if (sd.len <= k) continue; // explore.rs:187-195
            const uint64_t nchunks = (sd.len + k - 1) / k;
            // chunks whose first byte lies in this tile
            const uint64_t c0 = (tbeg + k - 1) / k;
            uint64_t c1 = (tend + k - 1) / k;
            if (c1 > nchunks) c1 = nchunks;
            for (uint64_t i = c0 + lane; i < c1; i += 32) {
                const bool raw = chunk_pair_equal(t, sd.len, k, i);
                const bool rawp = i > 0 && chunk_pair_equal(t, sd.len, k, i - 1);
                if (raw == rawp) continue;
                RunEvent ev;
                ev.slice = sl; ev.k = (uint8_t)k; ev.pad = 0; ev.chunk = i;
                if (raw) { // run starts here
                    ev.type = 0;
                    ev.joins = (i > 0 && chunk_pair_equal_upper(t, sd.len, k, i - 1)) ? 1 : 0;
                } else {   // chunk i is the last entry of a run
                    ev.type = 1; ev.joins = 0;
                }
                unsigned long long idx = atomicAdd(job.n_events, 1ull);
                if (idx < job.cap_events) {
                    if (job.packed) ((uint64_t *)job.events)[idx] = pack_event(job.fmt, ev.slice, ev.k, ev.chunk, ev.type, ev.joins);
                    else job.events[idx] = ev;
                }
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, nonascii != 0) && lane == 0) atomicOr(job.err_flags, 1u);
}

inline uint64_t split_dist(uint64_t len, double d) { // explore.rs:156
    double v = std::ceil((double)len * d);
    if (!(v > 0.0)) return 0;
    return (uint64_t)v;
}

// Host half: events -> runs (calculate_indexes, explore.rs:289-347, plus the
// count > threshold half of filter_by_frequency, explore.rs:265-274).
inline void events_to_runs(std::vector<RunEvent> &ev, uint64_t threshold, std::vector<tidk_run> &out) {
    std::sort(ev.begin(), ev.end(), [](const RunEvent &a, const RunEvent &b) {
        if (a.k != b.k) return a.k < b.k;
        if (a.slice != b.slice) return a.slice < b.slice;
        if (a.chunk != b.chunk) return a.chunk < b.chunk;
        return a.type < b.type;
    });
    size_t i = 0;
    const size_t n = ev.size();
    while (i < n) {
        const uint32_t sl = ev[i].slice;
        const uint8_t k = ev[i].k;
        bool have = false, first = true;
        uint64_t rs = 0, re = 0; // current (merged) run: entry chunks rs..re
        auto flush = [&]() {
            if (!have) return;
            uint64_t start = first ? 0 : rs * k; // explore.rs:297
            uint64_t end = (re + 1) * k;
            if ((end - start) / k > threshold) {
                tidk_run r;
                r.record = sl >> 1; r.slice = (uint8_t)(sl & 1); r.k = k;
                r.first_in_slice = first ? 1 : 0; r.pad = 0;
                r.start = start; r.end = end; r.label_pos = rs * k;
                out.push_back(r);
            }
            first = false;
            have = false;
        };
        while (i < n && ev[i].slice == sl && ev[i].k == k) {
            // events alternate start / end within a (slice, k)
            const RunEvent &s = ev[i];
            const RunEvent &e = ev[i + 1]; // always present: every raw run is closed
            if (have && s.joins && s.chunk == re + 1) {
                re = e.chunk; // same run after upper-casing (explore.rs:324)
            } else {
                flush();
                rs = s.chunk; re = e.chunk; have = true;
            }
            i += 2;
        }
        flush();
    }
}

} // namespace tidk
