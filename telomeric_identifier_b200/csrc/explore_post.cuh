// explore_post.cuh -- events -> runs on the device (packed-key path).
//
// calculate_indexes (explore.rs:289-347) + the count half of filter_by_frequency (explore.rs:265-274)
// for millions of events: a radix sort of the 64-bit keys puts them into (k, record, slice, position)
// order -- the reference's FASTA-order arrival order -- where they alternate start / end per
// (slice, k), beginning with a start, so keys 2i and 2i + 1 are one raw run; one thread per raw run decides whether it heads a (merged) run, walks the chain of
// raw runs that abut and agree after upper-casing, applies the first-run rule (start = 0,
// explore.rs:297) and the threshold; an order-preserving compaction keeps the survivors.
// The sort and the compaction are CUB device primitives (plumbing, not the hot path).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <mutex>
#include <vector>

#include "explore.cuh"

namespace tidk {

// Run lists leave the device into pinned host blocks (a 6 MB list: 0.15 ms by DMA against 0.5 ms through the
// driver's staging of a pageable destination).  Pinning costs more than the copy, so the blocks are pooled: tidk_cuda_free_runs
// hands a block back, the next call takes the smallest one that fits.  The pool belongs to its CONTEXT (no process-wide
// state): every run list, pinned or plain, sits behind a 64-byte header that names its pool, so tidk_cuda_free_runs --
// which takes only the pointer -- finds its way home.  tidk_cuda_destroy unpins the idle blocks; a list still held by the
// caller at that point keeps the (then closed) pool alive until it is freed.
struct RunBlockPool;
struct RunListHeader {
    uint64_t magic;
    RunBlockPool *pool; // nullptr: plain malloc
    uint64_t pad[6];
};
static_assert(sizeof(RunListHeader) == 64, "run lists stay 64-byte aligned behind their header");
constexpr uint64_t kRunListMagic = 0x7469646b52756e73ull; // "tidkRuns"

struct RunBlockPool {
    struct Block { void *p; size_t cap; bool in_use; };
    std::mutex mu;
    std::vector<Block> blocks;
    bool closed = false; // the context is gone: returned blocks are unpinned, the last one deletes the pool
    tidk_run *take(size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        const size_t need = bytes + sizeof(RunListHeader);
        Block *best = nullptr;
        for (Block &b : blocks)
            if (!b.in_use && b.cap >= need && (!best || b.cap < best->cap)) best = &b;
        if (!best) {
            void *p = nullptr;
            const size_t cap = need + need / 4 + 4096;
            if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
            blocks.push_back({p, cap, false});
            best = &blocks.back();
        }
        best->in_use = true;
        RunListHeader *h = (RunListHeader *)best->p;
        h->magic = kRunListMagic;
        h->pool = this;
        return (tidk_run *)(h + 1);
    }
    void give_back(RunListHeader *h) {
        bool last = false;
        {
            std::lock_guard<std::mutex> g(mu);
            size_t idle = 0, busy = 0;
            for (Block &b : blocks) { idle += !b.in_use; busy += b.in_use; }
            for (size_t i = 0; i < blocks.size(); i++) {
                if (blocks[i].p != (void *)h) continue;
                if (closed || idle >= 4) { // enough idle blocks already (or nobody left to reuse it): unpin this one
                    cudaFreeHost(blocks[i].p);
                    blocks.erase(blocks.begin() + (long)i);
                } else {
                    blocks[i].in_use = false;
                }
                busy--;
                break;
            }
            last = closed && busy == 0;
        }
        if (last) delete this;
    }
    // tidk_cuda_destroy: unpin what is idle; lists the caller still holds stay valid
    static void close(RunBlockPool *pool) {
        if (!pool) return;
        bool last = false;
        {
            std::lock_guard<std::mutex> g(pool->mu);
            pool->closed = true;
            for (size_t i = pool->blocks.size(); i-- > 0;)
                if (!pool->blocks[i].in_use) {
                    cudaFreeHost(pool->blocks[i].p);
                    pool->blocks.erase(pool->blocks.begin() + (long)i);
                }
            last = pool->blocks.empty();
        }
        if (last) delete pool;
    }
};

// a run list outside any pool (small lists, host-built lists)
inline tidk_run *run_list_alloc_plain(size_t bytes) {
    RunListHeader *h = (RunListHeader *)malloc(bytes + sizeof(RunListHeader));
    if (!h) return nullptr;
    h->magic = kRunListMagic;
    h->pool = nullptr;
    return (tidk_run *)(h + 1);
}
inline void run_list_free(tidk_run *runs) {
    if (!runs) return;
    RunListHeader *h = (RunListHeader *)runs - 1;
    if (h->magic != kRunListMagic) return; // not one of ours: leave it alone rather than corrupt the heap
    if (h->pool) h->pool->give_back(h); else { h->magic = 0; free(h); }
}

__global__ void __launch_bounds__(256) runs_from_keys_kernel(const uint64_t *keys, KeyFmt fmt, uint64_t n_pairs, uint64_t threshold,
                                                            tidk_run *runs, uint8_t *flags, uint32_t *err_flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    uint32_t sl, k, ty, jn, sl2, k2, ty2, jn2;
    uint64_t s, e;
    unpack_event(fmt, keys[2 * i], sl, k, s, ty, jn);
    unpack_event(fmt, keys[2 * i + 1], sl2, k2, e, ty2, jn2);
    // (the type bits are advisory: the plane kernel only fills them in mixed-case blocks; the pairing is by rank)
    if (sl != sl2 || k != k2 || e <= s) { atomicOr(err_flags, 4u); flags[i] = 0; return; }
    bool first = true, head = true;
    if (i > 0) {
        uint32_t psl, pk, pty, pjn; uint64_t pe;
        unpack_event(fmt, keys[2 * i - 1], psl, pk, pe, pty, pjn);
        if (psl == sl && pk == k) {
            first = false;
            head = !(jn && s == pe + 1); // same run after upper-casing (explore.rs:324)
        }
    }
    if (!head) { flags[i] = 0; return; }
    uint64_t re = e;
    for (uint64_t j = i + 1; j < n_pairs; j++) { // absorb the raw runs that join this one
        uint32_t nsl, nk, nty, njn; uint64_t ns, ne;
        unpack_event(fmt, keys[2 * j], nsl, nk, ns, nty, njn);
        if (nsl != sl || nk != k || !njn || ns != re + 1) break;
        unpack_event(fmt, keys[2 * j + 1], nsl, nk, ne, nty, njn);
        re = ne;
    }
    const uint64_t start = first ? 0 : s * k; // explore.rs:297
    const uint64_t end = (re + 1) * k;
    const bool keep = (end - start) / k > threshold;
    flags[i] = keep ? 1 : 0;
    if (keep) {
        tidk_run r;
        r.record = sl >> 1; r.slice = (uint8_t)(sl & 1); r.k = (uint8_t)k;
        r.first_in_slice = first ? 1 : 0; r.pad = 0;
        r.start = start; r.end = end; r.label_pos = s * k;
        runs[i] = r;
    }
}

// keys: n_events packed keys in sc.events.  On success *out_runs (a block of sc.run_pool, or a plain allocation) / *n_runs hold the result.
// n_sorted >= n_events keys sit in sc.events; the surplus are all-ones fillers, which the sort sends behind every event
inline int explore_post_device(ExploreScratch &sc, cudaStream_t st, uint64_t n_events, uint64_t n_sorted, KeyFmt fmt, int key_bits, uint64_t threshold,
                               tidk_run **out_runs, uint64_t *n_runs, uint32_t *launches, std::string &err) {
    auto cuerr = [&](cudaError_t e, const char *what) {
        err = std::string(what) + " failed: " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? TIDK_ENOMEM : TIDK_ECUDA;
    };
    auto grow = [&](void *&p, size_t &cap, size_t need) -> cudaError_t {
        if (p && need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e2 = cudaMalloc(&p, need + need / 8 + 256);
        if (e2 == cudaSuccess) cap = need + need / 8 + 256; else p = nullptr;
        return e2;
    };
    *out_runs = nullptr; *n_runs = 0;
    if (n_events == 0) return TIDK_OK;
    if (n_events & 1) { err = "internal: odd number of run events"; return TIDK_ECUDA; }
    const uint64_t n_pairs = n_events / 2;
    if (n_sorted > 0x7FFFFFFFull) { err = "too many run events for the device path"; return TIDK_EARG; }
    cudaError_t e;
    if ((e = grow(sc.keys2, sc.keys2_cap, n_sorted * 8)) != cudaSuccess) return cuerr(e, "cudaMalloc(keys)");
    size_t temp_sort = 0, temp_sel = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, temp_sort, (const uint64_t *)sc.events, (uint64_t *)sc.keys2, (int)n_sorted, 0, key_bits, st);
    cub::DeviceSelect::Flagged(nullptr, temp_sel, (const tidk_run *)nullptr, (const uint8_t *)nullptr, (tidk_run *)nullptr,
                               (unsigned long long *)nullptr, (int)n_pairs, st);
    if ((e = grow(sc.temp, sc.temp_cap, std::max(temp_sort, temp_sel))) != cudaSuccess) return cuerr(e, "cudaMalloc(temp)");
    if ((e = grow(sc.runs, sc.runs_cap, n_pairs * sizeof(tidk_run))) != cudaSuccess) return cuerr(e, "cudaMalloc(runs)");
    if ((e = grow(sc.runs_out, sc.runs_out_cap, n_pairs * sizeof(tidk_run))) != cudaSuccess) return cuerr(e, "cudaMalloc(runs_out)");
    if ((e = grow(sc.flags, sc.flags_cap, n_pairs)) != cudaSuccess) return cuerr(e, "cudaMalloc(flags)");
    size_t tb = sc.temp_cap;
    if ((e = cub::DeviceRadixSort::SortKeys(sc.temp, tb, (const uint64_t *)sc.events, (uint64_t *)sc.keys2, (int)n_sorted, 0, key_bits, st)) != cudaSuccess)
        return cuerr(e, "radix sort");
    uint32_t *d_err = (uint32_t *)((char *)sc.misc + 128);
    unsigned long long *d_nsel = (unsigned long long *)sc.misc + 4;
    runs_from_keys_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>((const uint64_t *)sc.keys2, fmt, n_pairs, threshold,
                                                                          (tidk_run *)sc.runs, (uint8_t *)sc.flags, d_err);
    tb = sc.temp_cap;
    if ((e = cub::DeviceSelect::Flagged(sc.temp, tb, (const tidk_run *)sc.runs, (const uint8_t *)sc.flags, (tidk_run *)sc.runs_out,
                                        d_nsel, (int)n_pairs, st)) != cudaSuccess)
        return cuerr(e, "select");
    if (launches) *launches += 3; // sort (several kernels inside CUB), run builder, compaction
    unsigned long long h_nsel = 0;
    uint32_t h_err = 0;
    if ((e = cudaMemcpyAsync(&h_nsel, d_nsel, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuerr(e, "D2H count");
    if ((e = cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuerr(e, "D2H flags");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuerr(e, "explore post-processing");
    if (h_err & 4u) { err = "internal: run events do not alternate start / end"; return TIDK_ECUDA; }
    if (h_nsel) {
        const size_t bytes = h_nsel * sizeof(tidk_run);
        const bool pooled = bytes >= (256u << 10); // small lists: a plain allocation and a staged copy are as fast
        if (pooled && !sc.run_pool) sc.run_pool = new RunBlockPool;
        *out_runs = pooled ? sc.run_pool->take(bytes) : nullptr;
        if (!*out_runs) *out_runs = run_list_alloc_plain(bytes);
        if (!*out_runs) { err = "out of host memory"; return TIDK_ENOMEM; }
        if ((e = cudaMemcpyAsync(*out_runs, sc.runs_out, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess) {
            run_list_free(*out_runs);
            *out_runs = nullptr;
            return cuerr(e, "D2H runs");
        }
    }
    *n_runs = h_nsel;
    return TIDK_OK;
}

} // namespace tidk
