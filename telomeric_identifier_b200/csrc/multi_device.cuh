// multi_device.cuh -- a context over several devices (tidk_cuda_create with n_devices > 1).
//
// the reference's drivers are one sequential loop in one process (search.rs:57, finder.rs:85,
// explore.rs:97-100), so a host that is to use all GPUs of a box through the C ABI needs ONE context that fans a call
// out.  The context holds one ordinary single-device context per device id (the same id may be listed more than once:
// two contexts on one device, which is what a one-GPU test box can check) and no device state of its own.  There is no
// exchange step between devices:
//   window counts, host pointers   the global window sequence is cut into equal contiguous ranges -- whole records where
//                                  possible, long records at multiples of the window, so no halo (an occurrence never
//                                  spans two windows, search.rs:98,105-110) -- each device counts its range straight from
//                                  the caller's buffer on its own thread, the counts are placed into [record][query][window]
//   explore, host pointers         whole records in contiguous groups balanced by scanned bases; a record's runs do not
//                                  depend on other records, each device returns (k, record, slice, start) order, so the
//                                  reference's arrival order is the groups in record order within each k
//   resident batches               whole records in contiguous groups balanced by bases (the window is not known at upload
//                                  time); every later call runs on all parts, counts land in place (a group's counts are
//                                  one contiguous block of the [record][query][window] arrays)
// included by tidk_cuda.cu (host code only).
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace multi {

// One parked host thread per device but the first (which the caller's thread serves).  A call is fanned out by bumping a
// generation counter; the workers spin on it for a short while after each job -- a loop of asynchronous passes then finds
// them awake and a dispatch costs about a microsecond -- and sleep on a condition variable otherwise.  (Spawning the
// threads per call cost 100-150 us on an 8-device context: three times the kernel of an eighth of the C4 genome.)
struct Workers {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> gen{0};
    std::atomic<int> pending{0};
    std::atomic<bool> quit{false};
    std::function<void(int)> job;
    explicit Workers(int n) {
        for (int g = 1; g < n; g++) th.emplace_back([this, g] { loop(g); });
    }
    ~Workers() {
        quit.store(true);
        { std::lock_guard<std::mutex> l(mu); gen.fetch_add(1); }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    void loop(int g) {
        uint64_t seen = 0;
        for (;;) {
            uint64_t cur = gen.load(std::memory_order_acquire);
            if (cur == seen) { // spin briefly, then sleep
                const auto t0 = std::chrono::steady_clock::now();
                while ((cur = gen.load(std::memory_order_acquire)) == seen) {
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                        std::unique_lock<std::mutex> l(mu);
                        cv.wait(l, [&] { return gen.load(std::memory_order_acquire) != seen; });
                        cur = gen.load(std::memory_order_acquire);
                        break;
                    }
                }
            }
            seen = cur;
            if (quit.load()) return;
            job(g);
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    // f(g) for g = 0 .. n - 1, g = 0 on the calling thread; returns when all are done.  Not re-entrant (a tidk_ctx is single-caller).
    void run(const std::function<void(int)> &f) {
        if (th.empty()) { f(0); return; }
        job = f;
        pending.store((int)th.size(), std::memory_order_release);
        { std::lock_guard<std::mutex> l(mu); gen.fetch_add(1, std::memory_order_release); }
        cv.notify_all();
        f(0);
        while (pending.load(std::memory_order_acquire) != 0) { /* the others finish within microseconds of the caller's share */
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
};

struct Piece { uint32_t rec; uint64_t first, n; };

inline uint64_t n_windows(uint64_t len, uint64_t window) { return len / window + (len % window != 0); }

// equal contiguous ranges of the global window sequence
inline std::vector<std::vector<Piece>> plan_shards(const uint64_t *offsets, uint32_t n_records, uint64_t window, int world) {
    std::vector<uint64_t> base((size_t)n_records + 1, 0);
    for (uint32_t r = 0; r < n_records; r++) base[r + 1] = base[r] + n_windows(offsets[r + 1] - offsets[r], window);
    const uint64_t total = base[n_records];
    std::vector<std::vector<Piece>> plans((size_t)world);
    size_t rec = 0;
    for (int g = 0; g < world; g++) {
        uint64_t lo = total / world * g + std::min<uint64_t>(total % world, (uint64_t)g);
        const uint64_t hi = total / world * (g + 1) + std::min<uint64_t>(total % world, (uint64_t)g + 1);
        while (lo < hi) {
            while (base[rec + 1] <= lo) rec++;
            const uint64_t n = std::min(hi, base[rec + 1]) - lo;
            plans[(size_t)g].push_back({(uint32_t)rec, lo - base[rec], n});
            lo += n;
        }
    }
    return plans;
}

// contiguous groups of whole records, balanced by cost(r) (+ 1 per record: empty records still spread out)
template <class Cost>
inline std::vector<uint32_t> cut_records(uint32_t n_records, int world, Cost cost) {
    std::vector<uint64_t> acc((size_t)n_records + 1, 0);
    for (uint32_t r = 0; r < n_records; r++) acc[r + 1] = acc[r] + cost(r) + 1;
    std::vector<uint32_t> cut((size_t)world + 1, n_records);
    cut[0] = 0;
    for (int g = 1; g < world; g++) {
        const uint64_t want = acc[n_records] / (uint64_t)world * (uint64_t)g;
        uint32_t c = (uint32_t)(std::lower_bound(acc.begin(), acc.end(), want) - acc.begin());
        if (c < cut[(size_t)g - 1]) c = cut[(size_t)g - 1];
        if (c > n_records) c = n_records;
        cut[(size_t)g] = c;
    }
    return cut;
}

// f(g, sub) on one thread per device; the first failure is the call's result, its message the context's
template <class F>
inline int on_all(tidk_ctx *ctx, F f) {
    const size_t n = ctx->subs.size();
    std::vector<int> rc(n, TIDK_OK);
    if (!ctx->workers) ctx->workers = new Workers((int)n);
    ctx->workers->run([&](int g) { rc[(size_t)g] = f(g, ctx->subs[(size_t)g]); });
    for (size_t g = 0; g < n; g++)
        if (rc[g] != TIDK_OK) {
            ctx->err = "device " + std::to_string(ctx->subs[g]->device) + ": " + ctx->subs[g]->err;
            return rc[g];
        }
    return TIDK_OK;
}

inline int window_counts(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint32_t n_records, uint64_t window,
                         const char *const *queries, const uint32_t *qlens, uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev) {
    if (window == 0 || nq == 0 || !offsets || !queries || !qlens || !out_fwd || !out_rev || n_records == 0) { // the single-device path words the error
        const int rc = tidk_cuda_window_counts(ctx->subs[0], seq, offsets, n_records, window, queries, qlens, nq, out_fwd, out_rev);
        if (rc) ctx->err = ctx->subs[0]->err;
        return rc;
    }
    const int world = (int)ctx->subs.size();
    const auto plans = plan_shards(offsets, n_records, window, world);
    std::vector<uint64_t> nw(n_records), obase((size_t)n_records + 1, 0);
    for (uint32_t r = 0; r < n_records; r++) {
        nw[r] = n_windows(offsets[r + 1] - offsets[r], window);
        obase[r + 1] = obase[r] + nw[r] * nq;
    }
    return on_all(ctx, [&](int g, tidk_ctx *sub) -> int {
        const std::vector<Piece> &pieces = plans[(size_t)g];
        sub->last_h2d = sub->last_d2h = 0;
        sub->last_packed = 0;
        if (pieces.empty()) return TIDK_OK;
        std::vector<uint64_t> loc_off(pieces.size() + 1, 0);
        uint64_t a0 = 0, n_loc = 0;
        for (size_t i = 0; i < pieces.size(); i++) {
            const Piece &p = pieces[i];
            const uint64_t o = offsets[p.rec], e = offsets[p.rec + 1];
            const uint64_t beg = o + p.first * window, end = std::min(o + (p.first + p.n) * window, e);
            if (i == 0) a0 = beg;
            loc_off[i + 1] = end - a0; // consecutive window ranges of consecutive records: one contiguous stretch
            n_loc += p.n * nq;
        }
        std::vector<uint64_t> lf(n_loc + 1), lr(n_loc + 1);
        const int rc = tidk_cuda_window_counts(sub, seq + a0, loc_off.data(), (uint32_t)pieces.size(), window, queries, qlens, nq, lf.data(), lr.data());
        if (rc) return rc;
        uint64_t pos = 0;
        for (const Piece &p : pieces)
            for (uint32_t q = 0; q < nq; q++) {
                const uint64_t dst = obase[p.rec] + q * nw[p.rec] + p.first;
                memcpy(out_fwd + dst, &lf[pos], p.n * 8);
                memcpy(out_rev + dst, &lr[pos], p.n * 8);
                pos += p.n;
            }
        return TIDK_OK;
    });
}

// The groups' run lists, in record order within each k.  Every group's list is sorted by (k, record, slice, start) already
// (explore_post.cuh), so the merged list is, for k = kmin .. kmax, the groups' k-segments one behind the other: segment
// boundaries by binary search, then every group copies its segments to their places on its own thread and shifts the record
// numbers on the way (a stable sort of the concatenation took 13 ms on 400 000 runs: longer than the scan it followed).
// The lists are handed back to their contexts' pools here.
struct PartRuns { tidk_run *runs = nullptr; uint64_t n = 0; uint32_t r0 = 0; };
inline int merge_runs(tidk_ctx *ctx, std::vector<PartRuns> &part, tidk_run **out_runs, uint64_t *n_runs) {
    auto release = [&]() {
        for (PartRuns &p : part) { if (p.runs) tidk_cuda_free_runs(p.runs); p.runs = nullptr; }
    };
    uint64_t total = 0;
    for (auto &p : part) total += p.n;
    *n_runs = total;
    *out_runs = nullptr;
    if (total == 0) { release(); return TIDK_OK; }
    const size_t bytes = total * sizeof(tidk_run);
    tidk_run *all = nullptr;
    if (bytes >= (256u << 10)) { // a pinned block of the first device's pool: reused from call to call, no fresh pages to fault in
        tidk::ExploreScratch &sc0 = ctx->subs[0]->ex;
        if (!sc0.run_pool) sc0.run_pool = new tidk::RunBlockPool;
        cudaSetDevice(ctx->subs[0]->device);
        all = sc0.run_pool->take(bytes);
    }
    if (!all) all = tidk::run_list_alloc_plain(bytes);
    if (!all) { *n_runs = 0; release(); return fail(ctx, TIDK_ENOMEM, "out of host memory"); }
    const size_t P = part.size();
    // seg[g][k] = first run of group g with key k (k = 0 .. 256)
    std::vector<std::array<uint64_t, 257>> seg(P);
    for (size_t g = 0; g < P; g++) {
        const tidk_run *r = part[g].runs;
        for (uint32_t k = 0; k <= 256; k++)
            seg[g][k] = (uint64_t)(std::lower_bound(r, r + part[g].n, k, [](const tidk_run &x, uint32_t kk) { return (uint32_t)x.k < kk; }) - r);
    }
    std::vector<std::array<uint64_t, 256>> dst(P); // where group g's k-segment goes
    uint64_t at = 0;
    for (uint32_t k = 0; k < 256; k++)
        for (size_t g = 0; g < P; g++) { dst[g][k] = at; at += seg[g][k + 1] - seg[g][k]; }
    auto copy_group = [&](size_t g) {
        const tidk_run *r = part[g].runs;
        const uint32_t r0 = part[g].r0;
        for (uint32_t k = 0; k < 256; k++) {
            tidk_run *o = all + dst[g][k];
            for (uint64_t i = seg[g][k]; i < seg[g][k + 1]; i++) { tidk_run t = r[i]; t.record += r0; *o++ = t; }
        }
    };
    std::vector<std::thread> th;
    for (size_t g = 1; g < P; g++) th.emplace_back(copy_group, g);
    copy_group(0);
    for (auto &t : th) t.join();
    release();
    *out_runs = all;
    return TIDK_OK;
}

inline int explore_runs(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint32_t n_records, double distance, uint32_t kmin,
                        uint32_t kmax, uint64_t threshold, tidk_run **out_runs, uint64_t *n_runs) {
    if (!out_runs || !n_runs || !offsets || n_records == 0 || !(distance <= 0.5)) {
        const int rc = tidk_cuda_explore_runs(ctx->subs[0], seq, offsets, n_records, distance, kmin, kmax, threshold, out_runs, n_runs);
        if (rc) ctx->err = ctx->subs[0]->err;
        return rc;
    }
    *out_runs = nullptr; *n_runs = 0;
    const int world = (int)ctx->subs.size();
    const std::vector<uint32_t> cut = cut_records(n_records, world, [&](uint32_t r) {
        const double v = std::ceil((double)(offsets[r + 1] - offsets[r]) * distance);
        return 2 * (v > 0 ? (uint64_t)v : 0);
    });
    std::vector<PartRuns> part((size_t)world);
    const int rc = on_all(ctx, [&](int g, tidk_ctx *sub) -> int {
        const uint32_t r0 = cut[(size_t)g], r1 = cut[(size_t)g + 1];
        if (r1 <= r0) return TIDK_OK;
        std::vector<uint64_t> loc_off((size_t)(r1 - r0) + 1);
        for (uint32_t i = r0; i <= r1; i++) loc_off[i - r0] = offsets[i] - offsets[r0];
        tidk_run *runs = nullptr;
        uint64_t n = 0;
        const int rc2 = tidk_cuda_explore_runs(sub, seq + offsets[r0], loc_off.data(), r1 - r0, distance, kmin, kmax, threshold, &runs, &n);
        if (rc2) return rc2;
        part[(size_t)g].runs = runs; part[(size_t)g].n = n; part[(size_t)g].r0 = r0;
        return TIDK_OK;
    });
    if (rc) {
        for (PartRuns &p : part) if (p.runs) tidk_cuda_free_runs(p.runs);
        return rc;
    }
    return merge_runs(ctx, part, out_runs, n_runs);
}

inline int seqbuf_upload(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets, uint32_t n_records, tidk_seqbuf **out) {
    if (!out) return fail(ctx, TIDK_EARG, "out is NULL");
    *out = nullptr;
    if (!offsets) return fail(ctx, TIDK_EARG, "offsets is NULL");
    const int world = (int)ctx->subs.size();
    tidk_seqbuf *b = new (std::nothrow) tidk_seqbuf();
    if (!b) return fail(ctx, TIDK_ENOMEM, "out of host memory");
    b->n_records = n_records;
    b->total = offsets[n_records];
    b->h_off.assign(offsets, offsets + n_records + 1);
    b->cut = cut_records(n_records, world, [&](uint32_t r) { return offsets[r + 1] - offsets[r]; });
    b->parts.assign((size_t)world, nullptr);
    const int rc = on_all(ctx, [&](int g, tidk_ctx *sub) -> int {
        const uint32_t r0 = b->cut[(size_t)g], r1 = b->cut[(size_t)g + 1];
        std::vector<uint64_t> loc_off((size_t)(r1 - r0) + 1);
        for (uint32_t i = r0; i <= r1; i++) loc_off[i - r0] = offsets[i] - offsets[r0];
        return tidk_cuda_seqbuf_upload(sub, seq ? seq + offsets[r0] : nullptr, loc_off.data(), r1 - r0, &b->parts[(size_t)g]);
    });
    if (rc) {
        for (size_t g = 0; g < b->parts.size(); g++)
            if (b->parts[g]) tidk_cuda_seqbuf_free(ctx->subs[g], b->parts[g]);
        delete b;
        return rc;
    }
    *out = b;
    return TIDK_OK;
}

inline void seqbuf_free(tidk_ctx *ctx, tidk_seqbuf *b) {
    for (size_t g = 0; g < b->parts.size(); g++)
        if (b->parts[g]) tidk_cuda_seqbuf_free(g < ctx->subs.size() ? ctx->subs[g] : nullptr, b->parts[g]);
    delete b;
}

// first element of record r in the [record][query][window] arrays
// first output element of every record (cached per batch for the last (window, query count): a loop of passes asks again and again)
inline const std::vector<uint64_t> &out_base(const tidk_seqbuf *b, uint64_t window, uint32_t nq) {
    if (b->ob_window != window || b->ob_nq != nq || b->ob_cache.size() != (size_t)b->n_records + 1) {
        std::vector<uint64_t> &ob = b->ob_cache;
        ob.assign((size_t)b->n_records + 1, 0);
        for (uint32_t r = 0; r < b->n_records; r++) ob[r + 1] = ob[r] + n_windows(b->h_off[r + 1] - b->h_off[r], window) * nq;
        b->ob_window = window; b->ob_nq = nq;
    }
    return b->ob_cache;
}

inline bool is_multi_buf(tidk_ctx *ctx, const tidk_seqbuf *b) { return b && b->parts.size() == ctx->subs.size(); }

inline int window_counts_resident(tidk_ctx *ctx, const tidk_seqbuf *b, uint64_t window, const char *const *queries, const uint32_t *qlens,
                                  uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev, float *kernel_ms, uint32_t *kernel_launches, bool async) {
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    if (!is_multi_buf(ctx, b)) return fail(ctx, TIDK_EARG, "the batch was not uploaded through this multi-device context");
    if (window == 0 || nq == 0 || !out_fwd || !out_rev) return fail(ctx, TIDK_EARG, "window must be > 0, at least one query, outputs not NULL");
    const std::vector<uint64_t> &ob = out_base(b, window, nq);
    std::vector<float> ms(ctx->subs.size(), 0.f);
    std::vector<uint32_t> nl(ctx->subs.size(), 0);
    auto one = [&](int g, tidk_ctx *sub) -> int {
        const uint32_t r0 = b->cut[(size_t)g];
        if (b->cut[(size_t)g + 1] <= r0) return TIDK_OK;
        if (async) return tidk_cuda_window_counts_resident_async(sub, b->parts[(size_t)g], window, queries, qlens, nq, out_fwd + ob[r0], out_rev + ob[r0]);
        return tidk_cuda_window_counts_resident(sub, b->parts[(size_t)g], window, queries, qlens, nq, out_fwd + ob[r0], out_rev + ob[r0], &ms[(size_t)g], &nl[(size_t)g]);
    };
    int rc = TIDK_OK;
    // (asynchronous calls only enqueue, but an enqueue is 10-20 us of driver calls per device: on the parked threads too --
    // issued one after the other from the caller's thread, eight devices cost 165 us per pass against a 49 us kernel)
    rc = on_all(ctx, one);
    if (async) return rc;
    float mx = 0.f;
    uint32_t sum = 0;
    for (size_t g = 0; g < ms.size(); g++) { mx = std::max(mx, ms[g]); sum += nl[g]; }
    ctx->last_kernel_ms = mx;
    ctx->last_kernel_launches = sum;
    if (kernel_ms) *kernel_ms = mx;           // the devices run side by side: the slowest one
    if (kernel_launches) *kernel_launches = sum;
    return rc;
}

inline int wait(tidk_ctx *ctx, float *kernel_ms, uint32_t *kernel_launches) {
    float mx = 0.f;
    uint32_t sum = 0;
    int rc = TIDK_OK;
    for (tidk_ctx *sub : ctx->subs) {
        float ms = 0.f;
        uint32_t nl = 0;
        const int r = tidk_cuda_wait(sub, &ms, &nl);
        if (r && !rc) { rc = r; ctx->err = sub->err; }
        mx = std::max(mx, ms);
        sum += nl;
    }
    ctx->last_kernel_ms = mx;
    ctx->last_kernel_launches = sum;
    if (kernel_ms) *kernel_ms = mx;
    if (kernel_launches) *kernel_launches = sum;
    return rc;
}

inline int explore_runs_resident(tidk_ctx *ctx, const tidk_seqbuf *b, double distance, uint32_t kmin, uint32_t kmax, uint64_t threshold,
                                 tidk_run **out_runs, uint64_t *n_runs, float *kernel_ms, uint32_t *kernel_launches) {
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    if (!out_runs || !n_runs) return fail(ctx, TIDK_EARG, "NULL argument");
    *out_runs = nullptr; *n_runs = 0;
    if (!is_multi_buf(ctx, b)) return fail(ctx, TIDK_EARG, "the batch was not uploaded through this multi-device context");
    const size_t world = ctx->subs.size();
    std::vector<PartRuns> part(world);
    std::vector<float> ms(world, 0.f);
    std::vector<uint32_t> nl(world, 0);
    const int rc = on_all(ctx, [&](int g, tidk_ctx *sub) -> int {
        const uint32_t r0 = b->cut[(size_t)g];
        if (b->cut[(size_t)g + 1] <= r0) return TIDK_OK;
        tidk_run *runs = nullptr;
        uint64_t n = 0;
        const int rc2 = tidk_cuda_explore_runs_resident(sub, b->parts[(size_t)g], distance, kmin, kmax, threshold, &runs, &n, &ms[(size_t)g], &nl[(size_t)g]);
        if (rc2) return rc2;
        part[(size_t)g].runs = runs; part[(size_t)g].n = n; part[(size_t)g].r0 = r0;
        return TIDK_OK;
    });
    if (rc) {
        for (PartRuns &p : part) if (p.runs) tidk_cuda_free_runs(p.runs);
        return rc;
    }
    float mx = 0.f, scan = 0.f, post = 0.f;
    uint32_t sum = 0;
    for (size_t g = 0; g < world; g++) {
        mx = std::max(mx, ms[g]); sum += nl[g];
        scan = std::max(scan, ctx->subs[g]->explore_scan_ms);
        post = std::max(post, ctx->subs[g]->explore_post_ms);
    }
    ctx->explore_scan_ms = scan; ctx->explore_post_ms = post;
    if (kernel_ms) *kernel_ms = mx;
    if (kernel_launches) *kernel_launches = sum;
    return merge_runs(ctx, part, out_runs, n_runs);
}

} // namespace multi
