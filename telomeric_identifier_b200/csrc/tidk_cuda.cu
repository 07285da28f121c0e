// tidk_cuda.cu -- C ABI of libtidk_cuda.so (see include/tidk_cuda.h).
// Host-side glue only: argument checks, device buffers, launches, copies.
// There is no CPU implementation of any entry point in this library.
#include "../../include/tidk_cuda.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "explore_host.cuh"
#include "host_pack.hpp"
#include "planes_build.cuh"
#include "window_counts_flat.cuh"
#include "window_counts.cuh"

using namespace tidk;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

} // namespace

struct tidk_seqbuf {
    uint8_t *d_seq = nullptr;
    uint64_t *d_off = nullptr;
    size_t seq_cap = 0, off_cap = 0;
    std::vector<uint64_t> h_off;
    uint32_t n_records = 0;
    uint64_t total = 0;
    uint64_t id = 0; // changes whenever the contents change (invalidates ctx-side caches)
    // resident bit-planes (planes_build.cuh), built once behind the upload of a batch that stays resident
    uint32_t *d_planes = nullptr, *d_plainbits = nullptr, *d_validbits = nullptr;
    size_t planes_cap = 0, bits_cap = 0;
    bool planes_ready = false;
    bool planes_nonascii = false; // the conversion saw a byte >= 0x80
    bool planes_all_plain = false; // every word that holds a base is 32 (or, the last one, fewer) upper-case A C G T
    float planes_ms = 0.f;        // device time of the conversion
    // a batch of a multi-device context (multi_device.cuh): one part per device, records [cut[g], cut[g + 1]) in part g
    std::vector<tidk_seqbuf *> parts;
    std::vector<uint32_t> cut;
    mutable std::vector<uint64_t> ob_cache; // first output element of every record for (ob_window, ob_nq): multi_device.cuh out_base
    mutable uint64_t ob_window = 0;
    mutable uint32_t ob_nq = 0;
};

// Packer threads of the host-packed upload, kept parked between calls: creating and joining 15 threads per call
// cost a few hundred microseconds of a 5 ms call.  run(n, fn) executes fn(1) .. fn(n - 1) on the pool and fn(0)
// on the caller, and returns when all are done.
class PackPool {
  public:
    ~PackPool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : th_) t.join();
    }
    void run(int n, const std::function<void(int)> &fn) {
        if (n <= 1) { fn(0); return; }
        {
            std::lock_guard<std::mutex> g(mu_);
            while ((int)th_.size() < n - 1) {
                const int id = (int)th_.size() + 1;
                th_.emplace_back([this, id] { loop(id); });
            }
            fn_ = &fn; active_ = n; pending_ = n - 1; gen_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *fn = nullptr;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                if (id >= active_) continue; // this round uses fewer threads
                fn = fn_;
            }
            (*fn)(id);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> th_;
    const std::function<void(int)> *fn_ = nullptr;
    uint64_t gen_ = 0;
    int active_ = 0, pending_ = 0;
    bool stop_ = false;
};

namespace multi { struct Workers; }

struct tidk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    tidk_seqbuf scratch; // staging for the host-pointer entry points
    DevBuf win_base, out_fwd, out_rev, misc, qbytes, runs, items, exc, chunk_state;
    std::vector<uint32_t> h_chunk_state; // explore upload: per 4 MiB chunk, 0 = went up as ASCII, else 1 + number of raw-word records
    float explore_scan_ms = 0.f, explore_post_ms = 0.f, explore_host_ms[2] = {0.f, 0.f};
    uint64_t last_h2d = 0, last_d2h = 0; // bytes moved by the last host-pointer window-count call
    int last_packed = 0;
    float last_kernel_ms = 0.f;     // device time / launches of the counting kernels of the last window-count call
    uint32_t last_kernel_launches = 0;
    // host-packed upload path (tidk_cuda_window_counts): device planes + per-thread staging
    DevBuf packed;
    struct PackLane { cudaStream_t st = nullptr; cudaEvent_t ev[2] = {nullptr, nullptr}; uint32_t *pin = nullptr; };
    std::vector<PackLane> pack_lanes;
    PackPool pack_pool;
    cudaStream_t dma_st = nullptr;  // hybrid upload: the ASCII share of the sequence
    static constexpr int kDmaDepth = 4;
    cudaEvent_t dma_ev[kDmaDepth] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t counter_flip = 0;      // which of the two queue counters the next launch uses
    bool counters_dirty = false;    // a call failed between taking a counter and finishing: re-zero both
    uint32_t *h_flags = nullptr;    // pinned + mapped: the window kernels store the (rare) error flag straight into host memory
    uint32_t *d_flags = nullptr;    // device view of h_flags
    ExploreScratch ex;
    // window counts, asynchronous form: three device output buffers used in turn; the counts of call i leave on
    // the copy stream while the kernel of call i + 1 runs -- and zeroes the buffer of call i + 2 on the side (flat kernels)
    struct WinSlot { cudaEvent_t k0 = nullptr, k1 = nullptr, done = nullptr; bool busy = false; uint32_t launches = 0; };
    static constexpr int kWinSlots = 3;
    WinSlot wslot[kWinSlots];
    DevBuf out_dev[kWinSlots];
    size_t out_zero_bytes[kWinSlots] = {0, 0, 0}; // leading bytes of out_dev[s] known to be zero (written by the previous flat launch)
    cudaStream_t copy_st = nullptr;
    uint32_t wslot_next = 0;
    float pending_kernel_ms = 0.f;     // device time / launches of the calls retired since the last tidk_cuda_wait
    uint32_t pending_launches = 0;
    int pending_err = 0;               // first deferred failure (reported by tidk_cuda_wait)
    std::unordered_map<const void *, int> occupancy; // resident CTAs per SM, per kernel instantiation, on THIS device
    uint64_t next_buf_id = 1;
    // multi-device context: one single-device context per device, this one holds no device state of its own
    std::vector<tidk_ctx *> subs;
    multi::Workers *workers = nullptr; // parked host threads, one per device but the first (multi_device.cuh)
    // work items cached for (sequence batch, window, query count): several clades or output
    // formats over one uploaded genome reuse them
    uint64_t items_buf_id = 0, items_window = 0, items_n = 0, items_nwin = 0;
    bool items_built = false;     // the item table of (items_buf_id, items_window, items_nq) exists (calls served by the flat kernels never build it)
    DevBuf spans;                 // span descriptors of the flat window counter (window_counts_flat.cuh), cached like the items
    uint64_t spans_buf_id = 0, spans_window = 0, spans_nq = 0;
    uint32_t items_nq = 0;
};

namespace {

int fail(tidk_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define CU(ctx, call)                                                                    \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? TIDK_ENOMEM : TIDK_ECUDA, \
                        "%s failed: %s", #call, cudaGetErrorString(e_));                 \
    } while (0)

int ensure(tidk_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return TIDK_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, TIDK_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return TIDK_OK;
}

constexpr size_t kSeqPad = 8192; // zeroed bytes behind the data: block loads never leave the buffer
// 1 KiB blocks that K1 converts: the data and as many whole blocks of the zero padding as lie INSIDE it (the last block
// starts below total + 1024 - 1 + 6 KiB and ends inside the 8 KiB that were zeroed; one block more would read up to
// 1 023 bytes of whatever the allocation held before -- a stale byte >= 0x80 there made the batch "non-ASCII")
inline uint64_t planes_blocks(uint64_t total) { return (total + 1023) / 1024 + kSeqPad / 1024 - 1; }

static int pack_threads() {
    if (const char *e = getenv("TIDK_PACK_THREADS")) return atoi(e);
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    unsigned share = 1; // ranks of one torchrun job share the host cores
    if (const char *lws = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)atoi(lws) > 0 ? (unsigned)atoi(lws) : 1;
    unsigned t = hw / share;
    return (int)(t > 32 ? 32 : t);
}


// one stream, two events and a two-slot pinned staging ring (2 x 1.5 MiB: the lo / hi planes of a 4 MiB chunk) per
// host thread of the upload paths
static int ensure_lanes(tidk_ctx *ctx, int T) {
    const size_t ring_bytes = 2 * (3ull * kPackWords * 4);
    if ((int)ctx->pack_lanes.size() < T) {
        const size_t old = ctx->pack_lanes.size();
        ctx->pack_lanes.resize(T);
        for (size_t i = old; i < (size_t)T; i++) {
            tidk_ctx::PackLane &L = ctx->pack_lanes[i];
            CU(ctx, cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
            CU(ctx, cudaEventCreateWithFlags(&L.ev[0], cudaEventDisableTiming));
            CU(ctx, cudaEventCreateWithFlags(&L.ev[1], cudaEventDisableTiming));
            CU(ctx, cudaHostAlloc((void **)&L.pin, ring_bytes, cudaHostAllocDefault));
        }
    }
    return TIDK_OK;
}

// Host -> device copy of a sequence.  A pinned source goes by DMA straight from the caller's buffer.  From pageable
// memory (a parsed FASTA in a malloc'ed buffer, a plain numpy array) one cudaMemcpy is staged by the driver on a single
// thread at about 10 GB/s; here the host threads of the upload paths stage it themselves -- each copies 512 KiB pieces
// into its own pinned ring and sends them on its own stream -- which runs at the rate of the bus.
// Returns with the copy complete (pageable) or enqueued on ctx->stream (pinned / small).
static int upload_bytes(tidk_ctx *ctx, uint8_t *dst, const uint8_t *src, uint64_t n) {
    bool pageable = false;
    static const bool allow = !(getenv("TIDK_STAGED_UPLOAD") && atoi(getenv("TIDK_STAGED_UPLOAD")) == 0);
    const int T = pack_threads();
    if (allow && T >= 4 && n >= (64ull << 20)) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, src) == cudaSuccess) pageable = a.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (!pageable) {
        CU(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
        return TIDK_OK;
    }
    int rc;
    if ((rc = ensure_lanes(ctx, T))) return rc;
    constexpr uint64_t kPiece = 512u << 10;
    const uint64_t n_pieces = (n + kPiece - 1) / kPiece;
    std::atomic<uint64_t> next{0};
    std::atomic<int> cuda_err{0};
    auto worker = [&](int t) {
        cudaSetDevice(ctx->device);
        tidk_ctx::PackLane &L = ctx->pack_lanes[t];
        for (int slot = 0;; slot ^= 1) {
            const uint64_t c = next.fetch_add(1);
            if (c >= n_pieces) break;
            if (cudaEventSynchronize(L.ev[slot]) != cudaSuccess) { cuda_err = 1; break; } // the slot's previous copy is done
            uint8_t *stage = (uint8_t *)L.pin + (size_t)slot * kPiece;
            const uint64_t b0 = c * kPiece, len = n - b0 < kPiece ? n - b0 : kPiece;
            memcpy(stage, src + b0, len);
            if (cudaMemcpyAsync(dst + b0, stage, len, cudaMemcpyHostToDevice, L.st) != cudaSuccess ||
                cudaEventRecord(L.ev[slot], L.st) != cudaSuccess) { cuda_err = 1; break; }
        }
        if (cudaStreamSynchronize(L.st) != cudaSuccess) cuda_err = 1;
    };
    ctx->pack_pool.run(T, worker);
    if (cuda_err) return fail(ctx, TIDK_ECUDA, "staged upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return TIDK_OK;
}

// seq_mode 1: copy the sequence; 0: offsets only; 2: allocate the sequence buffer, copy nothing (the hybrid
// upload fills the part of it that travels as ASCII)
int seqbuf_fill(tidk_ctx *ctx, tidk_seqbuf *b, const uint8_t *seq, const uint64_t *offsets,
                uint32_t n_records, int seq_mode = 1) {
    const bool copy_seq = seq_mode == 1;
    if (n_records && (!offsets)) return fail(ctx, TIDK_EARG, "offsets is NULL");
    const uint64_t total = n_records ? offsets[n_records] : 0;
    if (n_records && offsets[0] != 0) return fail(ctx, TIDK_EARG, "offsets[0] must be 0");
    for (uint32_t r = 0; r < n_records; r++)
        if (offsets[r + 1] < offsets[r]) return fail(ctx, TIDK_EARG, "offsets must be non-decreasing");
    if (total && !seq) return fail(ctx, TIDK_EARG, "seq is NULL");
    const size_t need = seq_mode ? (size_t)total + kSeqPad : 0;
    if (need > b->seq_cap) {
        if (b->d_seq) cudaFree(b->d_seq);
        b->d_seq = nullptr; b->seq_cap = 0;
        cudaError_t e = cudaMalloc((void **)&b->d_seq, need);
        if (e != cudaSuccess) { b->d_seq = nullptr; return fail(ctx, TIDK_ENOMEM, "cudaMalloc(seq %zu) failed: %s", need, cudaGetErrorString(e)); }
        b->seq_cap = need;
    }
    const size_t off_need = ((size_t)n_records + 1) * sizeof(uint64_t);
    if (off_need > b->off_cap) {
        if (b->d_off) cudaFree(b->d_off);
        b->d_off = nullptr; b->off_cap = 0;
        cudaError_t e = cudaMalloc((void **)&b->d_off, off_need);
        if (e != cudaSuccess) { b->d_off = nullptr; return fail(ctx, TIDK_ENOMEM, "cudaMalloc(off) failed: %s", cudaGetErrorString(e)); }
        b->off_cap = off_need;
    }
    b->h_off.assign(1, 0);
    if (n_records) b->h_off.assign(offsets, offsets + n_records + 1);
    b->n_records = n_records;
    b->total = total;
    b->id = ctx->next_buf_id++;
    int rc_up = 0;
    if (copy_seq) {
        if (total && (rc_up = upload_bytes(ctx, b->d_seq, seq, total))) return rc_up;
        CU(ctx, cudaMemsetAsync(b->d_seq + total, 0, kSeqPad, ctx->stream));
    }
    CU(ctx, cudaMemcpyAsync(b->d_off, b->h_off.data(), off_need, cudaMemcpyHostToDevice, ctx->stream));
    return TIDK_OK;
}

void seqbuf_release(tidk_seqbuf *b) {
    if (b->d_seq) cudaFree(b->d_seq);
    if (b->d_off) cudaFree(b->d_off);
    for (uint32_t *p : {b->d_planes, b->d_plainbits, b->d_validbits})
        if (p) cudaFree(p);
    b->d_seq = nullptr; b->d_off = nullptr; b->seq_cap = b->off_cap = 0;
    b->d_planes = b->d_plainbits = b->d_validbits = nullptr;
    b->planes_cap = b->bits_cap = 0;
    b->planes_ready = false;
}

// K1 (planes_build.cuh) over the whole batch, in stream order behind its upload.  Blocks until done (the caller
// of an upload waits for the copy anyway) and records the device time.
int seqbuf_build_planes(tidk_ctx *ctx, tidk_seqbuf *b) {
    b->planes_ready = false;
    const uint64_t n_blocks = planes_blocks(b->total); // the data and the zero padding behind it
    const uint64_t n_chunks = (n_blocks * 32 + kPackWords - 1) / kPackWords;
    const size_t plane_bytes = (size_t)n_chunks * 3 * kPackWords * 4, bits_bytes = ((size_t)n_blocks + 2) * 4;
    auto grow = [&](uint32_t *&p, size_t &cap, size_t need) -> int {
        if (p && need <= cap) return TIDK_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const cudaError_t e = cudaMalloc((void **)&p, need);
        if (e != cudaSuccess) { p = nullptr; return fail(ctx, TIDK_ENOMEM, "cudaMalloc(planes %zu) failed: %s", need, cudaGetErrorString(e)); }
        cap = need;
        return TIDK_OK;
    };
    int rc;
    size_t cap2 = b->bits_cap;
    if ((rc = grow(b->d_planes, b->planes_cap, plane_bytes))) return rc;
    if ((rc = grow(b->d_plainbits, b->bits_cap, bits_bytes))) return rc;
    if ((rc = grow(b->d_validbits, cap2, bits_bytes))) return rc;
    if (!ctx->h_flags) {
        CU(ctx, cudaHostAlloc((void **)&ctx->h_flags, 64, cudaHostAllocMapped));
        memset(ctx->h_flags, 0, 64);
        CU(ctx, cudaHostGetDevicePointer((void **)&ctx->d_flags, ctx->h_flags, 0));
    }
    // (the plain / valid words behind the last converted block are read by the scans' look-ahead: zero)
    CU(ctx, cudaMemsetAsync(b->d_plainbits + n_blocks, 0, 8, ctx->stream));
    CU(ctx, cudaMemsetAsync(b->d_validbits + n_blocks, 0, 8, ctx->stream));
    PlaneBuildJob job{b->d_seq, b->d_planes, b->d_plainbits, b->d_validbits, n_blocks, b->total, ctx->d_flags + 1};
    const uint64_t want = (n_blocks + 7) / 8, cap = (uint64_t)ctx->sm_count * 8;
    {   // the module is loaded lazily at a kernel's first launch (about 20 ms for this library): not between the two events
        cudaFuncAttributes fa;
        (void)cudaFuncGetAttributes(&fa, planes_build_kernel);
    }
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    planes_build_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(job);
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaEventElapsedTime(&b->planes_ms, ctx->ev0, ctx->ev1));
    b->planes_nonascii = ((volatile uint32_t *)ctx->h_flags)[1] != 0;
    b->planes_all_plain = ((volatile uint32_t *)ctx->h_flags)[2] == 0;
    ctx->h_flags[1] = ctx->h_flags[2] = 0;
    b->planes_ready = true;
    return TIDK_OK;
}

inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// utils.rs:49-58 switch_base (host side of the boundary: tiny, per query)
inline uint8_t comp_base(uint8_t c) {
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'T': return 'A';
    case 'G': return 'C';
    default: return 'N';
    }
}

StrandPattern encode_strand(const uint8_t *pat, uint32_t L) {
    StrandPattern sp{0u, 0u};
    for (uint32_t s = 0; s < L; s++) {
        uint8_t c = pat[L - 1 - s];
        uint32_t lo = (c == 'C' || c == 'G'), hi = (c == 'G' || c == 'T');
        sp.plo |= lo << s;
        sp.phi |= hi << s;
    }
    return sp;
}

// ---- kernels specialised at compile time for known motifs ---------------------
// Every repeat of the curated clade table (clades/curated.csv, 23 motifs) plus the
// forms people type at `tidk search` (TTAGG, TTAGGG, ...).  A query whose reverse
// complement is listed uses that kernel with the two outputs swapped.  Anything
// else runs the run-time-pattern kernel.
#define TIDK_STATIC_MOTIFS(X)                                                                   \
    X("AAGGAC") X("AACCCT") X("AAACCCT") X("AACCC") X("AACCT") X("AACAGACCCG") X("ACCTG")       \
    X("AAACCACCCT") X("AACCATCCCT") X("AACCCAGACCT") X("AACCCCAACCT") X("AACCCGAACCT")          \
    X("ACCCAG") X("ACGGCAGCG") X("AAATGTGGAGG") X("AAAGAACCT") X("AAAATTGTCCGTCC") X("AAACCC")  \
    X("AACCCAGACCC") X("AACCCAGACGC") X("AACCCTGACGC") X("AAACCCC") X("AACCCTG")                \
    X("TTAGG") X("TTAGGG") X("TTTAGGG") X("CCCTAA") X("CCTAA") X("TTGGGG") X("TTAGGC")

constexpr int cx_len(const char *s) { int n = 0; while (s[n]) n++; return n; }
constexpr uint64_t cx_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
constexpr uint64_t cx_enc(const char *s) {
    uint64_t f = 0;
    for (int j = 0; s[j]; j++) f |= cx_code(s[j]) << (2 * j);
    return f;
}

// persistent launch: the grid is the number of CTAs that are resident at once (the kernels assign
// the first items statically and rely on every CTA running from the start).  Occupancy is looked up once
// per kernel instantiation and CONTEXT (= device): contexts of different devices run on different threads.
struct LaunchEnv {
    int sm_count;
    std::unordered_map<const void *, int> *occupancy;
    cudaStream_t st;
};

template <class K>
int resident_grid(K kernel, const LaunchEnv &env, uint64_t n_items) {
    int &per_sm = (*env.occupancy)[(const void *)kernel];
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * 32, 0) != cudaSuccess || per_sm < 1)
            per_sm = 1;
    }
    const uint64_t want = (n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const uint64_t cap = (uint64_t)env.sm_count * per_sm;
    const uint64_t g = want < cap ? want : cap;
    return g < 1 ? 1 : (int)g;
}

using LaunchFn = void (*)(const WindowJob &, const LaunchEnv &);
template <int L, uint64_t F, class Src>
void launch_static(const WindowJob &job, const LaunchEnv &env) {
    using M = StaticMatcher<L, F>;
    typename M::Params mp{0};
    const int grid = resident_grid(window_counts_kernel<M, 4, Src>, env, job.n_items);
    window_counts_kernel<M, 4, Src><<<grid, kWarpsPerBlock * 32, 0, env.st>>>(job, mp);
}
// flat form on resident planes (window_counts_flat.cuh)
#ifndef TIDK_FLAT_MINB
#define TIDK_FLAT_MINB 3
#endif
using FlatLaunchFn = void (*)(const FlatJob &, const LaunchEnv &);
template <int L, uint64_t F>
void launch_static_flat(const FlatJob &job, const LaunchEnv &env) {
    using M = StaticMatcher<L, F>;
    const int grid = resident_grid(window_counts_flat_kernel<M, TIDK_FLAT_MINB>, env, job.n_spans);
    window_counts_flat_kernel<M, TIDK_FLAT_MINB><<<grid, kWarpsPerBlock * 32, 0, env.st>>>(job, 0u);
}
template <int L>
void launch_runtime_flat(const FlatJob &job, const typename FlatRuntime<L>::Params &rp, const LaunchEnv &env) {
    using M = FlatRuntime<L>;
    const int grid = resident_grid(window_counts_flat_kernel<M, TIDK_FLAT_MINB>, env, job.n_spans);
    window_counts_flat_kernel<M, TIDK_FLAT_MINB><<<grid, kWarpsPerBlock * 32, 0, env.st>>>(job, rp);
}
void launch_flat_runtime(int L, const FlatJob &job, const FlatRuntime<1>::Params &rp0, const LaunchEnv &env) {
    switch (L) {
#define C_(n) case n: { FlatRuntime<n>::Params rp; memcpy(&rp, &rp0, sizeof(rp)); launch_runtime_flat<n>(job, rp, env); break; }
        C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14)
        C_(15) C_(16) C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27)
        C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
    default: break;
    }
}

struct StaticEntry { int L; uint64_t F; LaunchFn fn, fn_packed; FlatLaunchFn fn_flat; };
#define X_(s) {cx_len(s), cx_enc(s), &launch_static<cx_len(s), cx_enc(s), AsciiSource>, &launch_static<cx_len(s), cx_enc(s), PackedSource>, &launch_static_flat<cx_len(s), cx_enc(s)>},
const StaticEntry kStatic[] = {TIDK_STATIC_MOTIFS(X_)};
#undef X_

LaunchFn find_static(const uint8_t *pat, uint32_t L, bool packed = false) {
    if (L > 32) return nullptr;
    uint64_t f = 0;
    for (uint32_t j = 0; j < L; j++) f |= cx_code((char)pat[j]) << (2 * j);
    for (const StaticEntry &e : kStatic)
        if (e.L == (int)L && e.F == f) return packed ? e.fn_packed : e.fn;
    return nullptr;
}

// ---- whole clades in one pass ----------------------------------------------------
// The multi-repeat clades of the curated table (table order = output order).  `find --clade X`
// with exactly this query list runs ONE fused kernel instead of one pass per repeat.
#define MOTIF_(s) Motif<cx_len(s), cx_enc(s)>
using Coleoptera = StaticMultiMatcher<MOTIF_("AACCC"), MOTIF_("AACCT"), MOTIF_("AACAGACCCG"), MOTIF_("ACCTG")>;
using Hemiptera = StaticMultiMatcher<MOTIF_("AAACCACCCT"), MOTIF_("AACCATCCCT")>;
using Hymenoptera = StaticMultiMatcher<MOTIF_("AACCCAGACCT"), MOTIF_("AACCCCAACCT"), MOTIF_("AACCCGAACCT"), MOTIF_("AACCT"),
                                       MOTIF_("ACCCAG"), MOTIF_("AACCCT"), MOTIF_("ACGGCAGCG"), MOTIF_("AAATGTGGAGG"),
                                       MOTIF_("AAAGAACCT"), MOTIF_("AAAATTGTCCGTCC"), MOTIF_("AAACCC"), MOTIF_("AACCC"),
                                       MOTIF_("AACCCAGACCC"), MOTIF_("AACCCAGACGC"), MOTIF_("AACCCTGACGC")>;
#undef MOTIF_

template <class M, int MINB, class Src>
void launch_multi(const WindowJob &job, const LaunchEnv &env) {
    typename M::Params mp{0};
    const int grid = resident_grid(window_counts_kernel<M, MINB, Src>, env, job.n_items);
    window_counts_kernel<M, MINB, Src><<<grid, kWarpsPerBlock * 32, 0, env.st>>>(job, mp);
}

struct CladeSet { std::vector<const char *> motifs; LaunchFn fn, fn_packed; };
const CladeSet kCladeSets[] = {
    {{"AACCC", "AACCT", "AACAGACCCG", "ACCTG"}, &launch_multi<Coleoptera, 3, AsciiSource>, &launch_multi<Coleoptera, 3, PackedSource>},
    {{"AAACCACCCT", "AACCATCCCT"}, &launch_multi<Hemiptera, 4, AsciiSource>, &launch_multi<Hemiptera, 4, PackedSource>},
    {{"AACCCAGACCT", "AACCCCAACCT", "AACCCGAACCT", "AACCT", "ACCCAG", "AACCCT", "ACGGCAGCG", "AAATGTGGAGG", "AAAGAACCT",
      "AAAATTGTCCGTCC", "AAACCC", "AACCC", "AACCCAGACCC", "AACCCAGACGC", "AACCCTGACGC"},
     &launch_multi<Hymenoptera, 2, AsciiSource>, &launch_multi<Hymenoptera, 2, PackedSource>},
};

FlatLaunchFn find_static_flat(const uint8_t *pat, uint32_t L) {
    if (L > 32) return nullptr;
    uint64_t f = 0;
    for (uint32_t j = 0; j < L; j++) f |= cx_code((char)pat[j]) << (2 * j);
    for (const StaticEntry &e : kStatic)
        if (e.L == (int)L && e.F == f) return e.fn_flat;
    return nullptr;
}

LaunchFn find_clade_set(const char *const *queries, const uint32_t *qlens, uint32_t nq, bool packed = false) {
    for (const CladeSet &c : kCladeSets) {
        if (c.motifs.size() != nq) continue;
        bool same = true;
        for (uint32_t q = 0; q < nq && same; q++)
            same = strlen(c.motifs[q]) == qlens[q] && memcmp(c.motifs[q], queries[q], qlens[q]) == 0;
        if (same) return packed ? c.fn_packed : c.fn;
    }
    return nullptr;
}

template <int L>
void launch_runtime(const WindowJob &job, const RuntimeQueries &rq, const LaunchEnv &env) {
    const int grid = resident_grid(window_counts_kernel<RuntimeMatcher<L>, 3>, env, job.n_items);
    window_counts_kernel<RuntimeMatcher<L>, 3><<<grid, kWarpsPerBlock * 32, 0, env.st>>>(job, rq);
}

void launch_planes(int L, const WindowJob &job, const RuntimeQueries &rq, const LaunchEnv &env) {
    switch (L) {
#define C_(n) case n: launch_runtime<n>(job, rq, env); break;
        C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14)
        C_(15) C_(16) C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27)
        C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
    default: launch_runtime<0>(job, rq, env); break;
    }
}

void revcomp_host(const uint8_t *f, uint32_t L, uint8_t *out) {
    for (uint32_t j = 0; j < L; j++) out[j] = comp_base(f[L - 1 - j]);
}

} // namespace

#include "multi_device.cuh"

extern "C" {

uint32_t tidk_cuda_abi_version(void) { return 1; }

int tidk_cuda_create(const int *device_ids, int n_devices, tidk_ctx **out) {
    if (!out) return fail(nullptr, TIDK_EARG, "out is NULL");
    *out = nullptr;
    if (n_devices < 0 || n_devices > 64) return fail(nullptr, TIDK_EARG, "0 .. 64 devices per ctx (got %d)", n_devices);
    if (n_devices > 1) { // one single-device context per listed device (multi_device.cuh)
        if (!device_ids) return fail(nullptr, TIDK_EARG, "device_ids is NULL with %d devices", n_devices);
        tidk_ctx *top = new (std::nothrow) tidk_ctx();
        if (!top) return fail(nullptr, TIDK_ENOMEM, "out of host memory");
        top->device = device_ids[0];
        for (int g = 0; g < n_devices; g++) {
            tidk_ctx *sub = nullptr;
            const int rc = tidk_cuda_create(device_ids + g, 1, &sub);
            if (rc) {
                for (tidk_ctx *s2 : top->subs) tidk_cuda_destroy(s2);
                delete top;
                return rc;
            }
            top->subs.push_back(sub);
        }
        // the devices of one context share the host's cores (packer threads per single-device context)
        setenv("LOCAL_WORLD_SIZE", std::to_string(n_devices).c_str(), 0);
        *out = top;
        return TIDK_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TIDK_ECUDA, "no CUDA device: %s (libtidk_cuda has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    int dev = (device_ids && n_devices == 1) ? device_ids[0] : 0;
    if (dev < 0 || dev >= count) return fail(nullptr, TIDK_EARG, "device %d out of range (0..%d)", dev, count - 1);
    tidk_ctx *ctx = new (std::nothrow) tidk_ctx();
    if (!ctx) return fail(nullptr, TIDK_ENOMEM, "out of host memory");
    ctx->device = dev;
    if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        fail(nullptr, TIDK_ECUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete ctx;
        return TIDK_ECUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, dev);
    *out = ctx;
    return TIDK_OK;
}

void tidk_cuda_destroy(tidk_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->subs.empty()) {
        delete ctx->workers; // joins the parked threads
        for (tidk_ctx *sub : ctx->subs) tidk_cuda_destroy(sub);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    seqbuf_release(&ctx->scratch);
    for (DevBuf *b : {&ctx->win_base, &ctx->out_fwd, &ctx->out_rev, &ctx->misc, &ctx->qbytes, &ctx->runs, &ctx->items, &ctx->spans, &ctx->exc, &ctx->chunk_state})
        if (b->p) cudaFree(b->p);
    tidk::RunBlockPool::close(ctx->ex.run_pool); // lists the caller still holds stay valid (explore_post.cuh)
    explore_scratch_release(ctx->ex);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->packed.p) cudaFree(ctx->packed.p);
    for (tidk_ctx::PackLane &L : ctx->pack_lanes) {
        if (L.pin) cudaFreeHost(L.pin);
        if (L.ev[0]) cudaEventDestroy(L.ev[0]);
        if (L.ev[1]) cudaEventDestroy(L.ev[1]);
        if (L.st) cudaStreamDestroy(L.st);
    }
    for (int s = 0; s < tidk_ctx::kWinSlots; s++) {
        if (ctx->wslot[s].busy) cudaEventSynchronize(ctx->wslot[s].done);
        for (cudaEvent_t e : {ctx->wslot[s].k0, ctx->wslot[s].k1, ctx->wslot[s].done})
            if (e) cudaEventDestroy(e);
        if (ctx->out_dev[s].p) cudaFree(ctx->out_dev[s].p);
    }
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    for (cudaEvent_t e : ctx->dma_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->dma_st) cudaStreamDestroy(ctx->dma_st);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *tidk_cuda_last_error(const tidk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void *tidk_cuda_host_alloc(tidk_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { // (portable: every device of a multi-device context copies from it)
        fail(ctx, TIDK_ENOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void tidk_cuda_host_free(tidk_ctx *ctx, void *p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFreeHost(p);
}

int tidk_cuda_seqbuf_upload(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                            uint32_t n_records, tidk_seqbuf **out) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::seqbuf_upload(ctx, seq, offsets, n_records, out);
    if (!out) return fail(ctx, TIDK_EARG, "out is NULL");
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    tidk_seqbuf *b = new (std::nothrow) tidk_seqbuf();
    if (!b) return fail(ctx, TIDK_ENOMEM, "out of host memory");
    int rc = seqbuf_fill(ctx, b, seq, offsets, n_records);
    if (rc == TIDK_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, TIDK_ECUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    // a batch that stays resident is transposed once (K1): every later pass reads the planes
    const bool planes_on = !(getenv("TIDK_PLANES") && atoi(getenv("TIDK_PLANES")) == 0); // (read per call: measurements switch)
    if (rc == TIDK_OK && planes_on && b->total > 0) rc = seqbuf_build_planes(ctx, b);
    if (rc != TIDK_OK) { seqbuf_release(b); delete b; return rc; }
    *out = b;
    return TIDK_OK;
}

void tidk_cuda_seqbuf_free(tidk_ctx *ctx, tidk_seqbuf *buf) {
    if (!buf) return;
    if (!buf->parts.empty()) {
        if (ctx) multi::seqbuf_free(ctx, buf);
        return;
    }
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    seqbuf_release(buf);
    delete buf;
}

// packed != nullptr: the host-packed planes are the source for the sequence from byte `split` on; the bytes
// below `split` (a multiple of 32, 0 = none) are resident as ASCII in buf->d_seq (hybrid upload).
// The call that used window slot `s` has finished on the device: fold its kernel time into the pending totals.
static int retire_slot(tidk_ctx *ctx, int s) {
    tidk_ctx::WinSlot &S = ctx->wslot[s];
    if (!S.busy) return TIDK_OK;
    S.busy = false;
    cudaError_t e = cudaEventSynchronize(S.done);
    if (e != cudaSuccess) {
        if (!ctx->pending_err) ctx->pending_err = TIDK_ECUDA;
        return fail(ctx, TIDK_ECUDA, "asynchronous window count failed: %s", cudaGetErrorString(e));
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, S.k0, S.k1) == cudaSuccess) ctx->pending_kernel_ms += ms;
    else (void)cudaGetLastError();
    ctx->pending_launches += S.launches;
    return TIDK_OK;
}

static int drain_slots(tidk_ctx *ctx) {
    int rc = TIDK_OK;
    for (int s = 0; s < tidk_ctx::kWinSlots; s++) {
        const int r = retire_slot(ctx, s);
        if (r && !rc) rc = r;
    }
    return rc;
}

// async: the counts leave on the copy stream behind the kernels and the call returns at once (tidk_cuda_wait collects);
// otherwise the call returns with the counts in place.
static int window_counts_impl(tidk_ctx *ctx, const tidk_seqbuf *buf, const uint32_t *packed, uint64_t window,
                              const char *const *queries, const uint32_t *qlens,
                              uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev,
                              float *kernel_ms, uint32_t *kernel_launches, uint64_t split = 0, bool async = false) {
    if (!ctx) return TIDK_EARG;
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    if (!buf) return fail(ctx, TIDK_EARG, "buf is NULL");
    if (window == 0) return fail(ctx, TIDK_EARG, "window must be > 0 (slice::chunks(0) panics, search.rs:98)");
    if (nq && (!queries || !qlens)) return fail(ctx, TIDK_EARG, "queries/qlens is NULL");
    for (uint32_t q = 0; q < nq; q++) {
        if (qlens[q] == 0 || !queries[q]) return fail(ctx, TIDK_EARG, "query %u is empty (KMP::new panics, utils.rs:23)", q);
        for (uint32_t j = 0; j < qlens[q]; j++)
            if ((uint8_t)queries[q][j] >= 0x80) return fail(ctx, TIDK_ENONASCII, "query %u has a non-ASCII byte", q);
    }
    CU(ctx, cudaSetDevice(ctx->device));

    const uint32_t nrec = buf->n_records;
    const bool cached = ctx->items_buf_id == buf->id && ctx->items_window == window && ctx->items_nq == nq;
    uint64_t n_windows = ctx->items_nwin;
    std::vector<uint64_t> wb;
    if (!cached) {
        wb.assign((size_t)nrec + 1, 0);
        for (uint32_t r = 0; r < nrec; r++) {
            uint64_t len = buf->h_off[r + 1] - buf->h_off[r];
            uint64_t nw = len / window + (len % window != 0);
            if (nw > 0xFFFFFFFFull) return fail(ctx, TIDK_EARG, "record %u has more than 2^32 windows", r);
            wb[r + 1] = wb[r] + nw;
        }
        n_windows = wb[nrec];
    }
    const uint64_t n_out = n_windows * nq;
    if (n_out == 0) return TIDK_OK;
    if (!out_fwd || !out_rev) return fail(ctx, TIDK_EARG, "out_fwd/out_rev is NULL");

    uint64_t spw64 = (window + kSegBytes - 1) / kSegBytes;
    if (spw64 > 0xFFFFFFFFull) spw64 = 0xFFFFFFFFull; // records are < 2^64 bytes: never reached in practice
    const uint32_t spw = (uint32_t)spw64;
    const uint64_t n_items = n_windows * spw;
    if (n_items >= (1ull << 34)) return fail(ctx, TIDK_EARG, "more than 2^34 work items (windows x 32 KiB segments) in one call");

    int rc;
    // Two output buffers used alternately (the counts of call i leave the device while call i + 1 counts).  One
    // allocation each, forward counts then reverse counts, so that a caller whose two output arrays are adjacent
    // in memory gets both with a single copy.
    if (!async && (rc = drain_slots(ctx))) return rc; // a blocking call behind asynchronous ones: collect those first
    const int slot = (int)(ctx->wslot_next % (uint32_t)tidk_ctx::kWinSlots);
    tidk_ctx::WinSlot &S = ctx->wslot[slot];
    if ((rc = retire_slot(ctx, slot))) return rc;
    if (!S.k0) {
        CU(ctx, cudaEventCreate(&S.k0));
        CU(ctx, cudaEventCreate(&S.k1));
        CU(ctx, cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
    }
    if (async && !ctx->copy_st) CU(ctx, cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
    if (2 * n_out * 8 > ctx->out_dev[slot].cap) { // (re)allocation: nothing of this slot is in flight (retired above)
        ctx->out_zero_bytes[slot] = 0;
        if ((rc = ensure(ctx, ctx->out_dev[slot], 2 * n_out * 8))) return rc;
    }
    constexpr size_t kCounterBytes = 2 * (size_t)kQueues * kQueueStride * 8; // two sets of queue counters (each launch zeroes the other set)
    if (!ctx->misc.p) {
        if ((rc = ensure(ctx, ctx->misc, kCounterBytes))) return rc;
        CU(ctx, cudaMemsetAsync(ctx->misc.p, 0, kCounterBytes, ctx->stream));
        ctx->counter_flip = 0;
    }
    if (!ctx->h_flags) {
        CU(ctx, cudaHostAlloc((void **)&ctx->h_flags, 64, cudaHostAllocMapped));
        memset(ctx->h_flags, 0, 64);
        CU(ctx, cudaHostGetDevicePointer((void **)&ctx->d_flags, ctx->h_flags, 0));
    }
    if (ctx->counters_dirty) {
        CU(ctx, cudaMemsetAsync(ctx->misc.p, 0, kCounterBytes, ctx->stream));
        ctx->counter_flip = 0;
    }
    ctx->counters_dirty = true; // cleared when this call has completed all its launches
    uint64_t *d_fwd = (uint64_t *)ctx->out_dev[slot].p, *d_rev = d_fwd + n_out;
    // Flat form (window_counts_flat.cuh) for single queries whenever the source is the batch's own resident planes: it
    // adds into pre-zeroed counters.
    const bool flat_ok = packed && packed == buf->d_planes && buf->planes_ready && split == 0 && buf->d_validbits &&
                         !(getenv("TIDK_FLAT") && atoi(getenv("TIDK_FLAT")) == 0);
    if ((spw > 1 || flat_ok) && ctx->out_zero_bytes[slot] < 2 * n_out * 8) { // (a flat launch of the previous call may have zeroed it already)
        CU(ctx, cudaMemsetAsync(d_fwd, 0, 2 * n_out * 8, ctx->stream));
    }
    ctx->out_zero_bytes[slot] = 0; // this call writes into it
    bool next_zeroed_out = false;  // this call's first flat launch zeroes the NEXT slot's buffer on the side
    int zeroing_slot = -1;
    const int nslot = (slot + 1) % tidk_ctx::kWinSlots;
    bool nslot_ready = false;
    if (flat_ok) { // (before the first timed event: an allocation here must not sit between the events)
        if ((rc = retire_slot(ctx, nslot))) return rc; // its counts have left the device
        if (2 * n_out * 8 > ctx->out_dev[nslot].cap) {
            ctx->out_zero_bytes[nslot] = 0;
            if ((rc = ensure(ctx, ctx->out_dev[nslot], 2 * n_out * 8))) return rc;
        }
        nslot_ready = true;
    }

    uint32_t launches = 0;
    CU(ctx, cudaEventRecord(S.k0, ctx->stream));
    if (!cached) {
        ctx->items_buf_id = 0;
        if ((rc = ensure(ctx, ctx->win_base, wb.size() * 8))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->win_base.p, wb.data(), wb.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->items_buf_id = buf->id; ctx->items_window = window; ctx->items_nq = nq;
        ctx->items_n = n_items; ctx->items_nwin = n_windows;
        ctx->items_built = false;
    }
    // the per-window item table: built by the first launch of the per-window kernels that needs it
    auto need_items = [&]() -> int {
        if (ctx->items_built) return TIDK_OK;
        int rc2;
        if ((rc2 = ensure(ctx, ctx->items, n_items * sizeof(ItemDesc)))) return rc2;
        PrepJob pj{buf->d_off, (const uint64_t *)ctx->win_base.p, (ItemDesc *)ctx->items.p, window, n_items, nrec, spw, nq};
        window_items_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, ctx->stream>>>(pj);
        launches++;
        ctx->items_built = true;
        return TIDK_OK;
    };

    const uint64_t n_spans = (buf->total + kFlatSpanBases - 1) / kFlatSpanBases;
    auto need_spans = [&]() -> int { // the span table of the flat kernels: built by the first flat launch on (batch, window, query count)
        if (ctx->spans_buf_id == buf->id && ctx->spans_window == window && ctx->spans_nq == nq) return TIDK_OK;
        ctx->spans_buf_id = 0;
        int rc2;
        if ((rc2 = ensure(ctx, ctx->spans, n_spans * sizeof(SpanDesc)))) return rc2;
        FlatPrepJob fp{buf->d_off, (const uint64_t *)ctx->win_base.p, (SpanDesc *)ctx->spans.p, window, n_spans, 0, nrec, nq};
        flat_spans_kernel<<<(unsigned)((n_spans + 255) / 256), 256, 0, ctx->stream>>>(fp);
        launches++;
        ctx->spans_buf_id = buf->id; ctx->spans_window = window; ctx->spans_nq = nq;
        return TIDK_OK;
    };
    FlatJob fjob{};
    if (flat_ok) {
        fjob.planes = buf->d_planes; fjob.validbits = buf->d_validbits;
        fjob.rec_off = buf->d_off; fjob.win_base = (const uint64_t *)ctx->win_base.p;
        fjob.spans = nullptr; // set by the launch (need_spans)
        fjob.out_fwd = (unsigned long long *)d_fwd; fjob.out_rev = (unsigned long long *)d_rev;
        fjob.window = window; fjob.total = buf->total; fjob.n_spans = n_spans; fjob.span0 = 0;
        fjob.n_vblocks = planes_blocks(buf->total);
        fjob.n_records = nrec; fjob.nq_total = nq;
    }

    WindowJob job{};
    job.seq = buf->d_seq;
    job.packed = packed; // non-null: the host-packed planes are the source (every query has a static kernel)
    job.items = nullptr; // set by run_static / the run-time launches once the item table exists (need_items)
    job.out_fwd = d_fwd;
    job.out_rev = d_rev;
    job.err_flags = ctx->d_flags; // written only when a byte >= 0x80 is seen: no flag copy per call
    job.n_items = n_items;

    // split queries: plane kernels take A/C/G/T-only patterns up to 32 bases
    std::vector<uint32_t> fast, slow;
    for (uint32_t q = 0; q < nq; q++) {
        bool ok = qlens[q] <= (uint32_t)kMaxFastLen;
        for (uint32_t j = 0; ok && j < qlens[q]; j++) ok = is_acgt((uint8_t)queries[q][j]);
        (ok ? fast : slow).push_back(q);
    }
    std::vector<uint8_t> qb; // slow-query bytes (fwd then revcomp per query)
    std::vector<size_t> qoff;
    for (uint32_t q : slow) {
        qoff.push_back(qb.size());
        const uint8_t *f = (const uint8_t *)queries[q];
        qb.insert(qb.end(), f, f + qlens[q]);
        qb.resize(qb.size() + qlens[q]);
        revcomp_host(f, qlens[q], qb.data() + qb.size() - qlens[q]);
    }
    if (!qb.empty()) {
        if ((rc = ensure(ctx, ctx->qbytes, qb.size()))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->qbytes.p, qb.data(), qb.size(), cudaMemcpyHostToDevice, ctx->stream));
    }

    const uint64_t max_blocks = (uint64_t)ctx->sm_count * 8;
    const LaunchEnv env{ctx->sm_count, &ctx->occupancy, ctx->stream}; // launch helpers turn the SM count into the resident grid

    unsigned long long *counters = (unsigned long long *)ctx->misc.p; // two sets of queue counters, used alternately
    unsigned long long *next_zeroed = nullptr;
    auto next_counter = [&]() -> unsigned long long * {
        unsigned long long *c = counters + (ctx->counter_flip & 1u) * (kQueues * kQueueStride);
        ctx->counter_flip ^= 1u;
        next_zeroed = counters + (ctx->counter_flip & 1u) * (kQueues * kQueueStride);
        return c;
    };

    // queries that have a compile-time specialised kernel (directly or through their reverse
    // complement) run alone; the rest of the fast ones are fused, up to 16 per launch
    std::vector<uint32_t> fused;
    uint8_t rcbuf[kMaxFastLen];
    if (packed && !slow.empty()) return fail(ctx, TIDK_ECUDA, "internal: packed source with a literal query");
    // Hybrid upload: the windows that start below `split` are resident as ASCII, the rest as packed planes;
    // a static kernel then runs once per source over its share of the (position-ordered) items.
    uint64_t n_ascii_items = packed ? 0 : n_items;
    if (packed && split > 0) {
        if (spw != 1) return fail(ctx, TIDK_ECUDA, "internal: hybrid upload with multi-segment windows");
        for (uint32_t r = 0; r < nrec; r++) {
            const uint64_t rb = buf->h_off[r], re = buf->h_off[r + 1];
            if (rb >= split) break;
            const uint64_t nw = (re - rb) / window + ((re - rb) % window != 0);
            const uint64_t below = (split - rb + window - 1) / window; // windows of this record that start below split
            n_ascii_items += below < nw ? below : nw;
        }
    }
    int lazy_rc = TIDK_OK;
    auto run_static = [&](LaunchFn f_ascii, LaunchFn f_packed, const WindowJob &j0) {
        if ((lazy_rc = need_items())) return;
        WindowJob j = j0;
        j.items = (const ItemDesc *)ctx->items.p;
        if (n_ascii_items > 0) {
            WindowJob ja = j;
            ja.packed = nullptr;
            ja.n_items = n_ascii_items;
            ja.work_counter = next_counter(); ja.next_counter = next_zeroed;
            f_ascii(ja, env);
            launches++;
        }
        if (n_ascii_items < n_items) {
            WindowJob jp = j;
            jp.items = j.items + n_ascii_items;
            jp.n_items = n_items - n_ascii_items;
            jp.work_counter = next_counter(); jp.next_counter = next_zeroed;
            f_packed(jp, env);
            launches++;
        }
    };
    if (LaunchFn clade = (slow.empty() ? find_clade_set(queries, qlens, nq, false) : nullptr)) {
        WindowJob j2 = job; // the whole query list is a known clade: one fused pass
        j2.q0 = 0; j2.nq = nq;
        run_static(clade, packed ? find_clade_set(queries, qlens, nq, true) : nullptr, j2);
        if (lazy_rc) return lazy_rc;
        fast.clear();
    }
    for (uint32_t q : fast) {
        const uint8_t *f = (const uint8_t *)queries[q];
        revcomp_host(f, qlens[q], rcbuf);
        if (flat_ok) {
            if ((rc = need_spans())) return rc;
            FlatJob fj = fjob;
            fj.spans = (const SpanDesc *)ctx->spans.p;
            fj.q = q;
            if (!next_zeroed_out && nslot_ready) { // one memset node less per pass: 2-4 us of a 40 us pass on an eighth of the C4 genome
                fj.zero_ptr = (uint4 *)ctx->out_dev[nslot].p;
                fj.zero_n16 = n_out; // 2 * n_out * 8 bytes
                zeroing_slot = nslot;
                next_zeroed_out = true;
            }
            fj.work_counter = next_counter(); fj.next_counter = next_zeroed;
            if (FlatLaunchFn ff = find_static_flat(f, qlens[q])) {
                ff(fj, env);
            } else if ((ff = find_static_flat(rcbuf, qlens[q]))) { // the listed motif is this query's reverse complement: swap the outputs
                std::swap(fj.out_fwd, fj.out_rev);
                ff(fj, env);
            } else {
                FlatRuntime<1>::Params rp{};
                const StrandPattern pf = encode_strand(f, qlens[q]), pr = encode_strand(rcbuf, qlens[q]);
                for (uint32_t s2 = 0; s2 < qlens[q]; s2++) {
                    rp.flo[s2] = 0u - ((pf.plo >> s2) & 1u); rp.fhi[s2] = 0u - ((pf.phi >> s2) & 1u);
                    rp.rlo[s2] = 0u - ((pr.plo >> s2) & 1u); rp.rhi[s2] = 0u - ((pr.phi >> s2) & 1u);
                }
                launch_flat_runtime((int)qlens[q], fj, rp, env);
            }
            launches++;
            if (zeroing_slot >= 0) { ctx->out_zero_bytes[zeroing_slot] = 2 * n_out * 8; zeroing_slot = -1; } // (enqueued: true from here on in stream order)
            continue;
        }
        LaunchFn fn = find_static(f, qlens[q], false), fnp = packed ? find_static(f, qlens[q], true) : nullptr;
        bool swap = false;
        if (!fn && (fn = find_static(rcbuf, qlens[q], false))) {
            swap = true;
            fnp = packed ? find_static(rcbuf, qlens[q], true) : nullptr;
        }
        if (packed && (!fn || !fnp)) return fail(ctx, TIDK_ECUDA, "internal: packed source without a static kernel");
        if (!fn) { fused.push_back(q); continue; }
        WindowJob j2 = job;
        j2.q0 = q; j2.nq = 1;
        if (swap) { j2.out_fwd = d_rev; j2.out_rev = d_fwd; }
        run_static(fn, fnp, j2);
        if (lazy_rc) return lazy_rc;
    }
    // Run-time patterns: one unrolled single-query launch each (about 100 us + 6.5 us per base on C2),
    // unless there are many -- the fused run-time kernel is not unrolled and costs about 130 us per
    // query on top of one pass, so it only wins from about nine queries up.
    const bool fuse_runtime = fused.size() >= 9;
    for (size_t g0 = 0; g0 < fused.size();) {
        size_t g1 = g0 + 1; // consecutive query indices only (output addressing is q0 + i)
        while (fuse_runtime && g1 < fused.size() && g1 - g0 < (size_t)kMaxFastQueries && fused[g1] == fused[g1 - 1] + 1) g1++;
        RuntimeQueries rq{};
        for (size_t i = g0; i < g1; i++) {
            const uint32_t q = fused[i];
            const uint8_t *f = (const uint8_t *)queries[q];
            revcomp_host(f, qlens[q], rcbuf);
            rq.fwd[i - g0] = encode_strand(f, qlens[q]);
            rq.rev[i - g0] = encode_strand(rcbuf, qlens[q]);
            rq.len[i - g0] = (uint8_t)qlens[q];
        }
        rq.nq = (uint32_t)(g1 - g0);
        if ((rc = need_items())) return rc;
        WindowJob j2 = job;
        j2.items = (const ItemDesc *)ctx->items.p;
        j2.q0 = fused[g0]; j2.nq = rq.nq;
        j2.work_counter = next_counter(); j2.next_counter = next_zeroed;
        launch_planes(rq.nq == 1 ? (int)rq.len[0] : 0, j2, rq, env);
        launches++;
        g0 = g1;
    }
    for (size_t i = 0; i < slow.size(); i++) {
        const uint32_t q = slow[i];
        if (cached && !ctx->win_base.p) return fail(ctx, TIDK_ECUDA, "internal: window table missing");
        SlowJob sj{};
        sj.seq = buf->d_seq; sj.rec_off = buf->d_off; sj.win_base = (const uint64_t *)ctx->win_base.p;
        sj.out_fwd = d_fwd; sj.out_rev = d_rev;
        sj.work_counter = next_counter(); sj.next_counter = next_zeroed;
        sj.err_flags = job.err_flags;
        sj.window = window; sj.n_windows = n_windows; sj.n_records = nrec; sj.q = q; sj.nq_total = nq;
        sj.fwd = (const uint8_t *)ctx->qbytes.p + qoff[i];
        sj.rev = sj.fwd + qlens[q];
        sj.len = qlens[q];
        const uint64_t wbk = (n_windows + kWarpsPerBlock - 1) / kWarpsPerBlock;
        const int g2 = (int)(wbk < max_blocks ? wbk : max_blocks);
        window_counts_bytes_kernel<<<g2 < 1 ? 1 : g2, kWarpsPerBlock * 32, 0, ctx->stream>>>(sj);
        launches++;
    }
    CU(ctx, cudaEventRecord(S.k1, ctx->stream));
    CU(ctx, cudaGetLastError());

    cudaStream_t cst = ctx->stream;
    if (async) { // the copy runs beside the next call's kernels
        cst = ctx->copy_st;
        CU(ctx, cudaStreamWaitEvent(cst, S.k1, 0));
    }
    if (out_rev == out_fwd + n_out) {
        CU(ctx, cudaMemcpyAsync(out_fwd, d_fwd, 2 * n_out * 8, cudaMemcpyDeviceToHost, cst));
    } else {
        CU(ctx, cudaMemcpyAsync(out_fwd, d_fwd, n_out * 8, cudaMemcpyDeviceToHost, cst));
        CU(ctx, cudaMemcpyAsync(out_rev, d_rev, n_out * 8, cudaMemcpyDeviceToHost, cst));
    }
    CU(ctx, cudaEventRecord(S.done, cst));
    S.busy = true;
    S.launches = launches;
    ctx->wslot_next = (ctx->wslot_next + 1u) % (uint32_t)tidk_ctx::kWinSlots;
    ctx->counters_dirty = false; // every launch of this call is enqueued; the counters are handed on in stream order
    if (async) return TIDK_OK;
    const float ms0 = ctx->pending_kernel_ms;
    const uint32_t nl0 = ctx->pending_launches;
    if ((rc = retire_slot(ctx, slot))) return rc;
    ctx->last_kernel_ms = ctx->pending_kernel_ms - ms0;
    ctx->last_kernel_launches = launches;
    ctx->pending_kernel_ms = ms0; // a blocking call reports its own time; the pending totals belong to asynchronous calls
    ctx->pending_launches = nl0;
    const uint32_t h_err = *(volatile uint32_t *)ctx->h_flags;
    if (h_err) ctx->h_flags[0] = 0; // flags are sticky: clear after reporting (the stream is idle here)
    if (kernel_ms) *kernel_ms = ctx->last_kernel_ms;
    if (kernel_launches) *kernel_launches = launches;
    if (h_err & 1u)
        return fail(ctx, TIDK_ENONASCII, "sequence contains a byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107)");
    return TIDK_OK;
}

// every query runs on a compile-time plane kernel that reads packed planes (directly, as a known clade, or through
// its reverse complement)
static bool queries_packable(const char *const *queries, const uint32_t *qlens, uint32_t nq) {
    if (!queries || !qlens || nq == 0) return false;
    if (find_clade_set(queries, qlens, nq, true)) return true;
    uint8_t rcbuf[kMaxFastLen];
    for (uint32_t q = 0; q < nq; q++) {
        bool ok = queries[q] && qlens[q] >= 1 && qlens[q] <= (uint32_t)kMaxFastLen;
        for (uint32_t j = 0; ok && j < qlens[q]; j++) ok = is_acgt((uint8_t)queries[q][j]);
        if (!ok) return false;
        revcomp_host((const uint8_t *)queries[q], qlens[q], rcbuf);
        if (!find_static((const uint8_t *)queries[q], qlens[q], true) && !find_static(rcbuf, qlens[q], true)) return false;
    }
    return true;
}

// The planes of a resident batch are the source whenever every query has a plane-reading kernel (0.375 B per base
// instead of 1, nothing transposed); other queries scan the ASCII copy.
static const uint32_t *resident_planes(tidk_ctx *ctx, const tidk_seqbuf *buf, const char *const *queries, const uint32_t *qlens,
                                       uint32_t nq, int *rc) {
    *rc = TIDK_OK;
    const bool on = !(getenv("TIDK_PLANES_COUNT") && atoi(getenv("TIDK_PLANES_COUNT")) == 0); // (read per call: measurements switch)
    if (!on || !buf || !buf->planes_ready) return nullptr;
    const bool flat_on = !(getenv("TIDK_FLAT") && atoi(getenv("TIDK_FLAT")) == 0);
    bool all_fast = flat_on && queries && qlens && nq > 0; // the flat kernels take any A/C/G/T query up to 32 bases from the planes
    for (uint32_t q = 0; all_fast && q < nq; q++) {
        all_fast = queries[q] && qlens[q] >= 1 && qlens[q] <= (uint32_t)kMaxFastLen;
        for (uint32_t j = 0; all_fast && j < qlens[q]; j++) all_fast = is_acgt((uint8_t)queries[q][j]);
    }
    if (!all_fast && !queries_packable(queries, qlens, nq)) return nullptr;
    if (buf->planes_nonascii)
        *rc = fail(ctx, TIDK_ENONASCII, "sequence contains a byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107)");
    return buf->d_planes;
}

int tidk_cuda_window_counts_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, uint64_t window,
                                     const char *const *queries, const uint32_t *qlens,
                                     uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev,
                                     float *kernel_ms, uint32_t *kernel_launches) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::window_counts_resident(ctx, buf, window, queries, qlens, nq, out_fwd, out_rev, kernel_ms, kernel_launches, false);
    int rc;
    const uint32_t *planes = resident_planes(ctx, buf, queries, qlens, nq, &rc);
    if (rc) return rc;
    return window_counts_impl(ctx, buf, planes, window, queries, qlens, nq, out_fwd, out_rev, kernel_ms,
                              kernel_launches);
}

int tidk_cuda_window_counts_resident_async(tidk_ctx *ctx, const tidk_seqbuf *buf, uint64_t window,
                                           const char *const *queries, const uint32_t *qlens,
                                           uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::window_counts_resident(ctx, buf, window, queries, qlens, nq, out_fwd, out_rev, nullptr, nullptr, true);
    int rc;
    const uint32_t *planes = resident_planes(ctx, buf, queries, qlens, nq, &rc);
    if (rc) return rc;
    return window_counts_impl(ctx, buf, planes, window, queries, qlens, nq, out_fwd, out_rev, nullptr, nullptr, 0, /*async=*/true);
}

int tidk_cuda_wait(tidk_ctx *ctx, float *kernel_ms, uint32_t *kernel_launches) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::wait(ctx, kernel_ms, kernel_launches);
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = drain_slots(ctx);
    if (kernel_ms) *kernel_ms = ctx->pending_kernel_ms;
    if (kernel_launches) *kernel_launches = ctx->pending_launches;
    ctx->last_kernel_ms = ctx->pending_kernel_ms;
    ctx->last_kernel_launches = ctx->pending_launches;
    ctx->pending_kernel_ms = 0.f;
    ctx->pending_launches = 0;
    if (rc) return rc;
    if (ctx->h_flags) {
        const uint32_t h_err = *(volatile uint32_t *)ctx->h_flags;
        if (h_err) ctx->h_flags[0] = 0;
        if (h_err & 1u)
            return fail(ctx, TIDK_ENONASCII, "sequence contains a byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107)");
    }
    return TIDK_OK;
}

// ---- host-packed upload ----------------------------------------------------------------------
// Getting the bases to the device is the end-to-end bound of the host-pointer entry point (about 55 GB/s
// against a kernel that streams 5 TB/s), so when every query has a plane kernel the host packs the
// sequence into the lo / hi planes the matchers use (2 bits per base, host_pack.hpp) and reports the rare
// words whose validity plane is not all ones; only those cross the bus.  T worker threads each own a
// stream and a two-slot pinned staging ring; a worker packs one 4 MiB chunk of sequence (1 MiB of planes,
// contiguous on the device) into a free slot and issues its copy, so packing and DMA overlap without any
// cross-thread barrier.  See upload_packed() for the two-ended variant that is actually used.
// validity-plane exceptions of the host packer: (global word index << 32 | va) -> the va third of the word's chunk
__global__ void va_scatter_kernel(uint32_t *packed, const uint64_t *exc, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t e = exc[i], word = e >> 32;
    packed[(word >> tidk::kPackShift) * (3ull * tidk::kPackWords) + 2ull * tidk::kPackWords + (word & (tidk::kPackWords - 1))] = (uint32_t)e;
}

// ---- explore: planes back to ASCII ------------------------------------------------------------
// The explore scan compares raw bytes (explore.rs:208), so its host-pointer entry point uses the packers in
// STRICT form (valid = upper-case A C G T): an exception-free word is exactly the bytes its lo / hi planes say,
// and the words with anything else in them (lower case, N, IUPAC codes, the tail behind the last base) travel
// as records -- chunk-local word index + the 32 raw bytes -- in the third (validity) part of their chunk, written
// by the packer that found them and carried by the same copy as the planes.  On the device the packed chunks are
// expanded into the ordinary ASCII sequence buffer (one 16-byte store per thread, write-bound: 1.5 GB in about
// 0.3 ms) and the raw words laid over them; the scan then runs unchanged.  The bus carries 0.25 B per base
// instead of 1.
constexpr uint32_t kExcPerChunk = tidk::kPackWords / 16; // more exception words than this: the chunk goes up as ASCII
constexpr int kUnpackRounds = 16;

__global__ void __launch_bounds__(256) unpack_planes_kernel(const uint32_t *packed, const uint32_t *chunk_state, uint8_t *seq,
                                                            uint64_t half0, uint64_t n_half_words, uint64_t total) {
    __shared__ uint32_t lut[256]; // (hi nibble << 4 | lo nibble) -> four ASCII bases; (hi, lo): 00 A, 01 C, 11 G, 10 T
    {
        const uint32_t t = threadIdx.x, l4 = t & 15u, h4 = t >> 4;
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) {
            const uint32_t code = ((h4 >> b) & 1u) * 2u + ((l4 >> b) & 1u);
            v |= (uint32_t)("ACTG"[code]) << (8 * b);
        }
        lut[t] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kUnpackRounds; r++) { // 16 bases per thread and round, 64 KiB of sequence per CTA (the table is built once)
        const uint64_t i = half0 + ((uint64_t)blockIdx.x * kUnpackRounds + r) * 256 + threadIdx.x;
        if (i >= n_half_words) return;
        const uint64_t word = i >> 1;
        const uint64_t chunk = word >> tidk::kPackShift;
        if (chunk_state[chunk] == 0 || word * 32 >= total) continue; // the chunk went up as ASCII
        const uint32_t *pc = packed + chunk * (3ull * tidk::kPackWords) + (word & (tidk::kPackWords - 1));
        const uint32_t sh = (uint32_t)(i & 1) * 16u;
        const uint32_t lo = __ldg(pc) >> sh, hi = __ldg(pc + tidk::kPackWords) >> sh;
        uint4 o;
        o.x = lut[((hi & 0xFu) << 4) | (lo & 0xFu)];
        o.y = lut[(((hi >> 4) & 0xFu) << 4) | ((lo >> 4) & 0xFu)];
        o.z = lut[(((hi >> 8) & 0xFu) << 4) | ((lo >> 8) & 0xFu)];
        o.w = lut[(((hi >> 12) & 0xFu) << 4) | ((lo >> 12) & 0xFu)];
        *reinterpret_cast<uint4 *>(seq + i * 16) = o; // the buffer is 256-byte aligned and padded behind `total`
    }
}

// raw-word records of a packed chunk, in its validity third: [kExcPerChunk x u32 chunk-local word][n x 32 raw bytes].
// grid = (chunks from chunk0 on) x (2 * kExcPerChunk / 256); one thread per 16 bytes.
__global__ void __launch_bounds__(256) patch_words_kernel(const uint32_t *packed, const uint32_t *chunk_state, uint8_t *seq, uint64_t chunk0) {
    constexpr uint32_t kGroups = 2 * kExcPerChunk / 256;
    const uint64_t chunk = chunk0 + blockIdx.x / kGroups;
    const uint32_t st = chunk_state[chunk];
    if (st <= 1) return;
    const uint32_t i = (blockIdx.x % kGroups) * 256 + threadIdx.x, rec = i >> 1;
    if (rec >= st - 1) return;
    const uint32_t *third = packed + chunk * (3ull * tidk::kPackWords) + 2ull * tidk::kPackWords;
    const uint64_t word = chunk * (uint64_t)tidk::kPackWords + third[rec];
    *reinterpret_cast<uint4 *>(seq + word * 32 + (i & 1) * 16) = reinterpret_cast<const uint4 *>(third + kExcPerChunk)[i];
}

//
// Two-ended ("hybrid") upload.  The packers are bound by the host's memory bandwidth (about 76 GB/s of
// sequence on 16 vCPUs) while the bus would still have room, and a plain DMA of ASCII needs no CPU at
// all.  So the 4 MiB chunks are claimed from both ends: chunks of ASCII from the front of the caller's
// buffer go straight into the sequence buffer by DMA (a few copies in flight, issued by whichever packer
// passes by), the packers take chunks from the back, and they meet wherever the machine's pack rate and
// bus rate put the balance.  *split receives the first packed byte; the windows below it are counted
// from ASCII, the rest from planes.
// explore_mode: STRICT packing, chunks with many exception words sent as ASCII by their packer, and the packed
// chunks expanded into the sequence buffer afterwards (see unpack_planes_kernel): on return (stream order) the
// buffer holds the caller's bytes exactly.
static int upload_packed(tidk_ctx *ctx, const uint8_t *seq, uint64_t total, int T, bool allow_ascii, uint64_t *split,
                         bool explore_mode = false) {
    const uint64_t n_words = (total + 31) / 32;
    const uint64_t n_chunks = (n_words + kPackWords - 1) / kPackWords;
    const size_t chunk_bytes = 3ull * kPackWords * 4;
    const uint64_t chunk_seq = (uint64_t)kPackWords * 32; // sequence bytes per chunk
    const size_t plane2_bytes = 2ull * kPackWords * 4;   // lo + hi planes of a chunk: what a packer sends
    *split = 0;
    int rc;
    if ((rc = ensure(ctx, ctx->packed, (n_chunks + 1) * chunk_bytes))) return rc; // + one zero chunk: blocks read past the end
    if (!explore_mode) {
        CU(ctx, cudaMemsetAsync((char *)ctx->packed.p + n_chunks * chunk_bytes, 0, chunk_bytes, ctx->stream));
        // the validity plane (third part of every chunk) is all ones except where the packers say otherwise
        CU(ctx, cudaMemset2DAsync((char *)ctx->packed.p + 2ull * kPackWords * 4, chunk_bytes, 0xFF, (size_t)kPackWords * 4, n_chunks, ctx->stream));
    }
    std::vector<std::vector<uint64_t>> exc((size_t)T);
    std::vector<uint32_t> &chunk_state = ctx->h_chunk_state; // explore: 0 = the chunk went up as ASCII, else 1 + raw-word records
    chunk_state.assign((size_t)n_chunks, 0);
    std::vector<uint64_t> lane_bytes((size_t)T, 0);          // bytes each packer put on the bus
    if ((rc = ensure_lanes(ctx, T))) return rc;
    if (allow_ascii && !ctx->dma_st) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->dma_st, cudaStreamNonBlocking));
        for (int i = 0; i < tidk_ctx::kDmaDepth; i++)
            CU(ctx, cudaEventCreateWithFlags(&ctx->dma_ev[i], cudaEventDisableTiming));
    }
    // chunks [front, back] are unclaimed; the DMA thread takes `front++`, a packer `back--`
    std::atomic<int64_t> front{0}, back{(int64_t)n_chunks - 1};
    std::mutex claim_mu;
    auto claim = [&](bool from_front) -> int64_t {
        std::lock_guard<std::mutex> g(claim_mu);
        if (front.load() > back.load()) return -1;
        return from_front ? front.fetch_add(1) : back.fetch_sub(1);
    };
    std::atomic<int> cuda_err{0};
    std::atomic<uint32_t> nonascii{0};
    // The ASCII copies are issued by whichever packer passes by between two chunks (a dedicated thread would
    // be a 17th runnable thread on 16 busy cores and starve): retire finished copies, top the queue up to
    // kDmaDepth chunks in flight.  A packer comes by every ~50 us, one copy takes ~80 us.
    std::mutex dma_mu;
    int dma_head = 0, dma_tail = 0, dma_inflight = 0;
    auto pump = [&]() {
        if (!allow_ascii) return;
        std::unique_lock<std::mutex> g(dma_mu, std::try_to_lock);
        if (!g.owns_lock()) return;
        while (dma_inflight > 0 && cudaEventQuery(ctx->dma_ev[dma_tail]) == cudaSuccess) {
            dma_tail = (dma_tail + 1) % tidk_ctx::kDmaDepth;
            dma_inflight--;
        }
        while (dma_inflight < tidk_ctx::kDmaDepth) {
            const int64_t c = claim(true);
            if (c < 0) break;
            const uint64_t byte0 = (uint64_t)c * chunk_seq;
            const uint64_t n = total - byte0 < chunk_seq ? total - byte0 : chunk_seq;
            if (cudaMemcpyAsync(ctx->scratch.d_seq + byte0, seq + byte0, n, cudaMemcpyHostToDevice, ctx->dma_st) != cudaSuccess ||
                cudaEventRecord(ctx->dma_ev[dma_head], ctx->dma_st) != cudaSuccess) { cuda_err = 1; break; }
            dma_head = (dma_head + 1) % tidk_ctx::kDmaDepth;
            dma_inflight++;
        }
    };
    auto worker = [&](int t) {
        cudaSetDevice(ctx->device);
        tidk_ctx::PackLane &L = ctx->pack_lanes[t];
        uint32_t flag = 0;
        for (int slot = 0;; slot ^= 1) {
            pump();
            const int64_t c = claim(false);
            if (c < 0) break;
            if (cudaEventSynchronize(L.ev[slot]) != cudaSuccess) { cuda_err = 1; break; } // slot's previous copy done
            uint32_t *buf = L.pin + (size_t)slot * 3 * kPackWords;
            const uint64_t byte0 = (uint64_t)c * chunk_seq;
            const uint64_t have = total - byte0 < chunk_seq ? total - byte0 : chunk_seq;
            constexpr uint32_t kSub = 8; // packed in eighths, looking after the ASCII copy queue in between
            const size_t exc_before = exc[(size_t)t].size();
            for (uint32_t s8 = 0; s8 < kSub; s8++) {
                const uint32_t w0 = s8 * (kPackWords / kSub);
                const uint64_t b0 = (uint64_t)w0 * 32;
                if (explore_mode)
                    pack_words2<true>(seq + byte0 + b0, (size_t)(have > b0 ? have - b0 : 0), kPackWords / kSub, buf + w0,
                                      buf + kPackWords + w0, (uint64_t)c * kPackWords + w0, exc[(size_t)t], &flag);
                else
                    pack_words2(seq + byte0 + b0, (size_t)(have > b0 ? have - b0 : 0), kPackWords / kSub, buf + w0,
                                buf + kPackWords + w0, (uint64_t)c * kPackWords + w0, exc[(size_t)t], &flag);
                pump();
            }
            // explore: an exception word costs 36 bytes on the bus, a packed word saves 24 -- a chunk with more than
            // one exception word in sixteen (soft-masked stretches, N gaps) goes up as ASCII instead
            if (explore_mode) // words wholly behind the last base pack as invalid: nothing to reproduce there
                while (exc[(size_t)t].size() > exc_before && (exc[(size_t)t].back() >> 32) * 32 >= total) exc[(size_t)t].pop_back();
            size_t copy_bytes = plane2_bytes;
            if (explore_mode) {
                std::vector<uint64_t> &ev = exc[(size_t)t];
                const size_t n_new = ev.size() - exc_before;
                if (n_new > kExcPerChunk) {
                    ev.resize(exc_before);
                    if (cudaMemcpyAsync(ctx->scratch.d_seq + byte0, seq + byte0, have, cudaMemcpyHostToDevice, L.st) != cudaSuccess ||
                        cudaEventRecord(L.ev[slot], L.st) != cudaSuccess) { cuda_err = 1; break; }
                    lane_bytes[(size_t)t] += have;
                    continue; // chunk_state stays 0: ASCII
                }
                // the records ride in the validity third of the staging slot, behind the planes: one copy
                uint32_t *idx = buf + 2 * kPackWords;
                uint8_t *rawp = reinterpret_cast<uint8_t *>(idx + kExcPerChunk);
                for (size_t i = 0; i < n_new; i++) {
                    const uint64_t word = ev[exc_before + i] >> 32, w0b = word * 32;
                    const size_t nb = (size_t)(total - w0b < 32 ? total - w0b : 32);
                    idx[i] = (uint32_t)(word - (uint64_t)c * kPackWords);
                    memcpy(rawp + i * 32, seq + w0b, nb);
                    if (nb < 32) memset(rawp + i * 32 + nb, 0, 32 - nb);
                }
                ev.resize(exc_before);
                chunk_state[(size_t)c] = 1u + (uint32_t)n_new;
                if (n_new) copy_bytes += (size_t)kExcPerChunk * 4 + n_new * 32;
            }
            lane_bytes[(size_t)t] += copy_bytes;
            if (cudaMemcpyAsync((char *)ctx->packed.p + (uint64_t)c * chunk_bytes, buf, copy_bytes, cudaMemcpyHostToDevice, L.st) != cudaSuccess ||
                cudaEventRecord(L.ev[slot], L.st) != cudaSuccess) { cuda_err = 1; break; }
        }
        if (cudaStreamSynchronize(L.st) != cudaSuccess) cuda_err = 1;
        if (flag) nonascii = 1;
    };
    ctx->pack_pool.run(T, worker);
    if (allow_ascii && cudaStreamSynchronize(ctx->dma_st) != cudaSuccess) cuda_err = 1;
    if (cuda_err) return fail(ctx, TIDK_ECUDA, "packed upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (nonascii)
        return fail(ctx, TIDK_ENONASCII, "sequence contains a byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107)");
    const uint64_t n_ascii = (uint64_t)front.load() < n_chunks ? (uint64_t)front.load() : n_chunks; // chunks [0, n_ascii) went as ASCII
    uint64_t sp = n_ascii * chunk_seq;
    if (sp > total) sp = total;
    if (explore_mode) {
        uint64_t n_packed = 0;
        for (uint64_t c = 0; c < n_chunks; c++) n_packed += chunk_state[(size_t)c] != 0;
        ctx->last_h2d += sp + n_chunks * 4; // the front chunks by DMA + the chunk table
        for (uint64_t b : lane_bytes) ctx->last_h2d += b;
        if (n_packed) {
            if ((rc = ensure(ctx, ctx->chunk_state, (size_t)n_chunks * 4))) return rc;
            // (the host table is a ctx member: it outlives this copy; the next upload starts after a stream sync)
            CU(ctx, cudaMemcpyAsync(ctx->chunk_state.p, chunk_state.data(), (size_t)n_chunks * 4, cudaMemcpyHostToDevice, ctx->stream));
            const uint64_t n_half = n_chunks * (uint64_t)kPackWords * 2, half0 = n_ascii * (uint64_t)kPackWords * 2;
            unpack_planes_kernel<<<(unsigned)((n_half - half0 + 256 * kUnpackRounds - 1) / (256 * kUnpackRounds)), 256, 0, ctx->stream>>>(
                (const uint32_t *)ctx->packed.p, (const uint32_t *)ctx->chunk_state.p, ctx->scratch.d_seq, half0, n_half, total);
            CU(ctx, cudaGetLastError());
            bool any_rec = false;
            for (uint64_t c = n_ascii; c < n_chunks && !any_rec; c++) any_rec = chunk_state[(size_t)c] > 1;
            if (any_rec) {
                patch_words_kernel<<<(unsigned)((n_chunks - n_ascii) * (2 * kExcPerChunk / 256)), 256, 0, ctx->stream>>>(
                    (const uint32_t *)ctx->packed.p, (const uint32_t *)ctx->chunk_state.p, ctx->scratch.d_seq, n_ascii);
                CU(ctx, cudaGetLastError());
            }
        }
        CU(ctx, cudaMemsetAsync(ctx->scratch.d_seq + total, 0, kSeqPad, ctx->stream));
        *split = sp;
        return TIDK_OK;
    }
    if (n_ascii > 0) {
        // the last window below the split may reach up to a window (<= 32 KiB here) plus a block beyond it
        const uint64_t extra = total - sp < (64u << 10) ? total - sp : (64u << 10);
        if (extra) CU(ctx, cudaMemcpyAsync(ctx->scratch.d_seq + sp, seq + sp, extra, cudaMemcpyHostToDevice, ctx->stream));
        if (sp + extra == total) CU(ctx, cudaMemsetAsync(ctx->scratch.d_seq + total, 0, kSeqPad, ctx->stream));
        ctx->last_h2d += sp + extra;
    }
    ctx->last_h2d += (n_chunks - n_ascii) * plane2_bytes;
    size_t n_exc = 0;
    for (const auto &v : exc) n_exc += v.size();
    if (n_exc) {
        std::vector<uint64_t> all;
        all.reserve(n_exc);
        for (const auto &v : exc) all.insert(all.end(), v.begin(), v.end());
        if ((rc = ensure(ctx, ctx->exc, n_exc * 8))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->exc.p, all.data(), n_exc * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream)); // `all` is pageable and goes out of scope
        va_scatter_kernel<<<(unsigned)((n_exc + 255) / 256), 256, 0, ctx->stream>>>((uint32_t *)ctx->packed.p, (const uint64_t *)ctx->exc.p, n_exc);
        CU(ctx, cudaGetLastError());
        ctx->last_h2d += n_exc * 8;
    }
    *split = sp;
    return TIDK_OK;
}

int tidk_cuda_window_counts(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                            uint32_t n_records, uint64_t window, const char *const *queries,
                            const uint32_t *qlens, uint32_t nq, uint64_t *out_fwd,
                            uint64_t *out_rev) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::window_counts(ctx, seq, offsets, n_records, window, queries, qlens, nq, out_fwd, out_rev);
    if (window == 0) return fail(ctx, TIDK_EARG, "window must be > 0 (slice::chunks(0) panics, search.rs:98)");
    CU(ctx, cudaSetDevice(ctx->device));
    // host-packed path: worth it when there are enough cores to pack faster than PCIe moves ASCII,
    // and possible when every query runs on a compile-time plane kernel
    const uint64_t total = (n_records && offsets) ? offsets[n_records] : 0;
    const char *env = getenv("TIDK_HOST_PACK");
    const int T = pack_threads();
    // 12+ packer threads: 16 cores pack ~76 GB/s of sequence on the B200 hosts (tools/pack_bench.cpp).  With
    // fewer threads per rank (2-, 4-, 8-GPU jobs sharing one host) the box is bound by its memory bandwidth,
    // not by the bus (8 ranks: 184 GB/s of plain DMA in aggregate), and packing -- which reads the sequence
    // and writes + re-reads the planes -- only adds to that: measured 180 vs 184 Gbases/s on 8 GPUs.
    const char *eh = getenv("TIDK_HYBRID_UPLOAD");
    // the DMA front of the two-ended upload reads the caller's buffer directly: it has to be pinned (from pageable
    // memory each 4 MiB copy would be staged synchronously on the packer thread that issues it)
    bool pageable = true;
    {
        cudaPointerAttributes a{};
        if (seq && cudaPointerGetAttributes(&a, seq) == cudaSuccess) pageable = a.type == cudaMemoryTypeUnregistered;
        (void)cudaGetLastError();
    }
    const bool hybrid = !pageable && window <= kSegBytes && (eh ? atoi(eh) != 0 : true); // windows of one 32 KiB item can be split between the sources
    const bool packable = total >= (32u << 20) && seq && (env ? atoi(env) != 0 : T >= 12) && T >= 1 && queries_packable(queries, qlens, nq);
    uint64_t n_out = 0;
    for (uint32_t r = 0; r < n_records && offsets; r++) {
        const uint64_t len = offsets[r + 1] - offsets[r];
        n_out += (len / window + (len % window != 0)) * nq;
    }
    ctx->last_d2h = n_out * 16 + 4;
    ctx->last_packed = packable ? 1 : 0;
    ctx->last_h2d = ((uint64_t)n_records + 1) * 16 + (packable ? 0 : total);
    if (packable) {
        int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records, /*seq_mode=*/hybrid ? 2 : 0);
        if (rc) return rc;
        uint64_t split = 0;
        if ((rc = upload_packed(ctx, seq, total, T, hybrid, &split))) return rc;
        return window_counts_impl(ctx, &ctx->scratch, (const uint32_t *)ctx->packed.p, window, queries, qlens, nq,
                                  out_fwd, out_rev, nullptr, nullptr, split);
    }
    int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records);
    if (rc) return rc;
    return window_counts_impl(ctx, &ctx->scratch, nullptr, window, queries, qlens, nq, out_fwd,
                              out_rev, nullptr, nullptr);
}

int tidk_cuda_explore_runs_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, double distance,
                                    uint32_t kmin, uint32_t kmax, uint64_t threshold,
                                    tidk_run **out_runs, uint64_t *n_runs, float *kernel_ms,
                                    uint32_t *kernel_launches) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::explore_runs_resident(ctx, buf, distance, kmin, kmax, threshold, out_runs, n_runs, kernel_ms, kernel_launches);
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    if (!buf || !out_runs || !n_runs) return fail(ctx, TIDK_EARG, "NULL argument");
    *out_runs = nullptr; *n_runs = 0;
    if (!(distance <= 0.5)) return fail(ctx, TIDK_EARG, "distance must be <= 0.5 (explore.rs:41-43)");
    if (kmin == 0 || kmax < kmin) return fail(ctx, TIDK_EARG, "need 1 <= kmin <= kmax");
    if (kmax > (uint32_t)kMaxExploreK) return fail(ctx, TIDK_EARG, "k up to %d supported", kMaxExploreK);
    CU(ctx, cudaSetDevice(ctx->device));
    std::string err;
    int rc = explore_runs_device(ctx->ex, ctx->stream, ctx->ev0, ctx->ev1, ctx->sm_count, buf->d_seq,
                                 buf->d_off, buf->h_off.data(), buf->n_records, distance, kmin, kmax,
                                 threshold, out_runs, n_runs, kernel_ms, kernel_launches, err,
                                 &ctx->explore_scan_ms, &ctx->explore_post_ms, ctx->explore_host_ms, nullptr,
                                 buf->planes_ready ? buf->d_planes : nullptr, buf->planes_ready ? buf->d_plainbits : nullptr,
                                 buf->planes_ready && buf->planes_all_plain);
    if (rc) return fail(ctx, rc, "%s", err.c_str());
    return TIDK_OK;
}

int tidk_cuda_explore_runs(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                           uint32_t n_records, double distance, uint32_t kmin, uint32_t kmax,
                           uint64_t threshold, tidk_run **out_runs, uint64_t *n_runs) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) return multi::explore_runs(ctx, seq, offsets, n_records, distance, kmin, kmax, threshold, out_runs, n_runs);
    CU(ctx, cudaSetDevice(ctx->device));
    // Compacted upload: with --distance 0.01 only 2 % of a genome is ever looked at, so only the slices
    // cross the bus (C3: 10 MB instead of 500 MB), laid out back to back with room for the block
    // over-read; the slice table then points into that buffer.  Per-slice copies only pay for few, long
    // records -- read sets (d = 0.5 anyway) and small d over many records keep the whole-batch upload.
    if (out_runs && n_runs && seq && offsets && n_records > 0 && n_records <= 4096 && distance > 0.0 && distance < 0.4 &&
        kmin >= 1 && kmax >= kmin && kmax <= (uint32_t)kMaxExploreK) {
        *out_runs = nullptr; *n_runs = 0;
        if (offsets[0] != 0) return fail(ctx, TIDK_EARG, "offsets[0] must be 0");
        std::vector<SliceDesc> sl((size_t)n_records * 2);
        std::vector<uint64_t> src((size_t)n_records * 2);
        uint64_t cur = 0;
        for (uint32_t r = 0; r < n_records; r++) {
            if (offsets[r + 1] < offsets[r]) return fail(ctx, TIDK_EARG, "offsets must be non-decreasing");
            const uint64_t len = offsets[r + 1] - offsets[r];
            const uint64_t dist = split_dist(len, distance);
            if (dist > len) return fail(ctx, TIDK_EARG, "slice longer than record");
            for (int side = 0; side < 2; side++) {
                sl[2 * (size_t)r + side] = {cur, dist};
                src[2 * (size_t)r + side] = side == 0 ? offsets[r] : offsets[r] + (len - dist); // explore.rs:157-158
                cur += (dist + 2048 + 255) & ~255ull; // the scan reads whole 1 KiB blocks, up to kmax bases behind the slice
            }
        }
        tidk_seqbuf *b = &ctx->scratch;
        int rc = seqbuf_fill(ctx, b, seq, offsets, n_records, /*seq_mode=*/0); // offsets only (totals, record lengths)
        if (rc) return rc;
        const size_t need = (size_t)cur + kSeqPad;
        if (need > b->seq_cap) {
            if (b->d_seq) cudaFree(b->d_seq);
            b->d_seq = nullptr; b->seq_cap = 0;
            CU(ctx, cudaMalloc((void **)&b->d_seq, need));
            b->seq_cap = need;
        }
        for (size_t i = 0; i < sl.size(); i++)
            if (sl[i].len)
                CU(ctx, cudaMemcpyAsync(b->d_seq + sl[i].beg, seq + src[i], sl[i].len, cudaMemcpyHostToDevice, ctx->stream));
        std::string err;
        rc = explore_runs_device(ctx->ex, ctx->stream, ctx->ev0, ctx->ev1, ctx->sm_count, b->d_seq, b->d_off, b->h_off.data(),
                                 n_records, distance, kmin, kmax, threshold, out_runs, n_runs, nullptr, nullptr, err,
                                 &ctx->explore_scan_ms, &ctx->explore_post_ms, ctx->explore_host_ms, sl.data());
        b->id = ctx->next_buf_id++; // the buffer no longer holds the batch as uploaded: nothing cached may refer to it
        if (rc) return fail(ctx, rc, "%s", err.c_str());
        return TIDK_OK;
    }
    // Whole-batch upload (read sets, --distance near 0.5).  With enough host cores the sequence goes up two-ended
    // like the window counter's (ASCII chunks by DMA from the front, STRICT-packed planes from the back) and the
    // packed part is expanded back to ASCII on the device: the scan is unchanged, the bus carries a quarter.
    {
        const uint64_t total = (n_records && offsets) ? offsets[n_records] : 0;
        const char *env = getenv("TIDK_EXPLORE_PACK");
        const int T = pack_threads();
        const bool forced = env && atoi(env) != 0;
        if (seq && total > 0 && T >= 1 && (env ? forced : (T >= 12 && total >= (32u << 20)))) {
            int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records, /*seq_mode=*/2);
            if (rc) return rc;
            cudaPointerAttributes a;
            bool pageable = true;
            if (cudaPointerGetAttributes(&a, seq) == cudaSuccess) pageable = a.type == cudaMemoryTypeUnregistered;
            else (void)cudaGetLastError();
            const char *eh = getenv("TIDK_HYBRID_UPLOAD");
            const bool hybrid = !pageable && (eh ? atoi(eh) != 0 : true); // DMA straight from the caller's buffer needs it pinned
            ctx->last_h2d = ((uint64_t)n_records + 1) * 8;
            ctx->last_packed = 1;
            uint64_t split = 0;
            if ((rc = upload_packed(ctx, seq, total, T, hybrid, &split, /*explore_mode=*/true))) {
                ctx->scratch.id = ctx->next_buf_id++;
                return rc;
            }
            return tidk_cuda_explore_runs_resident(ctx, &ctx->scratch, distance, kmin, kmax, threshold,
                                                   out_runs, n_runs, nullptr, nullptr);
        }
    }
    int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records);
    if (rc) return rc;
    ctx->last_packed = 0;
    ctx->last_h2d = ((uint64_t)n_records + 1) * 8 + ((n_records && offsets) ? offsets[n_records] : 0);
    return tidk_cuda_explore_runs_resident(ctx, &ctx->scratch, distance, kmin, kmax, threshold,
                                           out_runs, n_runs, nullptr, nullptr);
}

void tidk_cuda_free_runs(tidk_run *runs) {
    if (!runs) return;
    tidk::run_list_free(runs); // the header in front of the list names its context's pool of pinned blocks, or none (explore_post.cuh)
}

int tidk_cuda_seqbuf_planes_info(const tidk_seqbuf *buf, int *ready, float *build_ms, uint64_t *bytes) {
    if (!buf) return TIDK_EARG;
    if (!buf->parts.empty()) { // multi-device batch: ready on every part, the slowest conversion, the sum of the bytes
        int all = 1;
        float mx = 0.f;
        uint64_t sum = 0;
        for (size_t g = 0; g < buf->parts.size(); g++) {
            if (buf->cut[g + 1] <= buf->cut[g] || !buf->parts[g]) continue;
            int r = 0; float ms = 0.f; uint64_t by = 0;
            tidk_cuda_seqbuf_planes_info(buf->parts[g], &r, &ms, &by);
            all &= r; mx = ms > mx ? ms : mx; sum += by;
        }
        if (ready) *ready = all;
        if (build_ms) *build_ms = mx;
        if (bytes) *bytes = sum;
        return TIDK_OK;
    }
    if (ready) *ready = buf->planes_ready ? 1 : 0;
    if (build_ms) *build_ms = buf->planes_ready ? buf->planes_ms : 0.f;
    if (bytes) *bytes = buf->planes_ready ? (uint64_t)buf->planes_cap + 2 * (uint64_t)buf->bits_cap : 0;
    return TIDK_OK;
}

int tidk_cuda_last_transfer(const tidk_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int *host_packed) {
    if (!ctx) return TIDK_EARG;
    if (!ctx->subs.empty()) { // the sum over the devices
        uint64_t h = 0, d = 0;
        int pk = 0;
        for (const tidk_ctx *sub : ctx->subs) { h += sub->last_h2d; d += sub->last_d2h; pk |= sub->last_packed; }
        if (h2d_bytes) *h2d_bytes = h;
        if (d2h_bytes) *d2h_bytes = d;
        if (host_packed) *host_packed = pk;
        return TIDK_OK;
    }
    if (h2d_bytes) *h2d_bytes = ctx->last_h2d;
    if (d2h_bytes) *d2h_bytes = ctx->last_d2h;
    if (host_packed) *host_packed = ctx->last_packed;
    return TIDK_OK;
}

int tidk_cuda_last_kernel_time(const tidk_ctx *ctx, float *kernel_ms, uint32_t *kernel_launches) {
    if (!ctx) return TIDK_EARG;
    if (kernel_ms) *kernel_ms = ctx->last_kernel_ms;
    if (kernel_launches) *kernel_launches = ctx->last_kernel_launches;
    return TIDK_OK;
}

int tidk_cuda_last_explore_timings(const tidk_ctx *ctx, float *out, int n) {
    if (!ctx || !out || n <= 0) return 0;
    const float v[4] = {ctx->explore_scan_ms, ctx->explore_post_ms, ctx->explore_host_ms[0], ctx->explore_host_ms[1]};
    const int m = n < 4 ? n : 4;
    for (int i = 0; i < m; i++) out[i] = v[i];
    return m;
}

} // extern "C"
