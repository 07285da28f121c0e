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
#include <string>
#include <vector>

#include "explore.cuh"
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
};

struct tidk_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    tidk_seqbuf scratch; // staging for the host-pointer entry points
    DevBuf win_base, out_fwd, out_rev, misc, qbytes, runs;
    ExploreScratch ex;
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

int seqbuf_fill(tidk_ctx *ctx, tidk_seqbuf *b, const uint8_t *seq, const uint64_t *offsets,
                uint32_t n_records) {
    if (n_records && (!offsets)) return fail(ctx, TIDK_EARG, "offsets is NULL");
    const uint64_t total = n_records ? offsets[n_records] : 0;
    if (n_records && offsets[0] != 0) return fail(ctx, TIDK_EARG, "offsets[0] must be 0");
    for (uint32_t r = 0; r < n_records; r++)
        if (offsets[r + 1] < offsets[r]) return fail(ctx, TIDK_EARG, "offsets must be non-decreasing");
    if (total && !seq) return fail(ctx, TIDK_EARG, "seq is NULL");
    const size_t need = (size_t)total + kSeqPad;
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
    if (total) CU(ctx, cudaMemcpyAsync(b->d_seq, seq, total, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(b->d_seq + total, 0, kSeqPad, ctx->stream));
    CU(ctx, cudaMemcpyAsync(b->d_off, b->h_off.data(), off_need, cudaMemcpyHostToDevice, ctx->stream));
    return TIDK_OK;
}

void seqbuf_release(tidk_seqbuf *b) {
    if (b->d_seq) cudaFree(b->d_seq);
    if (b->d_off) cudaFree(b->d_off);
    b->d_seq = nullptr; b->d_off = nullptr; b->seq_cap = b->off_cap = 0;
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

template <int L>
void launch_fixed(const WindowJob &job, const FastQueries &fq, int grid, cudaStream_t st) {
    window_counts_planes_kernel<L><<<grid, kWarpsPerBlock * 32, 0, st>>>(job, fq);
}

void launch_planes(int L, const WindowJob &job, const FastQueries &fq, int grid, cudaStream_t st) {
    switch (L) {
#define C_(n) case n: launch_fixed<n>(job, fq, grid, st); break;
        C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13) C_(14)
        C_(15) C_(16) C_(17) C_(18) C_(19) C_(20) C_(21) C_(22) C_(23) C_(24) C_(25) C_(26) C_(27)
        C_(28) C_(29) C_(30) C_(31) C_(32)
#undef C_
    default: launch_fixed<0>(job, fq, grid, st); break;
    }
}

} // namespace

extern "C" {

uint32_t tidk_cuda_abi_version(void) { return 1; }

int tidk_cuda_create(const int *device_ids, int n_devices, tidk_ctx **out) {
    if (!out) return fail(nullptr, TIDK_EARG, "out is NULL");
    *out = nullptr;
    if (n_devices < 0 || n_devices > 1)
        return fail(nullptr, TIDK_EARG, "one device per ctx (got %d); run one process per GPU", n_devices);
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
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    seqbuf_release(&ctx->scratch);
    for (DevBuf *b : {&ctx->win_base, &ctx->out_fwd, &ctx->out_rev, &ctx->misc, &ctx->qbytes, &ctx->runs})
        if (b->p) cudaFree(b->p);
    explore_scratch_release(ctx->ex);
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
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
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
    if (rc != TIDK_OK) { seqbuf_release(b); delete b; return rc; }
    *out = b;
    return TIDK_OK;
}

void tidk_cuda_seqbuf_free(tidk_ctx *ctx, tidk_seqbuf *buf) {
    if (!buf) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    seqbuf_release(buf);
    delete buf;
}

int tidk_cuda_window_counts_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, uint64_t window,
                                     const char *const *queries, const uint32_t *qlens,
                                     uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev,
                                     float *kernel_ms, uint32_t *kernel_launches) {
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
    std::vector<uint64_t> wb((size_t)nrec + 1, 0);
    for (uint32_t r = 0; r < nrec; r++) {
        uint64_t len = buf->h_off[r + 1] - buf->h_off[r];
        wb[r + 1] = wb[r] + (len / window + (len % window != 0));
    }
    const uint64_t n_windows = wb[nrec];
    const uint64_t n_out = n_windows * nq;
    if (n_out == 0) return TIDK_OK;
    if (!out_fwd || !out_rev) return fail(ctx, TIDK_EARG, "out_fwd/out_rev is NULL");

    int rc;
    if ((rc = ensure(ctx, ctx->win_base, wb.size() * 8))) return rc;
    if ((rc = ensure(ctx, ctx->out_fwd, n_out * 8))) return rc;
    if ((rc = ensure(ctx, ctx->out_rev, n_out * 8))) return rc;
    if ((rc = ensure(ctx, ctx->misc, 256))) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->win_base.p, wb.data(), wb.size() * 8, cudaMemcpyHostToDevice, ctx->stream));

    WindowJob job{};
    job.seq = buf->d_seq;
    job.rec_off = buf->d_off;
    job.win_base = (const uint64_t *)ctx->win_base.p;
    job.out_fwd = (uint64_t *)ctx->out_fwd.p;
    job.out_rev = (uint64_t *)ctx->out_rev.p;
    job.err_flags = (uint32_t *)((char *)ctx->misc.p + 128);
    job.window = window;
    job.n_windows = n_windows;
    job.n_records = nrec;
    job.segs_per_window = (uint32_t)((window + kSegBytes - 1) / kSegBytes);
    if (window > (uint64_t)kSegBytes * 0xFFFFFFFFull) job.segs_per_window = 0xFFFFFFFFu;
    job.nq_total = nq;
    CU(ctx, cudaMemsetAsync(ctx->misc.p, 0, 256, ctx->stream));
    if (job.segs_per_window > 1) {
        CU(ctx, cudaMemsetAsync(job.out_fwd, 0, n_out * 8, ctx->stream));
        CU(ctx, cudaMemsetAsync(job.out_rev, 0, n_out * 8, ctx->stream));
    }

    // split queries: plane kernels take A/C/G/T-only patterns up to 32 bases
    std::vector<uint32_t> fast, slow;
    for (uint32_t q = 0; q < nq; q++) {
        bool ok = qlens[q] <= (uint32_t)kMaxFastLen;
        for (uint32_t j = 0; ok && j < qlens[q]; j++) ok = is_acgt((uint8_t)queries[q][j]);
        (ok ? fast : slow).push_back(q);
    }
    // slow-query bytes (fwd then rev per query)
    std::vector<uint8_t> qb;
    std::vector<size_t> qoff;
    for (uint32_t q : slow) {
        qoff.push_back(qb.size());
        const uint8_t *f = (const uint8_t *)queries[q];
        qb.insert(qb.end(), f, f + qlens[q]);
        for (uint32_t j = 0; j < qlens[q]; j++) qb.push_back(comp_base(f[qlens[q] - 1 - j]));
    }
    if (!qb.empty()) {
        if ((rc = ensure(ctx, ctx->qbytes, qb.size()))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->qbytes.p, qb.data(), qb.size(), cudaMemcpyHostToDevice, ctx->stream));
    }

    const uint64_t n_items = n_windows * job.segs_per_window;
    uint64_t want_blocks = (n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    uint64_t max_blocks = (uint64_t)ctx->sm_count * 8;
    int grid = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
    if (grid < 1) grid = 1;

    uint32_t launches = 0;
    unsigned long long *counters = (unsigned long long *)ctx->misc.p; // 16 counters available
    uint32_t counter_i = 0;
    auto next_counter = [&]() -> unsigned long long * {
        if (counter_i == 16) { // recycle (stream-ordered)
            cudaMemsetAsync(ctx->misc.p, 0, 128, ctx->stream);
            counter_i = 0;
        }
        return counters + counter_i++;
    };
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (fast.size() == 1) {
        uint32_t q = fast[0];
        FastQueries fq{};
        std::vector<uint8_t> rc_pat(qlens[q]);
        const uint8_t *f = (const uint8_t *)queries[q];
        for (uint32_t j = 0; j < qlens[q]; j++) rc_pat[j] = comp_base(f[qlens[q] - 1 - j]);
        fq.fwd[0] = encode_strand(f, qlens[q]);
        fq.rev[0] = encode_strand(rc_pat.data(), qlens[q]);
        fq.len[0] = (uint8_t)qlens[q];
        job.nq = 1; job.q0 = q;
        job.work_counter = next_counter();
        launch_planes((int)qlens[q], job, fq, grid, ctx->stream);
        launches++;
    } else {
        for (size_t g0 = 0; g0 < fast.size();) {
            // consecutive query indices only (output addressing uses q0 + i)
            size_t g1 = g0 + 1;
            while (g1 < fast.size() && g1 - g0 < (size_t)kMaxFastQueries && fast[g1] == fast[g1 - 1] + 1) g1++;
            FastQueries fq{};
            for (size_t i = g0; i < g1; i++) {
                uint32_t q = fast[i];
                std::vector<uint8_t> rc_pat(qlens[q]);
                const uint8_t *f = (const uint8_t *)queries[q];
                for (uint32_t j = 0; j < qlens[q]; j++) rc_pat[j] = comp_base(f[qlens[q] - 1 - j]);
                fq.fwd[i - g0] = encode_strand(f, qlens[q]);
                fq.rev[i - g0] = encode_strand(rc_pat.data(), qlens[q]);
                fq.len[i - g0] = (uint8_t)qlens[q];
            }
            job.nq = (uint32_t)(g1 - g0); job.q0 = fast[g0];
            job.work_counter = next_counter();
            launch_planes(0, job, fq, grid, ctx->stream);
            launches++;
            g0 = g1;
        }
    }
    for (size_t i = 0; i < slow.size(); i++) {
        uint32_t q = slow[i];
        SlowQuery sq{(const uint8_t *)ctx->qbytes.p + qoff[i], (const uint8_t *)ctx->qbytes.p + qoff[i] + qlens[q], qlens[q]};
        WindowJob j2 = job;
        j2.nq = 1; j2.q0 = q;
        j2.work_counter = next_counter();
        uint64_t wbk = (n_windows + kWarpsPerBlock - 1) / kWarpsPerBlock;
        int g2 = (int)(wbk < max_blocks ? wbk : max_blocks);
        window_counts_bytes_kernel<<<g2 < 1 ? 1 : g2, kWarpsPerBlock * 32, 0, ctx->stream>>>(j2, sq);
        launches++;
    }
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU(ctx, cudaGetLastError());

    uint32_t h_err = 0;
    CU(ctx, cudaMemcpyAsync(out_fwd, job.out_fwd, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out_rev, job.out_rev, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&h_err, job.err_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (kernel_ms) CU(ctx, cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
    if (kernel_launches) *kernel_launches = launches;
    if (h_err & 1u)
        return fail(ctx, TIDK_ENONASCII, "sequence contains a byte >= 0x80 (the reference fails in str::from_utf8, search.rs:107)");
    return TIDK_OK;
}

int tidk_cuda_window_counts(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                            uint32_t n_records, uint64_t window, const char *const *queries,
                            const uint32_t *qlens, uint32_t nq, uint64_t *out_fwd,
                            uint64_t *out_rev) {
    if (!ctx) return TIDK_EARG;
    if (window == 0) return fail(ctx, TIDK_EARG, "window must be > 0 (slice::chunks(0) panics, search.rs:98)");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records);
    if (rc) return rc;
    return tidk_cuda_window_counts_resident(ctx, &ctx->scratch, window, queries, qlens, nq, out_fwd,
                                            out_rev, nullptr, nullptr);
}

int tidk_cuda_explore_runs_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, double distance,
                                    uint32_t kmin, uint32_t kmax, uint64_t threshold,
                                    tidk_run **out_runs, uint64_t *n_runs, float *kernel_ms,
                                    uint32_t *kernel_launches) {
    if (!ctx) return TIDK_EARG;
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
                                 threshold, out_runs, n_runs, kernel_ms, kernel_launches, err);
    if (rc) return fail(ctx, rc, "%s", err.c_str());
    return TIDK_OK;
}

int tidk_cuda_explore_runs(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                           uint32_t n_records, double distance, uint32_t kmin, uint32_t kmax,
                           uint64_t threshold, tidk_run **out_runs, uint64_t *n_runs) {
    if (!ctx) return TIDK_EARG;
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = seqbuf_fill(ctx, &ctx->scratch, seq, offsets, n_records);
    if (rc) return rc;
    return tidk_cuda_explore_runs_resident(ctx, &ctx->scratch, distance, kmin, kmax, threshold,
                                           out_runs, n_runs, nullptr, nullptr);
}

void tidk_cuda_free_runs(tidk_run *runs) { free(runs); }

} // extern "C"
