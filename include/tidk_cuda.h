/*
 * tidk_cuda.h -- C ABI of libtidk_cuda.so, the B200 (sm_100a) implementation of
 * the data-parallel hot path of tolkit/telomeric-identifier 0.2.7.
 *
 * The reference has no FFI or plugin interface (pure Rust, single crate).  The
 * seam is the set of private per-record functions its three subcommand drivers
 * call; each entry point below names the reference call site it replaces
 * (paths relative to the reference repo root).  INTEGRATION.md shows the Rust
 * `extern "C"` block and the patched call sites a maintainer would add in a
 * `tidk-cuda` crate.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary
 *   - every function returns 0 on success or a negative TIDK_E* code; the
 *     message for the last failure on a ctx is tidk_cuda_last_error(ctx)
 *   - nothing is printed by the library (the host owns all text, as in
 *     search.rs:22,71,73); no exception or abort crosses the boundary
 *   - a ctx is single-caller; inputs are borrowed for the duration of a call
 *   - there is NO CPU fallback: without a CUDA device tidk_cuda_create fails
 */
#ifndef TIDK_CUDA_H
#define TIDK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIDK_OK 0
#define TIDK_EARG (-1)      /* reference would panic: window == 0, empty query, distance > 0.5 ... */
#define TIDK_ECUDA (-2)     /* CUDA runtime / launch failure, or no device */
#define TIDK_ENOMEM (-3)    /* host or device allocation failed */
#define TIDK_ENONASCII (-4) /* sequence byte >= 0x80: the reference errors in str::from_utf8
                               (search.rs:107, finder.rs:146) or changes length in to_uppercase.
                               explore: STRICTER than the reference ON PURPOSE -- explore.rs:212 only
                               panics when such a byte sits in a chunk that equals its neighbour; the
                               library rejects it anywhere inside a scanned slice (a byte >= 0x80 is never
                               valid FASTA sequence, and one rule per slice keeps the scan free of a
                               per-chunk check) */

typedef struct tidk_ctx tidk_ctx;
typedef struct tidk_seqbuf tidk_seqbuf;

/* ABI version of this header (bumped on any signature change). */
uint32_t tidk_cuda_abi_version(void);

/* Create a context on the given CUDA devices (device_ids == NULL or n_devices == 0: device 0).
 * n_devices > 1: ONE context that fans every call out to all listed devices (one worker thread per device inside the
 * call, no data-path collective): host-pointer window counts are cut into equal contiguous window ranges (long records at
 * multiples of the window, search.rs:98,105-110), explore and resident batches into contiguous groups of whole records;
 * counts land in the usual [record][query][window] arrays, runs come back in the reference's arrival order.  This is what
 * a single-process host -- the reference's drivers are one sequential loop, search.rs:57 -- binds to use a whole box.
 * The same id may be listed twice (two contexts on one device).  One process per GPU (bench.py) remains possible. */
int tidk_cuda_create(const int *device_ids, int n_devices, tidk_ctx **out);
void tidk_cuda_destroy(tidk_ctx *ctx);
/* ctx-owned string, valid until the next call on ctx.  ctx == NULL returns the
 * message of the last failed tidk_cuda_create on this thread. */
const char *tidk_cuda_last_error(const tidk_ctx *ctx);

/* Pinned host memory so that the FASTA reader (lib.rs:32-49 stays on the host)
 * can parse straight into DMA-able buffers. */
void *tidk_cuda_host_alloc(tidk_ctx *ctx, size_t bytes);
void tidk_cuda_host_free(tidk_ctx *ctx, void *p);

/* ---------------------------------------------------------------- search / find
 * Replaces search::write_window_counts (search.rs:83-151, called at search.rs:62)
 * and finder::write_window_counts (finder.rs:111-178, called at finder.rs:90),
 * i.e. utils::find_motifs x2 per window (utils.rs:19-34) with the window rules
 * of search.rs:98,105-130:
 *
 *   window i of a record = seq[i*W .. min((i+1)*W, len))
 *   fwd[i] = #{p : window-local, upper(seq[p+j]) == Q[j] for all j < L}   (overlapping
 *            occurrences all count: utils::remove_overlapping_indexes is the identity)
 *   rev[i] = same for revcomp(Q)  (utils.rs:37-58: A<->T, C<->G, anything else -> 'N')
 *   an occurrence that straddles a window boundary is counted in neither window
 *
 * seq      concatenated record sequences (bio Record::seq(): ASCII, newline-free)
 * offsets  n_records+1 byte offsets into seq
 * queries  nq byte strings, used VERBATIM as the forward pattern: a `search`
 *          host passes query.to_uppercase() (search.rs:93), a `find` host passes
 *          the clade repeat unchanged (finder.rs:129-136)
 * out_fwd / out_rev  caller-allocated, layout [record][query][window] with
 *          ceil(len_r / W) windows per record -- the reference's row order
 *          record -> repeat -> window (finder.rs:121,144)
 */
int tidk_cuda_window_counts(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                            uint32_t n_records, uint64_t window, const char *const *queries,
                            const uint32_t *qlens, uint32_t nq, uint64_t *out_fwd,
                            uint64_t *out_rev);

/* Bytes that crossed PCIe in the last tidk_cuda_window_counts call on ctx (or the last tidk_cuda_explore_runs call
 * that uploaded the whole batch: there the planes are STRICT -- upper-case A C G T only -- words with anything else
 * travel as raw bytes, and the device expands the planes back to ASCII before the scan), and whether the sequence
 * went up (at least in part) as host-packed bit-planes -- 2 bits per base plus validity exceptions for
 * the chunks packed from the back, ASCII by DMA for the chunks taken from the front; taken when every
 * query has a compile-time plane kernel, the batch is >= 32 MiB and >= 12 host threads are available
 * (TIDK_HOST_PACK=0/1, TIDK_HYBRID_UPLOAD=0 and TIDK_PACK_THREADS override) -- or entirely as ASCII. */
int tidk_cuda_last_transfer(const tidk_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int *host_packed);

/* Device time (CUDA events around the counting kernels, ms) and number of kernel launches of the last
 * window-count call on ctx, host-pointer or resident (diagnostics: the drivers' --gpu-stats line). */
int tidk_cuda_last_kernel_time(const tidk_ctx *ctx, float *kernel_ms, uint32_t *kernel_launches);

/* Same computation on a sequence batch that is already resident in HBM
 * (uploaded once, queried many times: several clades, tsv + bedgraph, ...).
 * kernel_ms (nullable) receives the device time of the counting kernels of this
 * call, measured with CUDA events on the ctx stream; kernel_launches (nullable)
 * the number of kernels launched. */
int tidk_cuda_seqbuf_upload(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                            uint32_t n_records, tidk_seqbuf **out);
void tidk_cuda_seqbuf_free(tidk_ctx *ctx, tidk_seqbuf *buf);
/* A batch uploaded with tidk_cuda_seqbuf_upload is transposed once on the device, behind the copy, into the
 * bit-planes the scans compute on (lo = ASCII bit 1, hi = ASCII bit 2, validity, one "32 upper-case A C G T bytes"
 * bit per word; 0.38 B per base next to the ASCII copy).  Every resident call whose queries have plane kernels, and
 * the explore scan, then read 0.25 - 0.375 B per base instead of 1.  ready: planes exist (TIDK_PLANES=0 turns the
 * conversion off); build_ms: device time of the conversion; bytes: device memory they take. */
int tidk_cuda_seqbuf_planes_info(const tidk_seqbuf *buf, int *ready, float *build_ms, uint64_t *bytes);
int tidk_cuda_window_counts_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, uint64_t window,
                                     const char *const *queries, const uint32_t *qlens,
                                     uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev,
                                     float *kernel_ms, uint32_t *kernel_launches);

/* Asynchronous form of tidk_cuda_window_counts_resident: the kernels are enqueued and the call returns; the counts
 * leave the device on a second stream while the next call's kernels run (three device output buffers are used in turn:
 * a call first waits for the counts of the call two before it to have left the device; its kernel also zeroes the
 * counters of the call after it).  The caller needs two sets of output arrays, used alternately.  out_fwd / out_rev must stay valid -- and
 * should be tidk_cuda_host_alloc memory, or the copy is staged -- until tidk_cuda_wait returns.  A single-threaded
 * host (search.rs:57 is one sequential loop) keeps the device busy this way: the per-call cost of a blocking call
 * (copy of the counts + wake-up of the waiting thread, 40-80 us) is as long as a whole launch on a sharded genome. */
int tidk_cuda_window_counts_resident_async(tidk_ctx *ctx, const tidk_seqbuf *buf, uint64_t window,
                                           const char *const *queries, const uint32_t *qlens,
                                           uint32_t nq, uint64_t *out_fwd, uint64_t *out_rev);
/* Blocks until every asynchronous call on ctx has delivered its counts.  kernel_ms / kernel_launches (nullable):
 * device time of the counting kernels (CUDA events on the ctx stream) and launches summed over the calls collected
 * since the previous wait.  Returns the first deferred error (TIDK_ENONASCII: some call saw a byte >= 0x80). */
int tidk_cuda_wait(tidk_ctx *ctx, float *kernel_ms, uint32_t *kernel_launches);

/* ---------------------------------------------------------------------- explore
 * Replaces split_seq_by_distance + chunk_fasta + calculate_indexes for every
 * record, slice and k (explore.rs:151-160, 178-233, 289-347; call sites
 * explore.rs:65-74 and 105-120).  Returns every run with
 * (end - start) / k > threshold, `start` already carrying the reference's
 * first-run rule (start = 0 for the first run of a slice, explore.rs:297),
 * ordered by (k, record, slice, start) = FASTA-file arrival order.
 * The period filter (explore.rs:351-354), canonical key (utils.rs:147-166) and
 * estimator (explore.rs:377-431) stay on the host.
 */
typedef struct {
    uint32_t record;
    uint8_t slice;          /* 0 = left [0, dist), 1 = right [len - dist, len) */
    uint8_t k;
    uint8_t first_in_slice; /* 1: start was forced to 0 by explore.rs:297 */
    uint8_t pad;
    uint64_t start;         /* slice-relative */
    uint64_t end;           /* slice-relative, (last chunk + 1) * k */
    uint64_t label_pos;     /* slice-relative offset of a chunk of the run; label = upper(bytes) */
} tidk_run;

int tidk_cuda_explore_runs(tidk_ctx *ctx, const uint8_t *seq, const uint64_t *offsets,
                           uint32_t n_records, double distance, uint32_t kmin, uint32_t kmax,
                           uint64_t threshold, tidk_run **out_runs, uint64_t *n_runs);
int tidk_cuda_explore_runs_resident(tidk_ctx *ctx, const tidk_seqbuf *buf, double distance,
                                    uint32_t kmin, uint32_t kmax, uint64_t threshold,
                                    tidk_run **out_runs, uint64_t *n_runs, float *kernel_ms,
                                    uint32_t *kernel_launches);
void tidk_cuda_free_runs(tidk_run *runs);

/* Time breakdown of the last explore call on ctx, in ms.  Device time (CUDA events on the ctx stream):
 * out[0] = streaming scan kernel(s), out[1] = event sort + run building + compaction.  Host wall clock:
 * out[2] = slice table + uploads before the scan, out[3] = post-processing including the copy of the runs.
 * Writes min(n, 4) values and returns how many were written. */
int tidk_cuda_last_explore_timings(const tidk_ctx *ctx, float *out, int n);

#ifdef __cplusplus
}
#endif
#endif
