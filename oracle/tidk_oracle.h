/*
 * tidk_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the hot path of tolkit/telomeric-identifier 0.2.7
 * (`tidk search`, `tidk find`, `tidk explore`).  The Rust reference cannot be
 * built in this image (no cargo/rustc, un-vendored crates), so this file is the
 * parity oracle and the timed CPU baseline ("port").  It is pinned against every
 * golden vector in the reference's own unit tests (tests/test_oracle_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libtidk_cuda.so and the
 * host driver) never links or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repo root).
 */
#ifndef TIDK_ORACLE_H
#define TIDK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIDK_ORACLE_OK 0
#define TIDK_ORACLE_EARG (-1)     /* reference would panic (window 0, empty motif, ...) */
#define TIDK_ORACLE_ENONASCII (-4) /* byte >= 0x80: reference fails in str::from_utf8 / unicode upper */
#define TIDK_ORACLE_ENOMEM (-3)

/* utils.rs:37-58  reverse_complement + switch_base (anything but A/C/G/T -> 'N'). */
void tidk_oracle_revcomp(const uint8_t *s, size_t n, uint8_t *out);

/* utils.rs:19-34  find_motifs: all (overlapping) occurrence offsets.
 * KMP restated from rust-bio 2.0.3 bio::pattern_matching::kmp for L < 65,
 * exact all-occurrence scan for L >= 65 (BOM; result-equivalent, parity unpinned).
 * Writes up to cap offsets to out_idx (may be NULL); returns the total count. */
size_t tidk_oracle_find_motifs(const uint8_t *motif, size_t L, const uint8_t *text, size_t n,
                               uint64_t *out_idx, size_t cap);

/* search.rs:83-151 / finder.rs:111-178  write_window_counts for ONE record and
 * ONE query.  `uppercase_query` != 0 follows search.rs:93-94 (query upper-cased
 * first), 0 follows finder.rs:129-136 (query verbatim).  out_fwd/out_rev hold
 * ceil(len/window) entries.  Returns the number of windows, or <0. */
int64_t tidk_oracle_window_counts(const uint8_t *seq, uint64_t len, uint64_t window,
                                  const uint8_t *query, uint32_t qlen, int uppercase_query,
                                  uint64_t *out_fwd, uint64_t *out_rev);

/* Batched form with the same layout as tidk_cuda_window_counts:
 * [record][query][window].  `threads` > 1 shards by record (information only;
 * the reference is single-threaded, search.rs:57 / finder.rs:85). */
int tidk_oracle_window_counts_batch(const uint8_t *seq, const uint64_t *offsets, uint32_t n_records,
                                    uint64_t window, const uint8_t *const *queries,
                                    const uint32_t *qlens, uint32_t nq, int uppercase_query,
                                    uint64_t *out_fwd, uint64_t *out_rev, int threads);

/* explore.rs:438-445 check_repeats: smallest period of s. */
size_t tidk_oracle_period(const uint8_t *s, size_t n);

/* utils.rs:87-94 string_rotation. */
int tidk_oracle_string_rotation(const uint8_t *s1, size_t n1, const uint8_t *s2, size_t n2);

/* utils.rs:147-166 lex_min (Booth-style minimal_rotation at utils.rs:100-127). out has n bytes. */
void tidk_oracle_lex_min(const uint8_t *s, size_t n, uint8_t *out);

/* explore.rs:151-160 split_seq_by_distance: dist = ceil(len * d). */
uint64_t tidk_oracle_split_dist(uint64_t len, double d);

/* explore.rs:178-233 chunk_fasta.  Emits entry positions (slice-relative) into
 * out_pos (cap entries; may be NULL to count).  The label of an entry at
 * position p is upper(slice[p'..p'+k]) where p' is the position of the first
 * chunk of its raw run; label_pos receives p'.  Returns the entry count or <0. */
int64_t tidk_oracle_chunk_fasta(const uint8_t *slice, uint64_t m, uint32_t k,
                                uint64_t *out_pos, uint64_t *out_label_pos, size_t cap);

/* One retained run == explore.rs:235-241 RepeatPosition after
 * filter_by_frequency, plus where it came from. */
typedef struct {
    uint32_t record;
    uint8_t slice; /* 0 = left, 1 = right */
    uint8_t k;
    uint8_t first_in_slice;
    uint8_t pad;
    uint64_t start;     /* slice-relative; 0 for the first run of a slice (explore.rs:297) */
    uint64_t end;       /* slice-relative */
    uint64_t label_pos; /* slice-relative offset of the chunk the label was taken from */
} tidk_oracle_run;

/* explore.rs:289-347 calculate_indexes on one slice (chunk_fasta included).
 * `apply_period_filter` = 1 reproduces filter_by_frequency exactly
 * (count > threshold && period >= 3, explore.rs:265-274,351-354); 0 keeps every
 * run with count > threshold (what the CUDA boundary returns).
 * Appends to *runs (realloc'd; *n_runs / *cap updated).  Returns 0 or <0. */
int tidk_oracle_slice_runs(const uint8_t *slice, uint64_t m, uint32_t k, uint64_t threshold,
                           int apply_period_filter, uint32_t record, uint8_t which,
                           tidk_oracle_run **runs, uint64_t *n_runs, uint64_t *cap);

/* explore.rs:50-126 driver loop in FASTA file order: for k in kmin..=kmax, for
 * each record, left slice then right slice.  Caller frees *runs with free(). */
int tidk_oracle_explore_runs(const uint8_t *seq, const uint64_t *offsets, uint32_t n_records,
                             double distance, uint32_t kmin, uint32_t kmax, uint64_t threshold,
                             int apply_period_filter, int threads,
                             tidk_oracle_run **runs, uint64_t *n_runs);

/* explore.rs:377-431 get_telomeric_repeat_estimates + filter_count_vec
 * (explore.rs:455-464) on n labelled runs given in arrival order.
 * labels: n strings of length klen[i] at label_ptr[i]; counts[i] = get_count().
 * `literal` = 1 runs the reference's own pairwise loop (explore.rs:395-423);
 * 0 runs the O(n) closed form (valid for A/C/G/T/N labels).
 * Output rows sorted by (count desc, key asc); keys are written back to back
 * into out_keys (each out_klen[j] bytes).  Returns the row count or <0. */
int64_t tidk_oracle_estimates(const uint8_t *const *label_ptr, const uint32_t *klen,
                              const int64_t *counts, uint64_t n, int literal,
                              uint8_t *out_keys, size_t out_keys_cap, uint32_t *out_klen,
                              int64_t *out_counts, size_t out_cap);

#ifdef __cplusplus
}
#endif
#endif
