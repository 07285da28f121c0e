/*
 * tidk_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See tidk_oracle.h.  Plain C11 + pthreads, no other dependencies.
 *
 * Restates, function by function, the hot path of tolkit/telomeric-identifier
 * 0.2.7.  Citations are reference file:line.  Deliberately written the way the
 * reference computes (per-window upper-case copy, per-window KMP table build,
 * byte-serial scan) so that its timing is a fair stand-in for the reference's
 * CPU path, not an optimised CPU implementation.
 */
#include "tidk_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ utils */

/* utils.rs:49-58 switch_base */
static inline uint8_t switch_base(uint8_t c) {
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'T': return 'A';
    case 'G': return 'C';
    default: return 'N'; /* 'N' => 'N', _ => 'N' */
    }
}

/* utils.rs:37-46 reverse_complement */
void tidk_oracle_revcomp(const uint8_t *s, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[n - 1 - i] = switch_base(s[i]);
}

static inline uint8_t ascii_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* rust-bio 2.0.3 bio::pattern_matching::kmp (Cargo.lock:148-149), as used at
 * utils.rs:23-24: lps table, then an automaton step per text byte; a full
 * match falls back through lps so overlapping occurrences are reported. */
static void kmp_lps(const uint8_t *p, size_t m, size_t *lps) {
    size_t q = 0;
    lps[0] = 0;
    for (size_t i = 1; i < m; i++) {
        while (q > 0 && p[q] != p[i]) q = lps[q - 1];
        if (p[q] == p[i]) q++;
        lps[i] = q;
    }
}

static size_t kmp_find_all(const uint8_t *p, size_t m, const uint8_t *t, size_t n, uint64_t *out,
                           size_t cap) {
    size_t stack_lps[64];
    size_t *lps = stack_lps; /* m < 65 on this path */
    kmp_lps(p, m, lps);
    size_t q = 0, cnt = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t a = t[i];
        while (q == m || (p[q] != a && q > 0)) q = lps[q - 1];
        if (p[q] == a) q++;
        if (q == m) {
            if (out && cnt < cap) out[cnt] = (uint64_t)(i + 1 - m);
            cnt++;
        }
    }
    return cnt;
}

/* utils.rs:25-28: BOM for motifs >= 65.  Any exact all-occurrence matcher gives
 * the same index list; parity for this branch is unpinned by the reference. */
static size_t exact_find_all(const uint8_t *p, size_t m, const uint8_t *t, size_t n, uint64_t *out,
                             size_t cap) {
    size_t cnt = 0;
    if (n < m) return 0;
    for (size_t i = 0; i + m <= n; i++) {
        if (t[i] == p[0] && memcmp(t + i, p, m) == 0) {
            if (out && cnt < cap) out[cnt] = (uint64_t)i;
            cnt++;
        }
    }
    return cnt;
}

/* utils.rs:19-34 find_motifs */
size_t tidk_oracle_find_motifs(const uint8_t *motif, size_t L, const uint8_t *text, size_t n,
                               uint64_t *out_idx, size_t cap) {
    if (L == 0) return 0; /* reference panics; callers reject earlier */
    if (L < 65) return kmp_find_all(motif, L, text, n, out_idx, cap);
    return exact_find_all(motif, L, text, n, out_idx, cap);
}

/* ---------------------------------------------------------- search / find */

/* search.rs:83-151 and finder.rs:111-178 (one query).
 * utils::remove_overlapping_indexes (utils.rs:63-80) is the identity: its loop
 * breaks on the first test because index == 0 (utils.rs:71), so the count is
 * the number of (overlapping) occurrences. */
int64_t tidk_oracle_window_counts(const uint8_t *seq, uint64_t len, uint64_t window,
                                  const uint8_t *query, uint32_t qlen, int uppercase_query,
                                  uint64_t *out_fwd, uint64_t *out_rev) {
    if (window == 0 || qlen == 0) return TIDK_ORACLE_EARG; /* chunks(0) / KMP::new("") panic */
    uint8_t *fwd = (uint8_t *)malloc(qlen), *rev = (uint8_t *)malloc(qlen);
    uint8_t *up = (uint8_t *)malloc(window < len ? window : (len ? len : 1));
    if (!fwd || !rev || !up) { free(fwd); free(rev); free(up); return TIDK_ORACLE_ENOMEM; }
    for (uint32_t j = 0; j < qlen; j++) {
        if (query[j] >= 0x80) { free(fwd); free(rev); free(up); return TIDK_ORACLE_ENONASCII; }
        fwd[j] = uppercase_query ? ascii_upper(query[j]) : query[j]; /* search.rs:93 vs finder.rs:129 */
    }
    tidk_oracle_revcomp(fwd, qlen, rev); /* search.rs:94 / finder.rs:135 */
    int64_t nwin = 0;
    for (uint64_t beg = 0; beg < len; beg += window) { /* seq.chunks(window) */
        uint64_t wl = (len - beg < window) ? (len - beg) : window;
        for (uint64_t j = 0; j < wl; j++) { /* search.rs:107 from_utf8(..)?.to_uppercase() */
            uint8_t c = seq[beg + j];
            if (c >= 0x80) { free(fwd); free(rev); free(up); return TIDK_ORACLE_ENONASCII; }
            up[j] = ascii_upper(c);
        }
        out_fwd[nwin] = tidk_oracle_find_motifs(fwd, qlen, up, wl, NULL, 0); /* search.rs:109 */
        out_rev[nwin] = tidk_oracle_find_motifs(rev, qlen, up, wl, NULL, 0); /* search.rs:110 */
        nwin++;
    }
    free(fwd); free(rev); free(up);
    return nwin;
}

typedef struct {
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_records;
    uint64_t window;
    const uint8_t *const *queries;
    const uint32_t *qlens;
    uint32_t nq;
    int upper;
    uint64_t *out_fwd, *out_rev;
    const uint64_t *out_base; /* per record: first output index */
    int tid, nthreads;
    int rc;
} wc_job;

static void *wc_worker(void *arg) {
    wc_job *j = (wc_job *)arg;
    for (uint32_t r = (uint32_t)j->tid; r < j->n_records; r += (uint32_t)j->nthreads) {
        uint64_t len = j->offsets[r + 1] - j->offsets[r];
        uint64_t nwin = (len + j->window - 1) / j->window;
        for (uint32_t q = 0; q < j->nq; q++) { /* finder.rs:121-176: repeat outer, window inner */
            uint64_t o = j->out_base[r] + (uint64_t)q * nwin;
            int64_t rc = tidk_oracle_window_counts(j->seq + j->offsets[r], len, j->window,
                                                   j->queries[q], j->qlens[q], j->upper,
                                                   j->out_fwd + o, j->out_rev + o);
            if (rc < 0) { j->rc = (int)rc; return NULL; }
        }
    }
    return NULL;
}

int tidk_oracle_window_counts_batch(const uint8_t *seq, const uint64_t *offsets, uint32_t n_records,
                                    uint64_t window, const uint8_t *const *queries,
                                    const uint32_t *qlens, uint32_t nq, int uppercase_query,
                                    uint64_t *out_fwd, uint64_t *out_rev, int threads) {
    if (window == 0) return TIDK_ORACLE_EARG;
    for (uint32_t q = 0; q < nq; q++) if (qlens[q] == 0) return TIDK_ORACLE_EARG;
    uint64_t *base = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_records + 1));
    if (!base) return TIDK_ORACLE_ENOMEM;
    uint64_t acc = 0;
    for (uint32_t r = 0; r < n_records; r++) {
        base[r] = acc;
        uint64_t len = offsets[r + 1] - offsets[r];
        acc += ((len + window - 1) / window) * nq;
    }
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n_records && n_records > 0) threads = (int)n_records;
    wc_job *jobs = (wc_job *)calloc((size_t)threads, sizeof(wc_job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        jobs[t] = (wc_job){seq, offsets, n_records, window, queries, qlens, nq, uppercase_query,
                           out_fwd, out_rev, base, t, threads, 0};
        if (threads == 1) wc_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, wc_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs); free(th); free(base);
    return rc;
}

/* --------------------------------------------------------------- explore */

/* explore.rs:438-445 check_repeats: the BTreeMap of delayed iterators keeps
 * delay d alive iff s[i] == s[i-d] for every i >= d; the smallest surviving key
 * is the shortest period (n itself always survives). */
size_t tidk_oracle_period(const uint8_t *s, size_t n) {
    for (size_t d = 1; d < n; d++) {
        size_t j = 0;
        while (j + d < n && s[j] == s[j + d]) j++;
        if (j + d == n) return d;
    }
    return n;
}

/* utils.rs:87-94 string_rotation: s2+s2 contains s1 (equal lengths). */
int tidk_oracle_string_rotation(const uint8_t *s1, size_t n1, const uint8_t *s2, size_t n2) {
    if (n1 != n2) return 0;
    if (n1 == 0) return 1; /* "".contains("") */
    for (size_t r = 0; r < n2; r++) { /* start offsets 0..n2 in s2s2; offset n2 == offset 0 */
        size_t j = 0;
        while (j < n1 && s2[(r + j) % n2] == s1[j]) j++;
        if (j == n1) return 1;
    }
    return 0;
}

/* utils.rs:100-127 minimal_rotation, restated literally. */
static size_t minimal_rotation(const uint8_t *s, size_t n) {
    size_t i = 0, j = 1;
    for (;;) {
        size_t k = 0;
        uint8_t ci = s[i % n], cj = s[j % n];
        while (k < n) {
            ci = s[(i + k) % n];
            cj = s[(j + k) % n];
            if (ci != cj) break;
            k++;
        }
        if (k == n) return i < j ? i : j;
        if (ci > cj) { i += k + 1; i += (i == j); }
        else { j += k + 1; j += (i == j); }
    }
}

/* utils.rs:147-166 lex_min; natural_lexical_cmp on equal-length upper-case DNA
 * strings is plain byte order (lexical-sort 0.3.1; pinned by utils.rs:219-227). */
void tidk_oracle_lex_min(const uint8_t *s, size_t n, uint8_t *out) {
    if (n == 0) return;
    uint8_t stackbuf[3 * 64];
    uint8_t *buf = n <= 64 ? stackbuf : (uint8_t *)malloc(3 * n);
    uint8_t *rc = buf, *f = buf + n, *r = buf + 2 * n;
    tidk_oracle_revcomp(s, n, rc);
    size_t jf = minimal_rotation(s, n), jr = minimal_rotation(rc, n);
    for (size_t t = 0; t < n; t++) { f[t] = s[(jf + t) % n]; r[t] = rc[(jr + t) % n]; }
    memcpy(out, memcmp(f, r, n) <= 0 ? f : r, n);
    if (buf != stackbuf) free(buf);
}

/* explore.rs:156 */
uint64_t tidk_oracle_split_dist(uint64_t len, double d) {
    double v = ceil((double)len * d);
    if (!(v > 0.0)) return 0; /* `as usize` saturates negatives / NaN to 0 */
    return (uint64_t)v;
}

/* explore.rs:178-233 chunk_fasta */
int64_t tidk_oracle_chunk_fasta(const uint8_t *slice, uint64_t m, uint32_t k, uint64_t *out_pos,
                                uint64_t *out_label_pos, size_t cap) {
    if (k == 0) return TIDK_ORACLE_EARG; /* chunks(0) panics */
    if (m <= k) return 0;                /* explore.rs:187-195 */
    uint64_t nchunks = (m + k - 1) / k;
    uint64_t pos = 0;
    int is_first = 1;
    int64_t cnt = 0;
#define PUSH(P, LP)                                                            \
    do {                                                                       \
        if (out_pos && (size_t)cnt < cap) { out_pos[cnt] = (P); out_label_pos[cnt] = (LP); } \
        cnt++;                                                                 \
    } while (0)
    for (uint64_t i = 0; i + 1 < nchunks; i++) { /* chunks.zip(chunks.skip(1)) */
        const uint8_t *a = slice + i * k;
        const uint8_t *b = slice + (i + 1) * k;
        uint64_t blen = (m - (i + 1) * k < k) ? (m - (i + 1) * k) : k;
        int eq = (blen == k) && memcmp(a, b, k) == 0; /* slice equality incl. length */
        if (eq) {
            for (uint32_t t = 0; t < k; t++) if (a[t] >= 0x80) return TIDK_ORACLE_ENONASCII;
            if (is_first) {
                PUSH(pos, i * k);
                pos += k;
                PUSH(pos, i * k);
                is_first = 0;
            } else {
                pos += k;
                PUSH(pos, i * k);
            }
        } else {
            pos += k;
            is_first = 1;
        }
    }
#undef PUSH
    return cnt;
}

static int push_run(tidk_oracle_run **runs, uint64_t *n, uint64_t *cap, tidk_oracle_run r) {
    if (*n == *cap) {
        uint64_t nc = *cap ? *cap * 2 : 64;
        tidk_oracle_run *p = (tidk_oracle_run *)realloc(*runs, nc * sizeof(tidk_oracle_run));
        if (!p) return TIDK_ORACLE_ENOMEM;
        *runs = p; *cap = nc;
    }
    (*runs)[(*n)++] = r;
    return 0;
}

static int label_eq_upper(const uint8_t *s, uint64_t p1, uint64_t p2, uint32_t k) {
    for (uint32_t t = 0; t < k; t++)
        if (ascii_upper(s[p1 + t]) != ascii_upper(s[p2 + t])) return 0;
    return 1;
}

/* explore.rs:289-347 calculate_indexes (+ chunk_fasta, + filter_by_frequency) */
int tidk_oracle_slice_runs(const uint8_t *slice, uint64_t m, uint32_t k, uint64_t threshold,
                           int apply_period_filter, uint32_t record, uint8_t which,
                           tidk_oracle_run **runs, uint64_t *n_runs, uint64_t *cap) {
    int64_t ne = tidk_oracle_chunk_fasta(slice, m, k, NULL, NULL, 0);
    if (ne < 0) return (int)ne;
    if (ne == 0) return 0; /* explore.rs:336-342 -> None */
    uint64_t *pos = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ne * 2);
    if (!pos) return TIDK_ORACLE_ENOMEM;
    uint64_t *lab = pos + ne;
    tidk_oracle_chunk_fasta(slice, m, k, pos, lab, (size_t)ne);

    uint64_t start = 0; /* explore.rs:297 -- never re-initialised to the first run's position */
    int first = 1, rc = 0;
    uint8_t up[256];
    for (int64_t i = 0; i + 1 < ne && rc == 0; i++) { /* indexes.iter().zip(skip(1)).peekable() */
        uint64_t p1 = pos[i], p2 = pos[i + 1];
        int is_last = (i + 2 == ne); /* iter.peek().is_none() */
        int same = label_eq_upper(slice, lab[i], lab[i + 1], k) && (p2 - p1) == k; /* :324 */
        uint64_t end;
        if (is_last) end = p2 + k;          /* explore.rs:315-323 */
        else if (same) continue;            /* :324-325 */
        else end = p1 + k;                  /* :326-332 */
        /* filter_by_frequency, explore.rs:265-274 */
        uint64_t count = (end - start) / k;
        int keep = count > threshold;
        if (keep && apply_period_filter) {
            const uint8_t *l = slice + lab[i];
            size_t per;
            if (k <= sizeof up) {
                for (uint32_t t = 0; t < k; t++) up[t] = ascii_upper(l[t]);
                per = tidk_oracle_period(up, k);
            } else {
                uint8_t *tmp = (uint8_t *)malloc(k);
                for (uint32_t t = 0; t < k; t++) tmp[t] = ascii_upper(l[t]);
                per = tidk_oracle_period(tmp, k);
                free(tmp);
            }
            keep = per >= 3; /* !(period < REPEAT_PERIOD_THRESHOLD), explore.rs:16,351-354 */
        }
        if (keep) {
            tidk_oracle_run r = {record, which, (uint8_t)k, (uint8_t)first, 0, start, end, lab[i]};
            rc = push_run(runs, n_runs, cap, r);
        }
        first = 0;
        if (!is_last) start = p2; /* explore.rs:333 */
    }
    free(pos);
    return rc;
}

typedef struct {
    const uint8_t *seq;
    const uint64_t *offsets;
    uint32_t n_records;
    double distance;
    uint32_t k;
    uint64_t threshold;
    int period_filter;
    int tid, nthreads;
    tidk_oracle_run *runs;
    uint64_t n, cap;
    int rc;
} ex_job;

static void *ex_worker(void *arg) {
    ex_job *j = (ex_job *)arg;
    /* contiguous record blocks per thread so that concatenating thread outputs
     * in thread order reproduces FASTA file order */
    uint32_t per = (j->n_records + (uint32_t)j->nthreads - 1) / (uint32_t)j->nthreads;
    uint32_t r0 = (uint32_t)j->tid * per;
    uint32_t r1 = r0 + per < j->n_records ? r0 + per : j->n_records;
    for (uint32_t r = r0; r < r1 && j->rc == 0; r++) {
        const uint8_t *s = j->seq + j->offsets[r];
        uint64_t len = j->offsets[r + 1] - j->offsets[r];
        uint64_t dist = tidk_oracle_split_dist(len, j->distance); /* explore.rs:156 */
        if (dist > len) { j->rc = TIDK_ORACLE_EARG; break; }     /* slice index panic */
        /* explore.rs:157-158: [0..dist) then [len-dist..) ; :108 left before right */
        j->rc = tidk_oracle_slice_runs(s, dist, j->k, j->threshold, j->period_filter, r, 0,
                                       &j->runs, &j->n, &j->cap);
        if (j->rc) break;
        j->rc = tidk_oracle_slice_runs(s + (len - dist), dist, j->k, j->threshold,
                                       j->period_filter, r, 1, &j->runs, &j->n, &j->cap);
    }
    return NULL;
}

int tidk_oracle_explore_runs(const uint8_t *seq, const uint64_t *offsets, uint32_t n_records,
                             double distance, uint32_t kmin, uint32_t kmax, uint64_t threshold,
                             int apply_period_filter, int threads,
                             tidk_oracle_run **runs, uint64_t *n_runs) {
    *runs = NULL; *n_runs = 0;
    if (kmin == 0 || kmax < kmin || kmax > 255) return TIDK_ORACLE_EARG;
    if (distance > 0.5) return TIDK_ORACLE_EARG; /* explore.rs:41-43 bail */
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n_records && n_records > 0) threads = (int)n_records;
    uint64_t cap = 0;
    int rc = 0;
    ex_job *jobs = (ex_job *)calloc((size_t)threads, sizeof(ex_job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (uint32_t k = kmin; k <= kmax && rc == 0; k++) { /* explore.rs:88 */
        for (int t = 0; t < threads; t++) {
            jobs[t] = (ex_job){seq, offsets, n_records, distance, k, threshold, apply_period_filter,
                               t, threads, NULL, 0, 0, 0};
            if (threads == 1) ex_worker(&jobs[t]);
            else pthread_create(&th[t], NULL, ex_worker, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) {
            if (threads > 1) pthread_join(th[t], NULL);
            if (jobs[t].rc) rc = jobs[t].rc;
            for (uint64_t i = 0; i < jobs[t].n && rc == 0; i++)
                rc = push_run(runs, n_runs, &cap, jobs[t].runs[i]);
            free(jobs[t].runs);
        }
    }
    free(jobs); free(th);
    if (rc) { free(*runs); *runs = NULL; *n_runs = 0; }
    return rc;
}

/* ------------------------------------------------------------- estimator */

typedef struct {
    uint32_t klen;
    uint8_t *key; /* owned */
    int64_t count;
} est_row;

typedef struct {
    est_row *rows;
    size_t n, cap;
} est_map;

static est_row *map_find(est_map *m, const uint8_t *key, uint32_t klen) {
    for (size_t i = 0; i < m->n; i++)
        if (m->rows[i].klen == klen && memcmp(m->rows[i].key, key, klen) == 0) return &m->rows[i];
    return NULL;
}

static est_row *map_insert_new(est_map *m, const uint8_t *key, uint32_t klen, int64_t count) {
    if (m->n == m->cap) {
        size_t nc = m->cap ? m->cap * 2 : 16;
        m->rows = (est_row *)realloc(m->rows, nc * sizeof(est_row));
        m->cap = nc;
    }
    est_row *r = &m->rows[m->n++];
    r->klen = klen;
    r->key = (uint8_t *)malloc(klen ? klen : 1);
    memcpy(r->key, key, klen);
    r->count = count;
    return r;
}

/* explore.rs:359-371 test_repeats */
static int same_class(const uint8_t *a, const uint8_t *b, uint32_t k, uint8_t *tmp) {
    if (tidk_oracle_string_rotation(a, k, b, k)) return 1;
    tidk_oracle_revcomp(a, k, tmp);
    if (tidk_oracle_string_rotation(tmp, k, b, k)) return 1;
    tidk_oracle_revcomp(b, k, tmp);
    return tidk_oracle_string_rotation(a, k, tmp, k);
}

static int row_cmp(const void *pa, const void *pb) {
    const est_row *a = (const est_row *)pa, *b = (const est_row *)pb;
    if (a->count != b->count) return a->count > b->count ? -1 : 1; /* explore.rs:427 desc */
    uint32_t n = a->klen < b->klen ? a->klen : b->klen;          /* ties: HashMap order in the */
    int c = memcmp(a->key, b->key, n);                             /* reference; fixed here      */
    if (c) return c;
    return (a->klen > b->klen) - (a->klen < b->klen);
}

typedef struct {
    uint32_t klen;
    const uint8_t *key; /* canonical key */
    uint64_t gidx;      /* index inside its length group, arrival order */
    int64_t count;
} fast_item;

static int fast_cmp(const void *pa, const void *pb) {
    const fast_item *a = (const fast_item *)pa, *b = (const fast_item *)pb;
    if (a->klen != b->klen) return a->klen < b->klen ? -1 : 1;
    int c = memcmp(a->key, b->key, a->klen);
    if (c) return c;
    return (a->gidx > b->gidx) - (a->gidx < b->gidx);
}

int64_t tidk_oracle_estimates(const uint8_t *const *label_ptr, const uint32_t *klen,
                              const int64_t *counts, uint64_t n, int literal, uint8_t *out_keys,
                              size_t out_keys_cap, uint32_t *out_klen, int64_t *out_counts,
                              size_t out_cap) {
    est_map map = {NULL, 0, 0};
    uint32_t maxk = 0;
    for (uint64_t i = 0; i < n; i++) if (klen[i] > maxk) maxk = klen[i];
    uint8_t *tmp = (uint8_t *)malloc((size_t)maxk * 2 + 1);
    uint8_t *key = tmp + maxk;

    if (literal) {
        /* make_length_groups (explore.rs:277-284): arrival order inside a group */
        uint8_t *done = (uint8_t *)calloc(n ? n : 1, 1);
        uint64_t *g = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
        for (uint64_t i0 = 0; i0 < n; i0++) {
            if (done[i0]) continue;
            uint32_t k = klen[i0];
            uint64_t gn = 0;
            for (uint64_t i = i0; i < n; i++) if (!done[i] && klen[i] == k) { g[gn++] = i; done[i] = 1; }
            if (gn == 1) { /* explore.rs:388-393: map.insert overwrites */
                tidk_oracle_lex_min(label_ptr[g[0]], k, key);
                est_row *r = map_find(&map, key, k);
                if (r) r->count = (int32_t)counts[g[0]];
                else map_insert_new(&map, key, k, (int32_t)counts[g[0]]);
                continue;
            }
            uint8_t *tracker = (uint8_t *)calloc(gn, 1); /* Vec<usize> + contains (:395,:417) */
            for (uint64_t a = 0; a < gn; a++) {
                for (uint64_t b = a + 1; b < gn; b++) { /* combinations(2), lexicographic (:397) */
                    const uint8_t *s1 = label_ptr[g[a]], *s2 = label_ptr[g[b]];
                    if (same_class(s1, s2, k, tmp)) {
                        tidk_oracle_lex_min(s1, k, key); /* utils::lms -> lex_min(r1), utils.rs:140 */
                        int32_t add = (int32_t)counts[g[a]] + (int32_t)counts[g[b]];
                        est_row *r = map_find(&map, key, k);
                        if (!r) r = map_insert_new(&map, key, k, add); /* or_insert (:412-416) */
                        if (!tracker[a] && !tracker[b]) r->count = (int32_t)(r->count + add); /* :417-419 */
                    }
                    tracker[a] = 1; tracker[b] = 1; /* :421-422 */
                }
            }
            free(tracker);
        }
        free(done); free(g);
    } else {
        /* Closed form of the loop above for labels on which revcomp is an
         * involution (A/C/G/T/N): per class the first two members (m1 < m2 in
         * group arrival order) give c[m1] + c[m2], doubled iff (m1, m2) == (0, 1);
         * classes with one member in a group of >= 2 vanish; a group of one
         * inserts its own count. */
        fast_item *it = (fast_item *)malloc(sizeof(fast_item) * (n ? n : 1));
        uint8_t *keys = (uint8_t *)malloc((size_t)maxk * (n ? n : 1) + 1);
        /* group sizes / indices by klen */
        uint64_t gcount[256]; memset(gcount, 0, sizeof gcount);
        uint64_t *gsize_big = NULL; (void)gsize_big;
        for (uint64_t i = 0; i < n; i++) {
            uint32_t k = klen[i];
            if (k > 255) { free(it); free(keys); free(tmp); return TIDK_ORACLE_EARG; }
            uint8_t *kp = keys + (size_t)maxk * i;
            tidk_oracle_lex_min(label_ptr[i], k, kp);
            it[i] = (fast_item){k, kp, gcount[k]++, counts[i]};
        }
        qsort(it, n, sizeof(fast_item), fast_cmp);
        for (uint64_t i = 0; i < n;) {
            uint64_t j = i;
            while (j < n && it[j].klen == it[i].klen && memcmp(it[j].key, it[i].key, it[i].klen) == 0) j++;
            uint32_t k = it[i].klen;
            if (gcount[k] == 1) map_insert_new(&map, it[i].key, k, (int32_t)it[i].count);
            else if (j - i >= 2) {
                int32_t v = (int32_t)it[i].count + (int32_t)it[i + 1].count;
                if (it[i].gidx == 0 && it[i + 1].gidx == 1) v = (int32_t)(v * 2);
                map_insert_new(&map, it[i].key, k, v);
            }
            i = j;
        }
        free(it); free(keys);
    }

    qsort(map.rows, map.n, sizeof(est_row), row_cmp);
    int64_t nout = 0;
    size_t koff = 0;
    for (size_t i = 0; i < map.n; i++) {
        est_row *r = &map.rows[i];
        if (tidk_oracle_period(r->key, r->klen) > 3) { /* filter_count_vec, explore.rs:455-464 */
            if ((size_t)nout < out_cap && koff + r->klen <= out_keys_cap) {
                memcpy(out_keys + koff, r->key, r->klen);
                out_klen[nout] = r->klen;
                out_counts[nout] = r->count;
                koff += r->klen;
            }
            nout++;
        }
        free(r->key);
    }
    free(map.rows); free(tmp);
    return nout;
}
