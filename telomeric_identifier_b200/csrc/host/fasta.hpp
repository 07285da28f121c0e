// fasta.hpp -- FASTA ingest for the host driver (stays on the host, lib.rs:32-49).
// rust-bio 2.0.3 reader semantics as far as the reference relies on them (SURVEY.md 8c,
// unpinned by the reference's tests): record starts at a '>' line, id = first
// whitespace-delimited token, sequence = following lines with trailing whitespace removed.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace tidk_host {

// The concatenated sequence.  Not a std::vector: resize() would zero-fill (and page-fault) hundreds of megabytes on
// one thread before the parser's workers overwrite them; here the pages are first touched by the workers, in parallel.
class ByteBuf {
  public:
    ByteBuf() = default;
    ByteBuf(const ByteBuf &) = delete;
    ByteBuf &operator=(const ByteBuf &) = delete;
    ByteBuf(ByteBuf &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    ByteBuf &operator=(ByteBuf &&o) noexcept {
        if (this != &o) { free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~ByteBuf() { free(p_); }
    void resize_uninitialized(size_t n) {
        free(p_);
        p_ = n ? (uint8_t *)malloc(n) : nullptr;
        if (n && !p_) throw std::runtime_error("out of memory for the sequence");
        n_ = n;
    }
    uint8_t *data() { return p_; }
    const uint8_t *data() const { return p_; }
    size_t size() const { return n_; }
    const uint8_t *begin() const { return p_; }
    const uint8_t *end() const { return p_ + n_; }

  private:
    uint8_t *p_ = nullptr;
    size_t n_ = 0;
};

struct Fasta {
    std::vector<std::string> ids;
    std::vector<uint64_t> offsets{0};
    ByteBuf seq;
};

inline bool ends_with(const std::string &s, const char *suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

inline std::vector<char> slurp(const std::string &path) {
    std::vector<char> data;
    if (ends_with(path, ".gz")) { // lib.rs:40-44: by extension; MultiGzDecoder == gzread over members
        gzFile g = gzopen(path.c_str(), "rb");
        if (!g) throw std::runtime_error("cannot open " + path);
        gzbuffer(g, 1 << 20);
        std::vector<char> buf(1 << 22);
        int n;
        while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) data.insert(data.end(), buf.begin(), buf.begin() + n);
        gzclose(g);
        if (n < 0) throw std::runtime_error("gzip error in " + path);
    } else {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        data.resize(sz > 0 ? (size_t)sz : 0);
        if (sz > 0 && fread(data.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); throw std::runtime_error("short read on " + path); }
        fclose(f);
    }
    return data;
}

// Parse `data` into `fa` with `threads` workers.  Two passes over line-aligned chunks: pass 1
// finds the headers and counts the (right-trimmed) sequence bytes between them, a serial prefix sum
// turns that into record offsets and per-chunk write positions, pass 2 copies the sequence lines.
inline void parse_fasta_buffer(const char *data, size_t size, Fasta &fa, unsigned threads) {
    struct Header { size_t seq_before; std::string id; }; // sequence bytes of the chunk before this header
    struct Chunk { size_t beg = 0, end = 0, seq_total = 0, out = 0; std::vector<Header> headers; bool stray = false; };
    if (threads < 1) threads = 1;
    if (size < (1u << 20)) threads = 1;
    std::vector<Chunk> chunks(threads);
    size_t pos = 0;
    for (unsigned t = 0; t < threads; t++) {
        chunks[t].beg = pos;
        size_t want = size * (t + 1) / threads;
        if (t + 1 == threads) want = size;
        if (want < pos) want = pos;
        if (want < size) {
            const char *nl = (const char *)memchr(data + want, '\n', size - want);
            want = nl ? (size_t)(nl - data) + 1 : size;
        }
        chunks[t].end = pos = want;
    }
    auto trimmed_end = [](const char *p, const char *e) {
        while (e > p && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '\f' || e[-1] == '\v')) e--;
        return e;
    };
    auto pass1 = [&](Chunk &c) {
        const char *p = data + c.beg, *end = data + c.end;
        size_t acc = 0;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;
            if (p < le && *p == '>') {
                const char *q = p + 1, *e = trimmed_end(q, le);
                const char *w = q;
                while (w < e && *w != ' ' && *w != '\t') w++;
                c.headers.push_back({acc, std::string(q, w)});
            } else {
                acc += (size_t)(trimmed_end(p, le) - p);
            }
            p = nl ? nl + 1 : end;
        }
        c.seq_total = acc;
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < threads; t++) th.emplace_back(pass1, std::ref(chunks[t]));
        pass1(chunks[0]);
        for (auto &x : th) x.join();
    }
    // serial stitch: record offsets + where each chunk writes
    size_t total = 0;
    bool started = false;
    fa.ids.clear();
    fa.offsets.assign(1, 0);
    for (Chunk &c : chunks) {
        c.out = total;
        if (!started && (c.headers.empty() ? c.seq_total : c.headers[0].seq_before) > 0)
            throw std::runtime_error("Expected > at record start.");
        for (const Header &h : c.headers) {
            if (started) fa.offsets.push_back(total + h.seq_before);
            else fa.offsets[0] = 0;
            fa.ids.push_back(h.id);
            started = true;
        }
        total += c.seq_total;
    }
    if (started) fa.offsets.push_back(total);
    fa.seq.resize_uninitialized(total);
    auto pass2 = [&](const Chunk &c) {
        const char *p = data + c.beg, *end = data + c.end;
        uint8_t *out = fa.seq.data() + c.out;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;
            if (!(p < le && *p == '>')) {
                const char *e = trimmed_end(p, le);
                memcpy(out, p, (size_t)(e - p));
                out += e - p;
            }
            p = nl ? nl + 1 : end;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < threads; t++) th.emplace_back(pass2, std::cref(chunks[t]));
        pass2(chunks[0]);
        for (auto &x : th) x.join();
    }
}

inline Fasta read_fasta(const std::string &path, unsigned threads = 0) {
    if (threads == 0) {
        threads = std::thread::hardware_concurrency();
        if (threads == 0) threads = 4;
        if (threads > 32) threads = 32;
    }
    Fasta fa;
    if (ends_with(path, ".gz")) {
        const std::vector<char> data = slurp(path);
        parse_fasta_buffer(data.data(), data.size(), fa, threads);
        return fa;
    }
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); throw std::runtime_error("cannot stat " + path); }
    if (st.st_size == 0) { close(fd); return fa; }
    // (no MAP_POPULATE: it maps the whole file on this one thread; the workers' first reads fault their own parts in)
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { // pipes, odd filesystems: fall back to read()
        close(fd);
        const std::vector<char> data = slurp(path);
        parse_fasta_buffer(data.data(), data.size(), fa, threads);
        return fa;
    }
    try {
        parse_fasta_buffer((const char *)m, (size_t)st.st_size, fa, threads);
    } catch (...) {
        munmap(m, (size_t)st.st_size); close(fd);
        throw;
    }
    munmap(m, (size_t)st.st_size);
    close(fd);
    return fa;
}

} // namespace tidk_host
