// fasta.hpp -- FASTA ingest for the host driver (stays on the host, lib.rs:32-49).
// rust-bio 2.0.3 reader semantics as far as the reference relies on them (SURVEY.md 8c,
// unpinned by the reference's tests): record starts at a '>' line, id = first
// whitespace-delimited token, sequence = following lines with trailing whitespace removed.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace tidk_host {

struct Fasta {
    std::vector<std::string> ids;
    std::vector<uint64_t> offsets{0};
    std::vector<uint8_t> seq;
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

inline Fasta read_fasta(const std::string &path) {
    const std::vector<char> data = slurp(path);
    Fasta fa;
    fa.seq.reserve(data.size());
    const char *p = data.data(), *end = p + data.size();
    bool started = false;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        if (p < le && *p == '>') {
            if (started) fa.offsets.push_back(fa.seq.size());
            const char *q = p + 1, *e = le;
            while (e > q && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) e--;
            const char *w = q;
            while (w < e && *w != ' ' && *w != '\t') w++;
            fa.ids.emplace_back(q, w);
            started = true;
        } else {
            const char *e = le;
            while (e > p && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '\f' || e[-1] == '\v')) e--;
            if (!started) {
                if (e > p) throw std::runtime_error("Expected > at record start.");
            } else {
                fa.seq.insert(fa.seq.end(), (const uint8_t *)p, (const uint8_t *)e);
            }
        }
        p = nl ? nl + 1 : end;
    }
    if (started) fa.offsets.push_back(fa.seq.size());
    return fa;
}

} // namespace tidk_host
