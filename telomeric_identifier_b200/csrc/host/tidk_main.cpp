// tidk (B200 host driver) -- plays the role of the reference's Rust host for the hot path:
// same subcommands, flags and defaults as main.rs:18-146 for search / find / explore, same output
// files and rows (search.rs:35-54,132-147; finder.rs:66-78,169-172; explore.rs:140-143), same
// stderr progress lines.  All counting / run detection goes through libtidk_cuda.so (C ABI);
// there is no CPU path.  `--log` writes the reference's log files; `plot` and `build` are out of scope
// (SURVEY.md section 2).
#include <sys/stat.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <ctime>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/tidk_cuda.h"
#include "clades.hpp"
#include "fasta.hpp"
#include "hostlogic.hpp"

using namespace tidk_host;

namespace {

struct Args {
    std::string sub, fasta, string_, clade, output, dir, extension = "tsv", database;
    uint64_t window = 10000;
    long long length = -1, minimum = 5, maximum = 12, threshold = 100;
    bool have_min = false, have_max = false;
    double distance = 0.01;
    bool print = false, verbose = false, log = false, stats = false;
    int device = 0;
    int gpus = 1;         // shard the windows (search / find) or the records (explore) over this many devices (device, device + 1, ...)
    unsigned threads = 0; // FASTA parser threads (0 = hardware concurrency, at most 32)
};

[[noreturn]] void die(const std::string &msg, int code = 2) {
    fprintf(stderr, "error: %s\n", msg.c_str());
    exit(code);
}

void usage() {
    fprintf(stderr,
            "tidk (B200) -- A Telomere Identification Toolkit, hot path on NVIDIA B200\n\n"
            "Usage: tidk search  <FASTA> -s <STRING> [-w 10000] -o <OUTPUT> -d <DIR> [-e tsv|bedgraph]\n"
            "       tidk find    <FASTA> -c <CLADE>  [-w 10000] -o <OUTPUT> -d <DIR> [-p] [--database CSV]\n"
            "       tidk explore <FASTA> [-l LENGTH | -m 5 -x 12] [-t 100] [--distance 0.01] [-v]\n"
            "Extra: --device N, --gpus N (search / find: windows, explore: records sharded over N devices), --gpu-stats (kernel time and Gbases/s on stderr)\n");
}

Args parse(int argc, char **argv) {
    Args a;
    if (argc < 2) { usage(); exit(2); }
    a.sub = argv[1];
    if (a.sub == "-h" || a.sub == "--help") { usage(); exit(0); }
    if (a.sub == "plot" || a.sub == "build") die("`tidk " + a.sub + "` is outside the B200 hot path; use the reference binary");
    if (a.sub != "search" && a.sub != "find" && a.sub != "explore" && a.sub != "fasta-check")
        die("unrecognized subcommand '" + a.sub + "'");
    auto need = [&](int &i) -> std::string {
        if (i + 1 >= argc) die(std::string("a value is required for '") + argv[i] + "'");
        return argv[++i];
    };
    for (int i = 2; i < argc; i++) {
        std::string f = argv[i];
        std::string val;
        size_t eq = f.find('=');
        bool has_eq = f.rfind("--", 0) == 0 && eq != std::string::npos;
        if (has_eq) { val = f.substr(eq + 1); f = f.substr(0, eq); }
        auto v = [&]() { return has_eq ? val : need(i); };
        if (f == "-s" || f == "--string") a.string_ = v();
        else if (f == "-w" || f == "--window") a.window = strtoull(v().c_str(), nullptr, 10);
        else if (f == "-o" || f == "--output") a.output = v();
        else if (f == "-d" || f == "--dir") a.dir = v();
        else if (f == "-e" || f == "--extension") a.extension = v();
        else if (f == "-c" || f == "--clade") a.clade = v();
        else if (f == "-p" || f == "--print") a.print = true;
        else if (f == "-l" || f == "--length") a.length = atoll(v().c_str());
        else if (f == "-m" || f == "--minimum") { a.minimum = atoll(v().c_str()); a.have_min = true; }
        else if (f == "-x" || f == "--maximum") { a.maximum = atoll(v().c_str()); a.have_max = true; }
        else if (f == "-t" || f == "--threshold") a.threshold = atoll(v().c_str());
        else if (f == "--distance") a.distance = atof(v().c_str());
        else if (f == "-v" || f == "--verbose") a.verbose = true;
        else if (f == "--log") a.log = true;
        else if (f == "--database") a.database = v();
        else if (f == "--device") a.device = atoi(v().c_str());
        else if (f == "--gpus") a.gpus = atoi(v().c_str());
        else if (f == "--threads") a.threads = (unsigned)atoi(v().c_str());
        else if (f == "--gpu-stats") a.stats = true;
        else if (!f.empty() && f[0] == '-') die("unexpected argument '" + f + "'");
        else if (a.fasta.empty()) a.fasta = f;
        else die("unexpected argument '" + f + "'");
    }
    return a;
}

void mkdirs(const std::string &dir) { // create_dir_all
    std::string cur;
    for (size_t i = 0; i <= dir.size(); i++) {
        if (i == dir.size() || dir[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) die("cannot create directory " + cur, 1);
        }
        if (i < dir.size()) cur += dir[i];
    }
}

// CUDA context creation takes ~1 s on a cold process; it runs on a helper thread while the FASTA
// is being parsed and is joined when the device is first needed.
struct Ctx {
    tidk_ctx *h = nullptr;
    int rc = 0;
    std::string err;
    std::thread starter;
    explicit Ctx(int device) {
        starter = std::thread([this, device] {
            int ids[1] = {device};
            rc = tidk_cuda_create(ids, 1, &h);
            if (rc != 0) err = tidk_cuda_last_error(nullptr);
        });
    }
    void ready() {
        if (starter.joinable()) starter.join();
        if (rc != 0) die("tidk_cuda_create: " + err, 1);
    }
    ~Ctx() { // the context itself is left to process exit (see main)
        if (starter.joinable()) starter.join();
    }
    void check(int rc) const {
        if (rc != 0) die(std::string("tidk-cuda error ") + std::to_string(rc) + ": " + tidk_cuda_last_error(h), 1);
    }
};

// --log: the files SubCommand::log writes (lib.rs:65-242), same text; the date is local time in the
// reference's format "%Y-%m-%d: %H:%M:%S" (lib.rs:59).  The version line names the tidk release whose
// behaviour this driver reproduces.
const char *kTidkVersion = "0.2.7";

std::string now_string() {
    char buf[64];
    time_t t = time(nullptr);
    struct tm tmv;
    localtime_r(&t, &tmv);
    strftime(buf, sizeof buf, "%Y-%m-%d: %H:%M:%S", &tmv);
    return buf;
}

std::string rust_f64(double v) { // Rust `{}` on f64: shortest digits that round-trip, no exponent for these magnitudes
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

void write_log_file(const std::string &name, const std::string &text) {
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) die("cannot create " + name, 1);
    fputs(text.c_str(), f);
    fputc('\n', f); // writeln!
    fclose(f);
    fprintf(stderr, "[+]\tLog file written to: %s\n", name.c_str());
}

const char *kHeader = "id\twindow\tforward_repeat_number\treverse_repeat_number\ttelomeric_repeat\n";

// rows in the reference order record -> repeat -> window
void write_rows(FILE *out, const Fasta &fa, uint64_t W, const std::vector<std::string> &labels,
                const std::vector<uint64_t> &f, const std::vector<uint64_t> &r, bool bedgraph) {
    std::string buf;
    buf.reserve(1 << 22);
    char line[512];
    size_t p = 0;
    for (size_t rec = 0; rec < fa.ids.size(); rec++) {
        const uint64_t n = fa.offsets[rec + 1] - fa.offsets[rec];
        const uint64_t nw = n / W + (n % W != 0);
        for (const std::string &lab : labels) {
            for (uint64_t i = 0; i < nw; i++, p++) {
                const uint64_t end = (i + 1) * W < n ? (i + 1) * W : n;
                int m;
                if (!bedgraph)
                    m = snprintf(line, sizeof line, "%s\t%llu\t%llu\t%llu\t%s\n", fa.ids[rec].c_str(), (unsigned long long)end,
                                 (unsigned long long)f[p], (unsigned long long)r[p], lab.c_str());
                else
                    m = snprintf(line, sizeof line, "%s\t%llu\t%llu\t%llu\n", fa.ids[rec].c_str(), (unsigned long long)(i * W),
                                 (unsigned long long)end, (unsigned long long)(f[p] + r[p]));
                if (m >= (int)sizeof line) { // very long id: format into a string
                    std::string s = fa.ids[rec] + "\t" + (bedgraph ? std::to_string(i * W) + "\t" : std::string()) + std::to_string(end) + "\t";
                    s += bedgraph ? std::to_string(f[p] + r[p]) + "\n"
                                  : std::to_string(f[p]) + "\t" + std::to_string(r[p]) + "\t" + lab + "\n";
                    buf += s;
                } else buf.append(line, (size_t)m);
                if (buf.size() > (1 << 22) - 1024) { fwrite(buf.data(), 1, buf.size(), out); buf.clear(); }
            }
        }
        fprintf(stderr, "[+]\tChromosome %s processed\n", fa.ids[rec].c_str()); // search.rs:71 / finder.rs:99
    }
    fwrite(buf.data(), 1, buf.size(), out);
}

// Multi-GPU search / find: windows are independent (an occurrence never spans two windows,
// search.rs:98,105-110), so the global window sequence is cut into `world` contiguous, equally long ranges
// -- whole records where possible, long records split at multiples of W, hence no halo -- and every device
// counts its range through its own context on its own thread.  A range is one contiguous stretch of the
// concatenated sequence, so nothing is copied on the host; the per-range counts ([piece][query][window])
// are placed into the global [record][query][window] arrays.  No collective: the counts are 16 B per window.
struct Piece { uint32_t rec; uint64_t first, n; };

std::vector<std::vector<Piece>> plan_shards(const std::vector<uint64_t> &offsets, uint64_t window, int world) {
    const size_t nrec = offsets.size() - 1;
    std::vector<uint64_t> base(nrec + 1, 0);
    for (size_t r = 0; r < nrec; r++) {
        const uint64_t n = offsets[r + 1] - offsets[r];
        base[r + 1] = base[r] + n / window + (n % window != 0);
    }
    const uint64_t total = base[nrec];
    std::vector<std::vector<Piece>> plans((size_t)world);
    size_t rec = 0;
    for (int g = 0; g < world; g++) {
        uint64_t lo = total / world * g + std::min<uint64_t>(total % world, g);     // floor(total * g / world) without overflow
        const uint64_t hi = total / world * (g + 1) + std::min<uint64_t>(total % world, g + 1);
        while (lo < hi) {
            while (base[rec + 1] <= lo) rec++;
            const uint64_t n = std::min(hi, base[rec + 1]) - lo;
            plans[(size_t)g].push_back({(uint32_t)rec, lo - base[rec], n});
            lo += n;
        }
    }
    return plans;
}

// --gpus N means "up to N devices": a CUDA context costs about a quarter of a second to create and as much to tear down, per
// device and one after the other, while a device counts 3 Gb in 0.1 s end to end -- on the 3.1 Gb genome eight devices took
// 4.1 s where one takes 0.7.  One device per 4 Gb of work, at most N (TIDK_GPUS_EXACT=1, or the one-GPU test switch
// TIDK_GPUS_SAME_DEVICE=1, keep exactly N).
int devices_for(const Args &a, uint64_t work_bases) {
    if (a.gpus <= 1) return 1;
    const bool exact = (getenv("TIDK_GPUS_EXACT") && atoi(getenv("TIDK_GPUS_EXACT")) != 0) ||
                       (getenv("TIDK_GPUS_SAME_DEVICE") && atoi(getenv("TIDK_GPUS_SAME_DEVICE")) != 0);
    if (exact) return a.gpus;
    const uint64_t per_device = 4000000000ull;
    const uint64_t want = (work_bases + per_device - 1) / per_device;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)a.gpus, want));
}

void sharded_counts(const Args &a, const Fasta &fa, const std::vector<const char *> &qp, const std::vector<uint32_t> &ql,
                    std::vector<uint64_t> &f, std::vector<uint64_t> &r, float *kernel_ms_max) {
    const int world = a.gpus;
    const uint32_t nq = (uint32_t)qp.size();
    const auto plans = plan_shards(fa.offsets, a.window, world);
    std::vector<uint64_t> nw(fa.ids.size()), obase(fa.ids.size() + 1, 0);
    for (size_t i = 0; i < fa.ids.size(); i++) {
        const uint64_t n = fa.offsets[i + 1] - fa.offsets[i];
        nw[i] = n / a.window + (n % a.window != 0);
        obase[i + 1] = obase[i] + nw[i] * nq;
    }
    // TIDK_GPUS_SAME_DEVICE=1 (tests on a one-GPU box): every shard runs on --device
    const bool same = getenv("TIDK_GPUS_SAME_DEVICE") && atoi(getenv("TIDK_GPUS_SAME_DEVICE")) != 0;
    setenv("LOCAL_WORLD_SIZE", std::to_string(world).c_str(), 0); // the shards share the host cores (packer threads per context)
    std::vector<std::string> errs((size_t)world);
    std::vector<std::thread> th;
    for (int g = 0; g < world; g++) {
        th.emplace_back([&, g] {
            const std::vector<Piece> &pieces = plans[(size_t)g];
            if (pieces.empty()) return;
            int ids[1] = {same ? a.device : a.device + g};
            tidk_ctx *h = nullptr;
            if (tidk_cuda_create(ids, 1, &h) != 0) { errs[(size_t)g] = std::string("tidk_cuda_create: ") + tidk_cuda_last_error(nullptr); return; }
            std::vector<uint64_t> loc_off(pieces.size() + 1, 0);
            uint64_t a0 = 0, n_loc = 0;
            for (size_t i = 0; i < pieces.size(); i++) {
                const Piece &p = pieces[i];
                const uint64_t o = fa.offsets[p.rec], e = fa.offsets[p.rec + 1];
                const uint64_t beg = o + p.first * a.window, end = std::min(o + (p.first + p.n) * a.window, e);
                if (i == 0) a0 = beg;
                loc_off[i + 1] = end - a0; // consecutive window ranges of consecutive records: one contiguous stretch
                n_loc += p.n * nq;
            }
            std::vector<uint64_t> lf(n_loc + 1), lr(n_loc + 1);
            const int rc = tidk_cuda_window_counts(h, fa.seq.data() + a0, loc_off.data(), (uint32_t)pieces.size(), a.window, qp.data(),
                                                   ql.data(), nq, lf.data(), lr.data());
            if (rc != 0) errs[(size_t)g] = std::string("tidk-cuda error ") + std::to_string(rc) + ": " + tidk_cuda_last_error(h);
            else {
                uint64_t pos = 0;
                for (const Piece &p : pieces)
                    for (uint32_t q = 0; q < nq; q++) {
                        const uint64_t dst = obase[p.rec] + q * nw[p.rec] + p.first;
                        memcpy(&f[dst], &lf[pos], p.n * 8);
                        memcpy(&r[dst], &lr[pos], p.n * 8);
                        pos += p.n;
                    }
            }
            tidk_cuda_destroy(h);
        });
    }
    for (auto &t : th) t.join();
    for (const std::string &e : errs)
        if (!e.empty()) die(e, 1);
    if (kernel_ms_max) *kernel_ms_max = 0.f; // per-shard kernel times are not collected on this path
}

// explore --gpus N: whole records in N contiguous groups balanced by scanned bases (2 * ceil(d * len) per record,
// explore.rs:151-160), one context and one thread per device, each scanning its stretch of the parsed sequence in
// place.  A record's runs do not depend on other records and every shard returns them in (k, record, slice, start)
// order, so the reference's arrival order is the shards in record order within each k: a stable sort on k.
std::vector<tidk_run> sharded_explore(const Args &a, const Fasta &fa, uint32_t kmin, uint32_t kmax, uint64_t thr) {
    const int world = a.gpus;
    const size_t nrec = fa.ids.size();
    std::vector<uint64_t> cost(nrec + 1, 0);
    for (size_t i = 0; i < nrec; i++) {
        const double v = std::ceil((double)(fa.offsets[i + 1] - fa.offsets[i]) * a.distance);
        cost[i + 1] = cost[i] + 2 * (v > 0 ? (uint64_t)v : 0) + 1; // + 1: empty records still spread out
    }
    std::vector<size_t> cut((size_t)world + 1, nrec);
    cut[0] = 0;
    for (int g = 1; g < world; g++) {
        const uint64_t want = cost[nrec] / (uint64_t)world * (uint64_t)g;
        cut[(size_t)g] = (size_t)(std::lower_bound(cost.begin(), cost.end(), want) - cost.begin());
        if (cut[(size_t)g] < cut[(size_t)g - 1]) cut[(size_t)g] = cut[(size_t)g - 1];
        if (cut[(size_t)g] > nrec) cut[(size_t)g] = nrec;
    }
    const bool same = getenv("TIDK_GPUS_SAME_DEVICE") && atoi(getenv("TIDK_GPUS_SAME_DEVICE")) != 0;
    setenv("LOCAL_WORLD_SIZE", std::to_string(world).c_str(), 0);
    std::vector<std::string> errs((size_t)world);
    std::vector<std::vector<tidk_run>> part((size_t)world);
    std::vector<std::thread> th;
    for (int g = 0; g < world; g++) {
        th.emplace_back([&, g] {
            const size_t r0 = cut[(size_t)g], r1 = cut[(size_t)g + 1];
            if (r1 <= r0) return;
            int ids[1] = {same ? a.device : a.device + g};
            tidk_ctx *h = nullptr;
            if (tidk_cuda_create(ids, 1, &h) != 0) { errs[(size_t)g] = std::string("tidk_cuda_create: ") + tidk_cuda_last_error(nullptr); return; }
            std::vector<uint64_t> loc_off(r1 - r0 + 1);
            for (size_t i = r0; i <= r1; i++) loc_off[i - r0] = fa.offsets[i] - fa.offsets[r0];
            tidk_run *runs = nullptr;
            uint64_t n = 0;
            const int rc = tidk_cuda_explore_runs(h, fa.seq.data() + fa.offsets[r0], loc_off.data(), (uint32_t)(r1 - r0), a.distance, kmin,
                                                  kmax, thr, &runs, &n);
            if (rc != 0) errs[(size_t)g] = std::string("tidk-cuda error ") + std::to_string(rc) + ": " + tidk_cuda_last_error(h);
            else {
                part[(size_t)g].assign(runs, runs + n);
                for (tidk_run &r : part[(size_t)g]) r.record += (uint32_t)r0;
                tidk_cuda_free_runs(runs);
            }
            tidk_cuda_destroy(h);
        });
    }
    for (auto &t : th) t.join();
    for (const std::string &e : errs)
        if (!e.empty()) die(e, 1);
    std::vector<tidk_run> all;
    for (auto &p : part) all.insert(all.end(), p.begin(), p.end());
    std::stable_sort(all.begin(), all.end(), [](const tidk_run &x, const tidk_run &y) { return x.k < y.k; });
    return all;
}

int run_counts(const Args &a, const std::vector<std::string> &queries, const std::vector<std::string> &labels,
               const std::string &path, bool header, bool bedgraph) {
    if (a.window == 0) die("window must be greater than 0", 1);
    const auto t0 = std::chrono::steady_clock::now();
    if (a.gpus < 1 || a.gpus > 64) die("--gpus must be between 1 and 64", 1);
    Ctx ctx(a.device); // starts creating the CUDA context in the background
    Fasta fa = read_fasta(a.fasta, a.threads);
    const auto t1 = std::chrono::steady_clock::now();
    mkdirs(a.dir);
    FILE *out = fopen(path.c_str(), "wb");
    if (!out) die("cannot create " + path, 1);
    if (header) fputs(kHeader, out);
    ctx.ready();
    const auto t_ready = std::chrono::steady_clock::now();
    uint64_t nwin = 0;
    for (size_t r = 0; r < fa.ids.size(); r++) {
        uint64_t n = fa.offsets[r + 1] - fa.offsets[r];
        nwin += n / a.window + (n % a.window != 0);
    }
    std::vector<uint64_t> f(nwin * queries.size() + 1), r(nwin * queries.size() + 1);
    std::vector<const char *> qp;
    std::vector<uint32_t> ql;
    for (auto &q : queries) { qp.push_back(q.c_str()); ql.push_back((uint32_t)q.size()); }
    float ms = 0;
    uint32_t nl = 0;
    Args a_eff = a;
    a_eff.gpus = devices_for(a, (uint64_t)fa.seq.size() * queries.size());
    if (a_eff.gpus > 1) {
        sharded_counts(a_eff, fa, qp, ql, f, r, &ms);
        nl = (uint32_t)a_eff.gpus;
    } else {
        // host-pointer entry point: the two-ended upload (2-bit planes packed by the host cores from the back, ASCII
        // by DMA from the front) moves a parsed genome several times faster than one copy from pageable memory
        ctx.check(tidk_cuda_window_counts(ctx.h, fa.seq.data(), fa.offsets.data(), (uint32_t)fa.ids.size(), a.window, qp.data(),
                                          ql.data(), (uint32_t)qp.size(), f.data(), r.data()));
        tidk_cuda_last_kernel_time(ctx.h, &ms, &nl);
    }
    const auto t2 = std::chrono::steady_clock::now();
    write_rows(out, fa, a.window, labels, f, r, bedgraph);
    fclose(out);
    fprintf(stderr, "[+]\tFinished searching genome.\n");
    if (a.stats) {
        const auto t3 = std::chrono::steady_clock::now();
        auto s = [](auto x, auto y) { return std::chrono::duration<double>(y - x).count(); };
        fprintf(stderr, "{\"bases\": %zu, \"kernel_ms\": %.4f, \"kernel_launches\": %u, \"kernel_gbases_s\": %.1f, "
                        "\"parse_s\": %.3f, \"wait_for_context_s\": %.3f, \"upload_and_count_s\": %.3f, \"write_s\": %.3f, "
                        "\"end_to_end_gbases_s\": %.3f, \"devices_used\": %d}\n",
                fa.seq.size(), ms, nl, ms > 0 ? fa.seq.size() / (ms * 1e-3) / 1e9 : 0.0, s(t0, t1), s(t1, t_ready), s(t_ready, t2),
                s(t2, t3),
                fa.seq.size() / s(t0, t3) / 1e9, a_eff.gpus);
    }
    return 0;
}

int cmd_search(const Args &a) { // search::search, search.rs:10-79
    if (a.fasta.empty() || a.string_.empty() || a.output.empty() || a.dir.empty())
        die("search needs <FASTA> --string --output --dir");
    if (a.extension != "tsv" && a.extension != "bedgraph") die("invalid value '" + a.extension + "' for '--extension'");
    fprintf(stderr, "[+]\tSearching genome for telomeric repeat: %s\n", a.string_.c_str());
    const std::string fwd = ascii_upper(a.string_); // search.rs:93
    const std::string path = a.dir + "/" + a.output + "_telomeric_repeat_windows." + a.extension;
    const int rc = run_counts(a, {fwd}, {fwd}, path, a.extension == "tsv", a.extension != "tsv");
    if (rc == 0 && a.log) // lib.rs:192-239 (the raw string ends with a newline and twenty blanks)
        write_log_file(a.dir + "/" + a.output + ".log",
                       std::string("tidk version: ") + kTidkVersion + "\nLog information for output file: " + path +
                           "\nDate: " + now_string() + "\n`tidk search` was run with the following parameters:\n    Input fasta: " +
                           a.fasta + "\n    Telomeric repeat search string: " + a.string_ + "\n    Window size: " +
                           std::to_string(a.window) + "\n                    ");
    return rc;
}

int cmd_find(const Args &a) { // finder::finder, finder.rs:13-107
    const std::string db = a.database.empty() ? database_path() : a.database;
    CladeTable table = load_clades(db);
    if (a.print) { // finder.rs:15-18 prints the table and exits 1; the pretty `tabled` layout is out of scope
        for (const std::string &c : get_clades(table)) {
            std::string line = c + "\t";
            auto seqs = return_telomere_sequence(table, c);
            for (size_t i = 0; i < seqs.size(); i++) line += (i ? ", " : "") + seqs[i];
            fprintf(stderr, "%s\n", line.c_str());
        }
        return 1;
    }
    if (a.fasta.empty() || a.clade.empty() || a.output.empty() || a.dir.empty()) die("find needs <FASTA> --clade --output --dir");
    bool known = false;
    for (const std::string &c : get_clades(table)) known = known || c == a.clade;
    if (!known) die("invalid value '" + a.clade + "' for '--clade <CLADE>'");
    const std::vector<std::string> reps = return_telomere_sequence(table, a.clade);
    if (reps.size() == 1) fprintf(stderr, "[+]\tSearching genome for a single telomeric repeat: %s\n", reps[0].c_str());
    else if (reps.size() > 1) {
        fprintf(stderr, "[+]\tSearching genome for %zu telomeric repeats:\n", reps.size());
        for (auto &r : reps) fprintf(stderr, "[+]\t\t%s\n", r.c_str());
    }
    const std::string path = a.dir + "/" + a.output + "_telomeric_repeat_windows.tsv";
    const int rc = run_counts(a, reps, reps, path, true, false); // repeats verbatim, finder.rs:129-136
    if (rc == 0 && a.log) { // lib.rs:69-117 (the reference names the output "...windows.csv" in the log; kept)
        std::string joined;
        for (size_t i = 0; i < reps.size(); i++) joined += (i ? ", " : "") + reps[i];
        write_log_file(a.dir + "/" + a.output + ".log",
                       std::string("tidk version: ") + kTidkVersion + "\nLog information for output file: " + a.dir + "/" +
                           a.output + "_telomeric_repeat_windows.csv\nDate: " + now_string() +
                           "\n`tidk find` was run with the following parameters:\n    Input fasta: " + a.fasta +
                           "\n    Window size: " + std::to_string(a.window) + "\n    Clade chosen: " + a.clade +
                           "\n    Telomeric repeats queried: " + joined);
    }
    return rc;
}

int cmd_explore(const Args &a) { // explore::explore, explore.rs:20-149
    if (a.fasta.empty()) die("explore needs <FASTA>");
    // main.rs:71-88: --length is required unless --minimum and --maximum are both given, and
    // defaults to 0 when --minimum is present
    long long length = a.length;
    if (length < 0) {
        if (!(a.have_min && a.have_max)) die("the following required arguments were not provided: --length <LENGTH> (or both --minimum and --maximum)");
        length = 0;
    }
    if (a.distance > 0.5) die("Distance from chromosome end as a proportion can't be more than 0.5.", 1); // explore.rs:41-43
    uint32_t kmin, kmax;
    if (length > 0) {
        fprintf(stderr, "[+]\tExploring genome for potential telomeric repeats of length: %lld\n", length);
        kmin = kmax = (uint32_t)length;
    } else {
        fprintf(stderr, "[+]\tExploring genome for potential telomeric repeats between lengths %lld and %lld.\n", a.minimum, a.maximum);
        for (long long k = a.minimum; k <= a.maximum; k++) fprintf(stderr, "[+]\t\tFinding telomeric repeat length: %lld\n", k);
        kmin = (uint32_t)a.minimum; kmax = (uint32_t)a.maximum;
    }
    const auto t0 = std::chrono::steady_clock::now();
    Ctx ctx(a.device); // starts creating the CUDA context in the background
    Fasta fa = read_fasta(a.fasta, a.threads);
    const auto t1 = std::chrono::steady_clock::now();
    auto t_ready = t1, t_call = t1;
    std::vector<std::pair<std::string, long long>> rows;
    float ms = 0;
    uint32_t nl = 0;
    uint64_t n = 0;
    bool multi_dev = false;
    if (a.gpus < 1 || a.gpus > 64) die("--gpus must be between 1 and 64", 1);
    if (kmax >= kmin && kmin > 0) {
        ctx.ready();
        t_ready = std::chrono::steady_clock::now();
        // `threshold as usize` (explore.rs:71,116): a negative i32 wraps to a huge usize
        const uint64_t thr = a.threshold < 0 ? (uint64_t)(int64_t)a.threshold : (uint64_t)a.threshold;
        std::vector<tidk_run> sharded;
        tidk_run *runs = nullptr;
        Args a_eff = a; // (explore's work: the scanned slices, 2 * ceil(d * len) per record)
        a_eff.gpus = devices_for(a, (uint64_t)(2.0 * a.distance * (double)fa.seq.size()) + 1);
        multi_dev = a_eff.gpus > 1;
        if (multi_dev) {
            sharded = sharded_explore(a_eff, fa, kmin, kmax, thr);
            runs = sharded.data();
            n = sharded.size();
            nl = (uint32_t)a_eff.gpus;
        } else {
            // the host-pointer entry point uploads only the chromosome-end slices when --distance is small
            ctx.check(tidk_cuda_explore_runs(ctx.h, fa.seq.data(), fa.offsets.data(), (uint32_t)fa.ids.size(), a.distance, kmin, kmax,
                                             thr, &runs, &n));
            float tms[2] = {0.f, 0.f};
            tidk_cuda_last_explore_timings(ctx.h, tms, 2);
            ms = tms[0] + tms[1];
            nl = 1;
        }
        t_call = std::chrono::steady_clock::now();
        std::vector<std::string> labels;
        std::vector<long long> counts;
        for (uint64_t i = 0; i < n; i++) {
            const tidk_run &r = runs[i];
            const uint64_t o = fa.offsets[r.record], e = fa.offsets[r.record + 1], len = e - o;
            double v = std::ceil((double)len * a.distance); // explore.rs:156
            const uint64_t dist = v > 0 ? (uint64_t)v : 0;
            const uint64_t base = r.slice == 0 ? o : e - dist;
            std::string lab((const char *)fa.seq.data() + base + r.label_pos, r.k);
            lab = ascii_upper(lab);
            if (period(lab) < 3) continue; // filter_by_frequency's simple-repeat half, explore.rs:265-274,351-354
            labels.push_back(lab);
            counts.push_back((long long)((r.end - r.start) / r.k));
        }
        if (!multi_dev) tidk_cuda_free_runs(runs);
        rows = estimates(labels, counts);
    }
    fprintf(stderr, "[+]\tFinished searching genome\n[+]\tGenerating output\n");
    printf("canonical_repeat_unit\tcount_repeat_runs_gt_%lld\n", a.threshold); // explore.rs:140
    for (auto &r : rows) printf("%s\t%lld\n", r.first.c_str(), r.second);
    if (a.stats) {
        const auto t3 = std::chrono::steady_clock::now();
        auto s = [](auto x, auto y) { return std::chrono::duration<double>(y - x).count(); };
        fprintf(stderr, "{\"runs\": %llu, \"kernel_ms\": %.4f, \"kernel_launches\": %u, \"parse_s\": %.3f, \"wait_for_context_s\": %.3f, "
                        "\"upload_and_scan_s\": %.3f, \"labels_and_estimator_s\": %.3f}\n",
                (unsigned long long)n, ms, nl, s(t0, t1), s(t1, t_ready), s(t_ready, t_call), s(t_call, t3));
    }
    if (a.log) // lib.rs:119-190: written to the current directory
        write_log_file("tidk-explore.log",
                       std::string("tidk version: ") + kTidkVersion + "\nLog information for output files: printed to STDOUT\nDate: " +
                           now_string() + "\n`tidk explore` was run with the following parameters:\n    Input fasta: " + a.fasta +
                           "\n    Explored telomeric repeat units of length: " + std::to_string(length) + "\n    Or from length: " +
                           std::to_string(a.minimum) + "\n    To length: " + std::to_string(a.maximum) + "\n    Threshold: " +
                           std::to_string(a.threshold) + "\n    Searching at " + rust_f64(a.distance * 100.0) +
                           "% distance from chromosome end");
    return 0;
}

} // namespace

// host-only self check of the FASTA reader (no device): record table + FNV-1a of the sequence
int cmd_fasta_check(const Args &a) {
    const auto t0 = std::chrono::steady_clock::now();
    Fasta fa = read_fasta(a.fasta, a.threads);
    if (a.stats) fprintf(stderr, "{\"parse_s\": %.4f}\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    uint64_t h = 1469598103934665603ull;
    for (uint8_t c : fa.seq) { h ^= c; h *= 1099511628211ull; }
    printf("%zu\t%zu\t%016llx\n", fa.ids.size(), fa.seq.size(), (unsigned long long)h);
    for (size_t i = 0; i < fa.ids.size(); i++)
        printf("%s\t%llu\n", fa.ids[i].c_str(), (unsigned long long)(fa.offsets[i + 1] - fa.offsets[i]));
    return 0;
}

int main(int argc, char **argv) {
    try {
        Args a = parse(argc, argv);
        int rc;
        if (a.sub == "fasta-check") rc = cmd_fasta_check(a);
        else if (a.sub == "search") rc = cmd_search(a);
        else if (a.sub == "find") rc = cmd_find(a);
        else rc = cmd_explore(a);
        // every output file is closed by now: leave without unwinding the CUDA context and the parsed genome
        // (tearing them down cost a command-line run 0.3 - 1.5 s after its last byte was written)
        fflush(stdout);
        fflush(stderr);
        _exit(rc);
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
