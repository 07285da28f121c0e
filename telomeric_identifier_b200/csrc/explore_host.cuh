// explore_host.cuh -- host driver of the explore path: slice table, kernel choice, event
// buffer sizing / retry, events -> runs.
#pragma once
#include <chrono>

#include <cub/device/device_scan.cuh>

#include "explore.cuh"
#include "explore_planes.cuh"
#include "explore_post.cuh"
#include "explore_vertical.cuh"

namespace tidk {

// ExploreScratch::misc: +0 work counter of the byte-wise kernel, +8 event slots handed out, +16 fillers among them,
// +32 compaction count, +40 tiles the vertical scan left to the plane kernel, +48 the vertical scan's work counter,
// +128 error flags, +1024 the plane kernel's ticket queues (kQueues x 128 bytes)
constexpr size_t kExMiscQueues = 1024;
constexpr uint64_t kVerticalMinTiles = 131072; // (125 Mbases; measured cross-over: 30 Mbases of plain reads, 100+ Mbases of a soft-masked genome) // batches with fewer tiles than this are scanned by the plane kernel
constexpr size_t kExMiscBytes = kExMiscQueues + (size_t)kQueues * kQueueStride * 8;

// split_seq_by_distance (explore.rs:151-160) for every record, plus the tiles each slice is cut into
// (tile == 0: the tiles of the vertical scan, explore_vertical.cuh)
__global__ void __launch_bounds__(256) explore_slices_kernel(const uint64_t *off, uint32_t n_records, double d, uint64_t tile,
                                                             uint32_t extra, SliceDesc *slices, uint64_t *tiles) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) tiles[2 * (uint64_t)n_records] = 0; // the scan's last element is the total
    if (r >= n_records) return;
    const uint64_t beg = off[r], len = off[r + 1] - beg;
    const double v = ceil(__dmul_rn((double)len, d)); // = split_dist() on the host: one IEEE multiply, then ceil
    const uint64_t dist = v > 0.0 ? (uint64_t)v : 0;
    slices[2 * (uint64_t)r] = {beg, dist};                    // explore.rs:157
    slices[2 * (uint64_t)r + 1] = {beg + (len - dist), dist}; // explore.rs:158
    const uint64_t span = dist == 0 ? 0 : dist + extra;
    const uint64_t t = tile ? (span + tile - 1) / tile : vt_count(dist);
    tiles[2 * (uint64_t)r] = t;
    tiles[2 * (uint64_t)r + 1] = t;
}

// tiles per slice for a slice table the host laid out (compacted upload: only the slices are resident)
__global__ void __launch_bounds__(256) explore_tiles_kernel(const SliceDesc *slices, uint32_t n_slices, uint64_t tile, uint32_t extra,
                                                            uint64_t *tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) tiles[n_slices] = 0;
    if (i >= n_slices) return;
    const uint64_t dist = slices[i].len;
    const uint64_t span = dist == 0 ? 0 : dist + extra;
    tiles[i] = (span + tile - 1) / tile;
}

inline int explore_runs_device(ExploreScratch &sc, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1,
                               int sm_count, const uint8_t *d_seq, const uint64_t *d_off,
                               const uint64_t *h_off, uint32_t n_records, double distance,
                               uint32_t kmin, uint32_t kmax, uint64_t threshold, tidk_run **out_runs,
                               uint64_t *n_runs, float *kernel_ms, uint32_t *kernel_launches,
                               std::string &err, float *scan_ms = nullptr, float *post_ms = nullptr,
                               float *host_ms = nullptr /* [0] slice table + uploads, [1] post-processing wall incl. D2H */,
                               const SliceDesc *h_slices = nullptr /* slices laid out by the caller (compacted upload) */,
                               const uint32_t *d_planes = nullptr, const uint32_t *d_plainbits = nullptr /* resident bit-planes of the batch */,
                               bool all_plain = false /* ... and nothing in the batch but upper-case A C G T */) {
    if (scan_ms) *scan_ms = 0.f;
    if (post_ms) *post_ms = 0.f;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    if (host_ms) host_ms[0] = host_ms[1] = 0.f;
    auto cuerr = [&](cudaError_t e, const char *what) {
        err = std::string(what) + " failed: " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? TIDK_ENOMEM : TIDK_ECUDA;
    };
    if (n_records > 0x3FFFFFFFu) { err = "too many records"; return TIDK_EARG; } // (2 n + 1 slice counts go through a 32-bit CUB scan)
    const bool planes = kmax <= (uint32_t)kPlaneMaxK; // bit-plane kernel; byte-wise kernel otherwise
    const uint32_t n_slices = n_records * 2;
    // The slice table (explore.rs:151-160) and the tile prefix sums are built on the device from the
    // resident offsets; the host only needs the totals, which ONE pass over its copy of the offsets gives
    // (tiles counted for both item sizes: which one is used depends on the total).
    uint64_t scanned = 0, max_dist = 0, nt_main = 0, nt_small = 0, n_vtiles = 0;
    const uint32_t extra = planes ? kmax : 0; // the plane kernel also evaluates the kmax positions behind the slice end
    const uint64_t tile_main = planes ? kExTile : kExploreTile;
    for (uint32_t r = 0; r < n_records; r++) {
        const uint64_t len = h_off[r + 1] - h_off[r];
        const uint64_t dist = split_dist(len, distance);
        if (dist > len) { err = "slice longer than record"; return TIDK_EARG; }
        const uint64_t span = dist == 0 ? 0 : dist + extra;
        nt_main += 2 * ((span + tile_main - 1) / tile_main);
        nt_small += 2 * ((span + kExTileSmall - 1) / kExTileSmall);
        n_vtiles += 2 * vt_count(dist);
        scanned += 2 * dist;
        max_dist = std::max(max_dist, dist);
    }
    // smaller items when the big ones would leave most of the GPU idle (C3: 60 slices)
    const bool small_items = planes && scanned / kExTile < (uint64_t)sm_count * 32;
    const uint64_t tile = small_items ? kExTileSmall : tile_main;
    const uint64_t n_tiles = small_items ? nt_small : nt_main;
    // events as sortable 64-bit keys + device-side run building when the batch fits the key layout
    const uint64_t max_chunks = max_dist / kmin + 2;
    // TIDK_EXPLORE_HOST_POST=1 forces the struct events + host stitching path (tests exercise both)
    KeyFmt fmt{bits_for(max_chunks), bits_for(n_slices > 0 ? n_slices - 1 : 0)};
    const int key_bits = (int)fmt.total_bits(kmax);
    const bool packed = !getenv("TIDK_EXPLORE_HOST_POST") && key_bits <= 64;
    // Resident planes: plain tiles (upper-case A C G T only) are scanned in vertical form (explore_vertical.cuh), the rest
    // of the batch by the plane kernel from the ASCII copy.  TIDK_VERTICAL=0 keeps everything on the plane kernel.
    const bool vertical_on = !(getenv("TIDK_VERTICAL") && atoi(getenv("TIDK_VERTICAL")) == 0); // (read per call: the tests switch)
    // (small batches stay on the plane kernel: the vertical scan works in items of 32 tiles -- 30 000 bases -- per warp)
    const uint64_t vertical_min = getenv("TIDK_VERTICAL_MIN_TILES") ? strtoull(getenv("TIDK_VERTICAL_MIN_TILES"), nullptr, 10) : kVerticalMinTiles;
    const bool vertical = vertical_on && planes && packed && d_planes && d_plainbits && !h_slices && kmax <= (uint32_t)kVMaxK &&
                          n_vtiles > 0 && n_vtiles >= vertical_min && (n_records == 0 || h_off[n_records] < (1ull << 36));
    std::vector<tidk_run> runs;
    if (n_tiles > 0) {
        cudaError_t e;
        auto grow = [&](void *&p, size_t &cap, size_t need) -> cudaError_t {
            if (p && need <= cap) return cudaSuccess;
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            cudaError_t e2 = cudaMalloc(&p, need + need / 8 + 256);
            if (e2 == cudaSuccess) cap = need + need / 8 + 256; else p = nullptr;
            return e2;
        };
        if ((e = grow(sc.slices, sc.slices_cap, (size_t)n_slices * sizeof(SliceDesc))) != cudaSuccess) return cuerr(e, "cudaMalloc(slices)");
        if ((e = grow(sc.tile_base, sc.tile_cap, ((size_t)n_slices + 1) * 8)) != cudaSuccess) return cuerr(e, "cudaMalloc(tile_base)");
        if ((e = grow(sc.tiles, sc.tiles_cap, ((size_t)n_slices + 1) * 8)) != cudaSuccess) return cuerr(e, "cudaMalloc(tiles)");
        if (planes && (e = grow(sc.items, sc.items_cap, (vertical ? n_vtiles : n_tiles) * sizeof(ExItem))) != cudaSuccess) return cuerr(e, "cudaMalloc(items)");
        if (vertical && (e = grow(sc.vtiles, sc.vtiles_cap, (size_t)n_vtiles * sizeof(VTile))) != cudaSuccess) return cuerr(e, "cudaMalloc(vtiles)");
        if (!sc.misc && (e = cudaMalloc(&sc.misc, kExMiscBytes)) != cudaSuccess) { sc.misc = nullptr; return cuerr(e, "cudaMalloc(misc)"); }
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)(n_slices + 1), st);
        if ((e = grow(sc.temp, sc.temp_cap, scan_bytes)) != cudaSuccess) return cuerr(e, "cudaMalloc(temp)");
        uint64_t cap = std::max<uint64_t>(1u << 16, scanned / 64);
        std::vector<RunEvent> events;
        float ms_total = 0.f;
        uint32_t launches = 0;
        bool items_ready = false, prep_ready = false;
        for (int attempt = 0; attempt < 3; attempt++) {
            if ((e = grow(sc.events, sc.events_cap, cap * (packed ? sizeof(uint64_t) : sizeof(RunEvent)))) != cudaSuccess) return cuerr(e, "cudaMalloc(events)");
            if ((e = cudaMemsetAsync(sc.misc, 0, kExMiscBytes, st)) != cudaSuccess) return cuerr(e, "memset");
            const uint64_t want = (n_tiles + 7) / 8, maxb = (uint64_t)sm_count * 8;
            int grid = (int)std::max<uint64_t>(1, std::min(want, maxb));
            if (host_ms && attempt == 0) host_ms[0] = since(t_begin);
            cudaEventRecord(ev0, st);
            if (!prep_ready) { // slice table, tiles per slice, prefix sums: inside the timed region
                if (h_slices) {
                    if ((e = cudaMemcpyAsync(sc.slices, h_slices, (size_t)n_slices * sizeof(SliceDesc), cudaMemcpyHostToDevice, st)) != cudaSuccess)
                        return cuerr(e, "H2D slices");
                    explore_tiles_kernel<<<(n_slices + 255) / 256, 256, 0, st>>>((const SliceDesc *)sc.slices, n_slices, tile, extra, (uint64_t *)sc.tiles);
                } else {
                    explore_slices_kernel<<<(n_records + 255) / 256, 256, 0, st>>>(d_off, n_records, distance, vertical ? 0 : tile, extra,
                                                                                   (SliceDesc *)sc.slices, (uint64_t *)sc.tiles);
                }
                size_t tb0 = sc.temp_cap;
                if ((e = cub::DeviceScan::ExclusiveSum(sc.temp, tb0, (const uint64_t *)sc.tiles, (uint64_t *)sc.tile_base, (int)(n_slices + 1), st)) != cudaSuccess)
                    return cuerr(e, "tile prefix sums");
                launches += 2;
                prep_ready = true;
            }
            if (vertical) {
                VJob vj{};
                vj.planes = d_planes; vj.plainbits = d_plainbits;
                vj.slices = (const SliceDesc *)sc.slices;
                if (!items_ready) {
                    vt_items_kernel<<<(n_slices + 255) / 256, 256, 0, st>>>((const uint64_t *)sc.tile_base, n_slices, (VTile *)sc.vtiles);
                    launches++;
                    items_ready = true;
                }
                vj.vtiles = (const VTile *)sc.vtiles;
                vj.n_vtiles = n_vtiles;
                vj.kmin = kmin; vj.kmax = kmax;
                vj.events = (uint64_t *)sc.events;
                vj.n_events = (unsigned long long *)sc.misc + 1;
                vj.cap_events = cap;
                vj.work_counter = (uint32_t *)((char *)sc.misc + 48);
                vj.fmt = fmt;
                vj.slow_items = (ExItem *)sc.items;
                vj.n_slow = (unsigned long long *)sc.misc + 5;
                auto vlaunch = [&](auto kernel) {
                    int per_sm = 0;
                    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VShared));
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kVWarps * 32, sizeof(VShared)) != cudaSuccess || per_sm < 1) per_sm = 1;
                    if (const char *cap_env = getenv("TIDK_V_PERSM")) per_sm = std::max(1, std::min(per_sm, atoi(cap_env))); // (measurements)
                    const uint64_t vwant = ((n_vtiles + 31) / 32 + kVWarps - 1) / kVWarps;
                    kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>(vwant, (uint64_t)sm_count * per_sm)), kVWarps * 32, sizeof(VShared), st>>>(vj);
                };
                if (kmin == 5 && kmax == 12) vlaunch(explore_vertical_kernel<5, 12>);
                else vlaunch(explore_vertical_kernel<0, 0>);
                if (!all_plain) {
                    // what the vertical scan leaves behind (tiles with anything but upper-case A C G T): the plane kernel, from
                    // the ASCII copy, on the items written for it (their number is read on the device)
                    ExJob job{};
                    job.seq = d_seq;
                    job.slices = (const SliceDesc *)sc.slices;
                    job.items = (const ExItem *)sc.items;
                    job.n_items = 0;
                    job.n_items_dev = (const unsigned long long *)sc.misc + 5;
                    job.kmin = kmin; job.kmax = kmax;
                    job.events = (RunEvent *)sc.events;
                    job.cap_events = cap;
                    job.work_counter = (unsigned long long *)((char *)sc.misc + kExMiscQueues);
                    job.n_events = (unsigned long long *)sc.misc + 1;
                    job.n_unused = (unsigned long long *)sc.misc + 2;
                    job.zero = 0;
                    job.err_flags = (uint32_t *)((char *)sc.misc + 128);
                    job.packed = 1u;
                    job.fmt = fmt;
                    auto launch = [&](auto kernel) {
                        int per_sm = 0;
                        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
                        const uint64_t gwant = (n_vtiles + 7) / 8;
                        kernel<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>(gwant, (uint64_t)sm_count * per_sm)), 256, 0, st>>>(job);
                    };
                    if (kmin == 5 && kmax == 12) launch(explore_planes_kernel<5, 12, false>);
                    else launch(explore_planes_kernel<0, 0, false>);
                    launches++;
                }
            } else if (planes) {
                if (!items_ready) {
                    ExPrepJob pj{(const SliceDesc *)sc.slices, (const uint64_t *)sc.tile_base, (ExItem *)sc.items, n_tiles, n_slices, kmax, (uint32_t)tile};
                    explore_items_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(pj);
                    launches++;
                    items_ready = true;
                }
                ExJob job{};
                job.seq = d_seq;
                job.slices = (const SliceDesc *)sc.slices;
                job.items = (const ExItem *)sc.items;
                job.n_items = n_tiles;
                job.kmin = kmin; job.kmax = kmax;
                job.events = (RunEvent *)sc.events;
                job.cap_events = cap;
                job.work_counter = (unsigned long long *)((char *)sc.misc + kExMiscQueues);
                job.n_events = (unsigned long long *)sc.misc + 1;
                job.n_unused = (unsigned long long *)sc.misc + 2;
                job.zero = 0;
                job.err_flags = (uint32_t *)((char *)sc.misc + 128);
                job.packed = packed ? 1u : 0u;
                job.fmt = fmt;
                job.planes = d_planes;
                job.plainbits = d_plainbits;
                const bool from_planes = d_planes && d_plainbits && !h_slices;
                // persistent launch: the grid is the resident set (the first items are assigned statically)
                auto launch = [&](auto kernel) {
                    int per_sm = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
                    grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sm_count * per_sm));
                    kernel<<<grid, 256, 0, st>>>(job);
                };
                if (kmin == 5 && kmax == 12) {
                    if (from_planes) launch(explore_planes_kernel<5, 12, true>);
                    else launch(explore_planes_kernel<5, 12, false>);
                } else {
                    if (from_planes) launch(explore_planes_kernel<0, 0, true>);
                    else launch(explore_planes_kernel<0, 0, false>);
                }
            } else {
                ExploreJob job{};
                job.seq = d_seq;
                job.slices = (const SliceDesc *)sc.slices;
                job.tile_base = (const uint64_t *)sc.tile_base;
                job.n_tiles = n_tiles;
                job.n_slices = n_slices;
                job.kmin = kmin; job.kmax = kmax;
                job.events = (RunEvent *)sc.events;
                job.cap_events = cap;
                job.work_counter = (unsigned long long *)sc.misc;
                job.n_events = (unsigned long long *)sc.misc + 1;
                job.err_flags = (uint32_t *)((char *)sc.misc + 128);
                job.packed = packed ? 1u : 0u;
                job.fmt = fmt;
                explore_events_kernel<<<grid, 256, 0, st>>>(job);
            }
            cudaEventRecord(ev1, st);
            launches++;
            if ((e = cudaGetLastError()) != cudaSuccess) return cuerr(e, "explore kernel launch");
            unsigned long long h_counts[6] = {0, 0, 0, 0, 0, 0}; // [1] event slots, [2] fillers among them, [5] tiles the vertical scan left to the plane kernel
            uint32_t h_err = 0;
            auto read_counts = [&]() -> cudaError_t {
                cudaError_t e2;
                if ((e2 = cudaMemcpyAsync(h_counts, sc.misc, sizeof h_counts, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e2;
                if ((e2 = cudaMemcpyAsync(&h_err, (char *)sc.misc + 128, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e2;
                if ((e2 = cudaStreamSynchronize(st)) != cudaSuccess) return e2;
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev0, ev1);
                ms_total += ms;
                return cudaSuccess;
            };
            if ((e = read_counts()) != cudaSuccess) return cuerr(e, "explore kernel");
            if (h_err & 1u) { err = "sequence contains a byte >= 0x80 (str::from_utf8 panics, explore.rs:212)"; return TIDK_ENONASCII; }
            const uint64_t ne = h_counts[1];               // slots handed out: events + fillers of unused reservations
            const uint64_t n_real = ne - h_counts[2];
            if (ne > cap) { cap = ne + ne / 16 + 65536; continue; } // event buffer too small: rerun (reservations make the count vary a little)
            // (CUB's device-wide primitives take 32-bit item counts: a batch with more events than that -- k = 1 or 2 on
            // gigabases -- has its keys stitched on the host instead)
            const uint64_t device_post_max = getenv("TIDK_EXPLORE_POST_MAX") ? strtoull(getenv("TIDK_EXPLORE_POST_MAX"), nullptr, 10) : 0x7FFFFFFFull;
            if (packed && ne > device_post_max) {
                std::vector<uint64_t> keys(ne);
                if ((e = cudaMemcpy(keys.data(), sc.events, ne * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuerr(e, "D2H event keys");
                events.clear();
                events.reserve(n_real);
                for (uint64_t key : keys) {
                    if (key == ~0ull) continue; // filler of an unused slot
                    RunEvent ev;
                    uint32_t k32, ty, jn;
                    unpack_event(fmt, key, ev.slice, k32, ev.chunk, ty, jn);
                    ev.k = (uint8_t)k32; ev.type = (uint8_t)ty; ev.joins = (uint8_t)jn; ev.pad = 0;
                    events.push_back(ev);
                }
                if (events.size() != n_real) { err = "internal: event count does not match"; return TIDK_ECUDA; }
                break;
            }
            if (packed) { // sort, build runs, filter and compact on the device
                const auto t_post = std::chrono::steady_clock::now();
                cudaEventRecord(ev0, st);
                const int prc = explore_post_device(sc, st, n_real, ne, fmt, key_bits, threshold, out_runs, n_runs, &launches, err);
                cudaEventRecord(ev1, st);
                cudaStreamSynchronize(st);
                if (host_ms) host_ms[1] = since(t_post);
                float pms = 0.f;
                cudaEventElapsedTime(&pms, ev0, ev1);
                if (kernel_ms) *kernel_ms = ms_total + pms;
                if (scan_ms) *scan_ms = ms_total;
                if (post_ms) *post_ms = pms;
                if (kernel_launches) *kernel_launches = launches;
                return prc;
            }
            events.resize(ne);
            if (ne && (e = cudaMemcpy(events.data(), sc.events, ne * sizeof(RunEvent), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuerr(e, "D2H events");
            events.erase(std::remove_if(events.begin(), events.end(), [](const RunEvent &ev) { return ev.k == 0xFF && ev.type == 0xFF; }), events.end()); // fillers
            if (events.size() != n_real) { err = "internal: event count does not match"; return TIDK_ECUDA; }
            break;
        }
        if (kernel_ms) *kernel_ms = ms_total;
        if (scan_ms) *scan_ms = ms_total;
        if (kernel_launches) *kernel_launches = launches;
        events_to_runs(events, threshold, runs);
    }
    *n_runs = runs.size();
    if (!runs.empty()) {
        *out_runs = run_list_alloc_plain(runs.size() * sizeof(tidk_run));
        if (!*out_runs) { err = "out of host memory"; *n_runs = 0; return TIDK_ENOMEM; }
        memcpy(*out_runs, runs.data(), runs.size() * sizeof(tidk_run));
    }
    return TIDK_OK;
}

} // namespace tidk
