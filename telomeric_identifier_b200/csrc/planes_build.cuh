// planes_build.cuh -- K1: ASCII -> resident bit-planes, once per uploaded sequence batch.
//
// BASELINE.json's north star splits the device work into "ASCII -> 2-bit codes plus an N / case mask" and the
// scans that read the packed form.  An upload is bound by the bus (55 GB/s) while this kernel streams at the
// rate of HBM, so a batch that stays resident (tidk_cuda_seqbuf_upload: several clades, tsv + bedgraph, explore
// after search, ...) is transposed ONCE here and every later pass reads 0.25 - 0.375 B per base instead of 1:
//
//   planes      [lo words][hi words][va words] per 4 MiB of sequence (the layout of the host-packed upload,
//               window_counts.cuh PackedSource): lo = ASCII bit 1, hi = ASCII bit 2, va = byte is one of ACGTacgt
//   plainbits   one bit per 32-base word: all 32 bytes are UPPER-CASE A C G T -- the words whose lo / hi planes
//               say everything there is to say; the explore scan (raw-byte equality, explore.rs:208) takes any
//               block with another word in it from the ASCII copy, which stays resident beside the planes
//   validbits   one bit per word: va is all ones (the window counter skips the va plane of such blocks)
//
// Algorithmic bytes: 1 B read + 0.375 B written per base.  The ASCII bytes behind the last base are zero
// (kSeqPad), so the words behind the end come out as invalid / not plain without any masking.
#pragma once
#include "planes.cuh"
#include "window_counts.cuh" // kPackShift / kPackWords

namespace tidk {

struct PlaneBuildJob {
    const uint8_t *seq;   // 256-B aligned, zero padded behind the data
    uint32_t *planes;     // chunked [lo][hi][va]
    uint32_t *plainbits;  // [n_blocks + 1]
    uint32_t *validbits;  // [n_blocks + 1]
    uint64_t n_blocks;    // 1 KiB blocks to convert (covers the data and the padded tail)
    uint64_t total;       // bases in the batch: the word that holds the last base counts as plain when its bases are
    uint32_t *err_flags;  // two words of mapped host memory: [0] = 1 when a byte >= 0x80 is seen, [1] = 1 when a word that holds a base is not plain
};

__device__ __forceinline__ uint32_t *plane_word_ptr(uint32_t *planes, uint64_t word) {
    return planes + (word >> kPackShift) * (3ull * kPackWords) + (word & (kPackWords - 1));
}
__device__ __forceinline__ const uint32_t *plane_word_ptr(const uint32_t *planes, uint64_t word) {
    return planes + (word >> kPackShift) * (3ull * kPackWords) + (word & (kPackWords - 1));
}

__global__ void __launch_bounds__(256, 4) planes_build_kernel(const PlaneBuildJob job) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_warps = (uint64_t)gridDim.x * 8;
    uint64_t blk = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= job.n_blocks) {
        return;
    }
    uint32_t nonascii = 0;
    bool other = false; // a word of the data (not of the padding) that is not plain
    const uint64_t n_words = (job.total + 31) >> 5;
    uint32_t w[8];
    ld256_stream(job.seq + (blk << 10) + (uint32_t)lane * 32u, w);
    for (;;) {
        const Nibbles nib = to_nibbles(w); // w is dead from here: refill it
        const uint64_t nxt = blk + n_warps;
        if (nxt < job.n_blocks) {
            ld256_stream(job.seq + (nxt << 10) + (uint32_t)lane * 32u, w);
            if (nxt + n_warps < job.n_blocks) asm volatile("prefetch.global.L2 [%0];" ::"l"(job.seq + ((nxt + n_warps) << 10) + (uint32_t)lane * 32u));
        }
        const PlaneCore core = extract_core(nib);
        const Planes P = finish_planes(nib, core, __any_sync(0xFFFFFFFFu, core.bad != 0u), nonascii);
        const uint32_t va = (P.b0 ^ (P.hi & ~P.lo)) & P.ok;
        const uint32_t any_l = nib.Y[0] | nib.Y[1] | nib.Y[2] | nib.Y[3]; // ASCII bit 5 = bit 1 of the high nibbles
        const bool valid = va == 0xFFFFFFFFu;
        bool plain = valid && (any_l & 0x22222222u) == 0u;
        if (blk == (job.total >> 10) && (job.total & 31u) != 0u && (uint32_t)lane == ((uint32_t)(job.total >> 5) & 31u)) {
            // the last, partial word of the batch: only its bases count (what follows is padding, never inside a slice)
            const uint32_t tail = (1u << (job.total & 31u)) - 1u;
            plain = (va & tail) == tail && (nibble_plane<1>(nib.Y) & tail) == 0u;
        }
        other |= !plain && (blk << 5) + (uint32_t)lane < n_words;
        uint32_t *p = plane_word_ptr(job.planes, (blk << 5) + (uint32_t)lane);
        p[0] = P.lo;
        p[kPackWords] = P.hi;
        p[2 * kPackWords] = va;
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, plain), vm = __ballot_sync(0xFFFFFFFFu, valid);
        if (lane == 0) {
            job.plainbits[blk] = pm;
            job.validbits[blk] = vm;
        }
        if (nxt >= job.n_blocks) break;
        blk = nxt;
    }
    // Plain posted stores, one word per flag: an atomicOr on mapped host memory is a PCIe atomic per warp, and on a genome
    // with N gaps or soft-masked stretches every warp has something to report -- 4 736 of them in a row took the kernel
    // from 0.8 to 24.7 ms on the C4 genome (3.1 Gb).
    const bool f0 = __any_sync(0xFFFFFFFFu, nonascii != 0), f1 = __any_sync(0xFFFFFFFFu, other);
    if (lane == 0) {
        if (f0) *(volatile uint32_t *)&job.err_flags[0] = 1u;
        if (f1) *(volatile uint32_t *)&job.err_flags[1] = 1u;
    }
}

} // namespace tidk
