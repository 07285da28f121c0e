// Test shim: C entry points around host_pack.hpp (the host half of the two-ended upload) so that the
// CPU suite can check the packers against a numpy model without a GPU.
#include <cstdint>
#include <vector>

#include "../telomeric_identifier_b200/csrc/host_pack.hpp"

extern "C" {

// three planes, as PackedSource reads them
void shim_pack3(const uint8_t *src, uint64_t n_bytes, uint64_t n_words, uint32_t *lo, uint32_t *hi, uint32_t *va,
                uint32_t *nonascii, int force_scalar) {
    if (force_scalar) tidk::pack_words_scalar(src, n_bytes, n_words, lo, hi, va, nonascii);
    else tidk::pack_words(src, n_bytes, n_words, lo, hi, va, nonascii);
}

// two planes + validity exceptions; returns the number of exceptions (at most cap are written)
uint64_t shim_pack2(const uint8_t *src, uint64_t n_bytes, uint64_t n_words, uint32_t *lo, uint32_t *hi, uint64_t word0,
                    uint64_t *exc, uint64_t cap, uint32_t *nonascii, int force_scalar) {
    std::vector<uint64_t> v;
    // 0: the library's dispatch (AVX-512 where the machine has it), 1: scalar, 2: the AVX2 form
    if (force_scalar == 1) tidk::pack_words2_scalar(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    else if (force_scalar == 2 && __builtin_cpu_supports("avx2")) tidk::pack_words2_avx2(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    else tidk::pack_words2(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    for (uint64_t i = 0; i < v.size() && i < cap; i++) exc[i] = v[i];
    return v.size();
}

// the explore upload's STRICT form: valid = upper-case A C G T only (exception-free words determine their bytes)
uint64_t shim_pack2_strict(const uint8_t *src, uint64_t n_bytes, uint64_t n_words, uint32_t *lo, uint32_t *hi, uint64_t word0,
                           uint64_t *exc, uint64_t cap, uint32_t *nonascii, int force_scalar) {
    std::vector<uint64_t> v;
    if (force_scalar == 1) tidk::pack_words2_scalar<true>(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    else if (force_scalar == 2 && __builtin_cpu_supports("avx2")) tidk::pack_words2_avx2<true>(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    else tidk::pack_words2<true>(src, n_bytes, n_words, lo, hi, word0, v, nonascii);
    for (uint64_t i = 0; i < v.size() && i < cap; i++) exc[i] = v[i];
    return v.size();
}
}
