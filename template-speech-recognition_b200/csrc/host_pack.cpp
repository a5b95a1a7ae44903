// host_pack.cpp -- AVX-512 body of tsr_host_pack_edges (util.cu): 64 input bytes (eight (t,f) cells x 8 orientation
// bytes) -> one 64-bit mask of the non-zero bytes = the eight packed bytes, one load + one test + one store.  Compiled by
// g++ (not nvcc) with a per-function target attribute and chosen at run time; the SSE2 loop in util.cu stays the
// fallback.  Internal symbols: not part of the C ABI.
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>

extern "C" __attribute__((visibility("hidden"))) int tsr_hostpack_has_avx512(void) {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
}

extern "C" __attribute__((visibility("hidden"), target("avx512f,avx512bw"))) void tsr_hostpack_range_avx512(const uint8_t* X, uint8_t* bits,
                                                                                                        long long c0, long long c1) {
    long long i = c0;
    for (; i + 32 <= c1; i += 32) {                    // 256 input bytes -> 32 packed bytes
        const __m512i v0 = _mm512_loadu_si512(X + (i + 0) * 8), v1 = _mm512_loadu_si512(X + (i + 8) * 8);
        const __m512i v2 = _mm512_loadu_si512(X + (i + 16) * 8), v3 = _mm512_loadu_si512(X + (i + 24) * 8);
        const uint64_t m[4] = {(uint64_t)_mm512_test_epi8_mask(v0, v0), (uint64_t)_mm512_test_epi8_mask(v1, v1),
                               (uint64_t)_mm512_test_epi8_mask(v2, v2), (uint64_t)_mm512_test_epi8_mask(v3, v3)};
        memcpy(bits + i, m, 32);
    }
    for (; i + 8 <= c1; i += 8) {
        const __m512i v = _mm512_loadu_si512(X + i * 8);
        const uint64_t m = (uint64_t)_mm512_test_epi8_mask(v, v);
        memcpy(bits + i, &m, 8);
    }
    for (; i < c1; ++i) {
        uint32_t b = 0;
        for (int e = 0; e < 8; ++e) b |= (uint32_t)(X[i * 8 + e] != 0) << e;
        bits[i] = (uint8_t)b;
    }
}
#else
extern "C" int tsr_hostpack_has_avx512(void) { return 0; }
extern "C" void tsr_hostpack_range_avx512(const uint8_t*, uint8_t*, long long, long long) {}
#endif
