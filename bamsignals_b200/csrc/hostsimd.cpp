#include "hostsimd.h"

#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace bsg {

#if defined(__x86_64__)
namespace {
__attribute__((target("avx2"))) void widen_avx2(int32_t* dst, const uint8_t* src, int64_t n) {
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = int32_t(src[i]); ++i; }
    for (; i + 16 <= n; i += 16) {
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_cvtepu8_epi32(b));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(b, 8)));
    }
    for (; i < n; ++i) dst[i] = int32_t(src[i]);
    _mm_sfence();
}
__attribute__((target("avx2"))) void copy_avx2(int32_t* dst, const int32_t* src, int64_t n) {
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8)
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
}  // namespace
#endif

void widen_u8_to_i32(int32_t* dst, const uint8_t* src, int64_t n) {
#if defined(__x86_64__)
    if (have_avx2() && n >= 64) { widen_avx2(dst, src, n); return; }
#endif
    for (int64_t i = 0; i < n; ++i) dst[i] = int32_t(src[i]);
}

void copy_i32_stream(int32_t* dst, const int32_t* src, int64_t n) {
#if defined(__x86_64__)
    if (have_avx2() && n >= 1024) { copy_avx2(dst, src, n); return; }
#endif
    memcpy(dst, src, size_t(n) * 4);
}

}  // namespace bsg
