// helper of tools/h2d_dirty_test.py: copy src -> dst (cache-dirtying stores), optionally followed by a write-back of the lines
#include <immintrin.h>
#include <stdint.h>
#include <string.h>
#include <cpuid.h>
int has_clwb(void) { unsigned a, b, c, d; if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return 0; return ((b >> 24) & 1) | (((b >> 23) & 1) << 1); }
void fill(char* dst, const char* src, size_t bytes, int mode) {
    memcpy(dst, src, bytes);
    if (mode == 1) { for (size_t o = 0; o < bytes; o += 64) _mm_clwb(dst + o); _mm_sfence(); }
    if (mode == 2) { for (size_t o = 0; o < bytes; o += 64) _mm_clflushopt(dst + o); _mm_sfence(); }
    if (mode == 3) { for (size_t o = 0; o < bytes; o += 64) _mm_clflush(dst + o); _mm_sfence(); }
}
void fill_nt(char* dst, const char* src, size_t bytes) {
    for (size_t o = 0; o < bytes; o += 32) _mm256_stream_si256((__m256i*)(dst + o), _mm256_loadu_si256((const __m256i*)(src + o)));
    _mm_sfence();
}
