// pf_pack_simd.cpp -- AVX2 body of the host-side 2-bit packer (pf_pack.cu); compiled by g++ with -mavx2 and only
// called after pf_cpu_has_avx2().  Same result as the table-driven scalar loop: base j of a read at bits
// [2j, 2j+2) of its little-endian word stream (A,C,G,T = 0..3), false if any byte is not upper-case A/C/G/T.
#include <immintrin.h>

#include <cstdint>

namespace pf {

bool cpu_has_avx2() {
    static const bool has = __builtin_cpu_supports("avx2");
    return has;
}

// Packs the first n32*32 bases of s into dst[0 .. 2*n32).  Returns false if a byte outside {A,C,G,T} was seen (dst is
// then unspecified; the caller clears it).
bool pack_groups_avx2(const uint8_t *s, uint64_t n32, uint32_t *dst) {
    const __m256i vA = _mm256_set1_epi8('A'), vC = _mm256_set1_epi8('C'), vG = _mm256_set1_epi8('G'), vT = _mm256_set1_epi8('T');
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i w1 = _mm256_set1_epi16(0x0401);      // bytes (1, 4): c0 + 4*c1
    const __m256i w2 = _mm256_set1_epi32(0x00100001);  // words (1, 16): t0 + 16*t1
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    __m256i all_ok = _mm256_set1_epi8(-1);
    for (uint64_t g = 0; g < n32; ++g) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + g * 32));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, vA), _mm256_cmpeq_epi8(v, vC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, vG), _mm256_cmpeq_epi8(v, vT)));
        all_ok = _mm256_and_si256(all_ok, ok);
        // A=0x41 C=0x43 G=0x47 T=0x54: ((c >> 1) ^ (c >> 2)) & 3 = 0,1,2,3 (bits shifted in from the neighbouring byte
        // land in bits 6-7 and are masked away)
        const __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), m3);
        const __m256i t = _mm256_maddubs_epi16(c, w1);  // 16-bit: c0 | c1 << 2
        const __m256i b = _mm256_madd_epi16(t, w2);     // 32-bit: one packed byte (4 bases) per lane
        const __m256i p = _mm256_shuffle_epi8(b, gather);
        dst[2 * g] = (uint32_t)_mm256_extract_epi32(p, 0);
        dst[2 * g + 1] = (uint32_t)_mm256_extract_epi32(p, 4);
    }
    return _mm256_movemask_epi8(all_ok) == -1;
}

}  // namespace pf
