// Host-side packing of bases into the device layout: per 32 bases one (plane0, plane1) pair of 32-bit words holding
// the NEGATED rank bits of A0 C1 G2 T3 (BitProfile::build for b, pa-bitpacking/src/profile.rs:124-131, split in 32-row
// halves). Packing on the host cuts the H2D traffic 4x (0.25 B/base instead of 1 B/base); it is also where the input is
// validated (the reference panics on bytes outside ACGT, profile.rs:113).
#include <immintrin.h>
#include <stdint.h>

// AVX2 (and AVX-512BW when the CPU has it, chosen at run time). Packs half-words [hw_begin, hw_end) of `seq` (len bases) into out[2*hw], out[2*hw+1]. Half-words past the end of the
// sequence are zero (padding). Returns 0, or 1 if a byte outside ACGT was seen.
// AVX-512BW body for whole 64-base groups (two half-words per iteration): rank bits straight into mask registers.
// Returns the first half-word it did not pack.
__attribute__((target("avx512f,avx512bw"))) static int64_t pack_planes_avx512(const uint8_t* seq, int64_t len, int64_t hw_begin,
                                                                             int64_t hw_end, uint32_t* out, int* bad) {
    int64_t hw = hw_begin;
    if (hw & 1) return hw;  // callers start on even half-words except for tails
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T');
    const __m512i b1 = _mm512_set1_epi8(2), b2 = _mm512_set1_epi8(4);
    __mmask64 all_ok = ~0ull;
    for (; hw + 2 <= hw_end && (hw + 2) * 32 <= len; hw += 2) {
        const __m512i v = _mm512_loadu_si512(seq + hw * 32);
        const uint64_t c1 = _mm512_test_epi8_mask(v, b1), c2 = _mm512_test_epi8_mask(v, b2);  // c.bit1, c.bit2 per base
        const uint64_t p0 = ~(c1 ^ c2), p1 = ~c2;  // negated rank bits: rank.bit0 = c.bit1 ^ c.bit2, rank.bit1 = c.bit2
        out[2 * hw] = (uint32_t)p0;
        out[2 * hw + 1] = (uint32_t)p1;
        out[2 * hw + 2] = (uint32_t)(p0 >> 32);
        out[2 * hw + 3] = (uint32_t)(p1 >> 32);
        all_ok &= _mm512_cmpeq_epi8_mask(v, cA) | _mm512_cmpeq_epi8_mask(v, cC) | _mm512_cmpeq_epi8_mask(v, cG) | _mm512_cmpeq_epi8_mask(v, cT);
    }
    if (all_ok != ~0ull) *bad = 1;
    return hw;
}

extern "C" int apa_pack_planes_host(const uint8_t* seq, int64_t len, int64_t hw_begin, int64_t hw_end, uint32_t* out) {
    int bad = 0;
    static const bool has512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
    if (has512) hw_begin = pack_planes_avx512(seq, len, hw_begin, hw_end, out, &bad);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    for (int64_t hw = hw_begin; hw < hw_end; hw++) {
        const int64_t j0 = hw * 32;
        uint32_t p0 = 0, p1 = 0;
        if (j0 + 32 <= len) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(seq + j0));
            // x = (c >> 1) & 3 -> A0 C1 T2 G3; rank = x ^ (x >> 1): bit0 = c.bit1 ^ c.bit2, bit1 = c.bit2
            const __m256i s5 = _mm256_slli_epi16(v, 5);  // c.bit2 -> bit7 of each byte
            const __m256i s6 = _mm256_slli_epi16(v, 6);  // c.bit1 -> bit7
            const uint32_t bit1 = (uint32_t)_mm256_movemask_epi8(s5);
            const uint32_t bit0 = (uint32_t)_mm256_movemask_epi8(_mm256_xor_si256(s5, s6));
            p0 = ~bit0;
            p1 = ~bit1;
            const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                               _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
            if ((uint32_t)_mm256_movemask_epi8(ok) != 0xffffffffu) bad = 1;
        } else {
            for (int t = 0; t < 32; t++) {
                const int64_t j = j0 + t;
                if (j < len) {
                    const uint32_t c = seq[j];
                    if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) bad = 1;
                    const uint32_t x = (c >> 1) & 3u, r = x ^ (x >> 1);
                    p0 |= ((r & 1u) ^ 1u) << t;
                    p1 |= ((r >> 1) ^ 1u) << t;
                }
            }
        }
        out[2 * hw] = p0;
        out[2 * hw + 1] = p1;
    }
    return bad;
}
