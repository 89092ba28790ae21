/* Synthetic sequence pairs for tests, bench.py and the pa-bin equivalent: stands in for the external `pa-generate` crate
 * (git 5eb506db..., not vendored in the reference tree; used at pa-test/src/lib.rs:43-63, pa-bin/src/lib.rs:118-126,
 * pa-bitpacking/benches/nw/main.rs:142-144). Host code only, built into libapa_generate.so - NOT part of libastarpa_c.so.
 * Stream: xoshiro256** seeded by splitmix64(seed); see csrc/generate.cpp for the mutation procedure. */
#ifndef APA_GENERATE_H
#define APA_GENERATE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Error models: 0 Uniform, 1 NoisyInsert, 2 NoisyDelete, 3 SymmetricRepeat (pa-test/src/lib.rs:42-47).
 * Returns |b| or -1 when b_cap is too small. */
int64_t apa_generate_pair(uint64_t n, double e, int model, uint64_t seed, uint8_t* a_out, uint8_t* b_out, uint64_t b_cap);
/* n_pairs pairs, pair p from seed0 + p, a's concatenated, b's at b_all + p * b_stride with lengths in b_len. */
int apa_generate_batch(uint64_t n_pairs, uint64_t n, double e, int model, uint64_t seed0, uint8_t* a_all, uint8_t* b_all,
                       uint64_t b_stride, int64_t* b_len, int n_threads);
/* FNV-1a (64 bit) of n_texts byte strings pool[off[p] .. off[p] + len[p]) - the digest tests and bench.py compare CIGAR texts by. */
void apa_fnv1a_batch(const char* pool, const int64_t* off, const int64_t* len, uint64_t n_texts, uint64_t* out);
#ifdef __cplusplus
}
#endif
#endif
