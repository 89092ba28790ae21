/* Drop-in C header for the A*PA2 entry points of libastarpa_c.
 *
 * Declares the same symbols, with the same argument meaning and ownership rules, as the reference's
 * cbindgen header astarpa-c/astarpa.h (implementation astarpa-c/src/lib.rs):
 *   astarpa2_simple      replaces astarpa-c/astarpa.h:15-20  (astarpa-c/src/lib.rs:8-24)
 *   astarpa2_full        replaces astarpa-c/astarpa.h:27-32  (astarpa-c/src/lib.rs:30-46)
 *   astarpa_free_cigar   replaces astarpa-c/astarpa.h:65     (astarpa-c/src/lib.rs:99-101)
 *   astarpa, astarpa_gcsh  (astarpa.h:41-60, A*PA v1) are exported so callers link unchanged; they are
 *                        served by the A*PA2 engine (same optimal cost, a valid CIGAR) — see INTEGRATION.md.
 * Inputs: raw bytes over ACGT, explicit lengths, not NUL-terminated, not retained. Output: the edit distance;
 * *cigar_ptr is callee-allocated NUL-terminated text ('=' 'X' 'I' 'D', count omitted when 1), *cigar_len its
 * strlen; release it with astarpa_free_cigar. Every call runs on the GPU (batch of one); there is no CPU
 * fallback: without a usable device the call aborts like the reference's panic does.
 */
#ifndef ASTARPA_H
#define ASTARPA_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

uint64_t astarpa2_simple(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                         uintptr_t* cigar_len);
uint64_t astarpa2_full(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                       uintptr_t* cigar_len);
uint64_t astarpa(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                 uintptr_t* cigar_len);
uint64_t astarpa_gcsh(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uintptr_t r, uintptr_t k,
                      bool prune_end, uint8_t** cigar_ptr, uintptr_t* cigar_len);
void astarpa_free_cigar(uint8_t* cigar);

#ifdef __cplusplus
}
#endif
#endif
