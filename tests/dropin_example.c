/* Drop-in check of libastarpa_c through the C header, in the spirit of the reference's astarpa-c/example.c
 * (same pair, cost 2 through every entry point). Built and run by tests/test_gpu_parity.py on the GPU box. */
#include <stdio.h>
#include <string.h>

#include "astarpa.h"

typedef uint64_t (*align_fn)(const uint8_t*, uintptr_t, const uint8_t*, uintptr_t, uint8_t**, uintptr_t*);

static int check(const char* name, uint64_t cost, uint8_t* cigar, uintptr_t len) {
    int ok = cost == 2 && cigar != NULL && strlen((const char*)cigar) == len && len > 0;
    printf("%s cost=%lu cigar=%s len=%lu %s\n", name, (unsigned long)cost, cigar ? (const char*)cigar : "(null)", (unsigned long)len,
           ok ? "ok" : "FAIL");
    astarpa_free_cigar(cigar);
    return ok;
}

int main(void) {
    const uint8_t* a = (const uint8_t*)"ACTCGCT";
    const uint8_t* b = (const uint8_t*)"AACTCGTT";
    const uintptr_t n = 7, m = 8;
    align_fn fns[3] = {astarpa2_simple, astarpa2_full, astarpa};
    const char* names[3] = {"astarpa2_simple", "astarpa2_full", "astarpa"};
    int ok = 1;
    for (int k = 0; k < 3; k++) {
        uint8_t* cigar = NULL;
        uintptr_t len = 0;
        uint64_t cost = fns[k](a, n, b, m, &cigar, &len);
        ok &= check(names[k], cost, cigar, len);
    }
    uint8_t* cigar = NULL;
    uintptr_t len = 0;
    uint64_t cost = astarpa_gcsh(a, n, b, m, 1, 15, false, &cigar, &len);
    ok &= check("astarpa_gcsh", cost, cigar, len);
    return ok ? 0 : 1;
}
