// Synthetic sequence-pair generator (host side, no GPU).
//
// Stands in for the external `pa-generate` crate (git 5eb506db…, not vendored in the reference tree;
// used at pa-test/src/lib.rs:43-63, pa-bin/src/lib.rs:118-126, pa-bitpacking/benches/nw/main.rs:142-144).
// Its RNG stream (ChaCha8) and exact mutation procedure are not reproducible here, so this is a
// statistically equivalent generator with a fixed, documented stream:
//   rng  = xoshiro256** seeded from splitmix64(seed)
//   a    = n i.i.d. uniform bases over ACGT
//   b    = a after floor(e*n) mutations; each mutation picks a uniform position of a and a uniform kind
//          {substitute by a different base, insert a uniform base before, delete}; applied in one
//          left-to-right pass (several mutations on one position are applied in draw order).
//   models: 0 Uniform, 1 NoisyInsert (+ one random block of floor(e*n/2)+1 bases inserted),
//           2 NoisyDelete (one block of floor(e*n/2)+1 bases deleted), 3 SymmetricRepeat (a tandem
//           repeat whose copy number differs between a and b), each on top of Uniform noise at rate e/2.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/apa_generate.h"

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto& v : s) v = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    uint64_t below(uint64_t n) { return n ? (uint64_t)(((__uint128_t)next() * n) >> 64) : 0; }
};

const char BASES[4] = {'A', 'C', 'G', 'T'};

struct Mut {
    uint32_t pos;
    uint32_t order;
    uint8_t kind;  // 0 sub, 1 ins, 2 del
    uint8_t base;
};

void mutate_uniform(const std::vector<uint8_t>& a, double e, Rng& rng, std::vector<uint8_t>& b) {
    size_t n = a.size();
    size_t k = (size_t)(e * (double)n);
    std::vector<Mut> muts(k);
    for (size_t t = 0; t < k; t++) {
        muts[t].pos = (uint32_t)rng.below(n);
        muts[t].order = (uint32_t)t;
        muts[t].kind = (uint8_t)rng.below(3);
        muts[t].base = (uint8_t)rng.below(4);
    }
    std::sort(muts.begin(), muts.end(), [](const Mut& x, const Mut& y) { return x.pos != y.pos ? x.pos < y.pos : x.order < y.order; });
    b.clear();
    b.reserve(n + k);
    size_t t = 0;
    for (size_t i = 0; i < n; i++) {
        bool present = true;
        uint8_t c = a[i];
        while (t < k && muts[t].pos == i) {
            const Mut& mu = muts[t++];
            if (mu.kind == 1) {
                b.push_back((uint8_t)BASES[mu.base]);
            } else if (mu.kind == 2) {
                present = false;
            } else if (present) {
                uint8_t nb = (uint8_t)BASES[mu.base];
                if (nb == c) nb = (uint8_t)BASES[(mu.base + 1) & 3];
                c = nb;
            }
        }
        if (present) b.push_back(c);
    }
}

void gen_pair(size_t n, double e, int model, uint64_t seed, std::vector<uint8_t>& a, std::vector<uint8_t>& b) {
    Rng rng(seed);
    a.resize(n);
    for (size_t i = 0; i < n; i++) a[i] = (uint8_t)BASES[rng.below(4)];
    if (model == 0 || n == 0) {
        mutate_uniform(a, e, rng, b);
        return;
    }
    size_t blk = (size_t)(e * (double)n / 2) + 1;
    if (model == 3) {
        // SymmetricRepeat: replace a window of a by copies of a short unit; b gets a different copy count.
        size_t unit = 1 + rng.below(std::min<size_t>(10, n));
        size_t copies_a = 1 + blk / unit, copies_b = copies_a + 1 + rng.below(3);
        size_t at = rng.below(n);
        std::vector<uint8_t> u(unit);
        for (auto& c : u) c = (uint8_t)BASES[rng.below(4)];
        std::vector<uint8_t> a2(a.begin(), a.begin() + at), b2;
        for (size_t c = 0; c < copies_a; c++) a2.insert(a2.end(), u.begin(), u.end());
        a2.insert(a2.end(), a.begin() + at, a.end());
        a2.resize(n);  // keep |a| == n
        a = a2;
        mutate_uniform(a, e / 2, rng, b2);
        size_t atb = std::min(at, b2.size());
        b.assign(b2.begin(), b2.begin() + atb);
        for (size_t c = 0; c < copies_b - copies_a; c++) b.insert(b.end(), u.begin(), u.end());
        b.insert(b.end(), b2.begin() + atb, b2.end());
        return;
    }
    std::vector<uint8_t> b2;
    mutate_uniform(a, e / 2, rng, b2);
    size_t at = rng.below(b2.size() + 1);
    if (model == 1) {
        b.assign(b2.begin(), b2.begin() + at);
        for (size_t t = 0; t < blk; t++) b.push_back((uint8_t)BASES[rng.below(4)]);
        b.insert(b.end(), b2.begin() + at, b2.end());
    } else {
        size_t end = std::min(b2.size(), at + blk);
        b.assign(b2.begin(), b2.begin() + at);
        b.insert(b.end(), b2.begin() + end, b2.end());
    }
}

}  // namespace

extern "C" {

int64_t apa_generate_pair(uint64_t n, double e, int model, uint64_t seed, uint8_t* a_out, uint8_t* b_out, uint64_t b_cap) {
    std::vector<uint8_t> a, b;
    gen_pair((size_t)n, e, model, seed, a, b);
    if (b.size() > b_cap) return -(int64_t)b.size();
    if (n) memcpy(a_out, a.data(), n);
    if (!b.empty()) memcpy(b_out, b.data(), b.size());
    return (int64_t)b.size();
}

// Batch: pair p uses seed seed0 + p (31415 is the reference's fixed seed, pa-test/src/lib.rs:51).
// a_all has n_pairs*n bytes; b_all has capacity n_pairs*b_stride; b_len[p] receives |b_p|.
int apa_generate_batch(uint64_t n_pairs, uint64_t n, double e, int model, uint64_t seed0, uint8_t* a_all, uint8_t* b_all,
                       uint64_t b_stride, int64_t* b_len, int n_threads) {
    std::atomic<uint64_t> next{0};
    std::atomic<int> bad{0};
    auto worker = [&]() {
        std::vector<uint8_t> a, b;
        for (;;) {
            uint64_t p = next.fetch_add(1);
            if (p >= n_pairs) break;
            gen_pair((size_t)n, e, model, seed0 + p, a, b);
            if (b.size() > b_stride) {
                bad = 1;
                b_len[p] = -1;
                continue;
            }
            if (n) memcpy(a_all + p * n, a.data(), n);
            if (!b.empty()) memcpy(b_all + p * b_stride, b.data(), b.size());
            b_len[p] = (int64_t)b.size();
        }
    };
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return bad.load();
}

}  // extern "C"

extern "C" void apa_fnv1a_batch(const char* pool, const int64_t* off, const int64_t* len, uint64_t n_texts, uint64_t* out) {
    for (uint64_t p = 0; p < n_texts; p++) {
        uint64_t h = 1469598103934665603ull;
        const unsigned char* s = (const unsigned char*)pool + off[p];
        for (int64_t k = 0; k < len[p]; k++) h = (h ^ s[k]) * 1099511628211ull;
        out[p] = h;
    }
}
