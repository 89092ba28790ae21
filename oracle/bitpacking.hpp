// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).
//
// CPU restatement of the `pa-bitpacking` crate parts on the hot path:
//   V, HEncoding(u64,u64)     pa-bitpacking/src/encoding.rs:4-74,143-169
//   BitProfile::build / eq    pa-bitpacking/src/profile.rs:112-147
//   myers::compute_block      pa-bitpacking/src/myers.rs:27-55
//   simd::compute / fill      pa-bitpacking/src/simd.rs:98-226,326-437 (semantics; any topological order of the
//                             same recurrence gives identical v/h/return — scalar::row, scalar.rs:37-46)
#pragma once
#include "types.hpp"

namespace oracle {

constexpr I W = 64;   // pa-bitpacking/src/lib.rs:45-48  (B = u64, W = 64)
constexpr I WI = 64;  // astarpa2/src/lib.rs:35

struct V {  // encoding.rs:4-38. bit r of p: D[j+r+1]-D[j+r] == +1 ; of m: == -1.
    uint64_t p, m;
    static V zero() { return {0, 0}; }
    static V one() { return {~0ull, 0}; }
    Cost value() const { return __builtin_popcountll(p) - __builtin_popcountll(m); }
    Cost value_of_prefix(I j) const {  // first j bits, 0 <= j < 64
        uint64_t mask = (1ull << j) - 1;
        return __builtin_popcountll(p & mask) - __builtin_popcountll(m & mask);
    }
    Cost value_of_suffix(I j) const {  // last j bits, 0 < j <= 64
        uint64_t mask = ~((1ull << (64 - j)) - 1);
        return __builtin_popcountll(p & mask) - __builtin_popcountll(m & mask);
    }
    bool operator==(const V& o) const { return p == o.p && m == o.m; }
};

struct H {  // HEncoding for (u64,u64): encoding.rs:143-169.  p, m are each 0 or 1.
    uint64_t p, m;
    static H zero() { return {0, 0}; }
    static H one() { return {1, 0}; }
    Cost value() const { return (Cost)p - (Cost)m; }
    bool operator==(const H& o) const { return p == o.p && m == o.m; }
};

struct Bits {  // profile.rs:89-111
    uint64_t b0, b1;
};

// RankTransform over "ACGT" (A0 C1 G2 T3); panics on anything else (profile.rs:113,117,126).
inline uint8_t rank_acgt(uint8_t c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
    }
    throw RefPanic("BitProfile::build: byte outside ACGT");
}

// profile.rs:112-133
inline void bitprofile_build(const uint8_t* a, size_t n, const uint8_t* b, size_t m, std::vector<Bits>& pa,
                             std::vector<Bits>& pb) {
    pa.resize(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t c = rank_acgt(a[i]);
        pa[i] = Bits{0ull - (c & 1), 0ull - ((c >> 1) & 1)};
    }
    pb.assign((m + W - 1) / W, Bits{0, 0});
    for (size_t j = 0; j < m; j++) {
        uint64_t c = rank_acgt(b[j]);
        pb[j / W].b0 |= ((c & 1) ^ 1) << (j % W);
        pb[j / W].b1 |= (((c >> 1) & 1) ^ 1) << (j % W);
    }
}
inline uint64_t profile_eq(const Bits& ca, const Bits& cb) { return (ca.b0 ^ cb.b0) & (ca.b1 ^ cb.b1); }
inline bool profile_is_match(const std::vector<Bits>& a, const std::vector<Bits>& b, I i, I j) {
    return (profile_eq(a[i], b[j / W]) >> (j % W)) & 1;
}

// myers.rs:27-55 — one 64-row x 1-column step. Note `eq |= h0.m` (input delta may be -1).
inline void myers_compute_block(H& h0, V& v, const Bits& ca, const Bits& cb) {
    uint64_t eq = profile_eq(ca, cb);
    uint64_t vp = v.p, vm = v.m;
    uint64_t vx = eq | vm;
    eq |= h0.m;
    uint64_t hx = (((eq & vp) + vp) ^ vp) | eq;
    uint64_t hp = vm | ~(hx | vp);
    uint64_t hm = vp & hx;
    uint64_t hpw = hp >> (W - 1);
    uint64_t hmw = hm >> (W - 1);
    hp = (hp << 1) | h0.p;
    hm = (hm << 1) | h0.m;
    h0 = H{hpw, hmw};
    v = V{hm | ~(vx | hp), hp & vx};
}

// simd::compute semantics (simd.rs:98-226): rectangle a[0..na) x b[0..nb) words.
// h: top deltas in, bottom deltas out; v: left deltas in, right deltas out. Returns sum of bottom h.
// (The reference's non-exact padded mode returns the same number and leaves h unspecified; every caller
//  that reads h uses exact mode, blocks.rs:740-746. We always compute exactly.)
inline Cost bp_compute(const Bits* a, size_t na, const Bits* b, size_t nb, H* h, V* v) {
    for (size_t j = 0; j < nb; j++) {
        V vj = v[j];
        const Bits cb = b[j];
        for (size_t i = 0; i < na; i++) myers_compute_block(h[i], vj, a[i], cb);
        v[j] = vj;
    }
    Cost s = 0;
    for (size_t i = 0; i < na; i++) s += h[i].value();
    return s;
}

// simd::fill semantics (simd.rs:326-437): as compute, but values[i][j] = v of word j after column i.
inline Cost bp_fill(const Bits* a, size_t na, const Bits* b, size_t nb, H* h, V* v, std::vector<std::vector<V>>& values) {
    for (size_t i = 0; i < na; i++) values[i].resize(nb);
    for (size_t j = 0; j < nb; j++) {
        V vj = v[j];
        const Bits cb = b[j];
        for (size_t i = 0; i < na; i++) {
            myers_compute_block(h[i], vj, a[i], cb);
            values[i][j] = vj;
        }
        v[j] = vj;
    }
    Cost s = 0;
    for (size_t i = 0; i < na; i++) s += h[i].value();
    return s;
}

}  // namespace oracle
