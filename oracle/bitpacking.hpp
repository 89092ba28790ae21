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

// ---- anti-diagonal SIMD strip, mirroring the reference's layout: L = 4 lanes of u64 x N = 2 vectors = 8 words per
// strip (pa-bitpacking/src/lib.rs:54, astarpa2/src/blocks.rs:721, simd.rs:229-315). Lane q of the strip holds word
// 7 - q, so that at "time" t the lanes touch the contiguous columns t-7 .. t of a and h. The corner triangles are done
// with the scalar step, as in the reference (simd.rs:243-247,310-314). Any topological order yields identical v / h.
typedef uint64_t v4u __attribute__((vector_size(32)));
struct Unzipped {  // "Unzip bits of a so we can directly use unaligned reads later" (simd.rs:137-139)
    std::vector<uint64_t> a0, a1, hp, hm;
};
static inline v4u loadu(const uint64_t* p) {
    v4u r;
    __builtin_memcpy(&r, p, 32);
    return r;
}
static inline void storeu(uint64_t* p, v4u v) { __builtin_memcpy(p, &v, 32); }
static inline void myers_step_v4(v4u& hp0, v4u& hm0, v4u& vp, v4u& vm, v4u eq) {  // myers.rs:61-91
    v4u vx = eq | vm;
    v4u eq2 = eq | hm0;
    v4u hx = (((eq2 & vp) + vp) ^ vp) | eq2;
    v4u hp = vm | ~(hx | vp);
    v4u hm = vp & hx;
    v4u hpw = hp >> 63;
    v4u hmw = hm >> 63;
    hp = (hp << 1) | hp0;
    hm = (hm << 1) | hm0;
    hp0 = hpw;
    hm0 = hmw;
    vp = hm | ~(vx | hp);
    vm = hp & vx;
}
// One strip of 8 words over all na columns. h lives in uz.hp / uz.hm (0/1 per column).
inline void strip8(const Bits* a, Unzipped& uz, size_t na, const Bits* b8, V* v8) {
    auto scalar = [&](size_t i, size_t j) {
        H h{uz.hp[i], uz.hm[i]};
        myers_compute_block(h, v8[j], a[i], b8[j]);
        uz.hp[i] = h.p;
        uz.hm[i] = h.m;
    };
    if (na < 16) {
        for (size_t j = 0; j < 8; j++)
            for (size_t i = 0; i < na; i++) scalar(i, j);
        return;
    }
    // top-left triangle: word j, columns 0 .. 6-j
    for (size_t j = 0; j < 8; j++)
        for (size_t i = 0; i + j < 7; i++) scalar(i, j);
    // lanes: q = 0..7 <-> word 7 - q ; vectors lo = lanes 0..3 (words 7..4), hi = lanes 4..7 (words 3..0)
    v4u b0lo, b1lo, b0hi, b1hi, vplo, vmlo, vphi, vmhi;
    for (int q = 0; q < 4; q++) {
        b0lo[q] = b8[7 - q].b0;
        b1lo[q] = b8[7 - q].b1;
        vplo[q] = v8[7 - q].p;
        vmlo[q] = v8[7 - q].m;
        b0hi[q] = b8[3 - q].b0;
        b1hi[q] = b8[3 - q].b1;
        vphi[q] = v8[3 - q].p;
        vmhi[q] = v8[3 - q].m;
    }
    const uint64_t* a0 = uz.a0.data();
    const uint64_t* a1 = uz.a1.data();
    uint64_t* hp = uz.hp.data();
    uint64_t* hm = uz.hm.data();
    for (size_t t = 7; t < na; t++) {  // lane q handles column t - 7 + q
        const size_t c = t - 7;
        v4u eqlo = (loadu(a0 + c) ^ b0lo) & (loadu(a1 + c) ^ b1lo);
        v4u eqhi = (loadu(a0 + c + 4) ^ b0hi) & (loadu(a1 + c + 4) ^ b1hi);
        v4u hplo = loadu(hp + c), hmlo = loadu(hm + c), hphi = loadu(hp + c + 4), hmhi = loadu(hm + c + 4);
        myers_step_v4(hplo, hmlo, vplo, vmlo, eqlo);
        myers_step_v4(hphi, hmhi, vphi, vmhi, eqhi);
        storeu(hp + c, hplo);
        storeu(hm + c, hmlo);
        storeu(hp + c + 4, hphi);
        storeu(hm + c + 4, hmhi);
    }
    for (int q = 0; q < 4; q++) {
        v8[7 - q] = V{vplo[q], vmlo[q]};
        v8[3 - q] = V{vphi[q], vmhi[q]};
    }
    // bottom-right triangle: word j, columns na-j .. na-1
    for (size_t j = 1; j < 8; j++)
        for (size_t i = na - j; i < na; i++) scalar(i, j);
}

// simd::compute semantics (simd.rs:98-226): rectangle a[0..na) x b[0..nb) words.
// h: top deltas in, bottom deltas out; v: left deltas in, right deltas out. Returns sum of bottom h.
// (The reference's non-exact padded mode returns the same number and leaves h unspecified; every caller
//  that reads h uses exact mode, blocks.rs:740-746. We always compute exactly.)
inline Cost bp_compute(const Bits* a, size_t na, const Bits* b, size_t nb, H* h, V* v) {
    size_t j = 0;
    if (nb >= 8 && na >= 16) {
        static thread_local Unzipped uz;
        uz.a0.resize(na);
        uz.a1.resize(na);
        uz.hp.resize(na);
        uz.hm.resize(na);
        for (size_t i = 0; i < na; i++) {
            uz.a0[i] = a[i].b0;
            uz.a1[i] = a[i].b1;
            uz.hp[i] = h[i].p;
            uz.hm[i] = h[i].m;
        }
        for (; j + 8 <= nb; j += 8) strip8(a, uz, na, b + j, v + j);
        for (size_t i = 0; i < na; i++) h[i] = H{uz.hp[i], uz.hm[i]};
    }
    for (; j < nb; j++) {  // remaining rows: scalar::row (scalar.rs:37-46)
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
