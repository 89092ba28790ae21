// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).
//
// CPU restatement of pa_bitpacking::search (semi-global search of a short pattern in a long text):
//   ScatterProfile::build / eq     pa-bitpacking/src/profile.rs:18-84   (text: acgtACGT only; pattern may hold N / * / Y / R)
//   scatter_profile::compute       pa-bitpacking/src/simd/scatter_profile.rs:6-120 (semantics: any topological order of
//                                  myers::compute_block over the rectangle gives the same h / v)
//   search                         pa-bitpacking/src/search.rs:46-118   (the `out` vector)
// Pinned by the reference's doc-test: search(b"AC", b"CTTACTTA", 0.0).out == [0,0,1,2,1,0,1,2,1,0,0] (search.rs:29-32).
#pragma once
#include <cmath>
#include <vector>

#include "bitpacking.hpp"

namespace oracle {

struct ScatterB {  // [B; 4]: per text base (A0 C1 T2 G3), the rows of the pattern word that match it
    uint64_t m[4];
};

inline uint8_t scatter_char(uint8_t c) {  // profile.rs:31-38
    switch (c) {
        case 'a': case 'A': return 0;
        case 'c': case 'C': return 1;
        case 't': case 'T': return 2;
        case 'g': case 'G': return 3;
    }
    throw RefPanic("ScatterProfile: text byte outside acgtACGT");
}
inline void scatter_mask(uint8_t c, uint64_t out[4]) {  // profile.rs:39-50
    static const uint8_t A[4] = {1, 0, 0, 0}, Cc[4] = {0, 1, 0, 0}, T[4] = {0, 0, 1, 0}, G[4] = {0, 0, 0, 1}, N[4] = {1, 1, 1, 1},
                         Y[4] = {0, 1, 1, 0}, R[4] = {1, 0, 0, 1};
    const uint8_t* s = nullptr;
    switch (c) {
        case 'a': case 'A': s = A; break;
        case 'c': case 'C': s = Cc; break;
        case 't': case 'T': s = T; break;
        case 'g': case 'G': s = G; break;
        case 'n': case 'N': case '*': s = N; break;
        case 'y': case 'Y': s = Y; break;
        case 'r': case 'R': s = R; break;
        default: throw RefPanic("ScatterProfile: unknown pattern base");
    }
    for (int i = 0; i < 4; i++) out[i] = s[i];
}
// profile.rs:28-66: a = text (columns), b = pattern (rows, 64 per word; the padding rows of the last word match everything)
inline void scatter_build(const uint8_t* a, size_t na, const uint8_t* b, size_t nb, std::vector<uint8_t>& pa, std::vector<ScatterB>& pb) {
    pa.resize(na);
    for (size_t i = 0; i < na; i++) pa[i] = scatter_char(a[i]);
    pb.assign((nb + 63) / 64, ScatterB{{0, 0, 0, 0}});
    for (size_t j = 0; j < nb; j++) {
        uint64_t mask[4];
        scatter_mask(b[j], mask);
        for (int i = 0; i < 4; i++) pb[j / 64].m[i] |= mask[i] << (j % 64);
    }
    for (size_t j = nb; j < (nb + 63) / 64 * 64; j++)
        for (int i = 0; i < 4; i++) pb[j / 64].m[i] |= 1ull << (j % 64);
}
inline void scatter_compute_block(H& h0, V& v, uint8_t ca, const ScatterB& cb) {  // myers.rs:27-55 with P = ScatterProfile
    uint64_t eq = cb.m[ca];
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

// search.rs:46-118. Returns `out`: the costs along the bottom row, then up the right column: |pattern| + |text| + 1 values.
inline std::vector<Cost> search(const uint8_t* pattern, size_t np, const uint8_t* text, size_t nt, float unmatched_cost) {
    std::vector<uint8_t> t;
    std::vector<ScatterB> p;
    scatter_build(text, nt, pattern, np, t, p);
    std::vector<H> h(nt, H::zero());
    std::vector<V> v0(p.size(), V::zero());
    const size_t padding = (np + 63) / 64 * 64 - np;
    if (!(unmatched_cost >= 0.0f && unmatched_cost <= 1.0f)) throw RefPanic("unmatched_cost must be in [0, 1]");
    if (unmatched_cost > 0.0f) {
        for (size_t i = 0;; i++) {
            size_t idx = (size_t)std::ceil((float)i / unmatched_cost);
            if (idx >= np) break;
            v0[idx / 64].p |= 1ull << (idx % 64);
        }
    }
    std::vector<V> v = v0;
    Cost bot_left = 0;
    for (const V& x : v) bot_left += x.value();
    for (size_t i = 0; i < nt; i++)
        for (size_t j = 0; j < p.size(); j++) scatter_compute_block(h[i], v[j], t[i], p[j]);
    Cost b = bot_left;
    std::vector<Cost> out{b};
    size_t skipped = 0;
    for (const H& x : h) {
        b += x.value();
        if (skipped < padding)
            skipped++;
        else
            out.push_back(b);
    }
    for (size_t w = v.size(); w-- > 0;) {  // "Fix since we round up to multiple of 64 chars."
        for (I j = 1; j <= 64; j++) {
            Cost delta = v[w].value_of_suffix(j);
            Cost unmatched = v0[w].value_of_suffix(j);
            Cost val = b - delta + unmatched;
            if (skipped < padding)
                skipped++;
            else
                out.push_back(val);
        }
        b -= v[w].value();
        b += v0[w].value();
    }
    if (out.size() != np + nt + 1) throw RefPanic("search: output length");
    return out;
}

// SearchResult::trace (search.rs:135-230): the alignment ending at out[idx]. The text window [end - width, end) is re-filled with
// every V column kept (width = 2 |pattern|, doubled until the cost at the end position equals the target: a window that starts
// past column 0 starts from all +1 deltas, an upper bound), then walked back greedily: matches first, then Del (a text base),
// Ins (a pattern base), Sub. Returns the CIGAR text; start receives the position the walk stopped at (poss[0] of the reference).
struct SearchTrace {
    std::string cigar;
    Pos start, end;
    Cost cost;
};
inline Cost v_value_to(const std::vector<V>& v, I j) {  // encoding.rs:54-63
    Cost s = 0;
    for (I w = 0; w < j / 64; w++) s += v[w].value();
    if (j % 64 != 0) s += v[j / 64].value_of_prefix(j % 64);
    return s;
}
inline Cost v_value_from(const std::vector<V>& v, I j) {  // encoding.rs:64-73
    Cost s = 0;
    if (j % 64 != 0) s += v[j / 64].value_of_suffix(64 - j % 64);
    for (size_t w = (size_t)((j + 63) / 64); w < v.size(); w++) s += v[w].value();
    return s;
}
inline SearchTrace search_trace(const uint8_t* pattern, size_t np, const uint8_t* text, size_t nt, float unmatched_cost, size_t idx) {
    const std::vector<Cost> out = search(pattern, np, text, nt, unmatched_cost);
    if (idx >= out.size()) throw RefPanic("trace: idx out of range");  // search.rs:122
    std::vector<uint8_t> t;
    std::vector<ScatterB> p;
    scatter_build(text, nt, pattern, np, t, p);
    std::vector<V> v0(p.size(), V::zero());
    if (unmatched_cost > 0.0f) {
        for (size_t i = 0;; i++) {
            size_t k = (size_t)std::ceil((float)i / unmatched_cost);
            if (k >= np) break;
            v0[k / 64].p |= 1ull << (k % 64);
        }
    }
    Pos pos = idx <= nt ? Pos{(I)idx, (I)np} : Pos{(I)nt, (I)(np - (idx - nt))};  // idx_to_pos, search.rs:121-132
    Cost target = out[idx];
    if ((size_t)pos.i == nt) target -= v_value_from(v0, pos.j);
    size_t width = 2 * np;
    const size_t end = (size_t)pos.i;
    size_t start;
    std::vector<std::vector<V>> fill;
    for (;;) {
        start = end > width ? end - width : 0;
        std::vector<H> h(end - start + 1, H::zero());
        std::vector<V> v = start == 0 ? v0 : std::vector<V>(v0.size(), V::one());
        fill.assign(h.size(), {});
        fill[0] = v;
        for (size_t i = start; i < end; i++) {
            for (size_t j = 0; j < p.size(); j++) scatter_compute_block(h[i - start + 1], v[j], t[i], p[j]);
            fill[i - start + 1] = v;
        }
        const Cost cost = v_value_to(v, pos.j);
        if (cost < target) throw RefPanic("trace: found a path cheaper than the target");  // assert, search.rs:176-179
        if (cost == target) break;
        if (start == 0) throw RefPanic("trace: target cost not reached with the whole text");
        width *= 2;
    }
    auto cost_at = [&](I i, I j) { return v_value_to(fill[(size_t)i - start], j); };
    auto is_match = [&](I i, I j) { return ((p[(size_t)j / 64].m[t[(size_t)i]] >> ((size_t)j % 64)) & 1ull) != 0; };  // profile.rs:72-74
    Cigar cigar;
    const Pos endpos = pos;
    Cost g = target;
    while (pos.i > (I)start && pos.j > 0) {
        I cnt = 0;
        while (pos.i > (I)start && pos.j > 0 && is_match(pos.i - 1, pos.j - 1)) cnt++, pos.i--, pos.j--;
        if (cnt > 0) {
            cigar.push_elem(CigarElem{OpMatch, cnt});
            continue;
        }
        if (cost_at(pos.i - 1, pos.j) == g - 1) {
            g--, pos.i--;
            cigar.push_elem(CigarElem{OpDel, 1});
            continue;
        }
        if (cost_at(pos.i, pos.j - 1) == g - 1) {
            g--, pos.j--;
            cigar.push_elem(CigarElem{OpIns, 1});
            continue;
        }
        if (cost_at(pos.i - 1, pos.j - 1) == g - 1) {
            g--, pos.i--, pos.j--;
            cigar.push_elem(CigarElem{OpSub, 1});
            continue;
        }
        throw RefPanic("Bad trace! Got stuck");
    }
    if (!(pos.i == 0 || g == 0)) throw RefPanic("trace: assert pos.0 == 0 || g == 0");  // search.rs:226
    cigar.reverse();
    return SearchTrace{cigar.to_string(), pos, endpos, target};
}

}  // namespace oracle
