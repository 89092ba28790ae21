// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
//
// CPU restatement of the external `pa-types` crate (git e02484de…, NOT vendored under /root/reference;
// Cargo.lock:1072-1074), reconstructed from its call sites in the reference:
//   Pos / I / Cost            astarpa2/src/band.rs:13, astarpa2/src/blocks/trace.rs:443
//   partial order on Pos      pa-heuristic/src/contour/rotate_to_front.rs:35,51
//   LexPos                    pa-heuristic/src/matches.rs:249-251, prune.rs:141
//   Cigar::{push_elem,reverse,to_string,verify}  astarpa2/src/blocks/trace.rs:31,129,132; pa-test/src/lib.rs:98
//   push_elem merge semantics: affine twin pa-affine-types/src/cigar.rs:140-149
//   to_string format pinned by astarpa-c/example.cpp:16 ("=I4=X=": count omitted when 1)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

using I = int32_t;
using Cost = int32_t;
constexpr I I_MAX = INT32_MAX;
constexpr I I_MIN = INT32_MIN;

// A reference `panic!`/`assert!` that would abort the Rust process.
struct RefPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define ORACLE_ASSERT(cond, msg)                                                         \
    do {                                                                                 \
        if (!(cond)) throw ::oracle::RefPanic(std::string("reference assert failed: ") + (msg)); \
    } while (0)

struct Pos {
    I i, j;  // Pos(i, j): i indexes `a` (columns), j indexes `b` (rows). astarpa2/src/domain.rs:70-73
    bool operator==(const Pos& o) const { return i == o.i && j == o.j; }
    bool operator!=(const Pos& o) const { return !(*this == o); }
};
// Partial order: p <= q  <=>  both coordinates <=.
inline bool pos_le(Pos p, Pos q) { return p.i <= q.i && p.j <= q.j; }
// LexPos total order.
inline bool lex_less(Pos p, Pos q) { return p.i != q.i ? p.i < q.i : p.j < q.j; }

// Rust integer helpers (semantics the reference relies on).
inline I next_multiple_of(I x, I m) {  // smallest multiple of m that is >= x (m > 0)
    I r = x % m;
    if (r < 0) r += m;
    return r == 0 ? x : x + (m - r);
}
inline I trunc_multiple(I x, I m) { return x / m * m; }  // C++ '/' truncates toward 0 like Rust
inline I div_ceil(I a, I b) {                            // signed div_ceil (int_roundings), b > 0
    I q = a / b, r = a % b;
    return (r > 0) ? q + 1 : q;
}

enum CigarOp : uint8_t { OpMatch = 0, OpSub = 1, OpDel = 2, OpIns = 3 };
inline char op_char(CigarOp op) {
    switch (op) {
        case OpMatch: return '=';
        case OpSub: return 'X';
        case OpIns: return 'I';
        case OpDel: return 'D';
    }
    return '?';
}
struct CigarElem {
    CigarOp op;
    I cnt;
};
struct Cigar {
    std::vector<CigarElem> ops;
    void push_elem(CigarElem e) {
        if (!ops.empty() && ops.back().op == e.op) {
            ops.back().cnt += e.cnt;
            return;
        }
        ops.push_back(e);
    }
    void reverse() { std::reverse(ops.begin(), ops.end()); }
    std::string to_string() const {
        std::string s;
        char buf[16];
        for (auto& e : ops) {
            if (e.cnt != 1) {
                snprintf(buf, sizeof buf, "%d", e.cnt);
                s += buf;
            }
            s += op_char(e.op);
        }
        return s;
    }
    // Cigar::verify(&CostModel::unit(), a, b): walks a and b; returns cost or -1 when invalid.
    // Ins consumes b (trace.rs:176-182), Del consumes a (trace.rs:198-204).
    Cost verify(const uint8_t* a, size_t n, const uint8_t* b, size_t m) const {
        size_t i = 0, j = 0;
        Cost cost = 0;
        for (auto& e : ops) {
            if (e.cnt <= 0) return -1;
            for (I c = 0; c < e.cnt; c++) {
                switch (e.op) {
                    case OpMatch:
                        if (i >= n || j >= m || a[i] != b[j]) return -1;
                        i++, j++;
                        break;
                    case OpSub:
                        if (i >= n || j >= m || a[i] == b[j]) return -1;
                        i++, j++, cost++;
                        break;
                    case OpIns:
                        if (j >= m) return -1;
                        j++, cost++;
                        break;
                    case OpDel:
                        if (i >= n) return -1;
                        i++, cost++;
                        break;
                }
            }
        }
        if (i != n || j != m) return -1;
        return cost;
    }
};

// ---------------------------------------------------------------------------------------------
// Independent ground truth (stands in for triple_accel::levenshtein_exp, pa-test/src/lib.rs:74):
// plain full-matrix Myers/Hyyrö bit-vector Levenshtein, O(n*m/64), no banding, no heuristics.
inline Cost levenshtein_bitvector(const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
    if (n == 0) return (Cost)m;
    if (m == 0) return (Cost)n;
    // pattern = b (rows), text = a (columns).
    size_t words = (m + 63) / 64;
    std::vector<uint64_t> peq(256 * words, 0);
    for (size_t j = 0; j < m; j++) peq[(size_t)b[j] * words + j / 64] |= 1ull << (j % 64);
    std::vector<uint64_t> pv(words, ~0ull), mv(words, 0);
    Cost score = (Cost)m;
    const uint64_t last_bit = 1ull << ((m - 1) % 64);
    for (size_t i = 0; i < n; i++) {
        const uint64_t* eqc = &peq[(size_t)a[i] * words];
        uint64_t hp_carry = 1, hm_carry = 0;  // top row: D[i][0] = i  => horizontal delta +1
        for (size_t w = 0; w < words; w++) {
            uint64_t eq = eqc[w];
            uint64_t vp = pv[w], vm = mv[w];
            uint64_t xv = eq | vm;
            eq |= hm_carry;
            uint64_t xh = (((eq & vp) + vp) ^ vp) | eq;
            uint64_t hp = vm | ~(xh | vp);
            uint64_t hm = vp & xh;
            if (w == words - 1) {
                if (hp & last_bit) score++;
                if (hm & last_bit) score--;
            }
            uint64_t hp_out = hp >> 63, hm_out = hm >> 63;
            hp = (hp << 1) | hp_carry;
            hm = (hm << 1) | hm_carry;
            hp_carry = hp_out;
            hm_carry = hm_out;
            pv[w] = hm | ~(xv | hp);
            mv[w] = hp & xv;
        }
    }
    return score;
}

// Textbook O(n*m) DP, for tiny inputs only (cross-check of the bit-vector ground truth).
inline Cost levenshtein_dp(const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
    std::vector<Cost> prev(m + 1), cur(m + 1);
    for (size_t j = 0; j <= m; j++) prev[j] = (Cost)j;
    for (size_t i = 1; i <= n; i++) {
        cur[0] = (Cost)i;
        for (size_t j = 1; j <= m; j++) {
            Cost s = prev[j - 1] + (a[i - 1] != b[j - 1]);
            cur[j] = std::min({s, prev[j] + 1, cur[j - 1] + 1});
        }
        std::swap(prev, cur);
    }
    return prev[m];
}

}  // namespace oracle
