// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).
//
// CPU restatement of the `pa-heuristic` slice used by astarpa2_{simple,full}:
//   HeuristicInstance (h, h_with_hint, prune_block, update_contours)  pa-heuristic/src/heuristic.rs:106-180
//   GapCost / NoCost                                                   heuristic/distances.rs:130-169
//   Seeds                                                              seeds.rs:20-157
//   QGrams                                                             matches/qgrams.rs:7-110
//   exact::hash_a / hash_to_smallvec                                   matches/exact.rs:15-69
//   MatchBuilder (push/sort/finish), CenteredVec                       matches.rs:93-333
//   preserve_for_local_pruning, extend_right{,_simd}                   matches/prepruning.rs:25-203
//   MatchPruner::{new,prune_block}, ActiveRange                        prune.rs:97-292
//   CSHI (GCSH): new, h, h_with_hint, distance, prune_block, update_contours   heuristic/csh.rs:152-554
//   HintContours<RotateToFrontContour>: new, score, score_with_hint, chain_score, update_layers
//                                                                      contour/hint_contours.rs:125-637
//   RotateToFrontContour                                               contour/rotate_to_front.rs:10-97
//   SplitVec (index/remove semantics = plain vector)                   split_vec.rs:15-128
#pragma once
#include <map>
#include <unordered_map>

#include "types.hpp"

namespace oracle {

using Layer = uint32_t;
constexpr Layer LAYER_MAX = UINT32_MAX;
constexpr Layer LAYER_MIN = 0;

struct Hint {  // hint_contours.rs:52-63
    Layer original_layer = LAYER_MAX;
};

struct HeuristicInstance {
    virtual ~HeuristicInstance() {}
    virtual Cost h(Pos pos) = 0;
    virtual std::pair<Cost, Hint> h_with_hint(Pos pos, Hint) { return {h(pos), Hint{}}; }  // heuristic.rs:126-128
    virtual void prune_block(I /*i_start*/, I /*i_end*/, I /*j_start*/, I /*j_end*/) {}    // heuristic.rs:149-151
    virtual void update_contours(Pos) {}                                                   // heuristic.rs:153-155
    // Introspection for tests.
    virtual size_t num_matches() const { return 0; }
    // Counters (not in the reference; for the CPU-baseline report).
    uint64_t n_h_calls = 0;
};

// heuristic/distances.rs:130-169
struct GapCostI : HeuristicInstance {
    Pos target;
    GapCostI(size_t n, size_t m) : target{(I)n, (I)m} {}
    static Cost gap(Pos from, Pos to) {
        int64_t d = (int64_t)(to.i - from.i) - (int64_t)(to.j - from.j);
        return (Cost)(d < 0 ? -d : d);
    }
    Cost h(Pos from) override {
        n_h_calls++;
        return gap(from, target);
    }
};
// NoCost (Dijkstra domain): h = 0.
struct NoCostI : HeuristicInstance {
    Cost h(Pos) override { return 0; }
};

// ------------------------------------------------------------------------------------------------ Seeds
using MatchCost = uint8_t;
struct Seed {
    I start, end;
    MatchCost seed_potential, seed_cost;
};
struct Seeds {  // seeds.rs:20-71
    std::vector<Seed> seeds;
    std::vector<I> seed_at;  // -1 = None
    std::vector<Cost> potential;
    std::vector<I> start_of_potential;
    Seeds() {}
    Seeds(size_t n, std::vector<Seed> sd) : seeds(std::move(sd)) {
        potential.assign(n + 1, 0);
        seed_at.assign(n + 1, -1);
        Cost cur = 0;
        I next = (I)seeds.size() - 1;  // iterate seeds in reverse
        start_of_potential.push_back((I)n);
        for (I i = (I)n; i >= 0; i--) {
            if (next >= 0) {
                const Seed& ns = seeds[next];
                if (i < ns.end) seed_at[i] = next;
                if (i == ns.start) {
                    cur += ns.seed_potential;
                    for (int t = 0; t < ns.seed_potential; t++) start_of_potential.push_back(i);
                    next--;
                }
            }
            potential[i] = cur;
        }
    }
    Cost pot(Pos p) const { return potential[p.i]; }
    Cost potential_distance(Pos from, Pos to) const {  // seeds.rs:84-88
        ORACLE_ASSERT(from.i <= to.i, "potential_distance");
        I end_i = seed_at[to.i] >= 0 ? seeds[seed_at[to.i]].start : to.i;
        return potential[from.i] - potential[end_i];
    }
    Pos transform(Pos pos) const {  // seeds.rs:140-143
        Cost p = pot(pos);
        return Pos{pos.i - pos.j - p, pos.j - pos.i - p};
    }
    Pos transform_back(Pos pos) const {  // seeds.rs:146-156
        if (pos.i == I_MAX && pos.j == I_MAX) return pos;
        I p = -(pos.i + pos.j) / 2;
        I i = start_of_potential[p];
        I diff = (pos.i - pos.j) / 2;
        return Pos{i, i - diff};
    }
    bool is_seed_start(Pos p) const { return seed_at[p.i] >= 0 && seeds[seed_at[p.i]].start == p.i; }
};

// ------------------------------------------------------------------------------------------------ QGrams
struct QGrams {  // matches/qgrams.rs
    const uint8_t* a;
    size_t n;
    const uint8_t* b;
    size_t m;
    static uint64_t char_to_bits(uint8_t c) { return (c >> 1) & 3; }  // A0 C1 T2 G3
    static uint64_t to_qgram(const uint8_t* s, I k) {                  // first char in the high-order bits
        uint64_t q = 0;
        for (I t = 0; t < k; t++) q = (q << 2) | char_to_bits(s[t]);
        return q;
    }
    std::vector<Seed> fixed_length_seeds(I k, MatchCost r) const {  // qgrams.rs:99-109
        std::vector<Seed> out;
        for (I i = 0; i < (I)n - k + 1; i += k) out.push_back(Seed{i, i + k, r, r});
        return out;
    }
};

// ------------------------------------------------------------------------------------------------ Matches
enum MatchStatus : uint8_t { Active, Pruned, PrePruned, Filtered };
struct Match {  // matches.rs:54-61
    Pos start, end;
    MatchCost match_cost, seed_potential;
    MatchStatus pruned;
    MatchCost score() const { return seed_potential - match_cost; }
    bool is_active() const { return pruned == Active; }
};

struct CenteredVec {  // matches.rs:94-127 — a vector centred on 0 that grows on demand; reads outside return the default.
    std::vector<I> vec;
    I def;
    explicit CenteredVec(I d) : vec(1, d), def(d) {}
    I index(I idx) const {
        int64_t k = (int64_t)idx + (int64_t)(vec.size() / 2);
        return (k < 0 || k >= (int64_t)vec.size()) ? def : vec[(size_t)k];
    }
    I& index_mut(I idx) {
        int64_t half = (int64_t)(vec.size() / 2);
        int64_t mag = idx < 0 ? -(int64_t)idx : (int64_t)idx;
        if (mag > half) {
            int64_t new_half = std::max<int64_t>(mag, (int64_t)vec.size());
            std::vector<I> nv((size_t)(2 * new_half + 1), def);
            std::copy(vec.begin(), vec.end(), nv.begin() + (new_half - half));
            vec.swap(nv);
            half = new_half;
        }
        return vec[(size_t)((int64_t)idx + half)];
    }
};

struct MatchConfig {  // matches.rs:388-399
    I k;
    MatchCost r;
    size_t local_pruning;
};

// prepruning.rs:25-32
inline bool extend_right(const uint8_t* a, size_t /*n*/, const uint8_t* b, size_t m, I& i, I j, I end_i) {
    while (i < end_i && j < (I)m && a[i] == b[j]) {
        i++;
        j++;
    }
    return i >= end_i;
}
// prepruning.rs:35-62 — note: the first-char test is bounded by |a| (not end_i) and the 32-wide loop may
// run past end_i; the return value is `i >= end_i` in every exit.
inline bool extend_right_simd(const uint8_t* a, size_t n, const uint8_t* b, size_t m, I& i, I j, I end_i) {
    if (i < (I)n && j < (I)m && a[i] == b[j]) {
        i++;
        j++;
    } else {
        return i >= end_i;
    }
    while (i < (I)n - 32 && j < (I)m - 32) {
        I cnt = 0;
        while (cnt < 32 && a[i + cnt] == b[j + cnt]) cnt++;
        i += cnt;
        j += cnt;
        if (cnt < 32) return i >= end_i;
        if (i >= end_i) return true;
    }
    return extend_right(a, n, b, m, i, j, end_i);
}

// prepruning.rs:95-203
inline bool preserve_for_local_pruning(const uint8_t* a, size_t n, const uint8_t* b, size_t m, const Seeds& seeds,
                                       const Match& mt, size_t p, std::vector<I>& fr, std::vector<I>& next_fr,
                                       CenteredVec& next_match_per_diag) {
    if (p == 0) return true;
    Pos s = mt.start, e = mt.end;
    Cost start_pot = seeds.pot(s);
    I seed_idx = seeds.seed_at[s.i];
    ORACLE_ASSERT(seed_idx >= 0, "match does not start in a seed");
    const Seed& last_seed = seeds.seeds[std::min((size_t)seed_idx + p - 1, seeds.seeds.size() - 1)];
    I end_i = last_seed.end;
    Cost end_pot = seeds.potential[end_i];
    size_t pd = (size_t)(start_pot - end_pot);

    // The reference `resize`s buffers that are reused between calls (stale contents survive); every entry
    // is reset to MIN or overwritten before it is read (boundary resets below + DT monotonicity), so keeping
    // the same reuse semantics is harmless. We mirror it literally.
    fr.resize(2 * pd + 1, I_MIN);
    next_fr.resize(2 * pd + 1, I_MIN);

    size_t d_lo = pd, d_hi = pd + 1;  // d_range = d_lo..d_hi (exclusive)
    fr[pd] = e.i;
    next_fr[pd] = I_MIN;

    if (extend_right_simd(a, n, b, m, fr[pd], e.j, end_i)) return true;
    if (next_match_per_diag.index(e.i - e.j) <= fr[pd]) return true;

    for (Cost g = 1 + mt.match_cost; g < (Cost)pd; g++) {
        fr[d_lo - 1] = I_MIN;
        fr[d_hi] = I_MIN;
        next_fr[d_lo - 1] = I_MIN;
        next_fr[d_hi] = I_MIN;
        // expand
        for (size_t d = d_lo; d < d_hi; d++) {
            next_fr[d - 1] = std::max(next_fr[d - 1], fr[d]);
            next_fr[d] = std::max(next_fr[d], fr[d] + 1);
            next_fr[d + 1] = std::max(next_fr[d + 1], fr[d] + 1);
        }
        std::swap(fr, next_fr);
        d_lo -= 1;
        d_hi += 1;
        // check & shrink
        while (d_lo < d_hi && g + seeds.potential[fr[d_lo]] >= start_pot) d_lo++;
        while (d_lo < d_hi && g + seeds.potential[fr[d_hi - 1]] >= start_pot) d_hi--;
        if (d_lo >= d_hi) return false;
        // extend
        for (size_t d = d_lo; d < d_hi; d++) {
            I& i = fr[d];
            I dd = e.i - e.j + ((I)d - (I)pd);
            I j = i - dd;
            I old_i = i;
            if (extend_right_simd(a, n, b, m, i, j, end_i)) return true;
            I nm = next_match_per_diag.index(dd);
            if (old_i <= nm && nm <= i) return true;
        }
    }
    return false;
}

struct Matches {
    Seeds seeds;
    std::vector<Match> matches;
    // stats (MatchStats, matches.rs:152-157)
    size_t pushed = 0, after_transform = 0, after_local_pruning = 0;
};

inline bool match_key_less(const Match& x, const Match& y) {  // matches.rs:249-251
    if (x.start != y.start) return lex_less(x.start, y.start);
    if (x.end != y.end) return lex_less(x.end, y.end);
    return x.match_cost < y.match_cost;
}

// find_matches -> exact::hash_a (matches.rs:17-39, exact.rs:15-69) with MatchBuilder (matches.rs:159-332).
inline Matches find_matches_hash_a(const uint8_t* a, size_t n, const uint8_t* b, size_t m, MatchConfig config,
                                   bool transform_filter) {
    ORACLE_ASSERT(config.r == 1, "hash_a requires r == 1");
    const I k = config.k;
    QGrams q{a, n, b, m};
    Matches out;
    out.seeds = Seeds(n, q.fixed_length_seeds(k, config.r));
    Seeds& seeds = out.seeds;
    const Pos transform_target = seeds.transform(Pos{(I)n, (I)m});
    CenteredVec next_match_per_diag(I_MAX);
    std::vector<I> fr, next_fr;

    // hash_to_smallvec: hash the chunked k-mers of a (SmallVec push order = increasing i). Flat open-addressing
    // table key -> first seed, duplicates chained in increasing i (stands in for FxHashMap<u32, SmallVec<[I; 2]>>).
    const size_t nseeds = n >= (size_t)k ? (n - k) / k + 1 : 0;
    size_t tsize = 16;
    while (tsize < 2 * nseeds + 2) tsize <<= 1;
    std::vector<uint32_t> tkey(tsize, 0);
    std::vector<I> thead(tsize, -1), ttail(tsize, -1), tnext(nseeds, -1);
    auto slot_of = [&](uint32_t key) { return (size_t)((key * 0x9E3779B1u) >> 7) & (tsize - 1); };
    for (size_t sidx = 0; sidx < nseeds; sidx++) {
        uint32_t key = (uint32_t)QGrams::to_qgram(a + sidx * k, k);
        size_t sl = slot_of(key);
        while (thead[sl] >= 0 && tkey[sl] != key) sl = (sl + 1) & (tsize - 1);
        if (thead[sl] < 0) {
            tkey[sl] = key;
            thead[sl] = ttail[sl] = (I)sidx;
        } else {
            tnext[ttail[sl]] = (I)sidx;
            ttail[sl] = (I)sidx;
        }
    }

    auto push = [&](Match mt) {  // MatchBuilder::push, matches.rs:205-247
        out.pushed++;
        if (transform_filter && !pos_le(seeds.transform(mt.start), transform_target)) return;
        out.after_transform++;
        if (config.local_pruning != 0 &&
            !preserve_for_local_pruning(a, n, b, m, seeds, mt, config.local_pruning, fr, next_fr, next_match_per_diag))
            return;
        out.after_local_pruning++;
        Seed& sd = seeds.seeds[seeds.seed_at[mt.start.i]];
        sd.seed_cost = std::min(sd.seed_cost, mt.match_cost);
        if (config.local_pruning != 0) {
            I d = mt.start.i - mt.start.j;
            I& old = next_match_per_diag.index_mut(d);
            ORACLE_ASSERT(old >= mt.start.i, "Matches should be added in reverse order on each diagonal.");
            old = mt.start.i;
        }
        out.matches.push_back(mt);
    };

    // b_qgrams_rev: all windows of b, right to left (qgrams.rs:81-97).
    {
        // rolling key, right to left: q >>= 2; q |= bits(c) << 2(k-1)  (qgrams.rs:81-97)
        uint64_t q = 0;
        const unsigned leftshift = 2 * (unsigned)(k - 1);
        for (I j = (I)m - 1; j >= 0; j--) {
            q = (q >> 2) | (QGrams::char_to_bits(b[j]) << leftshift);
            if (j > (I)m - k) continue;
            uint32_t key = (uint32_t)q;
            size_t sl = slot_of(key);
            while (thead[sl] >= 0 && tkey[sl] != key) sl = (sl + 1) & (tsize - 1);
            if (thead[sl] < 0) continue;
            for (I sidx = thead[sl]; sidx >= 0; sidx = tnext[sidx]) {
                I i = sidx * k;
                push(Match{Pos{i, j}, Pos{i + k, j + k}, 0, 1, Active});
            }
        }
    }
    // matches.sort(); finish(): sort again, dedup by (start,end) keeping the first (lowest cost).
    std::stable_sort(out.matches.begin(), out.matches.end(), match_key_less);
    out.matches.erase(std::unique(out.matches.begin(), out.matches.end(),
                                  [](const Match& x, const Match& y) { return x.start == y.start && x.end == y.end; }),
                      out.matches.end());
    return out;
}

// ------------------------------------------------------------------------------------------------ MatchPruner
struct ActiveRange {  // prune.rs:97-102
    I col;
    size_t before_start, before_end;
    bool has_after = false;
    size_t after_start = 0, after_end = 0;
};

struct MatchPruner {  // prune.rs:109-292 (Prune::Start only)
    std::vector<Match> by_start;
    std::map<std::pair<I, I>, std::pair<size_t, size_t>> start_index;
    std::vector<ActiveRange> active_range;

    MatchPruner() {}
    MatchPruner(std::vector<Match> matches_by_start, const Seeds& seeds) {
        by_start = std::move(matches_by_start);
        std::stable_sort(by_start.begin(), by_start.end(), [](const Match& x, const Match& y) {
            if (x.start != y.start) return lex_less(x.start, y.start);
            return x.match_cost < y.match_cost;
        });
        for (size_t idx = 0; idx < by_start.size();) {
            size_t e = idx;
            while (e < by_start.size() && by_start[e].start == by_start[idx].start) e++;
            start_index[{by_start[idx].start.i, by_start[idx].start.j}] = {idx, e};
            idx = e;
        }
        size_t idx = 0;
        for (const Seed& s : seeds.seeds) {
            ActiveRange ar{s.start, idx, idx};
            while (idx < by_start.size() && by_start[idx].start.i == s.start) {
                idx++;
                ar.before_end = idx;
            }
            active_range.push_back(ar);
        }
    }
    // matches_for_start (prune.rs:203-205): nullptr range when the position has no entry.
    bool matches_for_start(Pos p, size_t& lo, size_t& hi) const {
        auto it = start_index.find({p.i, p.j});
        if (it == start_index.end()) return false;
        lo = it->second.first;
        hi = it->second.second;
        return true;
    }
    // prune.rs:245-292. i_range = is..ie, j_range = js..je, both used as inclusive bounds.
    template <class F>
    void prune_block(I is, I ie, I js, I je, F&& f) {
        ORACLE_ASSERT(js <= je, "prune_block j_range");
        // binary_search_by_key(&(is+1), |ar| ar.col).unwrap_or_else(|idx| idx): first col >= is+1.
        size_t seed_idx = std::lower_bound(active_range.begin(), active_range.end(), is + 1,
                                           [](const ActiveRange& ar, I key) { return ar.col < key; }) -
                          active_range.begin();
        while (seed_idx < active_range.size() && active_range[seed_idx].col <= ie) {
            ActiveRange& ar = active_range[seed_idx];
            if (!ar.has_after) {
                ar.after_start = ar.after_end = ar.before_end;
                while (ar.after_start >= ar.before_start + 1 && by_start[ar.after_start - 1].start.j > je) {
                    ar.before_end -= 1;
                    ar.after_start -= 1;
                }
                ar.has_after = true;
            }
            while (ar.before_end > ar.before_start && by_start[ar.before_end - 1].start.j >= js) {
                Match& mm = by_start[ar.before_end - 1];
                mm.pruned = Pruned;
                f(mm);
                ar.before_end -= 1;
            }
            while (ar.after_start < ar.after_end && by_start[ar.after_start].start.j <= je) {
                Match& mm = by_start[ar.after_start];
                mm.pruned = Pruned;
                f(mm);
                ar.after_start += 1;
            }
            seed_idx++;
        }
    }
};

// ------------------------------------------------------------------------------------------------ Contours
struct Arrow {  // contour.rs:47-52
    Pos start, end;
    MatchCost score;
};

struct RotateToFrontContour {  // rotate_to_front.rs:10-97
    std::vector<Pos> points;
    void push(Pos p) { points.push_back(p); }
    bool contains(Pos q) {  // move-to-front on hit (rotate_right(1) of [0..=idx])
        for (size_t idx = 0; idx < points.size(); idx++) {
            if (pos_le(q, points[idx])) {
                if (idx > 0) std::rotate(points.begin(), points.begin() + idx, points.begin() + idx + 1);
                return true;
            }
        }
        return false;
    }
    size_t len() const { return points.size(); }
};

struct HintContours {  // hint_contours.rs:12-19 (C = RotateToFrontContour; SplitVec == vector semantically)
    std::vector<RotateToFrontContour> contours;
    Layer max_len = 1;
    Layer layers_removed = 0;

    // is_score_at_least, hint_contours.rs:125-133. Returns -1 for None.
    int64_t is_score_at_least(Pos q, Layer v) {
        Layer hi = (Layer)std::min<uint64_t>((uint64_t)v + max_len, contours.size());
        for (Layer w = v; w < hi; w++)
            if (contours[w].contains(q)) return w;
        return -1;
    }
    // score, hint_contours.rs:258-272
    Cost score(Pos q) {
        Layer low = 0, high = (Layer)contours.size();
        while (high - low > 1) {
            Layer mid = (low + high) / 2;
            int64_t v = is_score_at_least(q, mid);
            if (v >= 0)
                low = (Layer)v;
            else
                high = mid;
        }
        return (Cost)low;
    }
    // new_with_filter with filter == true, hint_contours.rs:213-255. Arrows arrive grouped by start.
    void build(const std::vector<Arrow>& arrows, Cost max_len_) {
        contours.clear();
        contours.resize(1);
        max_len = (Layer)max_len_;
        layers_removed = 0;
        contours[0].push(Pos{I_MAX, I_MAX});
        size_t idx = 0;
        while (idx < arrows.size()) {
            Pos start = arrows[idx].start;
            Layer v = 0;
            while (idx < arrows.size() && arrows[idx].start == start) {
                const Arrow& ar = arrows[idx];
                Cost nv = score(ar.end) + ar.score;
                v = std::max(v, (Layer)nv);
                idx++;
            }
            if (v == 0) continue;
            if (contours.size() <= v) contours.resize(v + 1);
            contours[v].push(start);
        }
    }
    // score_with_hint, hint_contours.rs:283-344
    std::pair<Cost, Hint> score_with_hint(Pos q, Hint hint) {
        Layer sub = hint.original_layer >= layers_removed ? hint.original_layer - layers_removed : 0;  // saturating_sub
        Layer v = std::min<Layer>(sub, (Layer)contours.size() - 1);
        const Layer SEARCH_RANGE = 5;
        int64_t found = is_score_at_least(q, v);
        if (found >= 0) {
            Layer vv = (Layer)found;
            Layer best = vv;
            Layer upper = (Layer)std::min<uint64_t>((uint64_t)vv + SEARCH_RANGE + 2, contours.size());
            for (Layer w = vv + 1; w <= upper; w++) {
                if (w < contours.size() && contours[w].contains(q)) best = w;
                if (w == contours.size() || w >= best + max_len) return {(Cost)best, Hint{best + layers_removed}};
            }
        } else {
            Layer lo = v >= SEARCH_RANGE ? v - SEARCH_RANGE : 0;
            Layer hi = v >= 1 ? v - 1 : 0;
            for (int64_t w = hi; w >= (int64_t)lo; w--) {
                if (contours[(Layer)w].contains(q)) return {(Cost)w, Hint{(Layer)w + layers_removed}};
            }
        }
        Cost w = score(q);
        return {w, Hint{(Layer)w + layers_removed}};
    }
};

// ------------------------------------------------------------------------------------------------ GCSH
struct GcshI : HeuristicInstance {  // CSHI with use_gap_cost = true, csh.rs:152-170
    Pos target, t_target;
    Seeds seeds;
    MatchPruner matches;
    HintContours contours;
    Layer lowest_modified_contour = LAYER_MAX;
    Layer highest_modified_contour = LAYER_MIN;
    size_t num_matches_ = 0;
    Cost h0 = 0;

    Arrow match_to_arrow(const Match& mt) const {
        return Arrow{seeds.transform(mt.start), seeds.transform(mt.end), mt.score()};
    }

    // CSHI::new, csh.rs:199-308
    GcshI(const uint8_t* a, size_t n, const uint8_t* b, size_t m, MatchConfig config) {
        Matches ms = find_matches_hash_a(a, n, b, m, config, /*transform_filter=*/true);
        seeds = std::move(ms.seeds);
        target = Pos{(I)n, (I)m};
        t_target = seeds.transform(target);
        std::vector<Match>& mv = ms.matches;
        mv.erase(std::remove_if(mv.begin(), mv.end(),
                                [&](const Match& x) { return !pos_le(seeds.transform(x.start), t_target); }),
                 mv.end());
        num_matches_ = mv.size();
        matches = MatchPruner(std::move(mv), seeds);
        std::vector<Arrow> arrows;
        for (size_t t = matches.by_start.size(); t-- > 0;) {
            const Match& x = matches.by_start[t];
            if (!x.is_active()) continue;
            Arrow ar = match_to_arrow(x);
            if (!pos_le(ar.end, t_target)) continue;
            arrows.push_back(ar);
        }
        contours.build(arrows, config.r);
        h0 = h(Pos{0, 0});
    }
    size_t num_matches() const override { return num_matches_; }

    Cost distance(Pos from, Pos to) const {  // csh.rs:176-186
        return std::max(GapCostI::gap(from, to), seeds.potential_distance(from, to));
    }
    Cost h(Pos pos) override {  // csh.rs:341-350
        n_h_calls++;
        Cost p = seeds.pot(pos);
        Cost val = contours.score(seeds.transform(pos));
        return val == 0 ? distance(pos, target) : p - val;
    }
    std::pair<Cost, Hint> h_with_hint(Pos pos, Hint hint) override {  // csh.rs:367-376
        n_h_calls++;
        Cost p = seeds.pot(pos);
        auto [val, new_hint] = contours.score_with_hint(seeds.transform(pos), hint);
        if (val == 0) return {distance(pos, target), new_hint};
        return {p - val, new_hint};
    }
    void prune_block(I is, I ie, I js, I je) override {  // csh.rs:472-493
        Hint hint{};
        Layer lo = lowest_modified_contour, hi = highest_modified_contour;
        matches.prune_block(is, ie, js, je, [&](const Match& mm) {
            auto [layer, new_hint] = contours.score_with_hint(seeds.transform(mm.start), hint);
            lo = std::min(lo, (Layer)layer);
            hi = std::max(hi, (Layer)layer);
            hint = new_hint;
        });
        lowest_modified_contour = lo;
        highest_modified_contour = hi;
    }

    // chain_score, hint_contours.rs:162-208. Returns 0 for None.
    Layer chain_score(Pos pos, Layer v) {
        Pos p = seeds.transform_back(pos);
        size_t lo, hi;
        if (!matches.matches_for_start(p, lo, hi)) return 0;
        Layer max_score = 0;
        for (size_t t = lo; t < hi; t++) {
            const Match& mt = matches.by_start[t];
            if (!mt.is_active()) continue;
            Arrow arrow = match_to_arrow(mt);
            if (!pos_le(arrow.end, t_target)) continue;
            Layer end_layer = v - 1;
            bool skip = false;
            while (!contours.contours[end_layer].contains(arrow.end)) {
                end_layer -= 1;
                if (end_layer + arrow.score <= max_score) {
                    skip = true;
                    break;
                }
            }
            if (skip) continue;
            Layer start_layer = end_layer + arrow.score;
            max_score = std::max(max_score, start_layer);
        }
        return max_score;
    }

    // update_contours (csh.rs:497-554) -> HintContours::update_layers(lowest_modified, Layer::MAX, arrows,
    // Some((pos.0, transform_back)))  (hint_contours.rs:460-637).
    void update_contours(Pos pos) override {
        update_layers(lowest_modified_contour, LAYER_MAX, pos.i);
        highest_modified_contour = LAYER_MIN;
    }
    enum ShiftKind { ShNone, ShLayers, ShInconsistent };
    struct Shift {
        ShiftKind k = ShNone;
        Layer s = 0;
        bool operator==(const Shift& o) const { return k == o.k && (k != ShLayers || s == o.s); }
        void merge(Shift o) {  // hint_contours.rs:72-88
            if (k == ShNone) {
                *this = o;
            } else if (k == ShLayers) {
                if (o.k == ShNone) {
                } else if (o.k == ShLayers && o.s == s) {
                } else
                    k = ShInconsistent;
            }
        }
    };
    void update_layers(Layer v, Layer last_change, I right_of) {
        auto& cs = contours.contours;
        v = std::max<Layer>(v, 1);
        last_change = std::max(last_change, v);
        Layer fully_shifted_layers = 0;
        Shift rolling_shift;
        v -= 1;
        for (;;) {
            v += 1;
            if (v >= cs.size()) break;
            RotateToFrontContour current = std::move(cs[v]);  // std::mem::take
            cs[v] = RotateToFrontContour{};
            Shift current_shift;
            bool changes = false;
            // prune_filter == retain(!f): iterate in order, evaluate f once per point.
            std::vector<Pos> kept;
            for (Pos pt : current.points) {
                Layer new_layer = chain_score(pt, v);
                bool prune;
                if (new_layer == 0) {
                    prune = true;  // no arrows left
                } else {
                    ORACLE_ASSERT(new_layer <= v, "New layer should never be larger than current layer");
                    if (new_layer == v) {
                        current_shift.k = ShInconsistent;
                        prune = false;
                    } else {
                        current_shift.merge(Shift{ShLayers, v - new_layer});
                        cs[new_layer].push(pt);
                        prune = true;
                    }
                }
                if (prune)
                    changes = true;
                else
                    kept.push_back(pt);
            }
            current.points = std::move(kept);
            cs[v] = std::move(current);

            if (changes) {
                last_change = std::max(last_change, v);
            } else {
                ORACLE_ASSERT(current_shift.k == ShNone || current_shift.k == ShInconsistent, "shift state");
            }
            // v >= last_change.saturating_add(max_len)
            {
                uint64_t lim = (uint64_t)last_change + contours.max_len;
                if (lim > LAYER_MAX) lim = LAYER_MAX;
                if (v >= lim) break;
            }
            if (cs[v].len() == 0 && current_shift.k != ShInconsistent) {
                if (rolling_shift.k == ShNone || current_shift.k == ShNone || rolling_shift == current_shift) {
                    fully_shifted_layers += 1;
                    if (rolling_shift.k == ShNone) rolling_shift = current_shift;
                }
            } else {
                fully_shifted_layers = 0;
                rolling_shift = Shift{};
            }
            if (rolling_shift.k == ShLayers && v >= last_change) {
                ORACLE_ASSERT(fully_shifted_layers > 0, "fully_shifted_layers");
                Layer shift = rolling_shift.s;
                if (fully_shifted_layers >= contours.max_len + shift - 1) {
                    for (Layer t = 0; t < shift; t++) {
                        ORACLE_ASSERT(cs[v].len() == 0, "removed contour must be empty");
                        cs.erase(cs.begin() + v);
                        contours.layers_removed += 1;
                        v -= 1;
                    }
                    break;
                }
            }
            // right_of = Some((pos.0, transform_back))
            {
                bool stop = true;
                Layer hi = (Layer)std::min<uint64_t>((uint64_t)v + 1 + contours.max_len, cs.size());
                for (Layer w = v + 1; w < hi; w++) {
                    for (Pos pt : cs[w].points)
                        if (seeds.transform_back(pt).i >= right_of) stop = false;
                    if (!stop) break;
                }
                if (stop) break;
            }
        }
    }
};

}  // namespace oracle
