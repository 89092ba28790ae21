// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).
//
// CPU restatement of the `astarpa2` crate (the A*PA2 driver):
//   IRange/JRange/rounding           astarpa2/src/ranges.rs:9-124
//   Block (index/get/get_diff)       astarpa2/src/block.rs:8-145
//   BlockParams / Blocks             astarpa2/src/blocks.rs:31-197
//   compute_next_block (+incremental doubling), compute_block, init_v_with_overlap{,_preserve_fixed}
//                                    astarpa2/src/blocks.rs:205-545,665-831
//   fill_with_blocks                 astarpa2/src/blocks.rs:572-662
//   trace / parent / dt_trace_block / extend_left{,_simd}   astarpa2/src/blocks/trace.rs:21-500
//   j_range / fixed_j_range / align_for_bounded_dist        astarpa2/src/domain.rs:77-541
//   exponential_search, DoublingStart::H0                   astarpa2/src/band.rs:13-23,100-141
//   presets simple()/full(), cost_or_align                  astarpa2/src/params.rs:70-128, lib.rs:122-175
// Unit-cost helpers of pa-affine-types (cost_model.rs:387-399,453-523) collapse to:
//   gap_cost = extend_cost = |di - dj|, max_{ins,del}_for_cost(s) = s, min_{ins,del}_extend = 1.
#pragma once
#include <cmath>
#include <memory>
#include <optional>

#include "bitpacking.hpp"
#include "heuristic.hpp"

namespace oracle {

// ------------------------------------------------------------------------------------------------ ranges.rs
struct IRange {
    I s, e;  // left-exclusive: columns s+1..=e computed from column s. (-1,0) = first column.
    I len() const { return e - s; }
    bool operator==(const IRange& o) const { return s == o.s && e == o.e; }
};
struct JRange {
    I s, e;  // inclusive
    bool is_empty() const { return s > e; }
    I len() const { return e - s + 1; }
    I exclusive_len() const { return e - s; }
    bool contains(I j) const { return s <= j && j <= e; }
    bool contains_range(JRange o) const { return s <= o.s && o.e <= e; }
    JRange union_(JRange o) const { return {std::min(s, o.s), std::max(e, o.e)}; }
    JRange intersection(JRange o) const { return {std::max(s, o.s), std::min(e, o.e)}; }
    JRange round_out() const { return {trunc_multiple(s, WI), next_multiple_of(e, WI)}; }
    JRange round_in() const { return {next_multiple_of(s, WI), trunc_multiple(e, WI)}; }
    bool operator==(const JRange& o) const { return s == o.s && e == o.e; }
    // v_range of a rounded range
    size_t v_start() const { return (size_t)(s / WI); }
    size_t v_end() const { return (size_t)(e / WI); }
};
inline JRange assert_rounded(JRange r) {
    ORACLE_ASSERT(r.s % WI == 0 && r.e % WI == 0, "assert_rounded");
    return r;
}

// ------------------------------------------------------------------------------------------------ block.rs
struct Block {
    std::vector<V> v;
    IRange i_range{-1, 0};
    JRange original_j_range{-WI, -WI};
    JRange j_range{-WI, -WI};  // rounded out
    std::optional<JRange> fixed_j_range;
    I offset = 0;
    Cost top_val = I_MAX;
    Cost bot_val = I_MAX;
    std::optional<I> j_h;

    static Block first_col(JRange original, JRange rounded) {  // block.rs:51-65
        ORACLE_ASSERT(rounded.s == 0, "first_col");
        Block b;
        b.v.assign((size_t)rounded.exclusive_len() / W, V::one());
        b.i_range = {-1, 0};
        b.original_j_range = original;
        b.j_range = rounded;
        b.fixed_j_range = original;
        b.offset = 0;
        b.top_val = 0;
        b.bot_val = rounded.exclusive_len();
        return b;
    }
    Cost index(I j) const {  // block.rs:69-122
        ORACLE_ASSERT(j_range.s <= j, "Cannot index block above its range");
        ORACLE_ASSERT(j_range.s - offset >= 0, "Offset too large");
        ORACLE_ASSERT(j_range.e - offset <= (I)v.size() * WI, "v not long enough");
        if (j > j_range.e) return bot_val + (j - j_range.e);
        if (j - j_range.s < j_range.e - j) {
            Cost val = top_val;
            I j0 = j_range.s;
            while (j0 + WI <= j) {
                val += v[(size_t)(j0 - offset) / W].value();
                j0 += WI;
            }
            return val + v[(size_t)(j0 - offset) / W].value_of_prefix(j - j0);
        } else {
            Cost val = bot_val;
            I j1 = j_range.e;
            while (j1 - WI > j) {
                val -= v[(size_t)(j1 - WI - offset) / W].value();
                j1 -= WI;
            }
            if (j1 > j) val -= v[(size_t)(j1 - WI - offset) / W].value_of_suffix(j1 - j);
            return val;
        }
    }
    std::optional<Cost> get(I j) const {  // block.rs:126-131
        if (j < j_range.s || j > j_range.e) return std::nullopt;
        return index(j);
    }
    std::optional<Cost> get_diff(I j) const {  // block.rs:134-145
        if (j < offset) return std::nullopt;
        size_t idx = (size_t)(j - offset) / W;
        if (idx >= v.size()) return std::nullopt;
        size_t bit = (size_t)(j - offset) % W;
        return (Cost)((v[idx].p >> bit) & 1) - (Cost)((v[idx].m >> bit) & 1);
    }
};

// ------------------------------------------------------------------------------------------------ blocks.rs
struct BlockParams {  // blocks.rs:31-74
    bool sparse = true;
    bool simd = true;
    bool no_ilp = false;
    bool incremental_doubling = true;
    bool dt_trace = false;
    Cost max_g = 40;
    I fr_drop = 20;
};
struct BlockStats {  // blocks.rs:76-84
    size_t num_blocks = 0, num_incremental_blocks = 0, computed_lanes = 0, unique_lanes = 0;
    uint64_t computed_cells = 0;  // sum of 64 * lanes * cols (SURVEY 8(d) "computed" numerator)
};
struct TraceStats {  // trace.rs:3-14
    size_t dt_trace_tries = 0, dt_trace_success = 0, dt_trace_fallback = 0;
    size_t fill_tries = 0, fill_success = 0, fill_fallback = 0;
};

enum class HMode { None, Input, Update, Output };

struct BlockElem {  // trace.rs:419-441
    I i = I_MAX;
    I ext = 0;
    I parent_d = 0;
};

// trace.rs:443-500. extend_left_simd is behaviourally extend_left (8-byte compares + overshoot fix-up).
inline I extend_left(I& i, I i0, I& j, const uint8_t* a, const uint8_t* b) {
    I cnt = 0;
    while (i > i0 && j > 0 && a[i - 1] == b[j - 1]) {
        i--;
        j--;
        cnt++;
    }
    return cnt;
}

struct Blocks {
    BlockParams params;
    bool trace;
    std::vector<Bits> a, b;
    std::vector<Block> blocks;
    size_t last_block_idx = 0;
    IRange i_range{-1, 0};
    std::vector<H> h;
    BlockStats stats;
    // When set, every incrementally computed block is recomputed from scratch and compared
    // (the reference's cfg!(test) differential check, blocks.rs:471-543).
    bool self_check = false;

    Blocks(BlockParams p, bool tr, const uint8_t* sa, size_t n, const uint8_t* sb, size_t m) : params(p), trace(tr) {
        bitprofile_build(sa, n, sb, m, a, b);  // blocks.rs:111-128
        if (params.incremental_doubling) h.assign(a.size(), H::zero());
    }

    void init(JRange initial_j_range) {  // blocks.rs:146-179
        ORACLE_ASSERT(initial_j_range.s == 0, "init");
        last_block_idx = 0;
        i_range = {-1, 0};
        JRange fixed = initial_j_range;
        if (!blocks.empty()) initial_j_range = initial_j_range.union_(blocks[0].j_range);
        JRange rounded = initial_j_range.round_out();
        Block block;
        if (trace) {
            block = Block::first_col(fixed, rounded);
        } else {
            block.v.assign(b.size(), V::one());
            block.i_range = {-1, 0};
            block.original_j_range = fixed;
            block.j_range = rounded;
            block.fixed_j_range = fixed;
            block.offset = 0;
            block.top_val = 0;
            block.bot_val = rounded.e;
        }
        if (blocks.empty())
            blocks.push_back(std::move(block));
        else
            blocks[0] = std::move(block);
    }
    void pop_last_block() {  // blocks.rs:182-185
        const IRange& o = blocks[last_block_idx].i_range;
        ORACLE_ASSERT(i_range.e == o.e, "Can not pop range");
        i_range.e = o.s;
        last_block_idx -= 1;
    }
    void i_range_push(IRange o) {
        ORACLE_ASSERT(i_range.e == o.s, "IRange::push");
        i_range.e = o.e;
    }
    void reuse_next_block(IRange ir, JRange jr) {  // blocks.rs:190-197
        i_range_push(ir);
        last_block_idx += 1;
        ORACLE_ASSERT(last_block_idx < blocks.size(), "reuse_next_block");
        ORACLE_ASSERT(blocks[last_block_idx].i_range == ir, "reuse i_range");
        ORACLE_ASSERT(blocks[last_block_idx].j_range == jr.round_out(), "reuse j_range");
    }
    const Block& last_block() const { return blocks[last_block_idx]; }
    std::optional<JRange> next_block_j_range() const {  // blocks.rs:551-553
        if (last_block_idx + 1 < blocks.size()) return blocks[last_block_idx + 1].j_range;
        return std::nullopt;
    }
    void set_last_block_fixed_j_range(std::optional<JRange> fixed) {  // blocks.rs:556-569
        auto& cur = blocks[last_block_idx].fixed_j_range;
        if (cur && fixed)
            cur = cur->union_(*fixed);
        else
            cur = fixed;
    }

    // blocks.rs:686-748 (free fn compute_block). v points at the slice for v_range.
    Cost compute_block(IRange ir, size_t vs, size_t ve, V* v, HMode mode) {
        if (ir.len() > 1) {
            stats.computed_lanes += ve - vs;
            stats.num_incremental_blocks += 1;
            stats.computed_cells += (uint64_t)64 * (ve - vs) * (uint64_t)ir.len();
        }
        const Bits* pa = a.data() + ir.s;
        size_t na = (size_t)ir.len();
        const Bits* pb = b.data() + vs;
        size_t nb = ve - vs;
        switch (mode) {
            case HMode::None: {
                static thread_local std::vector<H> tmp;
                tmp.assign(na, H::one());
                return bp_compute(pa, na, pb, nb, tmp.data(), v);
            }
            case HMode::Input: {
                static thread_local std::vector<H> tmp;
                tmp.assign(h.begin() + ir.s, h.begin() + ir.e);
                return bp_compute(pa, na, pb, nb, tmp.data(), v);
            }
            case HMode::Update: return bp_compute(pa, na, pb, nb, h.data() + ir.s, v);
            case HMode::Output: {
                std::fill(h.begin() + ir.s, h.begin() + ir.e, H::one());
                return bp_compute(pa, na, pb, nb, h.data() + ir.s, v);
            }
        }
        return 0;
    }

    // blocks.rs:753-767
    static void init_v_with_overlap(const Block& prev, Block& next) {
        ORACLE_ASSERT(next.offset == next.j_range.s, "init_v next offset");
        ORACLE_ASSERT(prev.offset == prev.j_range.s, "init_v prev offset");
        size_t pvs = prev.j_range.v_start();
        size_t vs = next.j_range.v_start(), ve = next.j_range.v_end();
        next.v.clear();
        next.v.resize(ve - vs, V::one());
        JRange ov = next.j_range.intersection(prev.j_range);
        // Range<usize> from (possibly crossed) bounds; an empty/negative overlap copies nothing.
        int64_t os = ov.s / WI, oe = ov.e / WI;
        if (os < oe) {
            for (int64_t w = os; w < oe; w++) next.v[(size_t)w - vs] = prev.v[(size_t)w - pvs];
        } else if (os > oe) {
            // Rust: slicing with start > end panics.
            throw RefPanic("init_v_with_overlap: slice index starts after end");
        }
    }
    // blocks.rs:774-831
    static void init_v_with_overlap_preserve_fixed(const Block& prev, const Block& old, Block& next) {
        auto& v = next.v;
        ORACLE_ASSERT(prev.offset == prev.j_range.s, "ipf prev offset");
        ORACLE_ASSERT(old.offset == old.j_range.s, "ipf old offset");
        ORACLE_ASSERT(next.offset == next.j_range.s, "ipf next offset");
        ORACLE_ASSERT(next.j_range.contains_range(old.j_range), "ipf contains");
        size_t pvs = prev.j_range.v_start(), pve = prev.j_range.v_end();
        size_t ovs = old.j_range.v_start();
        size_t vs = next.j_range.v_start(), ve = next.j_range.v_end();
        ORACLE_ASSERT(pvs <= vs, "ipf pvs<=vs");
        ORACLE_ASSERT(vs <= ovs, "ipf vs<=ovs");
        JRange pres = JRange{old.fixed_j_range->s - 1, *old.j_h}.round_in();
        size_t ps = pres.v_start(), pe = pres.v_end();
        ORACLE_ASSERT(ps < pe, "ipf preserve non-empty");
        v.resize(ve - vs, V::one());
        if (vs != ovs) {
            // copy_within(ps-ovs .. pe-ovs, ps - vs): memmove semantics
            std::vector<V> tmp(v.begin() + (ps - ovs), v.begin() + (pe - ovs));
            std::copy(tmp.begin(), tmp.end(), v.begin() + (ps - vs));
        }
        // prefix
        for (size_t w = vs; w < ps; w++) v[w - vs] = prev.v[w - pvs];
        // suffix
        size_t copy_end = std::min(ve, pve);
        ORACLE_ASSERT(pe <= copy_end, "ipf suffix slice");
        for (size_t w = pe; w < copy_end; w++) v[w - vs] = prev.v[w - pvs];
        for (size_t w = copy_end; w < ve; w++) v[w - vs] = V::one();
    }

    // blocks.rs:205-545
    void compute_next_block(IRange ir, JRange jr) {
        stats.num_blocks += 1;
        JRange original_j_range = jr;
        JRange j_range = jr.round_out();
        size_t vs = j_range.v_start(), ve = j_range.v_end();
        stats.unique_lanes += ve - vs;
        if (last_block_idx + 1 < blocks.size()) {
            const Block& nb = blocks[last_block_idx + 1];
            ORACLE_ASSERT(j_range.contains_range(nb.j_range), "j_range must grow");
            stats.unique_lanes -= (size_t)nb.j_range.exclusive_len() / W;
        }
        if (trace && !params.sparse) {
            fill_with_blocks(ir, original_j_range);
            return;
        }
        i_range_push(ir);
        Cost prev_top_val = last_block().index(j_range.s);
        Cost prev_bot_val = last_block().index(j_range.e);

        if (!trace && !params.incremental_doubling) {
            // Update the existing `v` vector of the single block in place (blocks.rs:258-285).
            Block& blk = blocks[last_block_idx];
            Cost top_val = prev_top_val + ir.len();
            Cost bot_val = prev_bot_val + compute_block(ir, vs, ve, blk.v.data() + vs, HMode::None);
            blk.i_range = ir;
            blk.original_j_range = original_j_range;
            blk.j_range = j_range;
            blk.top_val = top_val;
            blk.bot_val = bot_val;
            return;
        }
        ORACLE_ASSERT(params.sparse, "sparse");
        if (last_block_idx + 1 == blocks.size()) {
            blocks.push_back(Block{});
        } else {
            ORACLE_ASSERT(blocks[last_block_idx + 1].i_range == ir, "next block i_range");
        }
        Block& prev_block = blocks[last_block_idx];
        Block& next_block = blocks[last_block_idx + 1];
        last_block_idx += 1;

        Block old_block;  // settings only, not the vector
        old_block.i_range = next_block.i_range;
        old_block.original_j_range = next_block.original_j_range;
        old_block.j_range = next_block.j_range;
        old_block.fixed_j_range = next_block.fixed_j_range;
        old_block.offset = next_block.offset;
        old_block.top_val = next_block.top_val;
        old_block.bot_val = next_block.bot_val;
        old_block.j_h = next_block.j_h;

        next_block.i_range = ir;
        next_block.original_j_range = original_j_range;
        next_block.j_range = j_range;
        // fixed_j_range kept
        next_block.offset = j_range.s;
        next_block.top_val = prev_top_val + ir.len();
        next_block.bot_val = prev_bot_val;
        next_block.j_h = std::nullopt;

        if (!params.incremental_doubling || !prev_block.fixed_j_range) {
            init_v_with_overlap(prev_block, next_block);
            next_block.bot_val += compute_block(ir, vs, ve, next_block.v.data(), HMode::None);
            return;
        }

        JRange prev_fixed = prev_block.fixed_j_range->round_in();
        std::optional<JRange> old_fixed = old_block.fixed_j_range;
        next_block.j_h = prev_fixed.e;
        I new_j_h = prev_fixed.e;
        size_t offset = vs;

        if (old_block.j_h && old_fixed && next_multiple_of(old_fixed->s - 1, WI) < *old_block.j_h) {
            I old_j_h = *old_block.j_h;
            init_v_with_overlap_preserve_fixed(prev_block, old_block, next_block);
            JRange r0 = JRange{j_range.s, old_fixed->s - 1}.round_out();
            size_t v0s = r0.v_start(), v0e = r0.v_end();
            ORACLE_ASSERT(v0s <= v0e, "v_range_0");
            JRange r1 = assert_rounded(JRange{old_j_h, new_j_h});
            size_t v1s = r1.v_start(), v1e = r1.v_end();
            ORACLE_ASSERT(r1.s <= r1.e, "j_h may only increase!");
            JRange r2 = assert_rounded(JRange{new_j_h, j_range.e});
            size_t v2s = r2.v_start(), v2e = r2.v_end();
            ORACLE_ASSERT(r2.s <= r2.e, "v_range_2");
            compute_block(ir, v0s, v0e, next_block.v.data() + (v0s - offset), HMode::None);
            if (v1s < v1e) compute_block(ir, v1s, v1e, next_block.v.data() + (v1s - offset), HMode::Update);
            next_block.bot_val += compute_block(ir, v2s, v2e, next_block.v.data() + (v2s - offset), HMode::Input);
        } else {
            init_v_with_overlap(prev_block, next_block);
            JRange r01 = assert_rounded(JRange{j_range.s, new_j_h});
            ORACLE_ASSERT(r01.s <= r01.e, "v_range_01");
            size_t v01s = r01.v_start(), v01e = r01.v_end();
            JRange r2 = assert_rounded(JRange{new_j_h, j_range.e});
            ORACLE_ASSERT(r2.s <= r2.e, "v_range_2");
            size_t v2s = r2.v_start(), v2e = r2.v_end();
            compute_block(ir, v01s, v01e, next_block.v.data() + (v01s - offset), HMode::Output);
            next_block.bot_val += compute_block(ir, v2s, v2e, next_block.v.data() + (v2s - offset), HMode::Input);
        }

        if (self_check) {  // blocks.rs:471-543 (differential check, final part)
            Block nb2;
            nb2.i_range = next_block.i_range;
            nb2.j_range = next_block.j_range;
            nb2.offset = next_block.offset;
            init_v_with_overlap(prev_block, nb2);
            BlockStats saved = stats;
            std::vector<H> saved_h = h;
            Cost bot_diff = compute_block(ir, vs, ve, nb2.v.data(), HMode::None);
            stats = saved;
            h = saved_h;
            if (!(nb2.v == next_block.v) || prev_bot_val + bot_diff != next_block.bot_val)
                throw RefPanic("incremental doubling differs from from-scratch recompute");
        }
    }

    // blocks.rs:572-662
    void fill_with_blocks(IRange ir, JRange original_j_range) {
        JRange j_range = original_j_range.round_out();
        i_range_push(ir);
        size_t vs = j_range.v_start(), ve = j_range.v_end();
        const Block& prev_block = blocks[last_block_idx];
        ORACLE_ASSERT(prev_block.i_range.e == ir.s, "fill consecutive");
        Block next_block;
        next_block.i_range = {ir.s, ir.s};
        next_block.original_j_range = original_j_range;
        next_block.j_range = j_range;
        next_block.offset = j_range.s;
        next_block.fixed_j_range = std::nullopt;
        next_block.top_val = prev_block.index(j_range.s);
        next_block.bot_val = 0;
        next_block.j_h = std::nullopt;
        init_v_with_overlap(prev_block, next_block);
        for (I i = ir.s; i < ir.e; i++) {
            next_block.i_range = {i, i + 1};
            next_block.top_val += 1;
            last_block_idx += 1;
            if (last_block_idx == blocks.size())
                blocks.push_back(next_block);
            else
                blocks[last_block_idx] = next_block;
        }
        size_t len = (size_t)ir.len();
        std::vector<std::vector<V>> values(len);
        std::vector<H> hh(len, H::one());
        bp_fill(a.data() + ir.s, len, b.data() + vs, ve - vs, hh.data(), next_block.v.data(), values);
        Cost bot_val = blocks[last_block_idx - len].index(j_range.e);
        for (size_t t = 0; t < len; t++) {
            Block& blk = blocks[last_block_idx + 1 - len + t];
            blk.v = std::move(values[t]);
            bot_val += hh[t].value();
            blk.bot_val = bot_val;
        }
    }

    // ------------------------------------------------------------------------------------------- trace.rs
    // trace.rs:145-228
    std::pair<Pos, CigarElem> parent(Pos st, Cost& g) const {
        const Block& block = blocks[last_block_idx];
        ORACLE_ASSERT(block.i_range.e == st.i, "Parent of state but block.i differs");
        I cnt = 0;
        while (st.i > 0 && st.j > 0 && profile_is_match(a, b, st.i - 1, st.j - 1)) {
            cnt++;
            st.i--;
            st.j--;
        }
        if (cnt > 0) return {st, CigarElem{OpMatch, cnt}};
        auto vd = block.get_diff(st.j - 1);
        if (vd && *vd == 1) {
            g -= 1;
            return {Pos{st.i, st.j - 1}, CigarElem{OpIns, 1}};
        }
        ORACLE_ASSERT(last_block_idx >= 1, "parent: no previous block");
        const Block& prev_block = blocks[last_block_idx - 1];
        ORACLE_ASSERT(prev_block.i_range.e == st.i - 1, "parent prev block");
        Cost hd = st.j < prev_block.j_range.s ? 1 : g - prev_block.index(st.j);
        if (hd == 1) {
            g -= 1;
            return {Pos{st.i - 1, st.j}, CigarElem{OpDel, 1}};
        }
        Cost dd;
        if (st.j > prev_block.j_range.e) {
            ORACLE_ASSERT(st.j == prev_block.j_range.e + 1, "parent diag edge");
            dd = 1;
        } else {
            auto pd = prev_block.get_diff(st.j - 1);
            if (!pd) throw RefPanic("parent: get_diff unwrap on None");
            dd = *pd + hd;
        }
        if (dd == 1) {
            g -= 1;
            return {Pos{st.i - 1, st.j - 1}, CigarElem{OpSub, 1}};
        }
        throw RefPanic("ERROR: PARENT NOT FOUND IN TRACEBACK");
    }

    // trace.rs:231-416
    std::optional<Pos> dt_trace_block(const uint8_t* sa, const uint8_t* sb, Pos st, Cost& g_st, const Block& prev_block,
                                      Cigar& cigar, std::vector<BlockElem>& bl) const {
        const I block_start = prev_block.i_range.e;
        auto index = [](Cost g, I d) { return (size_t)(g * g + g + d); };
        bl[0] = BlockElem{st.i, 0, 0};

        auto do_trace = [&](Cost g, I d) -> Pos {  // inner fn trace(), trace.rs:266-308
            Pos new_st{block_start, st.j - (st.i - block_start) - d};
            g_st -= g;
            std::vector<CigarElem> ops;
            for (;;) {
                BlockElem fr = bl[index(g, d)];
                if (fr.ext > 0) ops.push_back(CigarElem{OpMatch, fr.ext});
                if (g == 0) break;
                g -= 1;
                d += fr.parent_d;
                CigarOp op;
                switch (fr.parent_d) {
                    case -1: op = OpIns; break;
                    case 0: op = OpSub; break;
                    case 1: op = OpDel; break;
                    default: throw RefPanic("dt trace parent_d");
                }
                ops.push_back(CigarElem{op, 1});
            }
            for (size_t t = ops.size(); t-- > 0;) cigar.push_elem(ops[t]);
            return new_st;
        };
        auto extend_and_check = [&](BlockElem& elem, I j, Cost target_g) -> bool {
            elem.ext += extend_left(elem.i, prev_block.i_range.e, j, sa, sb);
            if (elem.i != prev_block.i_range.e) return false;
            auto val = prev_block.get(j);
            return val && *val == target_g;
        };

        Cost g = 0;
        if (extend_and_check(bl[0], st.j, g_st)) return do_trace(0, 0);
        I d_lo = 0, d_hi = 0;
        for (;;) {
            Cost ng = g + 1;
            size_t end_idx = index(ng, d_hi + 1);
            if (bl.size() <= end_idx) bl.resize(end_idx + 1, BlockElem{});
            for (size_t t = index(ng, d_lo - 1); t <= end_idx; t++) bl[t] = BlockElem{};
            for (I d = d_lo; d <= d_hi; d++) {
                BlockElem fr = bl[index(g, d)];
                auto update = [](BlockElem& x, I y, I dd) {
                    if (y < x.i) {
                        x.i = y;
                        x.parent_d = dd;
                    }
                };
                update(bl[index(ng, d - 1)], fr.i - 1, 1);
                update(bl[index(ng, d)], fr.i - 1, 0);
                update(bl[index(ng, d + 1)], fr.i, -1);
            }
            g += 1;
            d_lo -= 1;
            d_hi += 1;
            I min_fr = I_MAX, min_i = I_MAX;
            for (I d = d_lo; d <= d_hi; d++) {
                BlockElem& fr = bl[index(g, d)];
                if (fr.i == I_MAX) continue;
                I j = st.j - (st.i - fr.i) - d;
                if (extend_and_check(fr, j, g_st - g)) return do_trace(g, d);
                min_fr = std::min(min_fr, 2 * fr.i - d);
                min_i = std::min(min_i, fr.i);
            }
            if (g == params.max_g / 2 && min_i > (block_start + st.i) / 2) return std::nullopt;
            if (g == params.max_g) return std::nullopt;
            if (params.fr_drop > 0) {
                auto w2 = [](I i, I d) { return (I)((uint32_t)2 * (uint32_t)i - (uint32_t)d); };  // release-mode wrapping
                auto thr = [&]() { return (I)((uint32_t)min_fr + (uint32_t)params.fr_drop); };
                while (d_lo < d_hi && (bl[index(g, d_lo)].i <= block_start || w2(bl[index(g, d_lo)].i, d_lo) > thr())) d_lo++;
                while (d_lo < d_hi && (bl[index(g, d_hi)].i <= block_start || w2(bl[index(g, d_hi)].i, d_hi) > thr())) d_hi--;
                if (d_lo > d_hi) return std::nullopt;
            }
        }
    }

    // trace.rs:21-135
    std::pair<Cigar, TraceStats> trace_path(const uint8_t* sa, const uint8_t* sb, Pos from, Pos to) {
        ORACLE_ASSERT(trace, "trace requires trace=true");
        ORACLE_ASSERT(blocks.back().i_range.e == to.i, "trace: last block");
        Cigar cigar;
        Cost g = blocks[last_block_idx].index(to.j);
        TraceStats ts;
        std::vector<BlockElem> dt_cache((size_t)(params.max_g + 1) * (params.max_g + 1));
        while (to != from) {
            while (last_block_idx > 0 && blocks[last_block_idx].i_range.s >= to.i) pop_last_block();
            if (params.dt_trace && to.i > 0) {
                const Block& prev_block = blocks[last_block_idx - 1];
                if (prev_block.i_range.e < to.i - 1) {
                    ts.dt_trace_tries++;
                    auto r = dt_trace_block(sa, sb, to, g, prev_block, cigar, dt_cache);
                    if (r) {
                        ts.dt_trace_success++;
                        to = *r;
                        continue;
                    }
                    ts.dt_trace_fallback++;
                }
            }
            if (params.sparse && to.i > 0) {
                const Block& block = blocks[last_block_idx];
                const Block& prev_block = blocks[last_block_idx - 1];
                ORACLE_ASSERT(prev_block.i_range.e < to.i && to.i <= block.i_range.e, "trace block bracket");
                if (prev_block.i_range.e < to.i - 1 || block.i_range.e > to.i) {
                    JRange prev_j_range = prev_block.j_range;
                    IRange ir{prev_block.i_range.e, to.i};
                    JRange jr{block.j_range.s, to.j};
                    pop_last_block();
                    I height = std::min(jr.exclusive_len(), ir.len() * 5 / 4);
                    for (;;) {
                        JRange jr2 = JRange{std::max(jr.e - height, prev_j_range.s), jr.e}.round_out();
                        ts.fill_tries++;
                        fill_with_blocks(ir, jr2);
                        if (blocks[last_block_idx].index(to.j) == g) {
                            ts.fill_success++;
                            break;
                        }
                        ts.fill_fallback++;
                        ORACLE_ASSERT(jr2.s != 0, "No trace found through block");
                        for (I t = ir.s; t < ir.e; t++) pop_last_block();
                        height *= 2;
                    }
                }
            }
            auto [par, elem] = parent(to, g);
            to = par;
            cigar.push_elem(elem);
        }
        ORACLE_ASSERT(g == 0, "trace must end at g == 0");
        cigar.reverse();
        return {cigar, ts};
    }
};

// ------------------------------------------------------------------------------------------------ params / driver
enum class DomainKind { Full, GapStart, GapGap, Astar };
enum class HeuristicKind { None, Gap, GCSH };

struct AstarPa2Params {  // params.rs:8-42 (subset reachable from the presets + test configurations)
    DomainKind domain = DomainKind::Astar;
    HeuristicKind heuristic = HeuristicKind::GCSH;
    I k = 12;
    MatchCost r = 1;
    size_t p = 14;
    bool doubling = true;  // BandDoubling{start: H0 | Gap, factor}
    bool doubling_start_gap = false;
    bool doubling_start_zero = false;  // DoublingStart::Zero (band.rs:16)
    bool linear = false;               // DoublingType::LinearSearch{start, delta} (band.rs:34-37, 142-190)
    Cost delta = 1;
    float factor = 2.0f;
    I block_width = 256;
    BlockParams front;
    bool sparse_h = true;
    bool prune = true;

    static AstarPa2Params simple() {  // params.rs:70-96
        AstarPa2Params q;
        q.domain = DomainKind::Astar;
        q.heuristic = HeuristicKind::Gap;
        q.block_width = 256;
        q.front = BlockParams{true, true, false, false, true, 40, 10};
        q.sparse_h = true;
        q.prune = false;
        return q;
    }
    static AstarPa2Params full() {  // params.rs:98-128
        AstarPa2Params q;
        q.domain = DomainKind::Astar;
        q.heuristic = HeuristicKind::GCSH;
        q.k = 12;
        q.r = 1;
        q.p = 14;
        q.block_width = 256;
        q.front = BlockParams{true, true, false, true, true, 40, 10};
        q.sparse_h = true;
        q.prune = true;
        return q;
    }
};

struct AstarPa2Stats {  // domain.rs:31-43
    BlockStats block_stats;
    TraceStats trace_stats;
    size_t f_max_tries = 0;
    uint64_t h_calls = 0;
    size_t num_matches = 0;
    Cost h0 = 0;
};

struct PassLog {  // test introspection: per pass, per block ranges (not in the reference)
    Cost f_max;
    std::vector<JRange> j_ranges;      // original (un-rounded) j_range per block incl. column 0
    std::vector<JRange> fixed_ranges;  // stored fixed_j_range after set_last_block_fixed_j_range
};

struct AstarPa2Instance {
    const uint8_t* a;
    size_t n;
    const uint8_t* b;
    size_t m;
    AstarPa2Params params;
    std::unique_ptr<HeuristicInstance> h;  // null unless Domain::Astar
    Hint hint;
    AstarPa2Stats stats;
    std::vector<PassLog>* log = nullptr;
    bool self_check = false;

    AstarPa2Instance(const uint8_t* a_, size_t n_, const uint8_t* b_, size_t m_, const AstarPa2Params& p)
        : a(a_), n(n_), b(b_), m(m_), params(p) {
        if (params.domain == DomainKind::Astar) {  // lib.rs:87-120 build()
            switch (params.heuristic) {
                case HeuristicKind::None: h.reset(new NoCostI()); break;
                case HeuristicKind::Gap: h.reset(new GapCostI(n, m)); break;
                case HeuristicKind::GCSH:
                    h.reset(new GcshI(a, n, b, m, MatchConfig{params.k, params.r, params.p}));
                    break;
            }
        }
    }

    Cost h_hint(Pos pos) {
        auto [val, nh] = h->h_with_hint(pos, hint);
        hint = nh;
        return val;
    }

    // domain.rs:77-246
    JRange j_range(IRange ir, std::optional<Cost> f_max_opt, const Block& prev, std::optional<JRange> old_range) {
        if (!f_max_opt) return JRange{0, (I)m};
        Cost f_max = *f_max_opt;
        I is = ir.s, ie = ir.e;
        JRange range{0, 0};
        switch (params.domain) {
            case DomainKind::Full: range = JRange{0, (I)m}; break;
            case DomainKind::GapStart: range = JRange{is + 1 - f_max, ie + f_max}; break;
            case DomainKind::GapGap: {
                I d = (I)m - (I)n;
                Cost s = f_max - GapCostI::gap(Pos{0, 0}, Pos{(I)n, (I)m});
                I extra = s / 2;  // Rust '/' truncates
                range = JRange{is + 1 + std::min(d, 0) - extra, ie + std::max(d, 0) + extra};
                break;
            }
            case DomainKind::Astar: {
                ORACLE_ASSERT(prev.fixed_j_range.has_value(), "With A* Domain, fixed_j_range should always be set.");
                I fixed_start = prev.fixed_j_range->s, fixed_end = prev.fixed_j_range->e;
                ORACLE_ASSERT(fixed_start <= fixed_end, "Fixed range must not be empty");
                Pos u{is, fixed_end};
                Cost gu = is < 0 ? 0 : prev.index(fixed_end);
                Pos v = u;
                auto f = [&](Pos vv) -> Cost {
                    ORACLE_ASSERT(vv.j - u.j >= vv.i - u.i, "f only valid below the diagonal of u");
                    return gu + GapCostI::gap(u, vv) + h_hint(vv);
                };
                if (!params.sparse_h) {
                    while (v.i < ie) {
                        v.i += 1;
                        v.j += 1;
                        v.j += 1;
                        while (v.j <= (I)m && f(v) <= f_max) v.j += 1;
                        v.j -= 1;
                    }
                } else {
                    v.i += 1;
                    v.j += 1;
                    v.j += params.block_width;
                    v.j = std::min(v.j, (I)m);
                    for (;;) {
                        if (v.j < v.i - u.i + u.j) {
                            v.j = v.i - u.i + u.j;
                            break;
                        }
                        Cost fv = f(v);
                        if (fv <= f_max) {
                            if (v.j == (I)m) break;
                            v.j += 8;
                            if (v.j >= (I)m) v.j = (I)m;
                        } else {
                            v.i += div_ceil(fv - f_max, 2);
                            if (v.i > ie) {
                                v.i = ie;
                                break;
                            }
                        }
                    }
                    v.i = ie;
                    for (;;) {
                        if (v.j < v.i - u.i + u.j) {
                            v.j = v.i - u.i + u.j;
                            break;
                        }
                        Cost fv = f(v);
                        if (fv <= f_max) break;
                        v.j -= div_ceil(fv - f_max, 2);
                    }
                }
                range = JRange{fixed_start, v.j};
                break;
            }
        }
        if (old_range) range = range.union_(*old_range);
        return range.intersection(JRange{0, (I)m});
    }

    // domain.rs:251-350
    std::optional<JRange> fixed_j_range(I i, std::optional<Cost> f_max_opt, std::optional<JRange> prev_fixed,
                                        const Block& block) {
        if (params.domain != DomainKind::Astar) return std::nullopt;
        if (!f_max_opt) return std::nullopt;
        Cost f_max = *f_max_opt;
        auto f = [&](I j) -> Cost { return block.index(j) + h_hint(Pos{i, j}); };
        ORACLE_ASSERT(prev_fixed.has_value(), "prev_fixed_j_range.unwrap()");
        ORACLE_ASSERT(block.j_range.s <= prev_fixed->s, "fixed_j_range: block starts below prev fixed start");
        I start = prev_fixed->s;
        I end = std::min(block.original_j_range.e, (I)m);
        while (start <= end) {
            Cost fv = f(start);
            if (fv <= f_max) break;
            start += params.sparse_h ? div_ceil(fv - f_max, 2) : 1;
        }
        while (end >= start) {
            Cost fv = f(end);
            if (fv <= f_max) break;
            end -= params.sparse_h ? div_ceil(fv - f_max, 2) : 1;
        }
        JRange fixed{start, end};
        if (block.fixed_j_range) {
            if (fixed.is_empty())
                fixed = *block.fixed_j_range;
            else
                fixed = fixed.union_(*block.fixed_j_range);
        }
        return fixed;
    }

    struct PassResult {
        Cost cost;
        std::optional<Cigar> cigar;
    };

    // domain.rs:356-541
    std::optional<PassResult> align_for_bounded_dist(std::optional<Cost> f_max, bool trace, Blocks* blocks_in) {
        stats.f_max_tries += 1;
        if (params.prune && params.domain == DomainKind::Astar) h->update_contours(Pos{0, 0});
        std::unique_ptr<Blocks> local;
        Blocks* blocks = blocks_in;
        if (!blocks) {
            local.reset(new Blocks(params.front, trace, a, n, b, m));
            local->self_check = self_check;
            blocks = local.get();
        }
        ORACLE_ASSERT(f_max.value_or(0) >= 0, "f_max >= 0");
        PassLog* pl = nullptr;
        if (log) {
            log->push_back(PassLog{f_max.value_or(-1), {}, {}});
            pl = &log->back();
        }
        Block first;
        first.fixed_j_range = JRange{-1, -1};
        JRange initial_j_range = j_range(IRange{-1, 0}, f_max, first, blocks->next_block_j_range());
        if (initial_j_range.is_empty() || initial_j_range.s > 0) return std::nullopt;
        blocks->init(initial_j_range);
        blocks->set_last_block_fixed_j_range(initial_j_range);
        if (pl) {
            pl->j_ranges.push_back(initial_j_range);
            pl->fixed_ranges.push_back(*blocks->last_block().fixed_j_range);
        }
        bool all_blocks_reused = true;
        for (I i = 0; i < (I)n; i += params.block_width) {
            IRange ir{i, std::min(i + params.block_width, (I)n)};
            JRange jr = j_range(ir, f_max, blocks->last_block(), blocks->next_block_j_range());
            if (jr.is_empty()) {
                ORACLE_ASSERT(!blocks->next_block_j_range().has_value(), "empty j_range with old range");
                return std::nullopt;
            }
            bool reuse = false;
            if (blocks->next_block_j_range() && *blocks->next_block_j_range() == jr && all_blocks_reused) reuse = true;
            all_blocks_reused &= reuse;
            std::optional<JRange> prev_fixed = blocks->last_block().fixed_j_range;
            if (reuse)
                blocks->reuse_next_block(ir, jr);
            else
                blocks->compute_next_block(ir, jr);
            std::optional<JRange> next_fixed = fixed_j_range(ir.e, f_max, prev_fixed, blocks->last_block());
            if (pl) pl->j_ranges.push_back(jr);
            if (next_fixed && next_fixed->is_empty()) return std::nullopt;
            blocks->set_last_block_fixed_j_range(next_fixed);
            if (pl) pl->fixed_ranges.push_back(blocks->last_block().fixed_j_range.value_or(JRange{0, -1}));
            if (params.prune && params.domain == DomainKind::Astar) {
                JRange inter = prev_fixed->intersection(*next_fixed);
                if (!inter.is_empty()) h->prune_block(ir.s, ir.e, inter.s, inter.e);
            }
        }
        auto dist = blocks->last_block().get((I)m);
        if (!dist) return std::nullopt;
        if (trace && *dist <= f_max.value_or(I_MAX)) {
            auto [cigar, ts] = blocks->trace_path(a, b, Pos{0, 0}, Pos{(I)n, (I)m});
            stats.trace_stats = ts;
            return PassResult{*dist, cigar};
        }
        return PassResult{*dist, std::nullopt};
    }
};

struct AlignResult {
    Cost cost = -1;
    bool has_cigar = false;
    Cigar cigar;
    AstarPa2Stats stats;
};

// lib.rs:122-175 cost_or_align + band.rs:100-141 exponential_search.
inline AlignResult cost_or_align(const uint8_t* a, size_t n, const uint8_t* b, size_t m, const AstarPa2Params& params,
                                 bool trace, std::vector<PassLog>* log = nullptr, bool self_check = false) {
    AstarPa2Instance nw(a, n, b, m, params);
    nw.log = log;
    nw.self_check = self_check;
    Cost h0 = nw.h ? nw.h->h(Pos{0, 0}) : 0;
    AlignResult out;
    if (!params.doubling) {
        ORACLE_ASSERT(params.domain == DomainKind::Full, "DoublingType::None requires Domain::Full");
        auto r = nw.align_for_bounded_dist(std::nullopt, trace, nullptr);
        ORACLE_ASSERT(r.has_value(), "unwrap on None");
        out.cost = r->cost;
        if (r->cigar) {
            out.has_cigar = true;
            out.cigar = *r->cigar;
        }
    } else {
        Cost start_f, start_increment;
        if (params.doubling_start_gap) {  // DoublingStart::Gap
            start_f = start_increment = GapCostI::gap(Pos{0, 0}, Pos{(I)n, (I)m});
        } else if (params.doubling_start_zero) {  // DoublingStart::Zero
            start_f = 0;
            start_increment = 1;
        } else {  // DoublingStart::H0
            start_f = h0;
            start_increment = 1;
        }
        start_increment = std::max(start_increment, (Cost)params.block_width);
        Blocks blocks(params.front, trace, a, n, b, m);
        blocks.self_check = self_check;
        // exponential_search(offset = start_f, s0 = start_increment, factor)
        Cost offset = start_f;
        Cost last_s = -1;
        Cost s = params.linear ? start_f : offset + start_increment;  // linear_search(s0 = start_f, delta), lib.rs:131-139
        Cost maxs = I_MAX;
        for (;;) {
            auto r = nw.align_for_bounded_dist(s, trace, &blocks);
            if (r) {
                ORACLE_ASSERT(r->cost <= maxs, "A solution was found for a previous s, but larger now");
                if (r->cost <= s) {
                    ORACLE_ASSERT(r->cost > last_s, "Cost should already have been found at last_s");
                    out.cost = r->cost;
                    if (r->cigar) {
                        out.has_cigar = true;
                        out.cigar = *r->cigar;
                    }
                    break;
                } else {
                    maxs = std::min(maxs, r->cost);
                }
            } else {
                ORACLE_ASSERT(maxs == I_MAX, "A solution was found for a previous s but not for current s");
            }
            last_s = s;
            if (params.linear) {  // band.rs:186
                s = std::min(s + params.delta, maxs);
                continue;
            }
            float grown = std::ceil(params.factor * (float)(s - offset));
            s = std::max((Cost)grown, 1) + offset;
            s = std::min(s, maxs);
        }
        nw.stats.block_stats = blocks.stats;
    }
    ORACLE_ASSERT(h0 <= out.cost, "Heuristic at start > final cost.");
    nw.stats.h0 = h0;
    if (nw.h) {
        nw.stats.h_calls = nw.h->n_h_calls;
        nw.stats.num_matches = nw.h->num_matches();
    }
    out.stats = nw.stats;
    return out;
}

}  // namespace oracle
