// Device-side A*PA2 driver for one pair (one warp): band selection, band doubling, block store.
// Replaces, per pair and entirely on the GPU:
//   AstarPa2Instance::{j_range, fixed_j_range, align_for_bounded_dist}   astarpa2/src/domain.rs:77-541
//   Blocks::{init, compute_next_block, set_last_block_fixed_j_range}     astarpa2/src/blocks.rs:146-569
//   band::exponential_search + AstarPa2::cost_or_align                   astarpa2/src/band.rs:100-141, lib.rs:122-175
// Incremental doubling and block reuse (blocks.rs:190-197,342-469) are built for the second and later passes of a pair
// (dev_pass<INC = true>): two block stores, the pair's h row, rows the previous pass had fixed are kept. The reference
// asserts that this gives the same V column as a fresh computation (blocks.rs:471-543); here the parity tests check cost,
// CIGAR and band log against the oracle and bound the count of computed cells by the oracle's BlockStats count.
#pragma once
#include "apa_batch.cuh"
#include "apa_blockdp.cuh"

namespace APA_NS {

// Everything one warp needs to know about its pair.
struct PairCtx {
    I n, m;
    const uint2* bprof;  // negated bit planes of b per 32 rows (profile.rs:124-131), zero padded to a multiple of 64 rows
    const uint2* aprof;  // the same packing of a (used by the diagonal extensions)
    uint8_t* arena;      // per-pair scratch
    uint32_t arena_size;
    BlkMeta* meta;       // nblk + 1 entries (entry 0 = column 0)
    int nblk;            // number of 256-column blocks
    int nblk_alloc;      // blocks.len() of the reference: entries [0, nblk_alloc) hold ranges of an earlier pass
    int last_idx;        // Blocks.last_block_idx: survives from one pass to the next (see dev_pass, column 0)
    uint32_t v_base;     // start of the V region in the arena (bytes)
    uint32_t v_top;      // bump pointer of block store A (odd passes), growing up from v_base; later the traceback scratch
    uint32_t hi_bot;     // lowest byte used from the arena end: block store B (even passes, incremental doubling only), growing
                         // down; later the CIGAR elements
    uint32_t cig_top;    // where the CIGAR elements grow down from: the arena end, or the bottom of store B when the final pass lives there
    uint32_t hrow_off;   // incremental doubling: the pair's row of horizontal deltas, one byte per column of a (Blocks.h, blocks.rs:106)
    int incremental;     // BlockParams.incremental_doubling (astarpa2_full: true, astarpa2_simple: false; params.rs:88,119)
    // exponential_search between the first-pass kernel and the continuation kernel (band.rs:100-141): last_s, s, maxs
    Cost bd_last_s, bd_s, bd_maxs;
    int more;            // 1: the first pass did not find the cost; the continuation kernel takes the pair over
    int status;          // ST_PENDING while healthy
    // stats
    DpCounters dpc;  // block-DP lane-steps: useful / issued
    unsigned long long computed_cells;
    int passes;
    int fill_blocks, dt_blocks;
    // phase timers (SM clock cycles of this warp): [0] heuristic build [1] block DP [2] passes total [3] traceback total
    // [4] DT-trace [5] CIGAR text [6] h() calls [7] prune_block/update_contours
    long long tphase[8];
    // optional band log (apa_debug_band_log): records of 7 ints {pass, f_max, block, j_s, j_e, fixed_s, fixed_e}
    int32_t* dbg;
    uint32_t dbg_cap, dbg_n;
#if APA_GENERAL
    RunParams par;
#endif
};

// Parameters of the search: compile-time constants of the two presets (params.rs:70-128), or RunParams in the general build.
enum : int { DOM_FULL = 0, DOM_GAPSTART = 1, DOM_GAPGAP = 2, DOM_ASTAR = 3 };
constexpr Cost F_MAX_NONE = -1;  // align_for_bounded_dist(f_max = None): DoublingType::None (lib.rs:126-130)
#if APA_GENERAL
#define P_BW(cx) ((cx).par.block_width)
#define P_ASTAR(cx) ((cx).par.domain == DOM_ASTAR)
#define P_SPARSE_H(cx) ((cx).par.sparse_h != 0)
#define P_PRUNE(cx) ((cx).par.prune != 0)
#define P_DT_TRACE(cx) ((cx).par.dt_trace != 0)
#define P_FR_DROP(cx) ((cx).par.fr_drop)
#define P_MAX_G(cx) ((cx).par.max_g)
#else
#define P_BW(cx) BLOCK_W
#define P_ASTAR(cx) true
#define P_SPARSE_H(cx) true
#define P_PRUNE(cx) true
#define P_DT_TRACE(cx) true
#define P_FR_DROP(cx) DT_FR_DROP
#define P_MAX_G(cx) DT_MAX_G
#endif
constexpr int DT_MAX_G = 40;    // BlockParams.max_g of both presets and of BlockParams::default() (params.rs:91,122, blocks.rs:66)
constexpr int DT_FR_DROP = 10;  // BlockParams.fr_drop of the presets (params.rs:92,123)

__device__ __forceinline__ void dbg_log(PairCtx& cx, Cost f_max, int t, JRange jr, JRange fx) {
    if (!cx.dbg) return;
    if (cx.dbg_n + 7 <= cx.dbg_cap && (threadIdx.x & 31) == 0) {
        int32_t* r = cx.dbg + cx.dbg_n;
        r[0] = cx.passes;
        r[1] = f_max;
        r[2] = t;
        r[3] = jr.s;
        r[4] = jr.e;
        r[5] = fx.s;
        r[6] = fx.e;
    }
    cx.dbg_n += 7;
}

__device__ __forceinline__ uint32_t arena_alloc(PairCtx& cx, uint32_t bytes) {
    bytes = (bytes + 15u) & ~15u;
    if ((uint64_t)cx.v_top + bytes > (uint64_t)cx.hi_bot) {  // 64-bit: arenas go up to 0xF0000000 bytes
        cx.status = ST_OVERFLOW;
        return 0xffffffffu;
    }
    uint32_t off = cx.v_top;
    cx.v_top += bytes;
    return off;
}

// Block store of the current pass: A grows up from v_base (odd passes, and every pass without incremental doubling), B grows
// down from the arena end (even passes with incremental doubling) - the store of the previous pass stays readable meanwhile.
__device__ __forceinline__ bool store_is_b(const PairCtx& cx) { return cx.incremental && (cx.passes & 1) == 0; }
__device__ __forceinline__ uint32_t store_alloc(PairCtx& cx, uint32_t bytes) {
    if (!store_is_b(cx)) return arena_alloc(cx, bytes);
    bytes = (bytes + 15u) & ~15u;
    if ((uint64_t)cx.v_top + bytes > (uint64_t)cx.hi_bot) {
        cx.status = ST_OVERFLOW;
        return 0xffffffffu;
    }
    cx.hi_bot -= bytes;
    return cx.hi_bot;
}

__device__ __forceinline__ BlkView view_of(const PairCtx& cx, const BlkMeta& mt) {
    BlkView v;
    v.js = mt.js;
    v.je = mt.je;
    v.top_val = mt.top_val;
    v.bot_val = mt.bot_val;
    v.ones = mt.ones;
    v.v = (const uint2*)(cx.arena + mt.v_off);
    int nhw = (mt.je - mt.js) >> 5;
    v.cum = (const int32_t*)(cx.arena + mt.v_off + (size_t)nhw * 8);
    return v;
}

// ---------------------------------------------------------------------------------------------- heuristics
// GapCost (pa-heuristic/src/heuristic/distances.rs:130-169): the whole heuristic of astarpa2_simple.
struct GapH {
    I n, m;
    static constexpr bool PRUNE = false;
    __device__ __forceinline__ Cost h(I i, I j, int = 3) const {
        I d = (n - i) - (m - j);
        return d < 0 ? -d : d;
    }
    __device__ __forceinline__ void prune_block(I, I, I, I) {}
    __device__ __forceinline__ void update_contours() {}
};

// NoCost (pa-heuristic/src/heuristic/distances.rs): h = 0, Dijkstra; also the placeholder of the non-A* domains.
struct NoneH {
    static constexpr bool PRUNE = false;
    __device__ __forceinline__ Cost h(I, I, int = 3) const { return 0; }
    __device__ __forceinline__ void prune_block(I, I, I, I) {}
    __device__ __forceinline__ void update_contours() {}
};

// ---------------------------------------------------------------------------------------------- band selection
// AstarPa2Instance::j_range for Domain::Astar with sparse_h (domain.rs:77-246). prev_fixed is the fixed range of
// the previous column, gu the value at its end (0 for the virtual column -1).
template <class Hh>
__device__ JRange dev_j_range(const PairCtx& cx, Hh& hh, I is, I ie, Cost f_max, JRange prev_fixed, const BlkView* prev,
                              bool has_old, JRange old_range) {
#if APA_GENERAL
    if (f_max == F_MAX_NONE) return JRange{0, cx.m};  // domain.rs:84-86
    if (!P_ASTAR(cx)) {  // domain.rs:93-112
        JRange range{0, cx.m};
        if (cx.par.domain == DOM_GAPSTART) {
            range = JRange{is + 1 - f_max, ie + f_max};
        } else if (cx.par.domain == DOM_GAPGAP) {
            const I d = cx.m - cx.n;
            const Cost s = f_max - (d < 0 ? -d : d);
            const I extra = s / 2;  // Rust '/' truncates toward zero, like C
            range = JRange{is + 1 + min(d, 0) - extra, ie + max(d, 0) + extra};
        }
        if (has_old) range = jr_union(range, old_range);
        return jr_inter(range, JRange{0, cx.m});
    }
#endif
    I fixed_start = prev_fixed.s, fixed_end = prev_fixed.e;
    I ui = is, uj = fixed_end;
    Cost gu = is < 0 ? 0 : blk_index(*prev, fixed_end);
    I vi = ui + 1, vj = uj + 1;
    if (P_SPARSE_H(cx)) {
        vj += P_BW(cx);
        vj = min(vj, cx.m);
        for (;;) {
            if (vj < vi - ui + uj) {
                vj = vi - ui + uj;
                break;
            }
            I dd = (vi - ui) - (vj - uj);
            Cost fv = gu + (dd < 0 ? -dd : dd) + hh.h(vi, vj, 0);
            if (fv <= f_max) {
                if (vj == cx.m) break;
                vj += 8;
                if (vj >= cx.m) vj = cx.m;
            } else {
                vi += div_ceil_pos(fv - f_max, 2);
                if (vi > ie) {
                    vi = ie;
                    break;
                }
            }
        }
        vi = ie;
        for (;;) {
            if (vj < vi - ui + uj) {
                vj = vi - ui + uj;
                break;
            }
            I dd = (vi - ui) - (vj - uj);
            Cost fv = gu + (dd < 0 ? -dd : dd) + hh.h(vi, vj, 0);
            if (fv <= f_max) break;
            vj -= div_ceil_pos(fv - f_max, 2);
        }
    } else {  // domain.rs:150-176: walk down the diagonal one column at a time, extending while f <= f_max
        vi = ui;
        vj = uj;
        while (vi < ie) {
            vi += 1;
            vj += 2;
            for (;;) {
                if (vj > cx.m) break;
                I dd = (vi - ui) - (vj - uj);
                Cost fv = gu + (dd < 0 ? -dd : dd) + hh.h(vi, vj, 0);
                if (fv > f_max) break;
                vj += 1;
            }
            vj -= 1;
        }
    }
    JRange range{fixed_start, vj};
    if (has_old) range = jr_union(range, old_range);
    return jr_inter(range, JRange{0, cx.m});
}

// AstarPa2Instance::fixed_j_range with sparse_h (domain.rs:251-350).
template <class Hh>
__device__ JRange dev_fixed_j_range(const PairCtx& cx, Hh& hh, I i, Cost f_max, JRange prev_fixed, const BlkView& blk, I orig_e,
                                    bool has_old_fixed, JRange old_fixed) {
    I start = prev_fixed.s;
    I end = min(orig_e, cx.m);
    while (start <= end) {
        Cost f = blk_index(blk, start) + hh.h(i, start, 1);
        if (f <= f_max) break;
        start += P_SPARSE_H(cx) ? div_ceil_pos(f - f_max, 2) : 1;
    }
    while (end >= start) {
        Cost f = blk_index(blk, end) + hh.h(i, end, 2);
        if (f <= f_max) break;
        end -= P_SPARSE_H(cx) ? div_ceil_pos(f - f_max, 2) : 1;
    }
    JRange fixed{start, end};
    if (has_old_fixed) {
        if (jr_empty(fixed))
            fixed = old_fixed;
        else
            fixed = jr_union(fixed, old_fixed);
    }
    return fixed;
}

constexpr Cost PASS_NONE = -1;

// One pass for a given f_max: AstarPa2Instance::align_for_bounded_dist (domain.rs:356-541), without the trace.
// Returns the distance found in the last column, or PASS_NONE.
// Right-edge column of the next block on one warp: stage the bases of the block's columns, then sweep (see run_block_dp in
// apa_coop.cuh for the variant where the warps of a CTA share the chunks of a tall band).
__device__ __forceinline__ Cost run_block_dp(WarpSmem& sm, PairCtx& cx, const BlkView& prev, I is, int ncols, I njs, I nje,
                                             uint2* vout, int32_t* cumout, Cost top_val, const uint8_t* h_in = nullptr,
                                             uint8_t* h_out = nullptr, int tap_hw = -1, uint8_t* h_tap = nullptr) {
    stage_amask(sm, cx.aprof, is, ncols, threadIdx.x & 31);
    return block_dp<false>(sm, cx.bprof, prev, ncols, njs, nje, vout, cumout, top_val, nullptr, cx.dpc, h_in, h_out, tap_hw, h_tap);
}

// INC: compiled with the code of incremental doubling and block reuse (second and later passes of a pair). The first-pass
// kernel is compiled without it: the single-pass path - the headline - keeps its instruction footprint (the pass kernel is
// sensitive to it: 45.7 -> 47.5 ms per 10 000 pairs with that code inlined, 48.2 ms with it out of line).
template <bool INC, class Hh, class SM>
__device__ Cost dev_pass(PairCtx& cx, SM& sm, Hh& hh, Cost f_max) {
    const int lane = threadIdx.x & 31;
    cx.passes++;
    if (Hh::PRUNE && P_PRUNE(cx)) {
        long long t_u0 = APA_TIC();
        hh.update_contours();
        APA_TOC(cx.tphase[7], t_u0);
    }
    // The block store of this pass starts empty; with incremental doubling the other store still holds the previous pass.
    if (!cx.incremental) {
        cx.v_top = cx.v_base;
    } else if (store_is_b(cx)) {
        cx.hi_bot = cx.arena_size;
    } else {
        cx.v_top = cx.v_base;
    }

    // Column 0 (domain.rs:395-413, blocks.rs:146-179).
    BlkMeta* meta = cx.meta;
    BlkMeta m0 = meta[0];
    bool had0 = cx.nblk_alloc > 0;
    // The old range handed to j_range for column 0 is next_block_j_range() = blocks[last_block_idx + 1] with
    // last_block_idx still where the PREVIOUS pass stopped (domain.rs:386-393 runs before Blocks::init resets it): it only
    // exists when that pass ended before an earlier one did. Blocks::init then unions with blocks[0] (blocks.rs:152-155).
    const bool had_next = cx.last_idx + 1 < cx.nblk_alloc;
    const BlkMeta mnext = had_next ? meta[cx.last_idx + 1] : BlkMeta{};
    JRange jr0 = dev_j_range(cx, hh, -1, 0, f_max, JRange{-1, -1}, nullptr, had_next, JRange{mnext.js, mnext.je});
    if (jr_empty(jr0) || jr0.s > 0) return PASS_NONE;
    {
        JRange init = jr0;
        if (had0) init = jr_union(init, JRange{m0.js, m0.je});
        JRange rounded = jr_round_out(init);
        BlkMeta nm;
        nm.orig_s = jr0.s;
        nm.orig_e = jr0.e;
        nm.js = rounded.s;
        nm.je = rounded.e;
        nm.fs = jr0.s;
        nm.fe = jr0.e;
        nm.has_fixed = 1;
        nm.ones = 1;
        nm.top_val = 0;
        nm.bot_val = rounded.e;
        nm.v_off = 0;
        nm.col_s = -1;
        nm.col_e = 0;
        nm.j_h = J_H_NONE;
        nm.store_pass = cx.passes;
        nm.pad_[0] = nm.pad_[1] = 0;
        __syncwarp();
        if (lane == 0) meta[0] = nm;
        __syncwarp();
        if (cx.nblk_alloc < 1) cx.nblk_alloc = 1;
        cx.last_idx = 0;
        dbg_log(cx, f_max, 0, jr0, jr0);
    }
    bool all_reused = true;
    for (int t = 1; t <= cx.nblk; t++) {
        const I is = (t - 1) * P_BW(cx);
        const I ie = min(is + P_BW(cx), cx.n);
        const BlkMeta pm = meta[t - 1];
        const BlkView prev = view_of(cx, pm);
        const bool existed = t < cx.nblk_alloc;
        const BlkMeta old = existed ? meta[t] : BlkMeta{};
        const JRange prev_fixed{pm.fs, pm.fe};
        JRange jr = dev_j_range(cx, hh, is, ie, f_max, prev_fixed, &prev, existed, JRange{old.js, old.je});
        if (jr_empty(jr)) return PASS_NONE;
        bool reuse = existed && old.js == jr.s && old.je == jr.e && all_reused;
        all_reused = all_reused && reuse;

        // Blocks::compute_next_block (blocks.rs:205-545). Without incremental doubling the block is computed from scratch. With
        // it (blocks.rs:342-469): rows the previous pass had fixed are kept and only the rest is computed, in up to three
        // ranges around the old and the new j_h - the row along which the pair's h row carries the horizontal deltas from one
        // pass to the next. A reused block (domain.rs:449-455, blocks.rs:190-197) is not computed at all: its column, fixed range
        // and j_h carry over (the column is copied into this pass's store).
        JRange rounded = jr_round_out(jr);
        int nhw = (rounded.e - rounded.s) >> 5;
        if (INC && cx.incremental && reuse && old.store_pass == cx.passes - 1) {
            rounded = JRange{old.js, old.je};
            nhw = (rounded.e - rounded.s) >> 5;
        }
        uint32_t off = store_alloc(cx, (uint32_t)nhw * 8u + (uint32_t)(nhw + 1) * 4u);
        if (cx.status != ST_PENDING) return PASS_NONE;
        Cost top_val = blk_index(prev, rounded.s) + (ie - is);
        uint2* vout = (uint2*)(cx.arena + off);
        int32_t* cumout = (int32_t*)(cx.arena + off + (size_t)nhw * 8);
        long long t_dp0 = APA_TIC();
        Cost bot_val = 0;
        I new_j_h = J_H_NONE;
        const bool old_v = existed && old.store_pass == cx.passes - 1;  // the old column is still in the other store
        bool computed = false;
        if constexpr (INC) {
            if (cx.incremental && reuse && old_v) {

                const uint2* ov = (const uint2*)(cx.arena + old.v_off);
                const int32_t* oc = (const int32_t*)(cx.arena + old.v_off + (size_t)nhw * 8);
                for (int k = lane; k < nhw; k += 32) vout[k] = ov[k];
                for (int k = lane; k <= nhw; k += 32) cumout[k] = oc[k];
                __syncwarp();
                top_val = old.top_val;
                bot_val = old.bot_val;
                new_j_h = old.j_h;
                computed = true;
            } else if (cx.incremental && pm.has_fixed && cx.passes > 1) {

                const JRange pfix = jr_round_in(JRange{pm.fs, pm.fe});
                new_j_h = pfix.e;
                uint8_t* hrow = cx.arena + cx.hrow_off + is;
                // the reference asserts these orders (blocks.rs:388-397,427-430); a violation is a panic there
                if (new_j_h < rounded.s || new_j_h > rounded.e) {
                    cx.status = ST_ASSERT;
                    return PASS_NONE;
                }
                const int ncols = ie - is;
                I start = rounded.s;        // first row of the sweep that carries the new h row
                Cost start_val = top_val;
                const uint8_t* h_in = nullptr;
                const bool three = old_v && old.j_h != J_H_NONE && old.has_fixed && next_mult64(old.fs - 1) < old.j_h;
                if (three) {
                    const I ps = next_mult64(old.fs - 1), pe = old.j_h;  // preserved rows [ps, pe): round_in(old_fixed.0 - 1 .. old_j_h)
                    if (pe > new_j_h || ps < rounded.s || rounded.s > old.js || pe > old.je) {  // "j_h may only increase!" and friends
                        cx.status = ST_ASSERT;
                        return PASS_NONE;
                    }
                    // range 0: everything above the preserved part, from the +1 top edge, h row untouched
                    if (ps > rounded.s) {
                        run_block_dp(sm, cx, prev, is, ncols, rounded.s, ps, vout, cumout, top_val);
                        cx.computed_cells += (unsigned long long)ncols * (unsigned long long)(ps - rounded.s);
                    }
                    // preserved part: the old column's words and running values
                    const int old_nhw = (old.je - old.js) >> 5;
                    const uint2* ov = (const uint2*)(cx.arena + old.v_off);
                    const int32_t* oc = (const int32_t*)(cx.arena + old.v_off + (size_t)old_nhw * 8);
                    const int o_new = (ps - rounded.s) >> 5, o_old = (ps - old.js) >> 5, cnt = (pe - ps) >> 5;
                    for (int k = lane; k < cnt; k += 32) {
                        vout[o_new + k] = ov[o_old + k];
                        cumout[o_new + k] = oc[o_old + k];
                    }
                    __syncwarp();
                    start = pe;
                    start_val = oc[(pe - old.js) >> 5];  // value at (ie, old_j_h): exact, the row was fixed
                    h_in = hrow;                         // the deltas along row old_j_h, left there by the previous pass
                }
                // Ranges 1 and 2 (or 01 and 2) of the reference are ONE sweep here, from `start` to the bottom of the band: what
                // range 1 (01) would write to the h row and range 2 read back are the deltas along row new_j_h inside the sweep,
                // recorded by the lane below that row (dp_chunk TAP). Cutting the sweep in two would halve the lanes at work.
                const int o = (start - rounded.s) >> 5;
                cx.computed_cells += (unsigned long long)ncols * (unsigned long long)(rounded.e - start);
                if (new_j_h == start) {  // an empty range 1 leaves the h row as it is; an empty range 01 leaves +1 deltas
                    if (!three) {
                        for (int k = lane; k < ncols; k += 32) hrow[k] = 1;
                        __syncwarp();
                    }
                    bot_val = run_block_dp(sm, cx, prev, is, ncols, start, rounded.e, vout + o, cumout + o, start_val, h_in);
                } else if (new_j_h == rounded.e) {
                    bot_val = run_block_dp(sm, cx, prev, is, ncols, start, rounded.e, vout + o, cumout + o, start_val, h_in, hrow);
                } else {
                    bot_val = run_block_dp(sm, cx, prev, is, ncols, start, rounded.e, vout + o, cumout + o, start_val, h_in, nullptr,
                                           (new_j_h - start) >> 5, hrow);
                }
        
                computed = true;
            }
        }
        if (!computed) {
            // From scratch: no incremental doubling, no fixed range to the left, or the FIRST pass of a pair - which writes no h
            // row (most pairs need one pass only - the headline averages 1.0 - and pay nothing; the second pass then recomputes
            // the rows the first had fixed, which the reference would have kept).
            bot_val = run_block_dp(sm, cx, prev, is, ie - is, rounded.s, rounded.e, vout, cumout, top_val);
            cx.computed_cells += (unsigned long long)(ie - is) * (unsigned long long)(rounded.e - rounded.s);
        }
        APA_TOC(cx.tphase[1], t_dp0);

        BlkMeta nm;
        nm.orig_s = reuse ? old.orig_s : jr.s;
        nm.orig_e = reuse ? old.orig_e : jr.e;
        nm.js = rounded.s;
        nm.je = rounded.e;
        nm.top_val = top_val;
        nm.bot_val = bot_val;
        nm.v_off = off;
        nm.ones = 0;
        nm.col_s = is;
        nm.col_e = ie;
        nm.j_h = new_j_h;
        nm.store_pass = cx.passes;
        nm.pad_[0] = nm.pad_[1] = 0;
        BlkView cur;
        cur.js = rounded.s;
        cur.je = rounded.e;
        cur.top_val = top_val;
        cur.bot_val = bot_val;
        cur.v = vout;
        cur.cum = cumout;
        cur.ones = 0;

        const bool has_old_fixed = existed && old.has_fixed;
        JRange next_fixed{0, -1};
        if (P_ASTAR(cx) && f_max != F_MAX_NONE) {
            next_fixed = dev_fixed_j_range(cx, hh, ie, f_max, prev_fixed, cur, nm.orig_e, has_old_fixed, JRange{old.fs, old.fe});
            // The block is stored (with its previous fixed range) even when the pass aborts right after it.
            JRange stored = next_fixed;
            bool store_fixed = true;
            if (jr_empty(next_fixed)) {
                stored = JRange{old.fs, old.fe};
                store_fixed = has_old_fixed;
            } else if (has_old_fixed) {
                stored = jr_union(JRange{old.fs, old.fe}, next_fixed);  // set_last_block_fixed_j_range, blocks.rs:556-563
            }
            nm.fs = stored.s;
            nm.fe = stored.e;
            nm.has_fixed = store_fixed ? 1 : 0;
        } else {  // fixed_j_range is None outside Domain::Astar (domain.rs:258-261): the stored range is cleared
            nm.fs = 0;
            nm.fe = -1;
            nm.has_fixed = 0;
        }
        __syncwarp();
        if (lane == 0) meta[t] = nm;
        __syncwarp();
        if (cx.nblk_alloc < t + 1) cx.nblk_alloc = t + 1;
        cx.last_idx = t;
        if (P_ASTAR(cx) && f_max != F_MAX_NONE && jr_empty(next_fixed)) return PASS_NONE;
        dbg_log(cx, f_max, t, jr, JRange{nm.fs, nm.fe});

        if (Hh::PRUNE && P_PRUNE(cx)) {
            long long t_p0 = APA_TIC();
            JRange inter = jr_inter(prev_fixed, next_fixed);
            if (!jr_empty(inter)) hh.prune_block(is, ie, inter.s, inter.e);
            APA_TOC(cx.tphase[7], t_p0);
        }
    }
    // dist = last_block.get(|b|) (domain.rs:520-522)
    const BlkMeta lm = meta[cx.nblk];
    if (cx.m < lm.js || cx.m > lm.je) return PASS_NONE;
    const BlkView last = view_of(cx, lm);
    return blk_index(last, cx.m);
}

// band::exponential_search / linear_search driven by cost_or_align (band.rs:100-190, lib.rs:122-175). The presets use
// BandDoubling{start: H0, factor: 2}: offset = h0, s0 = max(1, block_width) = 256. Returns the cost; the blocks of the
// final pass stay in the arena.
// After the last pass only its own block store matters: the traceback scratch grows up from behind store A (or from v_base when
// the final columns live in store B), the CIGAR elements grow down from the arena end (or from the bottom of store B).
__device__ __forceinline__ void arena_after_passes(PairCtx& cx) {
    cx.cig_top = cx.arena_size;
    if (!cx.incremental) return;
    if (store_is_b(cx)) {
        cx.v_top = cx.v_base;
        cx.cig_top = cx.hi_bot;
    } else {
        cx.hi_bot = cx.arena_size;
    }
}

constexpr Cost BD_MORE = -2;  // dev_band_doubling<BD_FIRST>: the first pass did not settle the pair
enum : int { BD_WHOLE = 0, BD_FIRST = 1, BD_CONTINUE = 2 };
// MODE BD_WHOLE: the whole search in one call (fused and general kernels). BD_FIRST: the first pass only - compiled without the
// incremental-doubling code; returns BD_MORE with the state of the search in cx.bd_* when another pass is needed. BD_CONTINUE:
// takes such a pair over from its second pass on (the continuation kernel).
template <int MODE, class Hh, class SM>
__device__ Cost dev_band_doubling(PairCtx& cx, SM& sm, Hh& hh, Cost h0) {
    Cost offset = h0;
    Cost s0 = BLOCK_W;
    float factor = 2.0f;
    bool linear = false;
    Cost delta = 0;
#if APA_GENERAL
    if (cx.par.doubling == 0) {  // DoublingType::None: one pass without a bound (lib.rs:126-130)
        Cost cost = dev_pass<true>(cx, sm, hh, F_MAX_NONE);
        if (cx.status == ST_PENDING && cost == PASS_NONE) cx.status = ST_ASSERT;  // .unwrap()
        arena_after_passes(cx);
        return cost;
    }
    {
        const I gap = cx.n > cx.m ? cx.n - cx.m : cx.m - cx.n;  // DoublingStart::initial_values (band.rs:13-23)
        Cost start_f = cx.par.start == 0 ? 0 : (cx.par.start == 1 ? gap : h0);
        Cost start_inc = cx.par.start == 1 ? gap : 1;
        offset = start_f;
        s0 = max(start_inc, (Cost)cx.par.block_width);
        factor = cx.par.factor;
        linear = cx.par.doubling == 2;
        delta = cx.par.delta;
    }
#endif
    Cost last_s = -1;
    Cost s = linear ? offset : offset + s0;
    Cost maxs = INT32_MAX;
    if (MODE == BD_CONTINUE) last_s = cx.bd_last_s, s = cx.bd_s, maxs = cx.bd_maxs;
    for (;;) {
        Cost cost = MODE == BD_FIRST ? dev_pass<false>(cx, sm, hh, s) : dev_pass<true>(cx, sm, hh, s);
        if (cx.status != ST_PENDING) return -1;
        if (cost != PASS_NONE) {
            if (cost > maxs) {
                cx.status = ST_ASSERT;
                return -1;
            }
            if (cost <= s) {
                if (cost <= last_s) {
                    cx.status = ST_ASSERT;
                    return -1;
                }
                arena_after_passes(cx);
                return cost;
            }
            maxs = min(maxs, cost);
        } else if (maxs != INT32_MAX) {
            cx.status = ST_ASSERT;
            return -1;
        }
        last_s = s;
        if (linear) {
            s = min(s + delta, maxs);
        } else {
            float grown = ceilf(factor * (float)(s - offset));
            s = max((Cost)grown, 1) + offset;
            s = min(s, maxs);
        }
        if (MODE == BD_FIRST) {
            cx.bd_last_s = last_s, cx.bd_s = s, cx.bd_maxs = maxs;
            return BD_MORE;
        }
    }
}

}  // namespace APA_NS
