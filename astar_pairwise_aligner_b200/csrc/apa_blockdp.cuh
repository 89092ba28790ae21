// K1 / K3: the Myers bit-vector block-DP wavefront (replaces pa_bitpacking::simd::{compute,fill},
// pa-bitpacking/src/simd.rs:98-226,326-437, and myers::compute_block, myers.rs:27-55).
//
// Mapping: one warp computes a (cols <= 256) x (rows = 32 * nhw) rectangle. Lane r owns the 32-row
// half-word r of the current 32-lane chunk and sweeps the columns left to right, one anti-diagonal per
// step: at step t lane r evaluates column t - r. The horizontal delta leaving the bottom row of the lane above
// (two 0/1 flags: +1, -1) arrives by __shfl_up_sync and is folded in as (hp << 1) | hp0 by an IMAD on the FMA pipe.
// Bands taller than 31 half-words are processed chunk by chunk; the bottom delta row of a chunk (2 bits per
// column) is handed to the next chunk through shared memory. The recurrence is bit-for-bit the reference's, evaluated in a different topological order, so the
// resulting V column is identical.
#pragma once
#include "apa_common.cuh"

namespace apa {

struct WarpSmem {
    uint2 amask[BLOCK_W];   // per column of the current block: (0 - rank bit0, 0 - rank bit1) of a[i]  (profile.rs:117-121)
    uint8_t hrow[BLOCK_W];  // bottom horizontal deltas of the previous chunk: bit0 = +1, bit1 = -1
    int32_t dt_i[2][96];    // DT-trace fronts of the current and previous level: column reached on diagonal d at [d + 48]
};

// The constant 2 as an operand the assembler cannot fold: (h << 1) | carry is issued as IMAD h, c[2], carry on the FMA
// pipe instead of a funnel shift on the ALU pipe, which is the saturated pipe of the block DP (profiles/README.md).
__constant__ uint32_t c_two = 2u;

// One 32-row x 1-column Myers step (myers.rs:27-55 on a 32-bit word).
// cp_in/cm_in: the horizontal delta entering this lane's top row, as 0/1 flags (+1 / -1), i.e. bit 31 of the hp/hm
// words of the lane above. cp_out/cm_out: the same for the delta leaving at the bottom.
__device__ __forceinline__ void myers_step(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t& vp, uint32_t& vm,
                                           uint32_t cp_in, uint32_t cm_in, uint32_t& cp_out, uint32_t& cm_out) {
    uint32_t eq = (a0 ^ b0) & (a1 ^ b1);  // BitProfile::eq, profile.rs:141-144 (b planes are stored negated)
    uint32_t vx = eq | vm;
    uint32_t eq2 = eq | cm_in;            // `eq |= h0.m`: the input delta may be -1 (myers.rs:31-32)
    uint32_t hx = (((eq2 & vp) + vp) ^ vp) | eq2;
    uint32_t hp = vm | ~(hx | vp);
    uint32_t hm = vp & hx;
    cp_out = hp >> 31;
    cm_out = hm >> 31;
    uint32_t hps, hms;  // (hp << 1) | h0.p, (hm << 1) | h0.m
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hps) : "r"(hp), "r"(c_two), "r"(cp_in));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hms) : "r"(hm), "r"(c_two), "r"(cm_in));
    vp = hms | ~(vx | hps);
    vm = hps & vx;
}

// Stage the a-masks of columns [col_s, col_s + ncols) into shared memory from the packed planes of a
// (col_s is a multiple of 256, so the block starts on a half-word boundary). Stored planes are negated rank bits:
// mask = 0 - rank_bit = stored_bit - 1.
__device__ __forceinline__ void stage_amask(WarpSmem& sm, const uint2* __restrict__ aprof, I col_s, int ncols, int lane) {
    const int hw0 = col_s >> 5;
    for (int k = 0; 32 * k < ncols; k++) {
        const uint2 pl = aprof[hw0 + k];
        sm.amask[32 * k + lane] = make_uint2(((pl.x >> lane) & 1u) - 1u, ((pl.y >> lane) & 1u) - 1u);
    }
    __syncwarp();
}

// One chunk of the wavefront: `nact` lanes take part, lane l working on column t - l at step t.
// FIRST (the chunk touches the top edge of the band, where +1 deltas enter): lane 0 is a FEEDER, not a row. Its state
// (vp, vm) = (0, ~0) is a fixed point of the step for every eq - hp = ~0, hm = 0, carries (1, 0) out, (vp, vm) unchanged -
// so lane 1 receives the +1 top-edge delta through the ordinary shuffle and no lane needs a per-step select. Rows are
// lanes 1 .. nact-1 (at most 31). Otherwise lane 0 is the first row and takes its incoming deltas from sm.hrow, where
// the last lane of the previous chunk left them. HAND_OFF: the last lane publishes its bottom deltas for the next chunk.
// The sweep is split into ramp-up / steady / ramp-down so the steady state (all lanes busy) runs without guards.
template <bool FILL, bool FIRST, bool HAND_OFF>
__device__ __forceinline__ void dp_chunk(WarpSmem& sm, int ncols, int nact, uint32_t b0, uint32_t b1, uint32_t& vp, uint32_t& vm,
                                         uint2* __restrict__ fillcol /* fillvals + hw of this lane */, int nhw) {
    const int lane = threadIdx.x & 31;
    const bool act_lane = lane < nact;
    const bool is_row = FIRST ? (act_lane && lane > 0) : act_lane;
    const bool is_top = lane == 0;
    const bool is_bot = lane == nact - 1;
    const int r = act_lane ? lane : 0;  // idle lanes shadow lane 0 on valid addresses; their results are never stored
    uint32_t cp_o = FIRST ? 1u : 0u, cm_o = 0u;  // the feeder's constant output (idle lanes carry it too, unused)
    auto step = [&](int t, bool guarded) {
        uint32_t cpi = __shfl_up_sync(FULL, cp_o, 1);
        uint32_t cmi = __shfl_up_sync(FULL, cm_o, 1);
        if (!FIRST) {
            uint32_t x = sm.hrow[min(t, ncols - 1)];
            cpi = is_top ? (x & 1u) : cpi;
            cmi = is_top ? (x >> 1) : cmi;
        }
        const int col = t - r;
        if (!guarded || (unsigned)col < (unsigned)ncols) {
            uint2 am = sm.amask[col];
            myers_step(am.x, am.y, b0, b1, vp, vm, cpi, cmi, cp_o, cm_o);
            if (HAND_OFF) {
                if (is_bot) sm.hrow[col] = (uint8_t)(cp_o | (cm_o << 1));
            }
            if (FILL) {
                if (is_row) fillcol[(size_t)col * nhw] = make_uint2(vp, vm);
            }
        }
    };
    int t = 0;
    const int t_steady = min(nact - 1, ncols);  // first step at which every active lane has a valid column
    for (; t < t_steady; t++) step(t, true);
    if (nact - 1 < ncols) {
#pragma unroll 4
        for (; t < ncols; t++) step(t, false);
    }
    const int T = ncols + nact - 1;
    for (; t < T; t++) step(t, true);
}

// Compute the right-edge column of a block.
//   prev      : stored column to the left (rows outside it start from +1 deltas: init_v_with_overlap, blocks.rs:753-767)
//   njs, nje  : rounded-out row range of the new block (multiples of 64)
//   vout/cum  : nhw (p,m) words and nhw+1 running values of the new column
//   fillvals  : if FILL, every column's V is also stored: fillvals[col * nhw + hw]   (simd::fill)
// Horizontal deltas along the top edge are +1 (HMode::None, blocks.rs:728-734).
// Returns the value at the bottom of the rounded range (bot_val); top_val is supplied by the caller.
template <bool FILL>
__device__ Cost block_dp(WarpSmem& sm, const uint2* __restrict__ bprof, const BlkView& prev, int ncols, I njs, I nje,
                         uint2* __restrict__ vout, int32_t* __restrict__ cumout, Cost top_val_new, uint2* __restrict__ fillvals,
                         unsigned long long& word_steps) {
    const int lane = threadIdx.x & 31;
    const int nhw = (nje - njs) >> 5;
    // chunks of at most 31 rows (the first chunk gives lane 0 to the feeder), evenly sized
    const int nchunks = (nhw + 30) / 31;
    const int per = nchunks ? (nhw + nchunks - 1) / nchunks : 0;
    Cost running = top_val_new;
    for (int c = 0; c < nchunks; c++) {
        const int nrow = min(per, nhw - per * c);
        const int rl = c == 0 ? lane - 1 : lane;  // row of this lane inside the chunk
        const bool is_row = rl >= 0 && rl < nrow;
        const int hw = per * c + (is_row ? rl : 0);
        const I j0 = njs + 32 * hw;
        // feeder / idle lanes: (0, ~0) is the feeder's fixed point; idle lanes never publish anything
        uint32_t vp = 0u, vm = ~0u, b0 = 0u, b1 = 0u;
        if (is_row) {
            vp = ~0u;
            vm = 0u;
            if (!prev.ones && j0 >= prev.js && j0 < prev.je) {
                uint2 pm = prev.v[(j0 - prev.js) >> 5];
                vp = pm.x;
                vm = pm.y;
            }
            uint2 bb = bprof[j0 >> 5];
            b0 = bb.x;
            b1 = bb.y;
        }
        const bool hand_off = (c + 1 < nchunks);
        uint2* fillcol = FILL ? fillvals + hw : nullptr;
        if (c == 0) {
            if (hand_off)
                dp_chunk<FILL, true, true>(sm, ncols, nrow + 1, b0, b1, vp, vm, fillcol, nhw);
            else
                dp_chunk<FILL, true, false>(sm, ncols, nrow + 1, b0, b1, vp, vm, fillcol, nhw);
        } else {
            if (hand_off)
                dp_chunk<FILL, false, true>(sm, ncols, nrow, b0, b1, vp, vm, fillcol, nhw);
            else
                dp_chunk<FILL, false, false>(sm, ncols, nrow, b0, b1, vp, vm, fillcol, nhw);
        }
        __syncwarp();
        if (is_row) vout[hw] = make_uint2(vp, vm);
        // running values: cum[hw] = value at the top of half-word hw.
        int val = is_row ? (__popc(vp) - __popc(vm)) : 0;
        int incl = val;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += y;
        }
        if (is_row) cumout[hw] = running + incl - val;
        running += __shfl_sync(FULL, incl, 31);
        word_steps += (unsigned long long)ncols * (unsigned long long)nrow;
    }
    if (lane == 0) cumout[nhw] = running;
    __syncwarp();
    return running;
}

}  // namespace apa
