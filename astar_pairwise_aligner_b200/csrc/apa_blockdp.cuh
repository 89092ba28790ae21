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

namespace APA_NS {

// APA_DP_V2 (default): the per-column equality word comes from a per-lane table in shared memory (4 words: this lane's 32
// rows of b against A, C, G, T) indexed by the column's base - one byte load, one IMAD for the address, one word load -
// instead of an a-mask load and 2 LOP3: the ALU pipe is the one the block DP saturates (profiles/README.md), the FMA pipe
// and the shared-memory pipe have room. Measured on B200: pass kernel of astarpa2_simple 150.1 -> 137.2 ms per 2 000 pairs,
// astarpa2_full 47.5 -> 47.1 ms per 10 000 pairs. APA_DP_V2=0 keeps the first formulation (make ab) for A/B measurements.
// APA_SMALL_CODE (default): the ramp-up / ramp-down loops of dp_chunk are not unrolled. The pass kernel stalls on
// instruction fetch (no_instruction is its #3 stall reason, profiles/r1c_*): 16.3k -> 10.3k SASS instructions, 47.8 -> 45.7 ms.
#ifndef APA_SMALL_CODE
#define APA_SMALL_CODE 1
#endif
#ifndef APA_DP_UNROLL
#define APA_DP_UNROLL 4
#endif
#ifndef APA_DP_V2
#define APA_DP_V2 1
#endif
// APA_TMA_STAGE=1: the planes of a for a block (64 bytes) reach shared memory by a TMA bulk copy (cp.async.bulk + mbarrier)
// instead of one coalesced load - the measurement north_star's "staged through shared memory via TMA" asks for. Measured on
// B200 (profiles/README.md, round 2): no gain - the operand is 64 bytes per 256 x 700 cells; kept as a switch, off.
#ifndef APA_TMA_STAGE
#define APA_TMA_STAGE 0
#endif

constexpr int DP_UNROLL = APA_DP_UNROLL;  // steady-state steps per loop iteration of dp_chunk

struct DpCounters {  // statistics of the block DP (bench.py: lane_utilisation = word_steps / issue_steps)
    unsigned long long word_steps;   // useful 32-row x 1-column lane-steps
    unsigned long long issue_steps;  // lane-steps issued: 32 lanes x anti-diagonals swept, ramps and idle lanes included
};

struct alignas(16) WarpSmem {  // (also the 1 040-byte landing zone of dev_pack_planes, before anything else of a pair is live)
#if APA_DP_V2
    uint32_t etab[4 * 32];   // etab[c * 32 + lane]: BitProfile::eq of base c against the 32 rows of b this lane owns
    uint8_t achar[BLOCK_W + 4];  // per column of the current block: rank of a[i] (A0 C1 G2 T3, profile.rs:113); +4: zero
                                 // padding, the steady loop of dp_chunk fetches one column ahead
#else
    uint2 amask[BLOCK_W];   // per column of the current block: (0 - rank bit0, 0 - rank bit1) of a[i]  (profile.rs:117-121)
#endif
    uint8_t hrow[BLOCK_W];  // bottom horizontal deltas of the previous chunk: bit0 = +1, bit1 = -1
    uint8_t htap[BLOCK_W];  // horizontal deltas along an interior row of the sweep (the new h row of incremental doubling)
#if APA_TMA_STAGE
    alignas(16) uint2 astage[8];      // landing zone of the bulk copy of a block's planes of a (8 half-words = 256 columns)
    alignas(8) unsigned long long mbar;  // its completion barrier
    uint32_t mbar_phase, mbar_ready;
#endif
    alignas(8) int32_t dt_i[2][96];  // (8-byte aligned: the build kernel keeps its uint2 plane windows here) DT-trace fronts of the current and previous level: column reached on diagonal d at [d + 48]
};

// APA_TMA_STAGE: shared memory is not initialised at launch; every kernel resets its warps' barrier flag before the first block.
__device__ __forceinline__ void tma_stage_reset(WarpSmem& sm) {
#if APA_TMA_STAGE
    if ((threadIdx.x & 31) == 0) sm.mbar_ready = 0u;
    __syncwarp();
#else
    (void)sm;
#endif
}

// Constants as operands the assembler cannot fold: (h << 1) | carry is issued as IMAD h, c[2], carry and the table
// address as IMAD c, c[128], base on the FMA pipe instead of shifts / LEAs on the ALU pipe.
__constant__ uint32_t c_two = 2u;
__constant__ uint32_t c_128 = 128u;

// One 32-row x 1-column Myers step (myers.rs:27-55 on a 32-bit word) for a given equality word.
// cp_in/cm_in: the horizontal delta entering this lane's top row, as 0/1 flags (+1 / -1), i.e. bit 31 of the hp/hm
// words of the lane above. cp_out/cm_out: the same for the delta leaving at the bottom.
__device__ __forceinline__ void myers_step_eq(uint32_t eq, uint32_t& vp, uint32_t& vm, uint32_t cp_in, uint32_t cm_in,
                                              uint32_t& cp_out, uint32_t& cm_out) {
    uint32_t vx = eq | vm;
    uint32_t eq2 = eq | cm_in;            // `eq |= h0.m`: the input delta may be -1 (myers.rs:31-32)
    uint32_t hx = (((eq2 & vp) + vp) ^ vp) | eq2;
    uint32_t hp = vm | ~(hx | vp);
    uint32_t hm = vp & hx;
    uint32_t hps, hms;  // (hp << 1) | h0.p, (hm << 1) | h0.m
    // Carry out = bit 31: a shift on the ALU pipe. Measured alternatives on B200 (profiles/README.md, r1c): IMAD.HI
    // (h * 2 >> 32) and IMAD.WIDE are slower than the shift they replace (multi-pass on the FMA pipe).
    cp_out = hp >> 31;
    cm_out = hm >> 31;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hps) : "r"(hp), "r"(c_two), "r"(cp_in));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hms) : "r"(hm), "r"(c_two), "r"(cm_in));
    vp = hms | ~(vx | hps);
    vm = hps & vx;
}

#if APA_DP_V2
// Stage the bases of columns [col_s, col_s + ncols) into shared memory from the packed planes of a (col_s is a multiple
// of 256, so the block starts on a half-word boundary; stored planes are negated rank bits). One lane per 4 columns:
// the 4 bits of each plane are spread to the low bits of 4 bytes with a multiply.
__device__ __forceinline__ void stage_amask(WarpSmem& sm, const uint2* __restrict__ aprof, I col_s, int ncols, int lane) {
    const int hw0 = col_s >> 5;
    const int nw = (ncols >> 2) + 1;  // one word past the last column: the steady loop of dp_chunk fetches one column ahead
#if APA_TMA_STAGE && !APA_GENERAL
    {   // TMA bulk copy of the block's plane words into sm.astage, completion on sm.mbar (one phase per block)
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sm.mbar);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.astage[0]);
        const uint32_t bytes = (uint32_t)(((ncols + 63) >> 6) * 16);  // whole 16-byte units: two half-words each (planes are padded)
        if (lane == 0) {
            if (sm.mbar_ready != 0x600DBA44u) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                sm.mbar_ready = 0x600DBA44u;
                sm.mbar_phase = 0u;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(aprof + hw0), "r"(bytes), "r"(bar)
                         : "memory");
        }
        __syncwarp();
        const uint32_t phase = sm.mbar_phase;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        __syncwarp();
        if (lane == 0) sm.mbar_phase = phase ^ 1u;
    }
#endif
    for (int w = lane; w < nw; w += 32) {
        uint32_t word = 0u;
        if (4 * w < ncols) {
#if APA_GENERAL  // block widths below 32: the block may start anywhere inside a half-word
            const uint2 pl = extract32(aprof, col_s + 4 * w);
            const int sh = 0;
            (void)hw0;
#else
#if APA_TMA_STAGE
            const uint2 pl = sm.astage[w >> 3];
#else
            const uint2 pl = aprof[hw0 + (w >> 3)];
#endif
            const int sh = (w & 7) * 4;
#endif
            const uint32_t n0 = (~pl.x >> sh) & 15u, n1 = (~pl.y >> sh) & 15u;
            word = ((n0 * 0x00204081u) & 0x01010101u) | (((n1 * 0x00204081u) & 0x01010101u) << 1);
        }
        ((uint32_t*)sm.achar)[w] = word;
    }
    __syncwarp();
}
// The equality words of this lane's rows (negated planes b0, b1 of 32 rows of b) against the four bases:
// eq = (a0 ^ b0) & (a1 ^ b1) with a0 / a1 = all-ones where the rank bit of a's base is set (profile.rs:141-144).
// Each lane reads back only what it wrote, so no synchronisation is needed.
__device__ __forceinline__ void stage_etab(WarpSmem& sm, uint32_t b0, uint32_t b1, int lane) {
    sm.etab[0 * 32 + lane] = b0 & b1;
    sm.etab[1 * 32 + lane] = ~b0 & b1;
    sm.etab[2 * 32 + lane] = b0 & ~b1;
    sm.etab[3 * 32 + lane] = ~b0 & ~b1;
}
#else
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
#endif

// One chunk of the wavefront: `nact` lanes take part, lane l working on column t - l at step t.
// FIRST (the chunk touches the top edge of the band, where +1 deltas enter): lane 0 is a FEEDER, not a row. Its state
// (vp, vm) = (0, ~0) is a fixed point of the step for every eq - hp = ~0, hm = 0, carries (1, 0) out, (vp, vm) unchanged -
// so lane 1 receives the +1 top-edge delta through the ordinary shuffle and no lane needs a per-step select. Rows are
// lanes 1 .. nact-1 (at most 31). Otherwise lane 0 is the first row and takes its incoming deltas from sm.hrow, where
// the last lane of the previous chunk left them. HAND_OFF: the last lane publishes its bottom deltas for the next chunk.
// The sweep is split into ramp-up / steady / ramp-down so the steady state (all lanes busy) runs without guards.
// CUSTOM_ETAB: the caller has filled sm.etab itself (match masks that are not plain base equality: apa_search).
// TAP: lane `tap_lane` also records the deltas ENTERING its top row, per column, in sm.htap: the horizontal deltas along the row
// boundary above it (HMode::Output / Update of an interior row without cutting the sweep in two, see dev_pass).
template <bool FILL, bool FIRST, bool HAND_OFF, bool CUSTOM_ETAB = false, bool TAP = false>
__device__ __forceinline__ void dp_chunk(WarpSmem& sm, int ncols, int nact, uint32_t b0, uint32_t b1, uint32_t& vp, uint32_t& vm,
                                         uint2* __restrict__ fillcol /* fillvals + hw of this lane */, int nhw, int tap_lane = -1) {
    const int lane = threadIdx.x & 31;
    const bool act_lane = lane < nact;
    const bool is_row = FIRST ? (act_lane && lane > 0) : act_lane;
    const bool is_top = lane == 0;
    const bool is_bot = lane == nact - 1;
    const int r = act_lane ? lane : 0;  // idle lanes shadow lane 0 on valid addresses; their results are never stored
    uint32_t cp_o = FIRST ? 1u : 0u, cm_o = 0u;  // the feeder's constant output (idle lanes carry it too, unused)
#if APA_DP_V2
    if (!CUSTOM_ETAB) stage_etab(sm, b0, b1, lane);
    __syncwarp();
    const uint32_t etab_lane = (uint32_t)__cvta_generic_to_shared(&sm.etab[lane]);
    // eq word of this lane's rows against column `col`: etab[achar[col] * 32 + lane]; the address is one IMAD
    auto load_eq = [&](int col) -> uint32_t {
        uint32_t eaddr;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(eaddr) : "r"((uint32_t)sm.achar[col]), "r"(c_128), "r"(etab_lane));
        return *(const uint32_t*)__cvta_shared_to_generic(eaddr);
    };
#else
    auto load_eq = [&](int col) -> uint32_t {
        uint2 am = sm.amask[col];
        return (am.x ^ b0) & (am.y ^ b1);  // BitProfile::eq, profile.rs:141-144 (b planes are stored negated)
    };
#endif
    auto carries_in = [&](int t, uint32_t& cpi, uint32_t& cmi) {
        cpi = __shfl_up_sync(FULL, cp_o, 1);
        cmi = __shfl_up_sync(FULL, cm_o, 1);
        if (!FIRST) {
            uint32_t x = sm.hrow[min(t, ncols - 1)];
            cpi = is_top ? (x & 1u) : cpi;
            cmi = is_top ? (x >> 1) : cmi;
        }
    };
    auto step_core = [&](int col, uint32_t eq, uint32_t cpi, uint32_t cmi) {
        if (TAP) {
            if (lane == tap_lane) sm.htap[col] = (uint8_t)(cpi | (cmi << 1));
        }
        myers_step_eq(eq, vp, vm, cpi, cmi, cp_o, cm_o);
        if (HAND_OFF) {
            if (is_bot) sm.hrow[col] = (uint8_t)(cp_o | (cm_o << 1));
        }
        if (FILL) {
            if (is_row) fillcol[(size_t)col * nhw] = make_uint2(vp, vm);
        }
    };
    // ramp-up / ramp-down step: lanes whose column is outside the block only take part in the shuffles
    auto step_guarded = [&](int t) {
        uint32_t cpi, cmi;
        carries_in(t, cpi, cmi);
        const int col = t - r;
        if ((unsigned)col < (unsigned)ncols) step_core(col, load_eq(col), cpi, cmi);
    };
    int t = 0;
    const int t_steady = min(nact - 1, ncols);  // first step at which every active lane has a valid column
#if APA_SMALL_CODE
#pragma unroll 1
#endif
    for (; t < t_steady; t++) step_guarded(t);
    if (nact - 1 < ncols) {
        // steady state: every lane has a valid column; the eq word of the next step is fetched one step ahead
        // (achar is padded, so the fetch past the last column of lane 0 stays in bounds)
        uint32_t eq_next = load_eq(t - r);
#pragma unroll DP_UNROLL
        for (; t < ncols; t++) {
            const uint32_t eq = eq_next;
            eq_next = load_eq(t + 1 - r);
            uint32_t cpi, cmi;
            carries_in(t, cpi, cmi);
            step_core(t - r, eq, cpi, cmi);
        }
    }
    const int T = ncols + nact - 1;
#if APA_SMALL_CODE
#pragma unroll 1
#endif
    for (; t < T; t++) step_guarded(t);
}

// Compute the right-edge column of a block.
//   prev      : stored column to the left (rows outside it start from +1 deltas: init_v_with_overlap, blocks.rs:753-767)
//   njs, nje  : rounded-out row range of the new block (multiples of 64)
//   vout/cum  : nhw (p,m) words and nhw+1 running values of the new column
//   fillvals  : if FILL, every column's V is also stored: fillvals[col * nhw + hw]   (simd::fill)
// Horizontal deltas along the top edge: +1 when h_in is null (HMode::None / Output, blocks.rs:728-734), else h_in[col] (one byte per
// column: bit0 = +1, bit1 = -1; HMode::Input / Update). When h_out is not null the deltas along the bottom edge are written there in
// the same form (HMode::Output / Update); h_in and h_out may be the same array.
// Returns the value at the bottom of the rounded range (bot_val); top_val is supplied by the caller.
template <bool FILL>
__device__ Cost block_dp(WarpSmem& sm, const uint2* __restrict__ bprof, const BlkView& prev, int ncols, I njs, I nje,
                         uint2* __restrict__ vout, int32_t* __restrict__ cumout, Cost top_val_new, uint2* __restrict__ fillvals,
                         DpCounters& dpc, const uint8_t* h_in = nullptr, uint8_t* h_out = nullptr, int tap_hw = -1, uint8_t* h_tap = nullptr) {
    // tap_hw in [0, nhw): the deltas along the top edge of half-word tap_hw (row njs + 32 tap_hw) go to h_tap as well
    const int lane = threadIdx.x & 31;
    const int nhw = (nje - njs) >> 5;
    if (h_in) {  // the first chunk takes its incoming deltas from shared memory like every later chunk does
        for (int k = lane; k < ncols; k += 32) sm.hrow[k] = h_in[k];
        __syncwarp();
    }
    // chunks of at most 31 rows (the first chunk gives lane 0 to the feeder), evenly sized
    const int nchunks = (nhw + 30) / 31;
    const int per = nchunks ? (nhw + nchunks - 1) / nchunks : 0;
    Cost running = top_val_new;
    for (int c = 0; c < nchunks; c++) {
        const int nrow = min(per, nhw - per * c);
        const bool feeder = c == 0 && !h_in;       // lane 0 of the first chunk feeds the +1 top edge
        const int rl = feeder ? lane - 1 : lane;  // row of this lane inside the chunk
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
        const bool hand_off = (c + 1 < nchunks) || h_out;
        uint2* fillcol = FILL ? fillvals + hw : nullptr;
        const int tl = tap_hw - per * c;  // the tapped half-word as a row of this chunk
        if (!FILL && tap_hw >= 0 && tl >= 0 && tl < nrow) {
            const int tap_lane = feeder ? tl + 1 : tl;
            if (feeder) {
                if (hand_off)
                    dp_chunk<false, true, true, false, true>(sm, ncols, nrow + 1, b0, b1, vp, vm, nullptr, nhw, tap_lane);
                else
                    dp_chunk<false, true, false, false, true>(sm, ncols, nrow + 1, b0, b1, vp, vm, nullptr, nhw, tap_lane);
            } else {
                if (hand_off)
                    dp_chunk<false, false, true, false, true>(sm, ncols, nrow, b0, b1, vp, vm, nullptr, nhw, tap_lane);
                else
                    dp_chunk<false, false, false, false, true>(sm, ncols, nrow, b0, b1, vp, vm, nullptr, nhw, tap_lane);
            }
        } else if (feeder) {
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
        dpc.word_steps += (unsigned long long)ncols * (unsigned long long)nrow;
    }
    // lane-steps issued: every chunk sweeps ncols + nact - 1 anti-diagonals on 32 lanes (nact = rows, + the feeder in chunk 0)
    if (nchunks) dpc.issue_steps += 32ull * (unsigned long long)(nchunks * (ncols - 1) + nhw + 1);
    if (lane == 0) cumout[nhw] = running;
    __syncwarp();
    if (h_tap && tap_hw >= 0) {
        for (int k = lane; k < ncols; k += 32) h_tap[k] = sm.htap[k];
        __syncwarp();
    }
    if (h_out) {  // what the last chunk left in shared memory is the bottom edge (an empty range passes its top edge through)
        for (int k = lane; k < ncols; k += 32) h_out[k] = nchunks ? sm.hrow[k] : (h_in ? h_in[k] : (uint8_t)1);
        __syncwarp();
    }
    return running;
}

}  // namespace APA_NS
