// Intra-pair parallel block DP: the W warps of a CTA share the chunks of ONE pair's tall band.
//
// A band taller than 31 half-words is swept chunk by chunk (apa_blockdp.cuh); chunk c needs the bottom horizontal deltas of
// chunk c - 1, column by column. On one warp the chunks run back to back. Here chunk c runs on warp c mod W and follows
// chunk c - 1 at a distance of a few steps: the producer publishes how many columns of its bottom row are final
// (`progress`), the consumer waits for the columns its next COOP_GROUP steps read. The delta rows live in a ring of W + 1 shared
// buffers: slot c mod (W + 1) is written by chunk c, read by chunk c + 1, and next written by chunk c + W + 1 - which runs
// on the warp of chunk c + 1, after it. Used when a batch has fewer pairs than the GPU has warp slots (few long pairs,
// memory-limited batches such as BASELINE configs[3], astarpa2_simple): the control code of the pair (band selection,
// heuristic, pruning) stays on warp 0, the leader; the other warps wait at a named barrier for the next block.
#pragma once
#include "apa_align.cuh"

namespace APA_NS {

constexpr int COOP_GROUP = 8;  // steps between two synchronisation points of consecutive chunks: chunk c runs nact - 1 + COOP_GROUP
                               // steps behind chunk c - 1, so 8 warps stay busy on 256-column blocks (7 * 39 + 39 < 256 + 31)
constexpr int COOP_MAX_CHUNKS = 512;  // bands up to 512 * 31 half-words (507 904 rows); taller blocks run on the leader alone

template <int W>
struct CoopSmem {
    WarpSmem lead;  // the leader's own scratch; lead.achar (bases of the block's columns) is read by every warp
    // job descriptor, written by the leader before the START barrier
    int quit;
    int h_first;  // 1: the top edge of the range is +1 (chunk 0 has the feeder lane); 0: chunk 0 takes its deltas from ring slot W
    int h_keep;   // 1: the last chunk publishes its bottom deltas too (they become the pair's h row)
    int tap_c, tap_lane;  // chunk and lane whose incoming deltas are recorded in htap (-1: none), see dp_chunk TAP
    uint8_t htap[BLOCK_W];
    int ncols, nhw, nchunks, per;
    I njs;
    Cost top_val;
    const uint2* bprof;
    const uint2* prev_v;
    I prev_js, prev_je;
    int prev_ones;
    uint2* vout;
    int32_t* cumout;
    int tot[COOP_MAX_CHUNKS];   // sum of the vertical deltas of each chunk
    volatile int progress[W + 1];  // per ring slot: chunk * 1024 + number of final columns of that chunk's bottom row
    alignas(8) uint8_t hrow[W + 1][BLOCK_W];  // bottom horizontal deltas: bit0 = +1, bit1 = -1
    uint32_t etab[W][4 * 32];      // per warp: equality words of the current chunk (see WarpSmem::etab)
};

template <int W>
__device__ __forceinline__ void coop_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(W * 32) : "memory");
}
enum : int { COOP_BAR_START = 1, COOP_BAR_MID = 2, COOP_BAR_END = 3 };

// One chunk on one warp, in groups of COOP_GROUP steps (guarded in the ramps, where some lanes only shuffle).
template <int W>
__device__ __forceinline__ void coop_chunk(CoopSmem<W>& cs, int wid, int c, int ncols, int nact, bool first, bool hand_off, uint32_t b0,
                                           uint32_t b1, uint32_t& vp, uint32_t& vm) {
    const int lane = threadIdx.x & 31;
    const bool act_lane = lane < nact;
    const bool is_top = lane == 0;
    const bool is_bot = lane == nact - 1;
    const int r = act_lane ? lane : 0;
    uint32_t cp_o = first ? 1u : 0u, cm_o = 0u;
    uint32_t* etab = cs.etab[wid];
    etab[0 * 32 + lane] = b0 & b1;
    etab[1 * 32 + lane] = ~b0 & b1;
    etab[2 * 32 + lane] = b0 & ~b1;
    etab[3 * 32 + lane] = ~b0 & ~b1;
    __syncwarp();
    const uint8_t* achar = cs.lead.achar;
    const uint8_t* hin = cs.hrow[(c + W) % (W + 1)];  // slot of chunk c - 1
    uint8_t* hout = cs.hrow[c % (W + 1)];
    volatile int* pin = &cs.progress[(c + W) % (W + 1)];
    volatile int* pout = &cs.progress[c % (W + 1)];
    const int T = ncols + nact - 1;
    const bool is_tap = c == cs.tap_c && lane == cs.tap_lane;
    const uint32_t etab_lane = (uint32_t)__cvta_generic_to_shared(&etab[lane]);
    auto load_eq = [&](int col) -> uint32_t {  // etab[achar[col] * 32 + lane], the address is one IMAD (see dp_chunk)
        uint32_t eaddr;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(eaddr) : "r"((uint32_t)achar[col]), "r"(c_128), "r"(etab_lane));
        return *(const uint32_t*)__cvta_shared_to_generic(eaddr);
    };
    for (int t0 = 0; t0 < T; t0 += COOP_GROUP) {
        const int t1 = min(t0 + COOP_GROUP, T);
        if (!first) {  // the next steps of lane 0 read columns t0 .. t1 - 1 of the row above
            const int need = (c - 1) * 1024 + min(t1, ncols);
            while (*pin < need) {
            }
            __threadfence_block();
        }
        if (t0 >= nact - 1 && t1 <= ncols) {  // every lane has a valid column during the whole group
            // All shared-memory operands of the group are fetched up front (the cooperative kernel runs at low occupancy, so
            // a load inside the dependent chain of a step costs its full latency): 8 equality words and the 8 incoming deltas.
            uint32_t eqs[COOP_GROUP];
#pragma unroll
            for (int k = 0; k < COOP_GROUP; k++) eqs[k] = load_eq(t0 + k - r);
            uint2 hv = make_uint2(0u, 0u);
            if (!first) hv = *(const uint2*)(hin + t0);  // t0 is a multiple of COOP_GROUP = 8
#pragma unroll
            for (int k = 0; k < COOP_GROUP; k++) {
                uint32_t cpi = __shfl_up_sync(FULL, cp_o, 1);
                uint32_t cmi = __shfl_up_sync(FULL, cm_o, 1);
                if (!first) {
                    const uint32_t x = ((k < 4 ? hv.x : hv.y) >> (8 * (k & 3))) & 0xffu;
                    cpi = is_top ? (x & 1u) : cpi;
                    cmi = is_top ? (x >> 1) : cmi;
                }
                if (is_tap) cs.htap[t0 + k - r] = (uint8_t)(cpi | (cmi << 1));
                myers_step_eq(eqs[k], vp, vm, cpi, cmi, cp_o, cm_o);
                if (hand_off && is_bot) hout[t0 + k - r] = (uint8_t)(cp_o | (cm_o << 1));
            }
        } else {
            for (int t = t0; t < t1; t++) {
                uint32_t cpi = __shfl_up_sync(FULL, cp_o, 1);
                uint32_t cmi = __shfl_up_sync(FULL, cm_o, 1);
                if (!first) {
                    const uint32_t x = hin[min(t, ncols - 1)];
                    cpi = is_top ? (x & 1u) : cpi;
                    cmi = is_top ? (x >> 1) : cmi;
                }
                const int col = t - r;
                if ((unsigned)col < (unsigned)ncols) {
                    if (is_tap) cs.htap[col] = (uint8_t)(cpi | (cmi << 1));
                    myers_step_eq(load_eq(col), vp, vm, cpi, cmi, cp_o, cm_o);
                    if (hand_off && is_bot) hout[col] = (uint8_t)(cp_o | (cm_o << 1));
                }
            }
        }
        if (hand_off) {  // columns up to t1 - 1 - (nact - 1) have left the bottom lane
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                *pout = c * 1024 + max(0, min(ncols, t1 - (nact - 1)));
            }
        }
    }
}

// The part of a block every warp of the CTA executes (the leader included), between the START and END barriers.
template <int W>
__device__ __forceinline__ Cost coop_work(CoopSmem<W>& cs, int wid) {
    const int lane = threadIdx.x & 31;
    const int ncols = cs.ncols, nhw = cs.nhw, nchunks = cs.nchunks, per = cs.per;
    const I njs = cs.njs;
    uint2* vout = cs.vout;
    int32_t* cumout = cs.cumout;
    for (int c = wid; c < nchunks; c += W) {
        const int nrow = min(per, nhw - per * c);
        const bool feeder = c == 0 && cs.h_first;
        const int rl = feeder ? lane - 1 : lane;
        const bool is_row = rl >= 0 && rl < nrow;
        const int hw = per * c + (is_row ? rl : 0);
        const I j0 = njs + 32 * hw;
        uint32_t vp = 0u, vm = ~0u, b0 = 0u, b1 = 0u;  // feeder / idle lanes: see block_dp
        if (is_row) {
            vp = ~0u;
            vm = 0u;
            if (!cs.prev_ones && j0 >= cs.prev_js && j0 < cs.prev_je) {
                const uint2 pm = cs.prev_v[(j0 - cs.prev_js) >> 5];
                vp = pm.x;
                vm = pm.y;
            }
            const uint2 bb = cs.bprof[j0 >> 5];
            b0 = bb.x;
            b1 = bb.y;
        }
        coop_chunk<W>(cs, wid, c, ncols, feeder ? nrow + 1 : nrow, feeder, c + 1 < nchunks || cs.h_keep, b0, b1, vp, vm);
        __syncwarp();
        if (is_row) vout[hw] = make_uint2(vp, vm);
        const int val = is_row ? (__popc(vp) - __popc(vm)) : 0;
        int incl = val;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += y;
        }
        if (is_row) cumout[hw] = incl - val;  // running value inside the chunk; the chunk's base is added below
        if (lane == 31) cs.tot[c] = incl;
    }
    coop_bar<W>(COOP_BAR_MID);
    // running values: base of chunk c = top_val + sum of the chunks above it
    int total = 0;
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
        int v = c0 + lane < nchunks ? cs.tot[c0 + lane] : 0;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
        total += v;
    }
    for (int c = wid; c < nchunks; c += W) {
        int base = 0;
        for (int c0 = 0; c0 < c; c0 += 32) {
            int v = c0 + lane < c ? cs.tot[c0 + lane] : 0;
#pragma unroll
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
            base += v;
        }
        base += cs.top_val;
        const int nrow = min(per, nhw - per * c);
        const int rl = (c == 0 && cs.h_first) ? lane - 1 : lane;
        if (rl >= 0 && rl < nrow) cumout[per * c + rl] += base;
    }
    if (wid == 0 && lane == 0) cumout[nhw] = cs.top_val + total;
    coop_bar<W>(COOP_BAR_END);
    return cs.top_val + total;
}

// Leader side of a block (the dev_pass hook): blocks with a single chunk, or too many, run on the leader alone.
template <int W>
__device__ __forceinline__ Cost run_block_dp(CoopSmem<W>& cs, PairCtx& cx, const BlkView& prev, I is, int ncols, I njs, I nje,
                                             uint2* vout, int32_t* cumout, Cost top_val, const uint8_t* h_in = nullptr,
                                             uint8_t* h_out = nullptr, int tap_hw = -1, uint8_t* h_tap = nullptr) {
    const int lane = threadIdx.x & 31;
    const int nhw = (nje - njs) >> 5;
    const int nchunks = (nhw + 30) / 31;
    stage_amask(cs.lead, cx.aprof, is, ncols, lane);
    if (nchunks < 2 || nchunks > COOP_MAX_CHUNKS)
        return block_dp<false>(cs.lead, cx.bprof, prev, ncols, njs, nje, vout, cumout, top_val, nullptr, cx.dpc, h_in, h_out, tap_hw, h_tap);
    if (h_in)  // the top-edge deltas of the range: ring slot W is what chunk 0 reads as "the chunk above" (its progress word
               // starts at -1, which is past everything chunk 0 asks for); the slot is next written by chunk W, on this warp
        for (int k = lane; k < ncols; k += 32) cs.hrow[W][k] = h_in[k];
    if (lane == 0) {
        cs.quit = 0;
        cs.h_first = h_in ? 0 : 1;
        cs.h_keep = h_out ? 1 : 0;
        cs.tap_c = -1;
        cs.tap_lane = -1;
        if (tap_hw >= 0) {  // chunks of `per` half-words; chunk 0 gives lane 0 to the feeder when the top edge is +1
            const int per = (nhw + nchunks - 1) / nchunks;
            cs.tap_c = tap_hw / per;
            cs.tap_lane = tap_hw - cs.tap_c * per + ((cs.tap_c == 0 && !h_in) ? 1 : 0);
        }
        cs.ncols = ncols;
        cs.nhw = nhw;
        cs.nchunks = nchunks;
        cs.per = (nhw + nchunks - 1) / nchunks;
        cs.njs = njs;
        cs.top_val = top_val;
        cs.bprof = cx.bprof;
        cs.prev_v = prev.v;
        cs.prev_js = prev.js;
        cs.prev_je = prev.je;
        cs.prev_ones = prev.ones;
        cs.vout = vout;
        cs.cumout = cumout;
    }
    if (lane <= W) cs.progress[lane] = -1;
    __syncwarp();
    coop_bar<W>(COOP_BAR_START);
    cx.dpc.word_steps += (unsigned long long)ncols * (unsigned long long)nhw;
    cx.dpc.issue_steps += 32ull * (unsigned long long)(((nhw + 30) / 31) * (ncols - 1) + nhw + 1);  // chunks of 31 half-words, as block_dp counts
    const Cost bot = coop_work<W>(cs, 0);
    if (h_tap && tap_hw >= 0) {  // (after the END barrier)
        for (int k = lane; k < ncols; k += 32) h_tap[k] = cs.htap[k];
        __syncwarp();
    }
    if (h_out) {  // (after the END barrier) the bottom deltas of the last chunk are the new h row of these columns
        const uint8_t* last = cs.hrow[(nchunks - 1) % (W + 1)];
        for (int k = lane; k < ncols; k += 32) h_out[k] = last[k];
        __syncwarp();
    }
    return bot;
}

// Warps 1 .. W-1 of a cooperative CTA: serve blocks until the leader says quit.
template <int W>
__device__ __forceinline__ void coop_worker_loop(CoopSmem<W>& cs, int wid) {
    for (;;) {
        coop_bar<W>(COOP_BAR_START);
        if (cs.quit) return;
        coop_work<W>(cs, wid);
    }
}
template <int W>
__device__ __forceinline__ void coop_release_workers(CoopSmem<W>& cs) {
    if ((threadIdx.x & 31) == 0) cs.quit = 1;
    __syncwarp();
    coop_bar<W>(COOP_BAR_START);
}

}  // namespace APA_NS
