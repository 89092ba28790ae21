// General-parameter kernel of the B200 A*PA2 engine: the device code of apa_*.cuh compiled with APA_GENERAL=1, i.e. with
// AstarPa2Params (astarpa2/src/params.rs:8-42) read at run time instead of being the constants of the two presets.
// Serves apa_batch_run_params(): Domain::{Full, GapStart, GapGap, Astar(NoCost | GapCost | GCSH(k, p))},
// DoublingType::{None, BandDoubling{start, factor}, LinearSearch{start, delta}}, block_width 1..=256, dt_trace on/off,
// fr_drop, sparse_h on/off, prune on/off - the configurations of the reference's own test matrix
// (astarpa2/src/tests.rs:19-119) except the SH heuristic. One fused persistent kernel (build + passes + trace per pair,
// per-warp arenas): this path is for coverage and cross-checks, the tuned path is the preset one in apa_engine.cu.
#define APA_GENERAL 1
#include "apa_gcsh.cuh"
#include "apa_trace.cuh"

using namespace apa_gen;

namespace {

__device__ __forceinline__ void general_body(const BatchDev& bd, const RunParams& par, WarpSmem* smem) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpSmem& sm = smem[wib];
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + wib;
    uint8_t* arena = bd.arena + (size_t)slot * bd.arena_size;
    unsigned long long acc_steps = 0, acc_issue = 0, acc_cells = 0, acc_pass = 0, acc_fill = 0, acc_dt = 0, acc_h = 0, acc_probe = 0;

    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(bd.queue, 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= bd.n_order) break;
        const uint32_t p = bd.order[q];

        PairCtx cx;
        cx.par = par;
        cx.n = (I)(bd.a_off[p + 1] - bd.a_off[p]);
        cx.m = (I)(bd.b_off[p + 1] - bd.b_off[p]);
        cx.bprof = bd.bprof + bd.bp_off[p];
        cx.aprof = bd.aprof + bd.ap_off[p];
        cx.arena = arena;
        cx.arena_size = bd.arena_size;
        cx.nblk = (cx.n + par.block_width - 1) / par.block_width;
        cx.nblk_alloc = 0;
        cx.last_idx = 0;
        cx.meta = (BlkMeta*)arena;
        const uint64_t meta_bytes64 = ((uint64_t)(cx.nblk + 1) * sizeof(BlkMeta) + 15u) & ~15ull;
        const uint32_t meta_bytes = (uint32_t)min(meta_bytes64, (uint64_t)0xffffffffu);
        cx.v_base = meta_bytes;
        cx.incremental = par.incremental != 0;
        cx.hrow_off = cx.v_base;
        if (cx.incremental) cx.v_base += ((uint32_t)cx.n + 31u) & ~15u;
        cx.v_top = cx.v_base;
        cx.hi_bot = bd.arena_size;
        cx.cig_top = bd.arena_size;
        cx.status = ST_PENDING;
        cx.dpc.word_steps = cx.dpc.issue_steps = cx.computed_cells = 0;
        cx.passes = 0;
        cx.fill_blocks = cx.dt_blocks = 0;
        for (int t = 0; t < 8; t++) cx.tphase[t] = 0;
        cx.dbg = bd.dbg;
        cx.dbg_cap = bd.dbg_cap;
        cx.dbg_n = 0;
        if (meta_bytes64 + (uint64_t)cx.n + 4096u > bd.arena_size) cx.status = ST_OVERFLOW;

        Cost cost = -1;
        long long cig_off = -1, cig_len = 0;
        if (cx.status == ST_PENDING) {
            // AstarPa2::build + cost_or_align (lib.rs:87-175): h0 = h(0, 0) for Domain::Astar, 0 otherwise.
            Cost h0 = 0;
            long long st_matches = 0, st_hcalls = 0;
            if (par.domain != DOM_ASTAR || par.heuristic == 0) {
                NoneH hh;
                cost = dev_band_doubling<BD_WHOLE>(cx, sm, hh, 0);
            } else if (par.heuristic == 1) {
                GapH hh{cx.n, cx.m};
                h0 = hh.h(0, 0);
                cost = dev_band_doubling<BD_WHOLE>(cx, sm, hh, h0);
            } else {
                GcshH hh;
                hh.k_ = par.k;
                hh.p_ = par.p;
                if (gcsh_build(cx, sm, hh)) {
                    h0 = hh.h(0, 0);
                    cost = dev_band_doubling<BD_WHOLE>(cx, sm, hh, h0);
                    acc_h += hh.h_calls;
                    acc_probe += hh.probes;
                    st_matches = hh.M;
                    st_hcalls = (long long)hh.h_calls + 1;  // + the h(0,0) inside CSHI::new (csh.rs:296)
                }
            }
            if (lane == 0) {
                long long* ps = bd.pair_stats + 8ull * p;
                ps[1] = h0, ps[2] = st_matches, ps[3] = st_hcalls;
            }
            if (cx.status == ST_PENDING && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
        }
        if (cx.status == ST_PENDING && bd.trace) {
            CigarWriter cw;
            cw.arena = arena;
            cw.arena_size = cx.cig_top;
            cw.count = 0;
            cw.pend_cnt = 0;
            cw.pend_op = 0;
            cw.nbuf = 0;
            cw.buf = 0;
            if (dev_trace(cx, sm, cw, cost)) {
                cig_off = emit_cigar_text(cw, bd.pool, bd.pool_cursor, bd.pool_cap, &cig_len);
                if (cig_off < 0) cx.status = ST_OVERFLOW;
            }
        }
        if (lane == 0) {
            bd.status[p] = cx.status == ST_PENDING ? ST_DONE : cx.status;
            bd.cost[p] = cost;
            bd.cig_off[p] = cig_off;
            bd.cig_len[p] = cig_len;
            long long* ps = bd.pair_stats + 8ull * p;
            ps[0] = cx.passes, ps[4] = (long long)cx.computed_cells, ps[5] = cx.dt_blocks, ps[6] = cx.fill_blocks, ps[7] = 0;
        }
        if (bd.dbg_n && lane == 0) *bd.dbg_n = cx.dbg_n;
        acc_steps += cx.dpc.word_steps;
        acc_issue += cx.dpc.issue_steps;
        acc_cells += cx.computed_cells;
        acc_pass += cx.passes;
        acc_fill += cx.fill_blocks;
        acc_dt += cx.dt_blocks;
    }
    if (lane == 0) {
        atomicAdd(&bd.stats[0], acc_steps);
        atomicAdd(&bd.stats[15], acc_issue);
        atomicAdd(&bd.stats[1], acc_cells);
        atomicAdd(&bd.stats[2], acc_pass);
        atomicAdd(&bd.stats[3], acc_fill);
        atomicAdd(&bd.stats[4], acc_dt);
        atomicAdd(&bd.stats[13], acc_h);
        atomicAdd(&bd.stats[14], acc_probe);
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 6) apa_general_kernel(BatchDev bd, RunParams par) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    general_body(bd, par, smem);
}

}  // namespace

cudaError_t apa_general_launch(const BatchDev& bd, const RunParams& par, unsigned grid, cudaStream_t st) {
    apa_general_kernel<<<grid, WARPS_PER_CTA * 32, 0, st>>>(bd, par);
    return cudaGetLastError();
}
