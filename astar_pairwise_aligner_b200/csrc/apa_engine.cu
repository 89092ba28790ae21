// Host engine + kernels of the B200 A*PA2 path, exported through the C-ABI in include/astarpa.h and
// include/astarpa_b200.h.  There is NO CPU fallback: every entry point fails loudly without a usable device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/astarpa.h"
#include "../../include/astarpa_b200.h"
#include "apa_gcsh.cuh"
#include "apa_trace.cuh"
#include "apa_coop.cuh"

using namespace apa;

extern "C" int apa_pack_planes_host(const uint8_t* seq, int64_t len, int64_t hw_begin, int64_t hw_end, uint32_t* out);

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
static int set_err(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            return set_err(APA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

// ------------------------------------------------------------------------------------------------ kernels
#include "apa_batch.cuh"


// Streaming upload: wait until the upload chunk of work-order position q has landed in HBM. Returns 1 when the pair's packed
// planes are there, 2 when its raw bases are (the caller packs them), 0 when nothing came: the wait is bounded (about 20 s)
// so that a copy that never comes - a failed upload, a profiler that serialises the launch ahead of the copies - ends in
// ST_ASSERT for the pair instead of a hung GPU. The fence orders the loads of the pair's bases after the observation.
__device__ __forceinline__ int wait_ready(const BatchDev& bd, unsigned long long q) {
    if (!bd.chunk_state) return bd.raw_a ? 2 : 1;
    const int lane = threadIdx.x & 31;
    uint32_t st = 0;
    if (lane == 0) {
        const volatile uint32_t* flag = bd.chunk_state + bd.pair_chunk[q];
        unsigned long long spins = 0;
        while ((st = *flag) == 0u) {
            __nanosleep(500);
            if (++spins > 40000000ull) break;
        }
        __threadfence();
    }
    return (int)__shfl_sync(FULL, st, 0);
}
// Device-side K0 for the pair a warp is about to align (BatchDev::raw_a / raw_b): returns true on a byte outside ACGT.
__device__ __forceinline__ bool pack_pair(const BatchDev& bd, uint32_t p, I n, I m, int mode, WarpSmem& sm) {
    if (mode != 2) return false;
    static_assert(sizeof(WarpSmem) >= 1040, "dev_pack_planes lands 65 16-byte words in the warp's shared memory");
    const int nhw_a = (int)(bd.ap_off[p + 1] - bd.ap_off[p]), nhw_b = (int)(bd.bp_off[p + 1] - bd.bp_off[p]);
    bool bad = dev_pack_planes(bd.raw_a + bd.a_off[p], n, bd.aprof + bd.ap_off[p], nhw_a, (uint32_t*)&sm);
    bad |= dev_pack_planes(bd.raw_b + bd.b_off[p], m, bd.bprof + bd.bp_off[p], nhw_b, (uint32_t*)&sm);
    __syncwarp();
    return bad;
}

// K1+K3 fused per pair: a persistent warp pulls pairs from the device work queue and runs the band-doubling
// search, the traceback and the CIGAR text emission for each.
__device__ __forceinline__ void apa_align_body(const BatchDev& bd, WarpSmem* smem) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpSmem& sm = smem[wib];
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + wib;
    uint8_t* arena = bd.arena + (size_t)slot * bd.arena_size;
    unsigned long long acc_steps = 0, acc_issue = 0, acc_cells = 0, acc_pass = 0, acc_fill = 0, acc_dt = 0;
    long long acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(bd.queue, 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= bd.n_order) break;
        const int landed = wait_ready(bd, q);  // streaming upload: this pair's bases are in HBM (1 planes, 2 raw, 0 never came)
        const uint32_t p = bd.order[q];

        PairCtx cx;
        cx.n = (I)(bd.a_off[p + 1] - bd.a_off[p]);
        cx.m = (I)(bd.b_off[p + 1] - bd.b_off[p]);
        cx.bprof = bd.bprof + bd.bp_off[p];
        cx.aprof = bd.aprof + bd.ap_off[p];
        cx.arena = arena;
        cx.arena_size = bd.arena_size;
        cx.nblk = (cx.n + BLOCK_W - 1) / BLOCK_W;
        cx.nblk_alloc = 0;
        cx.last_idx = 0;
        cx.meta = (BlkMeta*)arena;
        uint32_t meta_bytes = ((uint32_t)(cx.nblk + 1) * (uint32_t)sizeof(BlkMeta) + 15u) & ~15u;
        cx.v_base = meta_bytes;
        cx.incremental = bd.preset == APA_PRESET_FULL;  // params.rs:88,119
        cx.hrow_off = cx.v_base;
        if (cx.incremental) cx.v_base += ((uint32_t)cx.n + 31u) & ~15u;
        cx.v_top = cx.v_base;
        cx.hi_bot = bd.arena_size;
        cx.cig_top = bd.arena_size;
        cx.more = 0;
        cx.bd_last_s = cx.bd_s = cx.bd_maxs = 0;
        cx.status = ST_PENDING;
        cx.dpc.word_steps = cx.dpc.issue_steps = cx.computed_cells = 0;
        cx.passes = 0;
        cx.fill_blocks = cx.dt_blocks = 0;
        for (int t = 0; t < 8; t++) cx.tphase[t] = 0;
        cx.dbg = bd.dbg;
        cx.dbg_cap = bd.dbg_cap;
        cx.dbg_n = 0;
        if ((uint64_t)cx.v_base + 4096u > bd.arena_size) cx.status = ST_OVERFLOW;
        if (!landed) cx.status = ST_ASSERT;
        else if (pack_pair(bd, p, cx.n, cx.m, landed, sm)) cx.status = ST_BAD_INPUT;

        Cost cost = -1;
        long long cig_off = -1, cig_len = 0;
        long long st_h0 = 0, st_matches = 0, st_hcalls = 0;
        if (cx.status == ST_PENDING) {
            if (bd.preset == APA_PRESET_SIMPLE) {
                GapH hh{cx.n, cx.m};
                Cost h0 = hh.h(0, 0);
                st_h0 = h0;
                long long t0 = APA_TIC();
                cost = dev_band_doubling<BD_WHOLE>(cx, sm, hh, h0);
                APA_TOC(cx.tphase[2], t0);
                if (cx.status == ST_PENDING && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
            } else {
                GcshH hh;
                long long t0 = APA_TIC();
                bool built = gcsh_build(cx, sm, hh);
                APA_TOC(cx.tphase[0], t0);
                if (built) {
                    Cost h0 = hh.h(0, 0);
                    t0 = APA_TIC();
                    cost = dev_band_doubling<BD_WHOLE>(cx, sm, hh, h0);
                    APA_TOC(cx.tphase[2], t0);
                    cx.tphase[6] += hh.t_h;
                    st_h0 = h0;
                    st_matches = hh.M;
                    st_hcalls = (long long)hh.h_calls + 1;  // + the h(0,0) inside CSHI::new (csh.rs:296), not needed here
                    if (cx.status == ST_PENDING && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
                }
            }
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
            long long t0 = APA_TIC();
            bool traced = dev_trace(cx, sm, cw, cost);
            APA_TOC(cx.tphase[3], t0);
            if (traced) {
                t0 = APA_TIC();
                cig_off = emit_cigar_text(cw, bd.pool, bd.pool_cursor, bd.pool_cap, &cig_len);
                APA_TOC(cx.tphase[5], t0);
                if (cig_off < 0) cx.status = ST_OVERFLOW;
            }
        }
        if (lane == 0) {
            bd.status[p] = cx.status == ST_PENDING ? ST_DONE : cx.status;
            bd.cost[p] = cost;
            bd.cig_off[p] = cig_off;
            bd.cig_len[p] = cig_len;
        }
        if (lane == 0) {
            long long* ps = bd.pair_stats + 8ull * p;
            ps[0] = cx.passes, ps[1] = st_h0, ps[2] = st_matches, ps[3] = st_hcalls, ps[4] = (long long)cx.computed_cells;
            ps[5] = cx.dt_blocks, ps[6] = cx.fill_blocks, ps[7] = 0;
        }
        if (bd.dbg_n && lane == 0) *bd.dbg_n = cx.dbg_n;
        acc_steps += cx.dpc.word_steps;
        acc_issue += cx.dpc.issue_steps;
        acc_cells += cx.computed_cells;
        acc_pass += cx.passes;
        acc_fill += cx.fill_blocks;
        acc_dt += cx.dt_blocks;
        for (int t = 0; t < 8; t++) acc_t[t] += cx.tphase[t];
    }
    if (lane == 0) {
        atomicAdd(&bd.stats[0], acc_steps);
        atomicAdd(&bd.stats[15], acc_issue);
        atomicAdd(&bd.stats[1], acc_cells);
        atomicAdd(&bd.stats[2], acc_pass);
        atomicAdd(&bd.stats[3], acc_fill);
        atomicAdd(&bd.stats[4], acc_dt);
        for (int t = 0; t < 8; t++) atomicAdd(&bd.stats[5 + t], (unsigned long long)acc_t[t]);
    }
}

// ---- phase-split path ---------------------------------------------------------------------------------------------
// The same per-pair work cut into three persistent kernels — (0) heuristic build, (1) band-doubling passes, (2) traceback
// + CIGAR text — run back to back over a wave of pairs. Every pair of the wave owns an arena for the whole wave; the
// warp-uniform state (PairCtx, GcshH, cost) is parked in the arena header between phases. Compared with the fused kernel
// each phase kernel has a smaller instruction footprint (the fused one stalls on instruction fetch: 'no_instruction' is
// the #3 stall reason in profiles/r1_full_final_n20k_ncu_full.txt) and fewer live registers.
struct PairState {
    PairCtx cx;
    GcshH hh;
    Cost cost;
    Cost h0;
};
constexpr uint32_t ARENA_HEADER = 512;
static_assert(sizeof(PairState) <= ARENA_HEADER, "PairState must fit the arena header");

template <int PHASE, class SM>
__device__ __forceinline__ void apa_phase_body(const BatchDev& bd, SM& sm) {
    const int lane = threadIdx.x & 31;
    if (PHASE == 3 && *(volatile unsigned long long*)&bd.stats[16] == 0ull) return;  // no pair asked for a second pass
    unsigned long long acc_steps = 0, acc_issue = 0, acc_cells = 0, acc_pass = 0, acc_fill = 0, acc_dt = 0, acc_h = 0, acc_probe = 0;
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(bd.queue + 24 + PHASE, 1ull) + bd.q0;
        q = __shfl_sync(FULL, q, 0);
        if (q >= bd.n_order) break;
        int landed = 1;
        if (PHASE == 0) landed = wait_ready(bd, q);  // later phases run after the build kernel, which saw every pair land
        if ((PHASE == 1 || PHASE == 2) && bd.phase_flag) {  // overlapped kernels: the previous phase of this pair is done (bounded wait, as
                                                            // above); the continuation kernel follows the pass kernel on its stream
            int okf = 1;
            if (lane == 0) {
                unsigned long long spins = 0;
                while (bd.phase_flag[q] < (uint8_t)PHASE) {
                    __nanosleep(1000);
                    if (++spins > 20000000ull) {
                        okf = 0;
                        break;
                    }
                }
                __threadfence();
            }
            okf = __shfl_sync(FULL, okf, 0);
            if (!okf) {  // never came: report, and let the next phase pass through
                if (lane == 0) {
                    bd.status[bd.order[q]] = ST_ASSERT;
                    __threadfence();
                    bd.phase_flag[q] = 2;
                }
                continue;
            }
        }
        const uint32_t p = bd.order[q];
        uint8_t* arena = bd.arena + (size_t)(q - bd.q0) * bd.arena_size;
        PairState* ps = (PairState*)arena;
        PairCtx cx;
        if constexpr (PHASE == 0) {
            cx.n = (I)(bd.a_off[p + 1] - bd.a_off[p]);
            cx.m = (I)(bd.b_off[p + 1] - bd.b_off[p]);
            cx.bprof = bd.bprof + bd.bp_off[p];
            cx.aprof = bd.aprof + bd.ap_off[p];
            cx.arena = arena;
            cx.arena_size = bd.arena_size;
            cx.nblk = (cx.n + BLOCK_W - 1) / BLOCK_W;
            cx.nblk_alloc = 0;
        cx.last_idx = 0;
            cx.meta = (BlkMeta*)(arena + ARENA_HEADER);
            uint32_t meta_bytes = ((uint32_t)(cx.nblk + 1) * (uint32_t)sizeof(BlkMeta) + 15u) & ~15u;
            cx.v_base = ARENA_HEADER + meta_bytes;
            cx.incremental = bd.preset == APA_PRESET_FULL;  // params.rs:88,119
            cx.hrow_off = cx.v_base;
            if (cx.incremental) cx.v_base += ((uint32_t)cx.n + 31u) & ~15u;
            cx.v_top = cx.v_base;
            cx.hi_bot = bd.arena_size;
            cx.cig_top = bd.arena_size;
            cx.more = 0;
            cx.bd_last_s = cx.bd_s = cx.bd_maxs = 0;
            cx.status = ST_PENDING;
            cx.dpc.word_steps = cx.dpc.issue_steps = cx.computed_cells = 0;
            cx.passes = 0;
            cx.fill_blocks = cx.dt_blocks = 0;
            for (int t = 0; t < 8; t++) cx.tphase[t] = 0;
            cx.dbg = bd.dbg;
            cx.dbg_cap = bd.dbg_cap;
            cx.dbg_n = 0;
            if ((uint64_t)cx.v_base + 4096u > bd.arena_size) cx.status = ST_OVERFLOW;
            if (!landed) cx.status = ST_ASSERT;
            else if (pack_pair(bd, p, cx.n, cx.m, landed, sm)) cx.status = ST_BAD_INPUT;
            GcshH hh;
            if (cx.status == ST_PENDING && bd.preset == APA_PRESET_FULL) gcsh_build(cx, sm, hh);
            __syncwarp();
            if (lane == 0) {
                ps->cx = cx;
                if (bd.preset == APA_PRESET_FULL) ps->hh = hh;
                ps->cost = -1;
            }
            __syncwarp();
            if (bd.phase_flag) {  // everything this warp wrote for the pair is visible before the flag is
                __threadfence();
                __syncwarp();
                if (lane == 0) bd.phase_flag[q] = 1;
            }
            continue;
        }
        cx = ps->cx;
        Cost cost = ps->cost;
        // PHASE 1: the first pass of every pair (compiled without the incremental-doubling code). PHASE 3: the continuation
        // kernel - second and later passes of the pairs the first pass did not settle (cx.more), from the saved search state.
        if constexpr (PHASE == 1 || PHASE == 3) {
            if (PHASE == 3 && !cx.more) continue;
            if (cx.status == ST_PENDING) {
                long long* pst = bd.pair_stats + 8ull * p;
                constexpr int MODE = PHASE == 1 ? BD_FIRST : BD_CONTINUE;
                if (bd.preset == APA_PRESET_SIMPLE) {
                    GapH hh{cx.n, cx.m};
                    Cost h0 = hh.h(0, 0);
                    cost = dev_band_doubling<MODE>(cx, sm, hh, h0);
                    if (cx.status == ST_PENDING && cost != BD_MORE && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
                    if (lane == 0) pst[1] = h0, pst[2] = 0, pst[3] = 0;
                } else {
                    GcshH hh = ps->hh;
                    if (PHASE == 1) hh.h_calls = hh.probes = 0;
                    Cost h0 = ps->h0;
                    if (PHASE == 1) h0 = hh.h(0, 0);
                    cost = dev_band_doubling<MODE>(cx, sm, hh, h0);
                    if (cx.status == ST_PENDING && cost != BD_MORE && h0 > cost) cx.status = ST_ASSERT;
                    if (cost == BD_MORE) {  // the heuristic (pruned matches, search hints, counters) goes on in the continuation kernel
                        __syncwarp();
                        if (lane == 0) ps->hh = hh, ps->h0 = h0;
                    } else {
                        acc_h += hh.h_calls;
                        acc_probe += hh.probes;
                        if (lane == 0) pst[1] = h0, pst[2] = hh.M, pst[3] = (long long)hh.h_calls + 1;  // + CSHI::new's own h(0,0) (csh.rs:296)
                    }
                }
            }
            cx.more = (cx.status == ST_PENDING && cost == BD_MORE) ? 1 : 0;
            __syncwarp();
            if (lane == 0) {
                ps->cx = cx;
                ps->cost = cost;
            }
            __syncwarp();
            if (cx.more) {  // (PHASE 1 only) the pair's passes are not done yet: the continuation kernel has work
                if (lane == 0) atomicAdd(&bd.stats[16], 1ull);
                continue;
            }
            if (bd.phase_flag && bd.trace) {
                __threadfence();
                __syncwarp();
                if (lane == 0) bd.phase_flag[q] = 2;
            }
            if (bd.trace) continue;
        }
        // PHASE 2, or PHASE 1 / 3 of a cost-only run: finish the pair
        long long cig_off = -1, cig_len = 0;
        if constexpr (PHASE == 2) if (cx.status == ST_PENDING && bd.trace) {
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
            long long* pst = bd.pair_stats + 8ull * p;
            pst[0] = cx.passes, pst[4] = (long long)cx.computed_cells, pst[5] = cx.dt_blocks, pst[6] = cx.fill_blocks, pst[7] = 0;
        }
        if (bd.dbg_n && lane == 0) *bd.dbg_n = cx.dbg_n;
        acc_steps += cx.dpc.word_steps;
        acc_issue += cx.dpc.issue_steps;
        acc_cells += cx.computed_cells;
        acc_pass += cx.passes;
        acc_fill += cx.fill_blocks;
        acc_dt += cx.dt_blocks;
    }
    if ((PHASE == 1 || PHASE == 3) && lane == 0) {
        atomicAdd(&bd.stats[13], acc_h);
        atomicAdd(&bd.stats[14], acc_probe);
    }
    if (PHASE >= 1 && lane == 0) {
        atomicAdd(&bd.stats[0], acc_steps);
        atomicAdd(&bd.stats[15], acc_issue);
        atomicAdd(&bd.stats[1], acc_cells);
        atomicAdd(&bd.stats[2], acc_pass);
        atomicAdd(&bd.stats[3], acc_fill);
        atomicAdd(&bd.stats[4], acc_dt);
    }
}
// Each phase kernel exists in three register budgets (CTAs of 128 threads: 10 / 9 / 8 per SM = 48 / 56 / 64 registers);
// the host picks per phase (APA_BUILD_REGS / APA_PASS_REGS / APA_TRACE_REGS override the defaults).
#define APA_PHASE_KERNEL(NAME, PHASE, MINB)                                                     \
    __global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) NAME(BatchDev bd) {             \
        __shared__ WarpSmem smem[WARPS_PER_CTA];                                                \
        tma_stage_reset(smem[threadIdx.x >> 5]);                                                \
        apa_phase_body<PHASE>(bd, smem[threadIdx.x >> 5]);                                      \
    }
APA_PHASE_KERNEL(apa_phase_build_kernel, 0, 10)
APA_PHASE_KERNEL(apa_phase_build_kernel_r56, 0, 9)
APA_PHASE_KERNEL(apa_phase_build_kernel_r64, 0, 8)
APA_PHASE_KERNEL(apa_phase_pass_kernel, 1, 10)
APA_PHASE_KERNEL(apa_phase_pass_kernel_r56, 1, 9)
APA_PHASE_KERNEL(apa_phase_pass_kernel_r64, 1, 8)
APA_PHASE_KERNEL(apa_phase_cont_kernel, 3, 8)
APA_PHASE_KERNEL(apa_phase_trace_kernel, 2, 10)
APA_PHASE_KERNEL(apa_phase_trace_kernel_r56, 2, 9)
APA_PHASE_KERNEL(apa_phase_trace_kernel_r64, 2, 8)
// Pass kernel with the W warps of a CTA on one pair (apa_coop.cuh): warp 0 runs the pair, the others serve its tall blocks.
template <int W, int PHASE>
__global__ void __launch_bounds__(W * 32, 32 / W) apa_phase_pass_coop_kernel(BatchDev bd) {
    __shared__ CoopSmem<W> cs;
    const int wid = threadIdx.x >> 5;
    if (wid == 0) {
        tma_stage_reset(cs.lead);
        apa_phase_body<PHASE>(bd, cs);
        coop_release_workers<W>(cs);
    } else {
        coop_worker_loop<W>(cs, wid);
    }
}
typedef void (*phase_kernel_t)(BatchDev);
static phase_kernel_t phase_kernel(int phase, int regs) {
    static const phase_kernel_t tab[3][3] = {{apa_phase_build_kernel, apa_phase_build_kernel_r56, apa_phase_build_kernel_r64},
                                             {apa_phase_pass_kernel, apa_phase_pass_kernel_r56, apa_phase_pass_kernel_r64},
                                             {apa_phase_trace_kernel, apa_phase_trace_kernel_r56, apa_phase_trace_kernel_r64}};
    return tab[phase][regs >= 64 ? 2 : (regs >= 56 ? 1 : 0)];
}
static int phase_regs(int phase) {  // default register budget per phase, overridable for experiments
    static const char* names[3] = {"APA_BUILD_REGS", "APA_PASS_REGS", "APA_TRACE_REGS"};
    static const int defaults[3] = {56, 48, 48};  // measured on B200 (ms per 10 000 pairs at 48 / 56 / 64): build 35.8 / 34.9 / 37.3, pass 45.6 / 47.1 / 45.5, trace 20.4 / 20.6 / 21.3
    const char* ev = getenv(names[phase]);
    return ev ? atoi(ev) : defaults[phase];
}

// Two register budgets of the same kernel: 64 registers (8 CTAs = 32 warps per SM, few spills) and 40 registers
// (12 CTAs = 48 warps per SM). The host picks per batch: the number of warp slots is chosen so that the pairs fill
// whole waves, and the variant follows from the slots needed per SM.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 8) apa_align_kernel_r64(BatchDev bd) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    tma_stage_reset(smem[threadIdx.x >> 5]);
    apa_align_body(bd, smem);
}
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 10) apa_align_kernel_r48(BatchDev bd) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    tma_stage_reset(smem[threadIdx.x >> 5]);
    apa_align_body(bd, smem);
}
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 12) apa_align_kernel_r40(BatchDev bd) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    tma_stage_reset(smem[threadIdx.x >> 5]);
    apa_align_body(bd, smem);
}

// Stand-alone block-DP rectangle (apa_block_compute = pa_bitpacking::simd::compute with HMode::Update): one warp; h holds the
// top-edge deltas on entry and the bottom-edge deltas on return (one byte per column: bit0 = +1, bit1 = -1).
__global__ void apa_block_kernel(const uint2* aprof, int na, const uint2* bprof, int nhw, uint2* v, int32_t* cum, uint8_t* h) {
    __shared__ WarpSmem sm;
    tma_stage_reset(sm);
    const int lane = threadIdx.x & 31;
    BlkView prev;
    prev.js = 0;
    prev.je = nhw * 32;
    prev.top_val = 0;
    prev.bot_val = 0;
    prev.v = v;
    prev.cum = nullptr;
    prev.ones = 0;
    DpCounters ws{0, 0};
    // the rectangle may be wider than one block: sweep it in 256-column slabs, each slab's right column feeding the next
    for (int c0 = 0; c0 < na; c0 += BLOCK_W) {
        int nc = min(BLOCK_W, na - c0);
        stage_amask(sm, aprof, c0, nc, lane);
        block_dp<false>(sm, bprof, prev, nc, 0, nhw * 32, v, cum, 0, nullptr, ws, h + c0, h + c0);
        __syncwarp();
    }
}

// K0 as a kernel of its own (resident uploads from page-locked host memory: raw bases by DMA, packed here): CTA x takes every
// gridDim.x-th pair, its warps (and those of the gridDim.y CTAs sharing the pair) every (8 gridDim.y)-th group of 32 half-words
// of a, then of b. *bad is set when a byte outside ACGT was seen.
__global__ void __launch_bounds__(256) apa_pack_kernel(BatchDev bd, int* bad) {
    __shared__ uint4 land[8][66];
    const int wid = threadIdx.x >> 5, first = blockIdx.y * 8 + wid, stride = gridDim.y * 8;
    uint32_t* sbuf = (uint32_t*)land[wid];
    bool any_bad = false;
    for (uint64_t p = blockIdx.x; p < bd.n_pairs; p += gridDim.x) {
        const I n = (I)(bd.a_off[p + 1] - bd.a_off[p]), m = (I)(bd.b_off[p + 1] - bd.b_off[p]);
        any_bad |= dev_pack_planes(bd.raw_a + bd.a_off[p], n, bd.aprof + bd.ap_off[p], (int)(bd.ap_off[p + 1] - bd.ap_off[p]), sbuf, first, stride);
        any_bad |= dev_pack_planes(bd.raw_b + bd.b_off[p], m, bd.bprof + bd.bp_off[p], (int)(bd.bp_off[p + 1] - bd.bp_off[p]), sbuf, first, stride);
    }
    if (any_bad && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}

// ------------------------------------------------------------------------------------------------ host side
// pa_bitpacking::search (pa-bitpacking/src/search.rs:46-118, simd/scatter_profile.rs): the text runs along the columns in slabs
// of 256, the (short) pattern along the rows. Top deltas are 0 (a match may start anywhere in the text), so every chunk takes
// its incoming deltas from sm.hrow (zeroed for the first chunk) and leaves its bottom deltas there; what the last chunk leaves
// is the bottom row of the slab. The equality words are the pattern's match masks (wildcards N * Y R and the all-matching
// padding rows, profile.rs:39-66), which the per-lane table of the block DP takes as they are.
//
// One warp per TEXT SEGMENT. A cell (i, j) costs at most j (start in the top row above it), so an optimal path to it spans at
// most 2 j <= 2 rows columns: a warp that starts `warm` = 2 * rows columns before its segment from all-(+1) vertical deltas
// (an upper bound of the true column) has exact values from its segment's first column on. The same argument is what
// SearchResult::trace uses for its window (search.rs:141-186). Segment s covers columns [s * seg_len, (s + 1) * seg_len);
// hdelta is written for those columns only; the warp of the last segment leaves the final column in vfin.
// FILL: single segment starting at column 0 of `text` with v0 as given; every column's V is stored: fillvals[(c + 1) * nhw + hw]
// (column 0 of fillvals = v0), for SearchResult::trace.
template <bool FILL>
__global__ void __launch_bounds__(128) apa_search_kernel(const uint8_t* __restrict__ text, int nt, const uint4* __restrict__ pmask, int nhw,
                                                       const uint2* __restrict__ v0, uint2* __restrict__ vwork, uint2* __restrict__ vfin,
                                                       int8_t* __restrict__ hdelta, int seg_len, int warm, uint2* __restrict__ fillvals) {
    __shared__ WarpSmem smem[4];
    WarpSmem& sm = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * 4 + (threadIdx.x >> 5);
    const long long seg_start = (long long)seg * seg_len;
    if (seg_start >= nt && !(seg == 0)) return;
    const int seg_end = (int)min((long long)nt, seg_start + seg_len);
    const int c_begin = (int)max(0ll, seg_start - warm);
    uint2* v = vwork + (size_t)seg * nhw;
    for (int hw = lane; hw < nhw; hw += 32) {
        v[hw] = c_begin == 0 ? v0[hw] : make_uint2(~0u, 0u);
        if (FILL) fillvals[hw] = v0[hw];
    }
    __syncwarp();
    const int nchunks = (nhw + 31) / 32;
    for (int c0 = c_begin; c0 < seg_end; c0 += BLOCK_W) {
        const int nc = min(BLOCK_W, seg_end - c0);
        for (int k = lane; k < BLOCK_W + 4; k += 32) sm.achar[k] = k < nc ? (uint8_t)((text[c0 + k] >> 1) & 3u) : (uint8_t)0;  // A0 C1 T2 G3
        for (int k = lane; k < BLOCK_W; k += 32) sm.hrow[k] = 0;
        __syncwarp();
        for (int c = 0; c < nchunks; c++) {
            const int nrow = min(32, nhw - 32 * c);
            const bool is_row = lane < nrow;
            const int hw = 32 * c + (is_row ? lane : 0);
            uint32_t vp = 0u, vm = 0u;
            uint4 pm = make_uint4(0u, 0u, 0u, 0u);
            if (is_row) {
                const uint2 x = v[hw];
                vp = x.x, vm = x.y;
                pm = pmask[hw];
            }
            sm.etab[0 * 32 + lane] = pm.x;
            sm.etab[1 * 32 + lane] = pm.y;
            sm.etab[2 * 32 + lane] = pm.z;
            sm.etab[3 * 32 + lane] = pm.w;
            dp_chunk<FILL, false, true, true>(sm, nc, nrow, 0u, 0u, vp, vm, FILL ? fillvals + (size_t)(c0 + 1) * nhw + hw : nullptr, nhw);
            __syncwarp();
            if (is_row) v[hw] = make_uint2(vp, vm);
        }
        for (int k = lane; k < nc; k += 32) {
            if (hdelta && c0 + k >= seg_start) {
                const uint32_t x = sm.hrow[k];
                hdelta[c0 + k] = (int8_t)((int)(x & 1u) - (int)(x >> 1));
            }
        }
        __syncwarp();
    }
    if (seg_end == nt)
        for (int hw = lane; hw < nhw; hw += 32) vfin[hw] = v[hw];
}

// SearchResult::trace walk (search.rs:188-229) on the filled window: columns [start, end] of the text = fill columns 0 .. end - start,
// from (end, pj) with cost `target` back to column `start` or row 0. One warp: V::value_to is a strided popcount sum + shuffle
// reduction, everything else is warp-uniform. out: [0] status (0 walked, 1 the window's cost at the end position exceeds the
// target: re-fill a wider window, 2 cheaper than the target / stuck: a reference panic), [1] number of elements, [2] start i,
// [3] start j, [4] cost found at the end position. elems: CIGAR elements newest first, (op << 30) | count.
__global__ void apa_search_trace_kernel(const uint8_t* __restrict__ text, const uint4* __restrict__ pmask, int nhw, const uint2* __restrict__ fillvals,
                                        int start, int end, int pj, int target, uint32_t* __restrict__ elems, uint32_t elem_cap, int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    auto value_to = [&](int i, int j) -> int {  // encoding.rs:54-63 on 32-row half-words
        const uint2* col = fillvals + (size_t)(i - start) * nhw;
        int sum = 0;
        for (int hw = lane; 32 * hw < j; hw += 32) {
            const uint2 pm = col[hw];
            const int bits = j - 32 * hw;
            const uint32_t mask = bits >= 32 ? ~0u : ((1u << bits) - 1u);
            sum += __popc(pm.x & mask) - __popc(pm.y & mask);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
        return sum;
    };
    auto is_match = [&](int i, int j) -> bool {  // ScatterProfile::is_match, profile.rs:72-74
        const uint4 pm = pmask[j >> 5];
        const uint32_t c = (text[i] >> 1) & 3u;
        const uint32_t w = c == 0 ? pm.x : (c == 1 ? pm.y : (c == 2 ? pm.z : pm.w));
        return (w >> (j & 31)) & 1u;
    };
    const int cost_end = value_to(end, pj);
    if (lane == 0) out[4] = cost_end;
    if (cost_end != target) {
        if (lane == 0) out[0] = cost_end > target ? 1 : 2, out[1] = 0;
        return;
    }
    int pi = end, g = target;
    uint32_t n_el = 0, pend_op = 0, pend_cnt = 0;
    int status = 0;
    auto push = [&](uint32_t op, uint32_t cnt) {
        if (pend_cnt && pend_op == op) {
            pend_cnt += cnt;
            return;
        }
        if (pend_cnt) {
            if (n_el < elem_cap && lane == 0) elems[n_el] = cig_pack(pend_op, pend_cnt);
            n_el++;
        }
        pend_op = op, pend_cnt = cnt;
    };
    while (pi > start && pj > 0) {
        uint32_t cnt = 0;
        while (pi > start && pj > 0 && is_match(pi - 1, pj - 1)) cnt++, pi--, pj--;
        if (cnt > 0) {
            push(OP_MATCH, cnt);
            continue;
        }
        if (value_to(pi - 1, pj) == g - 1) {
            g--, pi--;
            push(OP_DEL, 1);
            continue;
        }
        if (value_to(pi, pj - 1) == g - 1) {
            g--, pj--;
            push(OP_INS, 1);
            continue;
        }
        if (value_to(pi - 1, pj - 1) == g - 1) {
            g--, pi--, pj--;
            push(OP_SUB, 1);
            continue;
        }
        status = 2;  // "Bad trace! Got stuck"
        break;
    }
    if (pend_cnt) {
        if (n_el < elem_cap && lane == 0) elems[n_el] = cig_pack(pend_op, pend_cnt);
        n_el++;
    }
    if (status == 0 && !(pi == 0 || g == 0)) status = 2;  // assert!(pos.0 == 0 || g == 0), search.rs:226
    if (lane == 0) out[0] = status, out[1] = (int)n_el, out[2] = pi, out[3] = pj;
}

struct apa_engine;
static cudaError_t eng_alloc(apa_engine* e, void** out, size_t bytes);
static void eng_release(apa_engine* e, void* p);

struct apa_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // streaming uploads overlap the persistent kernel: raw bases (DMA only)
    cudaStream_t copy_stream2 = nullptr;  // ... and host-packed planes, issued by the packing threads
    cudaStream_t st_pass = nullptr, st_trace = nullptr;  // overlapped phase kernels (lower priority than `stream`)
    cudaEvent_t ev_ov[3] = {};                           // start of the overlapped launch | pass kernel done | trace kernel done
    uint8_t* d_phase_flag = nullptr;
    size_t phase_flag_cap = 0;
    cudaEvent_t ev_ring[4] = {};          // pacing of the raw copies (two chunks in flight)
    uint32_t* d_ready = nullptr;  // [0, 256): per-chunk upload state (BatchDev::chunk_state), [300]: bad-input flag of apa_pack_kernel
    uint32_t* h_ready = nullptr;  // pinned constants the state flags are copied from: h_ready[1] = 1, h_ready[2] = 2
    uint32_t* h_stage = nullptr;  // pinned staging for the packed planes (grow-only)
    size_t h_stage_cap = 0;
    cudaEvent_t ev[6] = {};
    cudaEvent_t evp[4] = {};  // phase-split path: before build | after build | after passes | after trace
    // Device-buffer cache: cudaMalloc/cudaFree of multi-GB buffers costs milliseconds each and synchronises the device,
    // so buffers released by a batch are kept for the next one (grow-only, per engine).
    std::unordered_map<void*, size_t> live;
    std::vector<std::pair<void*, size_t>> free_blocks;
    unsigned long long* d_queue = nullptr;   // [0] queue head, [1] pool cursor, [2..16] stats, [24..26] phase-kernel queue heads
    uint8_t* d_arena = nullptr;
    size_t arena_total = 0;
};

struct apa_batch {
    uint64_t n_pairs = 0;
    std::vector<int64_t> a_off, b_off, bp_off, ap_off;
    uint64_t total_a = 0, total_b = 0, total_hw = 0, total_hw_a = 0;
    I max_n = 0, max_m = 0;
    int64_t *d_a_off = nullptr, *d_b_off = nullptr, *d_bp_off = nullptr, *d_ap_off = nullptr;
    uint2 *d_bprof = nullptr, *d_aprof = nullptr;
    uint8_t *d_araw = nullptr, *d_braw = nullptr;  // raw bases (device-side K0); kept for the whole run, see batch_run
    bool raw = false;                              // inputs are page-locked: upload raw bytes, pack on the device
    int32_t *d_status = nullptr, *d_cost = nullptr;
    int64_t *d_cig_off = nullptr, *d_cig_len = nullptr;
    long long* d_pair_stats = nullptr;  // 8 per pair (apa_pair_stats)
    uint32_t* d_order = nullptr;
    uint16_t* d_pair_chunk = nullptr;  // upload chunk of every work-order position (streaming upload)
    uint32_t chunks_raw = 0;           // chunks of the last streamed upload that went as raw bases
    char* d_pool = nullptr;
    uint64_t pool_cap = 0;
    // streaming upload (apa_align_batch): bases are copied chunk by chunk while the kernel already runs
    const uint8_t *h_a = nullptr, *h_b = nullptr;
    const uint32_t *h_ap = nullptr, *h_bp = nullptr;  // packed (2-bit plane) input of apa_align_batch_packed, in the device layout
    int64_t h_a_off0 = 0, h_b_off0 = 0;
    double pack_ms = 0;
    std::vector<uint32_t> chunk_end;  // pairs (in work order) available after each chunk
    std::vector<uint32_t> chunk_pair_end;  // pair index (exclusive) of each chunk: chunks are contiguous index ranges
    bool ran = false;
    int trace = 0;
    apa_batch_stats stats{};
    std::vector<int32_t> h_status;
    uint64_t pool_used = 0;
    int32_t* d_dbg = nullptr;  // set by apa_debug_band_log
    uint32_t dbg_cap = 0;
    uint32_t* d_dbg_n = nullptr;
};

static cudaError_t eng_alloc(apa_engine* e, void** out, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    int best = -1;
    for (size_t k = 0; k < e->free_blocks.size(); k++) {
        size_t cap = e->free_blocks[k].second;
        if (cap >= bytes && cap <= 2 * bytes + (1u << 20) && (best < 0 || cap < e->free_blocks[best].second)) best = (int)k;
    }
    if (best >= 0) {
        *out = e->free_blocks[best].first;
        e->live[*out] = e->free_blocks[best].second;
        e->free_blocks.erase(e->free_blocks.begin() + best);
        return cudaSuccess;
    }
    cudaError_t ce = cudaMalloc(out, bytes);
    if (ce != cudaSuccess) {  // drop the cache and retry once
        cudaGetLastError();
        for (auto& fb : e->free_blocks) cudaFree(fb.first);
        e->free_blocks.clear();
        ce = cudaMalloc(out, bytes);
    }
    if (ce != cudaSuccess && e->d_arena) {  // the scratch arena of an earlier batch may hold nearly all of HBM (wave-sized work lists):
        cudaGetLastError();                   // give it back; batch_run sizes a new one from what is free then
        cudaFree(e->d_arena);
        e->d_arena = nullptr;
        e->arena_total = 0;
        ce = cudaMalloc(out, bytes);
    }
    if (ce == cudaSuccess) e->live[*out] = bytes;
    return ce;
}
static void eng_release(apa_engine* e, void* p) {
    if (!p) return;
    auto it = e ? e->live.find(p) : decltype(e->live.find(p)){};
    if (!e || it == e->live.end()) {
        cudaFree(p);
        return;
    }
    e->free_blocks.emplace_back(p, it->second);
    e->live.erase(it);
}

extern "C" const char* apa_last_error(void) { return g_last_error.c_str(); }

// One process per GPU: keep the calling thread (and the packing threads it will spawn, and the page-locked buffers it will
// first touch) on the CPUs of the NUMA node the GPU hangs off, so that neither the DMA reads of the bases nor the packed planes
// cross the socket interconnect. Reads /sys/bus/pci/devices/<bus id>/local_cpulist; CPUs outside the caller's current affinity
// mask (cgroup cpuset) are ignored, and nothing changes when none is left. Returns the number of CPUs bound to, 0 when nothing
// was changed, < 0 on error.
#include <sched.h>
extern "C" int apa_bind_host_thread_to_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return set_err(APA_ERR_NO_DEVICE, "cudaDeviceGetPCIBusId failed");
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return 0;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!got) return 0;
    cpu_set_t cur, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof cur, &cur) != 0) return 0;
    int n = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-15,64-79"
        int lo = 0, hi = 0;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) {
        } else if (sscanf(tok, "%d", &lo) == 1) {
            hi = lo;
        } else {
            continue;
        }
        for (int c = lo; c <= hi && c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &cur)) {
                CPU_SET(c, &want);
                n++;
            }
    }
    if (n == 0 || n == CPU_COUNT(&cur)) return 0;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return 0;
    return n;
}

extern "C" int apa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void apa_engine_destroy(apa_engine* e);
extern "C" int apa_engine_create(int device, apa_engine** out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(APA_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                                              " (this library has no CPU fallback)");
    if (device < 0 || device >= n) return set_err(APA_ERR_NO_DEVICE, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return set_err(APA_ERR_NO_DEVICE, "built for sm_100a; device is older");
    struct EngineGuard {  // every CUDA_TRY below may return early
        apa_engine* e;
        ~EngineGuard() {
            if (e) apa_engine_destroy(e);
        }
    } guard{new apa_engine()};
    apa_engine* eng = guard.e;
    eng->device = device;
    eng->sm_count = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // (numerically lower = higher priority)
    CUDA_TRY(cudaStreamCreateWithPriority(&eng->stream, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&eng->st_pass, cudaStreamNonBlocking, std::min(prio_lo, prio_hi + 1)));
    CUDA_TRY(cudaStreamCreateWithPriority(&eng->st_trace, cudaStreamNonBlocking, std::min(prio_lo, prio_hi + 2)));
    for (auto& ev : eng->ev_ov) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->copy_stream2, cudaStreamNonBlocking));
    for (auto& ev : eng->ev_ring) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
    CUDA_TRY(cudaMalloc(&eng->d_ready, 2048));
    CUDA_TRY(cudaHostAlloc((void**)&eng->h_ready, 256 * sizeof(uint32_t), cudaHostAllocDefault));
    for (uint32_t k = 0; k < 256; k++) eng->h_ready[k] = k;
    for (auto& ev : eng->ev) CUDA_TRY(cudaEventCreate(&ev));
    for (auto& ev : eng->evp) CUDA_TRY(cudaEventCreate(&ev));
    CUDA_TRY(cudaMalloc(&eng->d_queue, 32 * sizeof(unsigned long long)));
    {   // load every kernel now: lazy first-use loading must never happen while a persistent kernel is spinning
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_align_kernel_r64));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_align_kernel_r48));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_align_kernel_r40));
        for (int ph = 0; ph < 3; ph++)
            for (int regs = 48; regs <= 64; regs += 8) CUDA_TRY(cudaFuncGetAttributes(&fa, phase_kernel(ph, regs)));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_block_kernel));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_pack_kernel));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_phase_pass_coop_kernel<4, 1>));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_phase_pass_coop_kernel<8, 1>));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_phase_pass_coop_kernel<4, 3>));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_phase_pass_coop_kernel<8, 3>));
        CUDA_TRY(cudaFuncGetAttributes(&fa, apa_phase_cont_kernel));
    }
    guard.e = nullptr;
    *out = eng;
    return APA_OK;
}

extern "C" void apa_engine_destroy(apa_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (auto& fb : e->free_blocks) cudaFree(fb.first);
    for (auto& lv : e->live) cudaFree(lv.first);
    if (e->d_arena) cudaFree(e->d_arena);
    if (e->d_queue) cudaFree(e->d_queue);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->evp)
        if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->copy_stream2) cudaStreamDestroy(e->copy_stream2);
    if (e->st_pass) cudaStreamDestroy(e->st_pass);
    if (e->st_trace) cudaStreamDestroy(e->st_trace);
    for (auto& ev : e->ev_ov)
        if (ev) cudaEventDestroy(ev);
    if (e->d_phase_flag) cudaFree(e->d_phase_flag);
    for (auto& ev : e->ev_ring)
        if (ev) cudaEventDestroy(ev);
    if (e->d_ready) cudaFree(e->d_ready);
    if (e->h_ready) cudaFreeHost(e->h_ready);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    delete e;
}

// CIGAR text pools are handed out in page-locked memory (D2H at full PCIe rate) and recycled: apa_free() returns a
// pool to a small process-wide cache instead of unpinning it (cudaHostAlloc / cudaFreeHost cost tens of milliseconds).
static std::mutex g_pin_mu;
static std::unordered_map<void*, size_t> g_pin_live;              // handed out
static std::vector<std::pair<void*, size_t>> g_pin_free;          // cached
static void* pinned_pool_get(size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    int best = -1;
    for (size_t k = 0; k < g_pin_free.size(); k++)
        if (g_pin_free[k].second >= bytes && (best < 0 || g_pin_free[k].second < g_pin_free[best].second)) best = (int)k;
    void* p = nullptr;
    size_t cap = 0;
    if (best >= 0) {
        p = g_pin_free[best].first;
        cap = g_pin_free[best].second;
        g_pin_free.erase(g_pin_free.begin() + best);
    } else {
        cap = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    }
    g_pin_live[p] = cap;
    return p;
}
extern "C" void apa_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        auto it = g_pin_live.find(p);
        if (it != g_pin_live.end()) {
            if (g_pin_free.size() < 4) {
                g_pin_free.emplace_back(p, it->second);
            } else {
                cudaFreeHost(p);
            }
            g_pin_live.erase(it);
            return;
        }
    }
    free(p);
}

// Pinned (page-locked) host memory for the end-to-end path.
extern "C" void* apa_pinned_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        set_err(APA_ERR_CUDA, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
extern "C" void apa_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" void apa_batch_free(apa_engine* e, apa_batch* b) {
    if (!b) return;
    if (e) cudaSetDevice(e->device);
    eng_release(e, b->d_a_off);
    eng_release(e, b->d_b_off);
    eng_release(e, b->d_bp_off);
    eng_release(e, b->d_bprof);
    eng_release(e, b->d_aprof);
    eng_release(e, b->d_araw);
    eng_release(e, b->d_braw);
    eng_release(e, b->d_ap_off);
    eng_release(e, b->d_status);
    eng_release(e, b->d_cost);
    eng_release(e, b->d_cig_off);
    eng_release(e, b->d_cig_len);
    eng_release(e, b->d_pair_stats);
    eng_release(e, b->d_order);
    eng_release(e, b->d_pair_chunk);
    eng_release(e, b->d_pool);
    delete b;
}

// ---- host packing tasks -------------------------------------------------------------------------------------------
struct PackTask {
    const uint8_t* seq;
    int64_t len;
    int64_t hw_begin, hw_end;  // half-words of this sequence (padding half-words included: they pack to zero)
    uint32_t* out;             // staging address of half-word 0 of this sequence
    uint32_t chunk;
};

struct Packer {  // packs all tasks with a few host threads; per-chunk completion counters let the caller pipeline
    std::vector<PackTask> tasks;
    std::vector<std::atomic<int>> remaining;  // per chunk
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    std::vector<std::thread> threads;
    explicit Packer(size_t n_chunks) : remaining(n_chunks) {
        for (auto& r : remaining) r.store(0);
    }
    void add(const PackTask& t) {
        tasks.push_back(t);
        remaining[t.chunk].fetch_add(1);
    }
    void start(int n_threads) {
        for (int t = 0; t < n_threads; t++)
            threads.emplace_back([this]() {
                for (;;) {
                    size_t k = next.fetch_add(1);
                    if (k >= tasks.size()) break;
                    const PackTask& tk = tasks[k];
                    if (apa_pack_planes_host(tk.seq, tk.len, tk.hw_begin, tk.hw_end, tk.out)) bad.store(1);
                    remaining[tk.chunk].fetch_sub(1, std::memory_order_release);
                }
            });
    }
    void wait_chunk(size_t c) {
        while (remaining[c].load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }
    void join() {
        for (auto& t : threads) t.join();
        threads.clear();
    }
    ~Packer() { join(); }
};

static int pack_threads() {
    // Host threads that pack bases for one engine. One process per GPU: under torchrun the ranks of a node share the host
    // cores, so each takes its share (LOCAL_WORLD_SIZE) instead of oversubscribing them; APA_PACK_THREADS overrides.
    if (const char* ev = getenv("APA_PACK_THREADS")) return std::max(1, atoi(ev));
    unsigned hc = std::thread::hardware_concurrency();
    unsigned n = std::max(1u, std::min(16u, hc ? hc : 4u));
    if (const char* ev = getenv("LOCAL_WORLD_SIZE")) {
        const unsigned ranks = (unsigned)std::max(1, atoi(ev));
        n = std::max(2u, std::min(n, (hc ? hc : 4u) / ranks));
    }
    cpu_set_t cur;  // a thread bound to part of the box (apa_bind_host_thread_to_device) does not start more workers than it has CPUs
    if (sched_getaffinity(0, sizeof cur, &cur) == 0 && CPU_COUNT(&cur) > 0) n = std::min<unsigned>(n, (unsigned)CPU_COUNT(&cur));
    return (int)n;
}

static int upload_planes(apa_engine* e, apa_batch* b, bool streaming);
static int upload_raw(apa_engine* e, apa_batch* b, bool streaming, cudaStream_t cs);
static int stream_upload(apa_engine* e, apa_batch* b, int mode);
static int upload_packed(apa_engine* e, apa_batch* b, cudaStream_t cs);
static void fill_batch_dev(apa_engine* e, apa_batch* b, BatchDev& bd);

// Frees a half-built batch on every early return of batch_prepare (CUDA_TRY returns from the middle of the function).
struct BatchGuard {
    apa_engine* e;
    apa_batch* b;
    ~BatchGuard() {
        if (b) apa_batch_free(e, b);
    }
    apa_batch* release() {
        apa_batch* t = b;
        b = nullptr;
        return t;
    }
};

// Page-locked host memory can be read by the copy engines directly: such inputs go to HBM as raw bytes and are packed
// there (device-side K0). Pageable inputs are packed by host threads into a pinned staging buffer (4x fewer bytes through
// the driver's pageable-copy path). APA_RAW=0 / 1 overrides the detection (1 is only safe for pinned buffers when streaming).
// Nsight Compute / Systems inject into the process and may serialise kernel launches: a kernel that waits for data the host sends
// after the launch (streamed upload) or for another kernel (overlapped phases) would then sit in its bounded wait. Under a
// profiler the engine therefore uploads first and runs the kernels back to back (the same as APA_STREAM=0 APA_OVERLAP=0).
static bool profiler_attached() {
    static const bool on = getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE");
    return on;
}

static bool host_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

static int batch_prepare(apa_engine* e, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all,
                         const int64_t* b_off, bool defer_data, apa_batch** out, const uint32_t* a_planes = nullptr,
                         const uint32_t* b_planes = nullptr) {
    *out = nullptr;
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    BatchGuard guard{e, new apa_batch()};
    apa_batch* b = guard.b;
    b->n_pairs = n_pairs;
    b->a_off.assign(a_off, a_off + n_pairs + 1);
    b->b_off.assign(b_off, b_off + n_pairs + 1);
    b->bp_off.resize(n_pairs + 1);
    b->ap_off.resize(n_pairs + 1);
    // Plane arrays: ceil(len / 64) * 2 half-words + 2 of padding (extract32 reads one half-word ahead), each pair starting on
    // a 128-byte line (16 half-word entries): no cache line ever holds bases of two pairs, so a line a warp reads after its
    // pair has landed cannot carry stale bytes of a pair that is still on its way (streaming upload).
    auto plane_hw = [](int64_t len) { return (uint64_t)(((len + 63) / 64) * 2 + 2 + 15) & ~(uint64_t)15; };
    uint64_t hw = 0, hwa = 0;
    for (uint64_t p = 0; p < n_pairs; p++) {
        int64_t n = a_off[p + 1] - a_off[p], m = b_off[p + 1] - b_off[p];
        if (n < 0 || m < 0 || n >= (1ll << 31) - 1024 || m >= (1ll << 31) - 1024)
            return set_err(APA_ERR_TOO_LARGE, "sequence length must be < 2^31 (I = i32)");
        b->max_n = std::max<I>(b->max_n, (I)n);
        b->max_m = std::max<I>(b->max_m, (I)m);
        b->bp_off[p] = (int64_t)hw;
        hw += plane_hw(m);
        b->ap_off[p] = (int64_t)hwa;
        hwa += plane_hw(n);
    }
    b->bp_off[n_pairs] = (int64_t)hw;
    b->ap_off[n_pairs] = (int64_t)hwa;
    b->total_hw = hw;
    b->total_hw_a = hwa;
    b->total_a = (uint64_t)(a_off[n_pairs] - a_off[0]);
    b->total_b = (uint64_t)(b_off[n_pairs] - b_off[0]);
    b->h_a = a_all;
    b->h_b = b_all;
    b->h_a_off0 = a_off[0];
    b->h_b_off0 = b_off[0];
    b->h_ap = a_planes;
    b->h_bp = b_planes;
    b->raw = !a_planes && host_pinned(a_all ? a_all + a_off[0] : nullptr) && host_pinned(b_all ? b_all + b_off[0] : nullptr);
    if (const char* ev = getenv("APA_RAW")) b->raw = atoi(ev) != 0;
    // Offsets rebased to the first pair (only lengths matter on the device).
    std::vector<int64_t> ao(b->a_off), bo(b->b_off);
    for (auto& x : ao) x -= a_off[0];
    for (auto& x : bo) x -= b_off[0];
    b->a_off = ao;
    b->b_off = bo;
    // Work order: largest estimated work first. With a streaming upload the batch is cut into chunks of
    // contiguous pairs (about 32 MB of bases each) that become available one after the other; pairs are sorted inside
    // each chunk only.
    std::vector<uint32_t> order(n_pairs);
    std::iota(order.begin(), order.end(), 0u);
    auto by_size = [&](uint32_t x, uint32_t y) {
        return (ao[x + 1] - ao[x]) + (bo[x + 1] - bo[x]) > (ao[y + 1] - ao[y]) + (bo[y + 1] - bo[y]);
    };
    if (n_pairs) {
        const uint64_t total = b->total_a + b->total_b;
        const bool want_stream = defer_data && !(getenv("APA_STREAM") && atoi(getenv("APA_STREAM")) == 0) && !profiler_attached();
        const uint64_t n_chunks = want_stream ? std::min<uint64_t>(200, std::max<uint64_t>(1, total / (32ull << 20))) : 1;
        const uint64_t per = (total + n_chunks - 1) / n_chunks;
        uint64_t acc = 0, start = 0;
        for (uint64_t p = 0; p < n_pairs; p++) {
            acc += (uint64_t)(ao[p + 1] - ao[p]) + (uint64_t)(bo[p + 1] - bo[p]);
            if ((acc >= per && b->chunk_pair_end.size() + 1 < n_chunks) || p + 1 == n_pairs) {
                std::stable_sort(order.begin() + start, order.begin() + p + 1, by_size);
                b->chunk_pair_end.push_back((uint32_t)(p + 1));
                acc = 0;
                start = p + 1;
            }
        }
    }

    cudaStream_t st = e->stream;
    CUDA_TRY(cudaEventRecord(e->ev[0], st));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_a_off, (n_pairs + 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_b_off, (n_pairs + 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_bp_off, (n_pairs + 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_bprof, std::max<uint64_t>(hw, 2) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_aprof, std::max<uint64_t>(hwa, 2) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_ap_off, (n_pairs + 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_status, std::max<uint64_t>(n_pairs, 1) * 4));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_cost, std::max<uint64_t>(n_pairs, 1) * 4));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_cig_off, std::max<uint64_t>(n_pairs, 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_cig_len, std::max<uint64_t>(n_pairs, 1) * 8));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_pair_stats, std::max<uint64_t>(n_pairs, 1) * 64));
    CUDA_TRY(eng_alloc(e, (void**)&b->d_order, std::max<uint64_t>(n_pairs, 1) * 8));  // [0,n): work order, [n,2n): retry list
    const bool streamed = defer_data && b->chunk_pair_end.size() >= 2;
    if (b->raw) {
        CUDA_TRY(eng_alloc(e, (void**)&b->d_araw, b->total_a + 64));
        CUDA_TRY(eng_alloc(e, (void**)&b->d_braw, b->total_b + 64));
    }
    if (streamed) {
        std::vector<uint16_t> pc(n_pairs);
        uint32_t q0 = 0;
        for (size_t c = 0; c < b->chunk_pair_end.size(); c++) {
            for (uint32_t q = q0; q < b->chunk_pair_end[c]; q++) pc[q] = (uint16_t)c;
            q0 = b->chunk_pair_end[c];
        }
        CUDA_TRY(eng_alloc(e, (void**)&b->d_pair_chunk, n_pairs * 2));
        CUDA_TRY(cudaMemcpyAsync(b->d_pair_chunk, pc.data(), n_pairs * 2, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));  // pc goes out of scope
    }
    if ((!b->raw || streamed) && !a_planes) {
        // pinned staging for the packed planes: [aprof | bprof]
        const size_t stage_words = (size_t)(hwa + hw) * 2 + 16;
        if (e->h_stage_cap < stage_words) {
            if (e->h_stage) cudaFreeHost(e->h_stage);
            e->h_stage = nullptr;
            e->h_stage_cap = 0;
            CUDA_TRY(cudaHostAlloc((void**)&e->h_stage, stage_words * 4, cudaHostAllocDefault));
            e->h_stage_cap = stage_words;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_a_off, ao.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_b_off, bo.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_bp_off, b->bp_off.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_ap_off, b->ap_off.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_pairs) CUDA_TRY(cudaMemcpyAsync(b->d_order, order.data(), n_pairs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->stats.h2d_bytes = (b->raw ? b->total_a + b->total_b : (hwa + hw) * 8) + 4 * (n_pairs + 1) * 8 + n_pairs * 4;  // streamed: recounted
    if (!defer_data) {
        int rc;
        if (b->raw) {  // raw bases by DMA, packed by apa_pack_kernel; the raw copy is not kept
            rc = upload_raw(e, b, /*streaming=*/false, st);
            if (rc == APA_OK && n_pairs) {
                BatchDev bd{};
                fill_batch_dev(e, b, bd);
                int* d_bad = (int*)(e->d_ready + 300);
                CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4, st));
                const unsigned gx = (unsigned)std::min<uint64_t>(n_pairs, 4096);
                const unsigned gy = n_pairs >= 4ull * e->sm_count ? 1u : (unsigned)((4ull * e->sm_count + n_pairs - 1) / n_pairs);
                apa_pack_kernel<<<dim3(gx, gy), 256, 0, st>>>(bd, d_bad);
                CUDA_TRY(cudaGetLastError());
                int h_bad = 0;
                CUDA_TRY(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (h_bad) rc = set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT (the reference panics here: pa-bitpacking/src/profile.rs:113)");
            }
            eng_release(e, b->d_araw);
            eng_release(e, b->d_braw);
            b->d_araw = b->d_braw = nullptr;
            b->raw = false;
        } else {
            rc = upload_planes(e, b, /*streaming=*/false);
        }
        if (rc != APA_OK) return rc;
    }
    CUDA_TRY(cudaEventRecord(e->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
    b->stats.h2d_ms = ms;
    *out = guard.release();
    return APA_OK;
}

// Plain (not streamed) upload of caller-packed planes (apa_align_batch_packed): the arrays are already in the device layout.
static int upload_packed(apa_engine* e, apa_batch* b, cudaStream_t cs) {
    (void)e;
    if (b->total_hw_a) CUDA_TRY(cudaMemcpyAsync(b->d_aprof, b->h_ap, (size_t)b->total_hw_a * 8, cudaMemcpyHostToDevice, cs));
    if (b->total_hw) CUDA_TRY(cudaMemcpyAsync(b->d_bprof, b->h_bp, (size_t)b->total_hw * 8, cudaMemcpyHostToDevice, cs));
    return APA_OK;
}

// Plain (not streamed) upload of raw bases from page-locked host memory: asynchronous DMA on stream cs.
static int upload_raw(apa_engine* e, apa_batch* b, bool /*streaming*/, cudaStream_t cs) {
    (void)e;
    if (b->total_a) CUDA_TRY(cudaMemcpyAsync(b->d_araw, b->h_a + b->h_a_off0, b->total_a, cudaMemcpyHostToDevice, cs));
    if (b->total_b) CUDA_TRY(cudaMemcpyAsync(b->d_braw, b->h_b + b->h_b_off0, b->total_b, cudaMemcpyHostToDevice, cs));
    return APA_OK;
}

// Pack the pairs [p0, p1) on the calling thread into the pinned staging buffer ([aprof | bprof], device layout).
static bool pack_pairs_host(apa_engine* e, apa_batch* b, uint32_t p0, uint32_t p1) {
    uint32_t* stage_a = e->h_stage;
    uint32_t* stage_b = e->h_stage + (size_t)b->total_hw_a * 2;
    bool bad = false;
    for (uint32_t p = p0; p < p1; p++) {
        const int64_t n = b->a_off[p + 1] - b->a_off[p], m = b->b_off[p + 1] - b->b_off[p];
        const int64_t nhw_a = b->ap_off[p + 1] - b->ap_off[p], nhw_b = b->bp_off[p + 1] - b->bp_off[p];
        bad |= apa_pack_planes_host(b->h_a + b->h_a_off0 + b->a_off[p], n, 0, nhw_a, stage_a + (size_t)b->ap_off[p] * 2) != 0;
        bad |= apa_pack_planes_host(b->h_b + b->h_b_off0 + b->b_off[p], m, 0, nhw_b, stage_b + (size_t)b->bp_off[p] * 2) != 0;
    }
    return bad;
}

// Plain (not streamed) upload of host-packed planes: host threads pack the bases into the pinned staging buffer (long
// sequences in segments), one copy per chunk on the engine's stream.
static int upload_planes(apa_engine* e, apa_batch* b, bool /*streaming*/) {
    const size_t n_chunks = b->chunk_pair_end.size();
    if (n_chunks == 0) return APA_OK;
    auto tp0 = std::chrono::steady_clock::now();
    uint32_t* stage_a = e->h_stage;
    uint32_t* stage_b = e->h_stage + (size_t)b->total_hw_a * 2;
    Packer pk(n_chunks);
    const int64_t SEG = 1 << 15;  // half-words per task (1 Mi bases)
    uint32_t p0 = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        const uint32_t p1 = b->chunk_pair_end[c];
        for (uint32_t p = p0; p < p1; p++) {
            const int64_t n = b->a_off[p + 1] - b->a_off[p], m = b->b_off[p + 1] - b->b_off[p];
            const int64_t nhw_a = b->ap_off[p + 1] - b->ap_off[p], nhw_b = b->bp_off[p + 1] - b->bp_off[p];
            const uint8_t* sa = b->h_a + b->h_a_off0 + b->a_off[p];
            const uint8_t* sb = b->h_b + b->h_b_off0 + b->b_off[p];
            for (int64_t h0 = 0; h0 < nhw_a; h0 += SEG)
                pk.add(PackTask{sa, n, h0, std::min(nhw_a, h0 + SEG), stage_a + (size_t)b->ap_off[p] * 2, (uint32_t)c});
            for (int64_t h0 = 0; h0 < nhw_b; h0 += SEG)
                pk.add(PackTask{sb, m, h0, std::min(nhw_b, h0 + SEG), stage_b + (size_t)b->bp_off[p] * 2, (uint32_t)c});
        }
        p0 = p1;
    }
    // small batches (the single-pair drop-in calls) are packed by the calling thread: spawning workers costs more than the work
    const bool inline_pack = b->total_a + b->total_b < (2u << 20);
    if (inline_pack) {
        for (const PackTask& tk : pk.tasks) {
            if (apa_pack_planes_host(tk.seq, tk.len, tk.hw_begin, tk.hw_end, tk.out)) pk.bad.store(1);
            pk.remaining[tk.chunk].fetch_sub(1, std::memory_order_release);
        }
    } else {
        pk.start(pack_threads());
    }
    cudaStream_t cs = e->stream;
    p0 = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        const uint32_t p1 = b->chunk_pair_end[c];
        pk.wait_chunk(c);
        const int64_t a0 = b->ap_off[p0], a1 = b->ap_off[p1], b0 = b->bp_off[p0], b1 = b->bp_off[p1];
        if (a1 > a0) CUDA_TRY(cudaMemcpyAsync(b->d_aprof + a0, stage_a + (size_t)a0 * 2, (size_t)(a1 - a0) * 8, cudaMemcpyHostToDevice, cs));
        if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(b->d_bprof + b0, stage_b + (size_t)b0 * 2, (size_t)(b1 - b0) * 8, cudaMemcpyHostToDevice, cs));
        p0 = p1;
    }
    pk.join();
    CUDA_TRY(cudaStreamSynchronize(cs));
    b->pack_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
    if (pk.bad.load()) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT (the reference panics here: pa-bitpacking/src/profile.rs:113)");
    return APA_OK;
}

// Streamed upload under the running build kernel by two producers that both take the next chunk in the kernel's order: the
// copy engines send RAW bases (pure DMA from page-locked memory, two chunks in flight; the kernel packs them: device-side K0),
// the host threads pack a chunk into 2-bit planes (4x fewer bytes over PCIe) and send those. Who takes how many follows from
// the box: 1 GPU with 16 cores to itself packs ~2/3 of the chunks, 8 ranks sharing the cores send most of them raw. Every
// chunk is followed, on the stream that carried it, by its state word (1 planes / 2 raw) that the kernel polls. (Measured
// dead end: packing threads working from the BACK of the batch finish the upload as early, but the kernel - which wants a
// pair for every resident warp at once - then only sees the front grow at the raw DMA rate: 115 ms per step against 106.)
// Pageable inputs cannot be read by DMA: the host threads pack all chunks. mode: 0 both, 1 raw only, 2 packed only.
static int stream_upload(apa_engine* e, apa_batch* b, int mode) {
    const uint32_t n_chunks = (uint32_t)b->chunk_pair_end.size();
    auto tp0 = std::chrono::steady_clock::now();
    const bool planes_in = mode == 3;  // caller-packed planes: DMA only, nothing to pack on either side
    const bool dma = mode != 2, pack = mode != 1 && !planes_in;
    std::mutex mu;
    uint32_t front = 0;
    const uint32_t back = n_chunks;  // unclaimed chunks: [front, back)
    // The packing threads share ONE chunk at a time, pair by pair, so that a chunk is ready after (pack time of a chunk) /
    // (threads) - the same order of time the copy engines need for one - and the two ends meet without one side idling.
    struct PackChunk {
        uint32_t c, p0, p1, next;
        std::atomic<uint32_t> done{0};
    };
    std::vector<std::unique_ptr<PackChunk>> pack_chunks;  // stable addresses: stragglers finish a chunk after the next one opened
    PackChunk* cur = nullptr;
    auto chunk_pairs = [&](uint32_t c, uint32_t& p0, uint32_t& p1) {
        p0 = c ? b->chunk_pair_end[c - 1] : 0u;
        p1 = b->chunk_pair_end[c];
    };
    auto claim_pair = [&](PackChunk*& pc, uint32_t& p) -> bool {  // next pair of the shared chunk, opening a new chunk when it is used up
        std::lock_guard<std::mutex> lk(mu);
        while (!cur || cur->next >= cur->p1) {
            if (front >= back) return false;
            const uint32_t c = front++;  // like the copy engines: the next chunk in the kernel's order
            pack_chunks.emplace_back(new PackChunk());
            cur = pack_chunks.back().get();
            cur->c = c;
            chunk_pairs(c, cur->p0, cur->p1);
            cur->next = cur->p0;
            if (cur->p0 == cur->p1) cur->done.store(0);  // (chunks are never empty: batch_prepare closes a chunk on a pair)
        }
        pc = cur;
        p = cur->next++;
        return true;
    };
    auto claim_front = [&](uint32_t& c) -> bool {
        std::lock_guard<std::mutex> lk(mu);
        if (front >= back) return false;
        c = front++;
        return true;
    };
    std::atomic<int> bad{0}, cuda_fail{0};
    std::atomic<uint64_t> bytes{0};
    uint32_t* stage_a = e->h_stage;
    uint32_t* stage_b = e->h_stage + (size_t)b->total_hw_a * 2;
    uint32_t n_raw = 0;
    cudaError_t ce_main = cudaSuccess;
    auto send_raw = [&](uint32_t c) {
        if (n_raw >= 2) ce_main = cudaEventSynchronize(e->ev_ring[(n_raw - 2) & 3]);  // two raw chunks in flight
        uint32_t p0, p1;
        chunk_pairs(c, p0, p1);
        if (planes_in) {
            const int64_t a0 = b->ap_off[p0], a1 = b->ap_off[p1], b0 = b->bp_off[p0], b1 = b->bp_off[p1];
            if (ce_main == cudaSuccess && a1 > a0)
                ce_main = cudaMemcpyAsync(b->d_aprof + a0, b->h_ap + (size_t)a0 * 2, (size_t)(a1 - a0) * 8, cudaMemcpyHostToDevice, e->copy_stream);
            if (ce_main == cudaSuccess && b1 > b0)
                ce_main = cudaMemcpyAsync(b->d_bprof + b0, b->h_bp + (size_t)b0 * 2, (size_t)(b1 - b0) * 8, cudaMemcpyHostToDevice, e->copy_stream);
            if (ce_main == cudaSuccess) ce_main = cudaMemcpyAsync(e->d_ready + c, &e->h_ready[1], 4, cudaMemcpyHostToDevice, e->copy_stream);
            if (ce_main == cudaSuccess) ce_main = cudaEventRecord(e->ev_ring[n_raw & 3], e->copy_stream);
            bytes.fetch_add((uint64_t)(a1 - a0 + b1 - b0) * 8);
            n_raw++;
            return;
        }
        const int64_t a0 = b->a_off[p0], a1 = b->a_off[p1], b0 = b->b_off[p0], b1 = b->b_off[p1];
        if (ce_main == cudaSuccess && a1 > a0)
            ce_main = cudaMemcpyAsync(b->d_araw + a0, b->h_a + b->h_a_off0 + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, e->copy_stream);
        if (ce_main == cudaSuccess && b1 > b0)
            ce_main = cudaMemcpyAsync(b->d_braw + b0, b->h_b + b->h_b_off0 + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, e->copy_stream);
        if (ce_main == cudaSuccess) ce_main = cudaMemcpyAsync(e->d_ready + c, &e->h_ready[2], 4, cudaMemcpyHostToDevice, e->copy_stream);
        if (ce_main == cudaSuccess) ce_main = cudaEventRecord(e->ev_ring[n_raw & 3], e->copy_stream);
        bytes.fetch_add((uint64_t)(a1 - a0 + b1 - b0));
        n_raw++;
    };
    uint32_t c;
    if (dma && claim_front(c)) send_raw(c);  // the first chunk is on its way before the packing threads exist
    std::vector<std::thread> workers;
    if (pack)
        for (int t = 0; t < pack_threads(); t++)
            workers.emplace_back([&]() {
                if (cudaSetDevice(e->device) != cudaSuccess) {
                    cuda_fail.store(1);
                    return;
                }
                PackChunk* pc;
                uint32_t p;
                while (claim_pair(pc, p)) {
                    if (pack_pairs_host(e, b, p, p + 1)) bad.store(1);  // the chunk is still sent: the kernel must not wait forever
                    if (pc->done.fetch_add(1, std::memory_order_acq_rel) + 1 != pc->p1 - pc->p0) continue;
                    // the last pair of the chunk: send its planes and its state word
                    const int64_t a0 = b->ap_off[pc->p0], a1 = b->ap_off[pc->p1], b0 = b->bp_off[pc->p0], b1 = b->bp_off[pc->p1];
                    cudaError_t ce = cudaSuccess;
                    if (a1 > a0) ce = cudaMemcpyAsync(b->d_aprof + a0, stage_a + (size_t)a0 * 2, (size_t)(a1 - a0) * 8, cudaMemcpyHostToDevice, e->copy_stream2);
                    if (ce == cudaSuccess && b1 > b0)
                        ce = cudaMemcpyAsync(b->d_bprof + b0, stage_b + (size_t)b0 * 2, (size_t)(b1 - b0) * 8, cudaMemcpyHostToDevice, e->copy_stream2);
                    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->d_ready + pc->c, &e->h_ready[1], 4, cudaMemcpyHostToDevice, e->copy_stream2);
                    if (ce != cudaSuccess) cuda_fail.store(1);
                    bytes.fetch_add((uint64_t)(a1 - a0 + b1 - b0) * 8);
                }
            });
    if (dma)
        while (ce_main == cudaSuccess && claim_front(c)) send_raw(c);
    const double ms_dma_issued = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
    for (auto& t : workers) t.join();
    if (getenv("APA_DEBUG_TIMING")) {
        const double ms_packed = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
        cudaStreamSynchronize(e->copy_stream);
        const double ms_raw_landed = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
        cudaStreamSynchronize(e->copy_stream2);
        const double ms_all_landed = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
        fprintf(stderr, "[stream_upload] %u chunks: %u raw (last issued at %.1f ms, landed by %.1f ms), %zu packed by %d threads (done at %.1f ms, landed by %.1f ms)\n",
                n_chunks, n_raw, ms_dma_issued, ms_raw_landed, pack_chunks.size(), pack ? pack_threads() : 0, ms_packed, ms_all_landed);
    }
    b->chunks_raw = planes_in ? 0u : n_raw;
    b->stats.h2d_bytes = bytes.load() + 4 * (b->n_pairs + 1) * 8 + b->n_pairs * 6;
    b->pack_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count();
    if (ce_main != cudaSuccess || cuda_fail.load())
        return set_err(APA_ERR_CUDA, std::string("streamed upload: ") + cudaGetErrorString(ce_main != cudaSuccess ? ce_main : cudaGetLastError()));
    if (bad.load()) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT (the reference panics here: pa-bitpacking/src/profile.rs:113)");
    return APA_OK;
}

extern "C" int apa_batch_upload(apa_engine* e, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all,
                                const int64_t* b_off, apa_batch** out) {
    return batch_prepare(e, n_pairs, a_all, a_off, b_all, b_off, false, out);
}


static uint32_t estimate_arena(const apa_batch* b, int preset, int trace, const RunParams* gp) {
    // meta + V columns of one pass + traceback scratch + CIGAR elements. Deliberately modest: pairs that do
    // not fit are re-run with a larger arena (ST_OVERFLOW).
    const uint64_t bw = gp ? (uint64_t)gp->block_width : BLOCK_W;
    const bool gcsh = gp ? (gp->domain == DOM_ASTAR && gp->heuristic == APA_HEURISTIC_GCSH) : preset == APA_PRESET_FULL;
    uint64_t nblk = (uint64_t)(b->max_n + bw - 1) / bw + 1;
    uint64_t meta = nblk * sizeof(BlkMeta);
    uint64_t band_rows = !gcsh ? std::min<uint64_t>((uint64_t)b->max_m + 64, std::max<uint64_t>(2048, (uint64_t)b->max_n / 8))
                               : std::min<uint64_t>((uint64_t)b->max_m + 64, 2048);
    if (gp && gp->domain == APA_DOMAIN_FULL) band_rows = (uint64_t)b->max_m + 64;
    uint64_t vcols = nblk * (band_rows / 32 * 12 + 16);
    const bool incremental = gp ? gp->incremental != 0 : preset == APA_PRESET_FULL;
    if (incremental) vcols = 2 * vcols + (uint64_t)b->max_n + 64;  // two block stores (this pass and the previous one) + the h row
    uint64_t tr = trace ? (DT_CACHE_ELEMS * 8 + 256 * (band_rows / 32) * 8 / 4 + (uint64_t)(b->max_n + b->max_m) / 4 * 4 + 65536) : 0;
    uint64_t heur = gcsh ? 24ull * (uint64_t)b->max_n + 65536 : 0;  // k-mer table, matches, contours
    uint64_t s = ARENA_HEADER + meta + vcols + tr + heur + 16384;
    s = (s + 1023) & ~1023ull;
    return (uint32_t)std::min<uint64_t>(s, 0xF0000000ull);
}

static int validate_params(const apa_params* q, RunParams* out) {
    if (!q) return set_err(APA_ERR_BAD_INPUT, "null params");
    auto bad = [](const char* what) { return set_err(APA_ERR_BAD_INPUT, std::string("unsupported AstarPa2Params: ") + what); };
    if (q->domain < APA_DOMAIN_FULL || q->domain > APA_DOMAIN_ASTAR) return bad("domain");
    if (q->domain == APA_DOMAIN_ASTAR) {
        if (q->heuristic < APA_HEURISTIC_NONE || q->heuristic > APA_HEURISTIC_GCSH) return bad("heuristic (NoCost, GapCost and GCSH are built)");
        if (q->heuristic == APA_HEURISTIC_GCSH) {
            if (q->r != 1) return bad("r (exact matches only)");
            if (q->k < 4 || q->k > 16) return bad("k must be 4..16");
            if (q->p < 0 || q->p > 15) return bad("p must be 0..15");
        }
    }
    if (q->doubling < APA_DOUBLING_NONE || q->doubling > APA_DOUBLING_LINEAR) return bad("doubling");
    if (q->doubling == APA_DOUBLING_NONE && q->domain != APA_DOMAIN_FULL) return bad("DoublingType::None requires Domain::Full (lib.rs:128)");
    if (q->doubling != APA_DOUBLING_NONE && (q->doubling_start < APA_START_ZERO || q->doubling_start > APA_START_H0)) return bad("doubling_start");
    if (q->doubling == APA_DOUBLING_BAND && !(q->factor > 1.0f)) return bad("factor must be > 1");
    if (q->doubling == APA_DOUBLING_LINEAR && q->delta < 1) return bad("delta must be >= 1");
    if (q->block_width < 1 || q->block_width > BLOCK_W) return bad("block_width must be 1..256");
    if (!q->sparse) return bad("sparse = false");
    if (q->max_g < 1 || q->max_g > DT_MAX_G) return bad("max_g must be 1..40");
    if (q->fr_drop < 0) return bad("fr_drop");
    out->domain = q->domain;
    out->heuristic = q->domain == APA_DOMAIN_ASTAR ? q->heuristic : APA_HEURISTIC_NONE;
    out->k = q->k;
    out->p = q->p;
    out->doubling = q->doubling;
    out->start = q->doubling_start;
    out->factor = q->factor;
    out->delta = q->delta;
    out->block_width = q->block_width;
    out->dt_trace = q->dt_trace != 0;
    out->max_g = q->max_g;
    out->fr_drop = q->fr_drop;
    out->sparse_h = q->sparse_h != 0;
    out->incremental = q->incremental_doubling != 0;
    out->prune = q->prune != 0;
    return APA_OK;
}

extern "C" int apa_params_preset(int preset, apa_params* out) {
    if (!out || (preset != APA_PRESET_SIMPLE && preset != APA_PRESET_FULL)) return set_err(APA_ERR_BAD_INPUT, "unknown preset");
    apa_params q{};
    q.domain = APA_DOMAIN_ASTAR;
    q.heuristic = preset == APA_PRESET_FULL ? APA_HEURISTIC_GCSH : APA_HEURISTIC_GAP;
    q.k = 12;  // params.rs:103-105
    q.r = 1;
    q.p = 14;
    q.doubling = APA_DOUBLING_BAND;
    q.doubling_start = APA_START_H0;
    q.factor = 2.0f;
    q.delta = 1;
    q.block_width = 256;
    q.sparse = 1;
    q.incremental_doubling = preset == APA_PRESET_FULL;
    q.dt_trace = 1;
    q.max_g = 40;
    q.fr_drop = 10;
    q.sparse_h = 1;
    q.prune = preset == APA_PRESET_FULL;
    *out = q;
    return APA_OK;
}

static void fill_batch_dev(apa_engine* e, apa_batch* b, BatchDev& bd) {
    bd.n_pairs = b->n_pairs;
    bd.a_off = b->d_a_off;
    bd.b_off = b->d_b_off;
    bd.bp_off = b->d_bp_off;
    bd.bprof = b->d_bprof;
    bd.ap_off = b->d_ap_off;
    bd.aprof = b->d_aprof;
    bd.status = b->d_status;
    bd.cost = b->d_cost;
    bd.cig_off = b->d_cig_off;
    bd.cig_len = b->d_cig_len;
    bd.pair_stats = b->d_pair_stats;
    bd.order = b->d_order;
    bd.n_order = (uint32_t)b->n_pairs;
    bd.queue = e->d_queue;
    bd.pool = b->d_pool;
    bd.pool_cursor = e->d_queue + 1;
    bd.pool_cap = b->pool_cap;
    bd.stats = e->d_queue + 2;
    bd.raw_a = b->raw ? b->d_araw : nullptr;
    bd.raw_b = b->raw ? b->d_braw : nullptr;
    bd.dbg = b->d_dbg;
    bd.dbg_cap = b->dbg_cap;
    bd.dbg_n = b->d_dbg_n;
}

static int batch_run(apa_engine* e, apa_batch* b, int preset, int trace, bool stream_data, const RunParams* gp = nullptr) {
    if (!e || !b) return set_err(APA_ERR_NO_DEVICE, "null engine/batch");
    if (!gp && preset != APA_PRESET_SIMPLE && preset != APA_PRESET_FULL) return set_err(APA_ERR_BAD_INPUT, "unknown preset");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    b->trace = trace;
    b->stats.kernel_launches = 0;
    b->stats.retries = 0;
    b->stats.phase_ms[0] = b->stats.phase_ms[1] = b->stats.phase_ms[2] = 0;
    if (b->n_pairs == 0) {
        b->ran = true;
        return APA_OK;
    }
    // CIGAR text pool: text length <= |a| + |b| per pair (+ NUL).
    uint64_t pool_need = trace ? (b->total_a + b->total_b + b->n_pairs + 64) : 16;
    if (b->pool_cap < pool_need) {
        eng_release(e, b->d_pool);
        b->d_pool = nullptr;
        CUDA_TRY(eng_alloc(e, (void**)&b->d_pool, pool_need));
        b->pool_cap = pool_need;
    }
    BatchDev bd{};
    fill_batch_dev(e, b, bd);
    bd.preset = preset;
    bd.trace = trace;

    CUDA_TRY(cudaEventRecord(e->ev[2], st));
    CUDA_TRY(cudaMemsetAsync(e->d_queue, 0, 32 * sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(b->d_status, 0, b->n_pairs * 4, st));
    uint32_t arena_size = estimate_arena(b, preset, trace, gp);
    // Few pairs (the regime of the cooperative pass kernel): HBM is plentiful per pair, and a pair that overflows its arena
    // is re-run from scratch, so start 8x above the modest estimate (BASELINE configs[3], n = 1 M at 15 %: ~240 MB per pair).
    if (!gp && b->n_pairs <= (uint64_t)e->sm_count * 8) arena_size = (uint32_t)std::min<uint64_t>((uint64_t)arena_size * 8, 0xF0000000ull);
    if (const char* ev = getenv("APA_ARENA_BYTES")) arena_size = (uint32_t)std::max<long long>(65536, atoll(ev));  // tests: force the overflow/retry path
    std::vector<uint32_t> pending;  // empty = all pairs in the uploaded order
    b->h_status.assign(b->n_pairs, 0);
    for (int attempt = 0; attempt < 8; attempt++) {
        // Slots (resident warps). Measured on B200 (profiles/README.md): throughput grows with resident warps up to
        // ~40 per SM even though the last wave is then only partly full; equalising the waves was slower.
        uint64_t n_work = attempt == 0 ? b->n_pairs : pending.size();
        const uint64_t max_slots = (uint64_t)e->sm_count * (gp ? 6 : 10) * WARPS_PER_CTA;  // the general kernel runs 6 CTAs per SM
        uint64_t slots = ((n_work + WARPS_PER_CTA - 1) / WARPS_PER_CTA) * WARPS_PER_CTA;
        slots = std::min<uint64_t>(slots, max_slots);
        if (const char* ev = getenv("APA_SLOTS")) slots = std::max<uint64_t>(WARPS_PER_CTA, ((uint64_t)atoll(ev) / WARPS_PER_CTA) * WARPS_PER_CTA);
        // Phase-split path: every pair of the batch keeps its own arena across the three phase kernels. Used when those
        // arenas fit in HBM (10 000 pairs x 3 MB = 30 GB of 180 GB); otherwise the fused kernel with per-warp arenas runs.
        // cudaMemGetInfo costs 1-30 ms with tens of GB allocated, so it is only asked when the arena has to grow.
        uint64_t budget = 0;  // bytes the arena may take; queried lazily
        auto query_budget = [&]() -> cudaError_t {
            if (budget) return cudaSuccess;
            size_t free_b = 0, total_b = 0;
            cudaError_t ce = cudaMemGetInfo(&free_b, &total_b);
            budget = (uint64_t)free_b + e->arena_total;
            budget = budget > (4ull << 30) ? budget - (2ull << 30) : std::max<uint64_t>(budget / 2, 1);
            if (const char* ev = getenv("APA_BUDGET_BYTES")) budget = std::max<long long>(1 << 20, atoll(ev));  // tests: force the wave path
            return ce;
        };
        uint64_t have = e->arena_total;  // an arena this large is already allocated
        if (const char* ev = getenv("APA_BUDGET_BYTES")) have = std::min<uint64_t>(have, (uint64_t)std::max<long long>(1 << 20, atoll(ev)));
        bool split = n_work * (uint64_t)arena_size <= have;
        if (!split) {
            CUDA_TRY(query_budget());
            split = n_work * (uint64_t)arena_size <= budget;
        }
        const bool split_allowed = !gp && !(getenv("APA_SPLIT") && atoi(getenv("APA_SPLIT")) == 0);  // the general kernel is fused
        split = split && split_allowed;
        // Memory-limited work lists (long, divergent pairs: BASELINE configs[3]): when only a few hundred per-pair arenas fit at
        // once, the phase-split path runs in waves of that many pairs, each pair on a whole CTA (apa_coop.cuh), instead of
        // the fused kernel with one warp per pair on a mostly empty GPU.
        uint64_t wave_n = n_work;
        if (!split && split_allowed) {
            CUDA_TRY(query_budget());
            const uint64_t fit = budget / arena_size;
            if (fit >= 1 && fit <= (uint64_t)e->sm_count * 8) {
                split = true;
                wave_n = std::min<uint64_t>(fit, n_work);
            }
        }
        // Warps per pair in the pass kernel: a whole CTA when the wave has fewer pairs than the GPU has CTA slots.
        int coop_w = 1;
        if (split) {
            if (wave_n <= (uint64_t)e->sm_count * 4)
                coop_w = 8;
            else if (wave_n <= (uint64_t)e->sm_count * 8)
                coop_w = 4;
            if (const char* ev = getenv("APA_COOP")) coop_w = atoi(ev) >= 8 ? 8 : (atoi(ev) >= 4 ? 4 : 1);
        }
        if (!split && slots * (uint64_t)arena_size > e->arena_total) {
            CUDA_TRY(query_budget());
            while (slots > WARPS_PER_CTA && slots * (uint64_t)arena_size > budget) slots = (slots / 2 / WARPS_PER_CTA) * WARPS_PER_CTA;
            if (slots * (uint64_t)arena_size > budget) return set_err(APA_ERR_TOO_LARGE, "scratch arena exceeds device memory");
        }
        const uint64_t n_arenas = split ? wave_n : slots;
        size_t need = (size_t)(split ? n_arenas : slots) * arena_size;
        if (e->arena_total < need) {
            if (e->d_arena) CUDA_TRY(cudaFree(e->d_arena));
            e->d_arena = nullptr;
            e->arena_total = 0;
            CUDA_TRY(cudaMalloc(&e->d_arena, need));
            e->arena_total = need;
        }
        bd.arena = e->d_arena;
        bd.arena_size = arena_size;
        bd.n_order = (uint32_t)n_work;
        if (attempt > 0) {
            CUDA_TRY(cudaMemcpyAsync(b->d_order + b->n_pairs, pending.data(), pending.size() * 4, cudaMemcpyHostToDevice, st));
            bd.order = b->d_order + b->n_pairs;
            CUDA_TRY(cudaMemsetAsync(e->d_queue, 0, sizeof(unsigned long long), st));
        }
        // The bases of a deferred batch (apa_align_batch) reach HBM here. A batch of several upload chunks streams under the
        // running kernel (stream_upload: raw bases by DMA from the front, host-packed planes from the back); one chunk, waves,
        // the general kernel and retries take the plain upload first: nothing to overlap with.
        const bool first_upload = stream_data && attempt == 0 && !b->chunk_pair_end.empty();
        const bool streaming = first_upload && !gp && b->chunk_pair_end.size() >= 2 && !(split && wave_n < n_work) && b->d_pair_chunk;
        int stream_mode = b->raw ? 0 : 2;  // page-locked inputs: both producers; pageable: host-packed only
        if (const char* ev = getenv("APA_RAW")) stream_mode = atoi(ev) != 0 ? (b->raw ? 1 : 2) : 2;
        if (b->h_ap) stream_mode = 3;  // caller-packed planes
        if (first_upload && !streaming) {
            int rc = b->h_ap ? upload_packed(e, b, st) : (b->raw ? upload_raw(e, b, /*streaming=*/false, st) : upload_planes(e, b, /*streaming=*/false));
            if (rc != APA_OK) return rc;
        }
        bd.chunk_state = streaming ? e->d_ready : nullptr;
        bd.pair_chunk = b->d_pair_chunk;
        if (attempt > 0) bd.raw_a = bd.raw_b = nullptr;  // a retried pair was packed when it was first opened: its planes are in HBM
        if (streaming) {
            CUDA_TRY(cudaMemsetAsync(e->d_ready, 0, 256 * 4, st));
            CUDA_TRY(cudaStreamSynchronize(st));  // offsets, order, zeroed queue and chunk states are in place before anything overlaps
        }
        const bool host_streaming = streaming;  // the host feeds the chunks while the build kernel already runs
        if (attempt == 0) {
            b->stats.upload_mode = !first_upload ? 0u : (b->h_ap ? (streaming ? 7u : 6u) : (streaming ? (stream_mode == 0 ? 5u : (stream_mode == 1 ? 4u : 2u)) : (b->raw ? 3u : 1u)));
            b->stats.upload_chunks = first_upload ? (uint32_t)b->chunk_pair_end.size() : 0u;
            b->stats.pass_warps_per_pair = split ? (uint32_t)coop_w : 0u;
            b->stats.waves = split ? (uint32_t)((n_work + wave_n - 1) / wave_n) : 0u;
        }
        // register variant: 64 registers when the slots fit 8 CTAs per SM
        int regs = slots <= (uint64_t)e->sm_count * 8 * WARPS_PER_CTA ? 64 : (slots <= (uint64_t)e->sm_count * 10 * WARPS_PER_CTA ? 48 : 40);
        if (const char* ev = getenv("APA_REGS")) regs = atoi(ev);
        bd.q0 = 0;
        const unsigned grid = (unsigned)(slots / WARPS_PER_CTA);
        // pass + trace kernels of the phase-split path, with the per-phase events (stats.phase_ms)
        // One warp per pair, the whole batch in one wave: the three phase kernels are launched together (see BatchDev::phase_flag).
        // APA_OVERLAP=0 runs them back to back (per-kernel timings: bench.py measures its roofline numbers that way).
        const bool overlap = split && coop_w == 1 && wave_n == n_work && attempt == 0 && trace &&
                             !(getenv("APA_OVERLAP") && atoi(getenv("APA_OVERLAP")) == 0) && !profiler_attached();
        bd.phase_flag = nullptr;
        if (overlap) {
            if (e->phase_flag_cap < n_work) {
                if (e->d_phase_flag) CUDA_TRY(cudaFree(e->d_phase_flag));
                e->d_phase_flag = nullptr;
                e->phase_flag_cap = 0;
                CUDA_TRY(cudaMalloc(&e->d_phase_flag, n_work + 256));
                e->phase_flag_cap = n_work;
            }
            CUDA_TRY(cudaMemsetAsync(e->d_phase_flag, 0, n_work, st));
            bd.phase_flag = e->d_phase_flag;
        }
        b->stats.overlapped = overlap ? 1u : 0u;
        auto launch_pass_trace = [&]() -> cudaError_t {
            const unsigned wave_pairs = (unsigned)(bd.n_order - bd.q0);
            if (overlap) {  // pass and trace kernels on their own streams, behind the setup of this launch only
                cudaError_t ce = cudaStreamWaitEvent(e->st_pass, e->ev_ov[0], 0);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->st_trace, e->ev_ov[0], 0);
                if (ce != cudaSuccess) return ce;
                phase_kernel(1, phase_regs(1))<<<grid, WARPS_PER_CTA * 32, 0, e->st_pass>>>(bd);
                apa_phase_cont_kernel<<<grid, WARPS_PER_CTA * 32, 0, e->st_pass>>>(bd);  // pairs that need a second pass (none on the headline)
                phase_kernel(2, phase_regs(2))<<<grid, WARPS_PER_CTA * 32, 0, e->st_trace>>>(bd);
                b->stats.kernel_launches += 3;
                ce = cudaEventRecord(e->ev_ov[1], e->st_pass);
                if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_ov[2], e->st_trace);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, e->ev_ov[1], 0);  // the engine's stream continues when all three are done
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, e->ev_ov[2], 0);
                if (ce == cudaSuccess) ce = cudaEventRecord(e->evp[2], st);
                if (ce == cudaSuccess) ce = cudaEventRecord(e->evp[3], st);
                return ce != cudaSuccess ? ce : cudaGetLastError();
            }
            if (coop_w == 8) {
                apa_phase_pass_coop_kernel<8, 1><<<std::min<unsigned>(wave_pairs, (unsigned)e->sm_count * 4), 256, 0, st>>>(bd);
                apa_phase_pass_coop_kernel<8, 3><<<std::min<unsigned>(wave_pairs, (unsigned)e->sm_count * 4), 256, 0, st>>>(bd);
            } else if (coop_w == 4) {
                apa_phase_pass_coop_kernel<4, 1><<<std::min<unsigned>(wave_pairs, (unsigned)e->sm_count * 8), 128, 0, st>>>(bd);
                apa_phase_pass_coop_kernel<4, 3><<<std::min<unsigned>(wave_pairs, (unsigned)e->sm_count * 8), 128, 0, st>>>(bd);
            } else {
                phase_kernel(1, phase_regs(1))<<<grid, WARPS_PER_CTA * 32, 0, st>>>(bd);
                apa_phase_cont_kernel<<<grid, WARPS_PER_CTA * 32, 0, st>>>(bd);
            }
            b->stats.kernel_launches += 2;
            cudaError_t ce = cudaEventRecord(e->evp[2], st);
            if (ce != cudaSuccess) return ce;
            if (trace) {
                phase_kernel(2, phase_regs(2))<<<grid, WARPS_PER_CTA * 32, 0, st>>>(bd);
                b->stats.kernel_launches++;
            }
            ce = cudaEventRecord(e->evp[3], st);
            return ce != cudaSuccess ? ce : cudaGetLastError();
        };
        auto add_phase_ms = [&]() -> cudaError_t {
            for (int k = 0; k < 3; k++) {
                float pms = 0;
                cudaError_t ce = cudaEventElapsedTime(&pms, e->evp[k], e->evp[k + 1]);
                if (ce != cudaSuccess) return ce;
                b->stats.phase_ms[k] += pms;
            }
            return cudaSuccess;
        };
        if (split) {
            for (uint64_t w0 = 0; w0 < n_work; w0 += wave_n) {
                bd.q0 = (uint32_t)w0;
                bd.n_order = (uint32_t)std::min<uint64_t>(w0 + wave_n, n_work);
                CUDA_TRY(cudaMemsetAsync(e->d_queue + 24, 0, 4 * sizeof(unsigned long long), st));
                CUDA_TRY(cudaEventRecord(e->evp[0], st));
                if (overlap) CUDA_TRY(cudaEventRecord(e->ev_ov[0], st));
                // Overlapped under a streamed upload: the build kernel waits for data for the first ~20 ms anyway, so it leaves
                // SM slots to the pass kernel from the start (APA_BUILD_CTAS per SM; the default was measured, profiles/README.md).
                unsigned build_grid = grid;
                if (overlap && streaming) {
                    int per_sm = 9;
                    if (const char* ev = getenv("APA_BUILD_CTAS")) per_sm = std::max(1, atoi(ev));
                    build_grid = std::min<unsigned>(grid, (unsigned)(e->sm_count * per_sm));
                }
                phase_kernel(0, phase_regs(0))<<<build_grid, WARPS_PER_CTA * 32, 0, st>>>(bd);
                b->stats.kernel_launches++;
                CUDA_TRY(cudaEventRecord(e->evp[1], st));
                // With a streaming upload the build kernel spins until the bases arrive: nothing that may synchronise with
                // the device (first-use module loading of another kernel, allocations) may be issued before the upload is
                // done, so the pass / trace kernels of a back-to-back run are launched after stream_upload() below. Overlapped
                // kernels go out at once: they wait for their pairs on the device, and every kernel was loaded at engine creation.
                if (!host_streaming || overlap) CUDA_TRY(launch_pass_trace());
                if (w0 + wave_n < n_work) {  // more waves follow: the arenas are reused
                    CUDA_TRY(cudaStreamSynchronize(st));
                    CUDA_TRY(add_phase_ms());
                }
            }
        } else if (gp) {
            CUDA_TRY(apa_general_launch(bd, *gp, (unsigned)(slots / WARPS_PER_CTA), st));
        } else if (regs >= 64)
            apa_align_kernel_r64<<<(unsigned)(slots / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(bd);
        else if (regs >= 48)
            apa_align_kernel_r48<<<(unsigned)(slots / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(bd);
        else
            apa_align_kernel_r40<<<(unsigned)(slots / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(bd);
        if (!split) b->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        if (host_streaming) {
            // the chunks travel while the persistent kernel already consumes the ones that have landed
            int rc = stream_upload(e, b, stream_mode);
            b->stats.upload_chunks_raw = b->chunks_raw;
            if (rc != APA_OK) {
                // let the kernel drain (every chunk state must become non-zero; chunks that were claimed have been sent), then report
                std::vector<uint32_t> ones(256, 1u);
                cudaStreamSynchronize(e->copy_stream);
                cudaStreamSynchronize(e->copy_stream2);
                cudaMemcpy(e->d_ready, ones.data(), 256 * 4, cudaMemcpyHostToDevice);
                cudaStreamSynchronize(st);
                return rc;
            }
            if (split && !overlap) CUDA_TRY(launch_pass_trace());
        }
        CUDA_TRY(cudaMemcpyAsync(b->h_status.data(), b->d_status, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (streaming) {  // the caller's buffers are no longer being read
            CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
            CUDA_TRY(cudaStreamSynchronize(e->copy_stream2));
        }
        if (split) CUDA_TRY(add_phase_ms());
        pending.clear();
        for (uint64_t p = 0; p < b->n_pairs; p++)
            if (b->h_status[p] == ST_OVERFLOW) pending.push_back((uint32_t)p);
        if (pending.empty()) break;
        b->stats.retries += pending.size();
        for (uint32_t p : pending) b->h_status[p] = ST_PENDING;
        // reset the status of the overflowed pairs and grow the arena
        std::vector<int32_t> zero(1, 0);
        for (uint32_t p : pending) CUDA_TRY(cudaMemcpyAsync(b->d_status + p, zero.data(), 4, cudaMemcpyHostToDevice, st));
        uint64_t grown = (uint64_t)arena_size * 4;
        if (grown > 0xF0000000ull) {
            if (arena_size >= 0xF0000000u) return set_err(APA_ERR_TOO_LARGE, "pair does not fit the largest scratch arena");
            grown = 0xF0000000ull;
        }
        arena_size = (uint32_t)grown;
    }
    if (!pending.empty()) return set_err(APA_ERR_TOO_LARGE, "pairs still overflow after 8 arena enlargements");
    CUDA_TRY(cudaEventRecord(e->ev[3], st));
    unsigned long long h_q[32];
    CUDA_TRY(cudaMemcpyAsync(h_q, e->d_queue, sizeof h_q, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]));
    b->stats.kernel_ms = ms;
    if (getenv("APA_DEBUG_TIMING") && b->stats.phase_ms[1] > 0) {
        float pre = 0, post = 0;
        cudaEventElapsedTime(&pre, e->ev[2], e->evp[0]);
        cudaEventElapsedTime(&post, e->evp[3], e->ev[3]);
        fprintf(stderr, "[batch_run] total %.2f ms: setup %.2f | build %.2f | pass %.2f | trace %.2f | tail %.2f\n", ms, pre,
                b->stats.phase_ms[0], b->stats.phase_ms[1], b->stats.phase_ms[2], post);
    }
    b->pool_used = h_q[1];
    b->stats.dp_word_steps = h_q[2];
    b->stats.computed_cells = h_q[3];
    b->stats.passes = h_q[4];
    b->stats.fill_blocks = h_q[5];
    b->stats.dt_blocks = h_q[6];
    for (int t = 0; t < 8; t++) b->stats.phase_cycles[t] = h_q[7 + t];
    b->stats.score_calls = h_q[15];
    b->stats.score_probes = h_q[16];
    b->stats.dp_issue_steps = h_q[17];
    b->ran = true;
    for (uint64_t p = 0; p < b->n_pairs; p++) {
        if (b->h_status[p] == ST_BAD_INPUT) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT in pair " + std::to_string(p));
        if (b->h_status[p] == ST_TOO_LARGE) return set_err(APA_ERR_TOO_LARGE, "CIGAR run of 2^30 or more operations in pair " + std::to_string(p));
        if (b->h_status[p] != ST_DONE)
            return set_err(APA_ERR_INTERNAL, "device assertion in pair " + std::to_string(p) + " (status " + std::to_string(b->h_status[p]) + ")");
    }
    return APA_OK;
}

extern "C" int apa_batch_run(apa_engine* e, apa_batch* b, int preset, int trace) { return batch_run(e, b, preset, trace, false); }

// dst != nullptr: the CIGAR text goes to the caller's (page-locked) buffer of at least pool_used bytes instead of a fresh pool.
static int batch_download(apa_engine* e, apa_batch* b, int64_t* costs, char** cigar_pool, int64_t* cigar_off, int64_t* cigar_len, char* dst) {
    if (!e || !b || !b->ran) return set_err(APA_ERR_BAD_INPUT, "batch has not been run");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (cigar_pool) *cigar_pool = nullptr;
    if (b->n_pairs == 0) return APA_OK;
    CUDA_TRY(cudaEventRecord(e->ev[4], st));
    std::vector<int32_t> c32(b->n_pairs);
    CUDA_TRY(cudaMemcpyAsync(c32.data(), b->d_cost, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
    uint64_t bytes = b->n_pairs * 4;
    char* pool = nullptr;
    if (b->trace && (cigar_pool || dst) && cigar_off && cigar_len) {
        pool = dst ? dst : (char*)pinned_pool_get(std::max<uint64_t>(b->pool_used, 1));
        if (!pool) return set_err(APA_ERR_TOO_LARGE, "host allocation of the CIGAR pool failed");
        CUDA_TRY(cudaMemcpyAsync(pool, b->d_pool, b->pool_used, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cigar_off, b->d_cig_off, b->n_pairs * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cigar_len, b->d_cig_len, b->n_pairs * 8, cudaMemcpyDeviceToHost, st));
        bytes += b->pool_used + b->n_pairs * 16;
    }
    CUDA_TRY(cudaEventRecord(e->ev[5], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (uint64_t p = 0; p < b->n_pairs; p++) costs[p] = c32[p];
    if (pool && cigar_pool) *cigar_pool = pool;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]));
    b->stats.d2h_ms = ms;
    b->stats.d2h_bytes = bytes;
    return APA_OK;
}
extern "C" int apa_batch_download(apa_engine* e, apa_batch* b, int64_t* costs, char** cigar_pool, int64_t* cigar_off, int64_t* cigar_len) {
    return batch_download(e, b, costs, cigar_pool, cigar_off, cigar_len, nullptr);
}

extern "C" int apa_batch_download_pair_stats(apa_engine* e, apa_batch* b, apa_pair_stats* out) {
    if (!e || !b || !b->ran || !out) return set_err(APA_ERR_BAD_INPUT, "batch has not been run");
    static_assert(sizeof(apa_pair_stats) == 64, "apa_pair_stats layout");
    CUDA_TRY(cudaSetDevice(e->device));
    if (b->n_pairs == 0) return APA_OK;
    CUDA_TRY(cudaMemcpyAsync(out, b->d_pair_stats, b->n_pairs * 64, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return APA_OK;
}

extern "C" int apa_batch_get_stats(apa_batch* b, apa_batch_stats* out) {
    if (!b || !out) return APA_ERR_BAD_INPUT;
    *out = b->stats;
    return APA_OK;
}

extern "C" int apa_align_batch(apa_engine* e, int preset, int trace, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off,
                               const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool, int64_t* cigar_off,
                               int64_t* cigar_len, apa_batch_stats* stats) {
    // Host buffers in, host buffers out: the bases stream to HBM in chunks while the persistent kernel is already
    // aligning the pairs that have arrived (H2D overlaps compute); results come back in one D2H at the end.
    apa_batch* b = nullptr;
    auto t0 = std::chrono::steady_clock::now();
    int rc = batch_prepare(e, n_pairs, a_all, a_off, b_all, b_off, true, &b);
    auto t1 = std::chrono::steady_clock::now();
    if (rc == APA_OK) rc = batch_run(e, b, preset, trace, true);
    auto t2 = std::chrono::steady_clock::now();
    if (rc == APA_OK) rc = apa_batch_download(e, b, costs, cigar_pool, cigar_off, cigar_len);
    auto t3 = std::chrono::steady_clock::now();
    if (getenv("APA_DEBUG_TIMING")) {
        auto ms = [](auto x, auto y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
        fprintf(stderr, "[apa_align_batch] prepare %.1f ms, run %.1f ms (pack+h2d %.1f ms), download %.1f ms\n", ms(t0, t1), ms(t1, t2),
                b ? b->pack_ms : 0.0, ms(t2, t3));
    }
    if (rc == APA_OK && stats) *stats = b->stats;
    apa_batch_free(e, b);
    return rc;
}

// ---- packed (2-bit) input
static uint64_t packed_halfwords(int64_t len) { return (uint64_t)(((len + 63) / 64) * 2 + 2 + 15) & ~(uint64_t)15; }
extern "C" int apa_packed_layout(uint64_t n, const int64_t* len, int64_t* off_out) {
    uint64_t hw = 0;
    for (uint64_t p = 0; p < n; p++) {
        if (len[p] < 0 || len[p] >= (1ll << 31) - 1024) return set_err(APA_ERR_TOO_LARGE, "sequence length must be < 2^31 (I = i32)");
        off_out[p] = (int64_t)hw;
        hw += packed_halfwords(len[p]);
    }
    off_out[n] = (int64_t)hw;
    return APA_OK;
}
extern "C" int apa_pack_sequences(uint64_t n, const uint8_t* seq_all, const int64_t* seq_off, uint32_t* planes_out, const int64_t* off, int n_threads) {
    std::atomic<uint64_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        for (;;) {
            const uint64_t p = next.fetch_add(1);
            if (p >= n) break;
            if (apa_pack_planes_host(seq_all + seq_off[p], seq_off[p + 1] - seq_off[p], 0, off[p + 1] - off[p], planes_out + (size_t)off[p] * 2)) bad.store(1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < std::max(1, n_threads); t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (bad.load()) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT (the reference panics here: pa-bitpacking/src/profile.rs:113)");
    return APA_OK;
}
extern "C" int apa_align_batch_packed(apa_engine* e, int preset, int trace, uint64_t n_pairs, const uint32_t* a_planes, const int64_t* a_len,
                                      const uint32_t* b_planes, const int64_t* b_len, int64_t* costs, char** cigar_pool, int64_t* cigar_off,
                                      int64_t* cigar_len, apa_batch_stats* stats) {
    if (n_pairs && (!a_planes || !b_planes || !a_len || !b_len)) return set_err(APA_ERR_BAD_INPUT, "apa_align_batch_packed: null input");
    std::vector<int64_t> a_off(n_pairs + 1, 0), b_off(n_pairs + 1, 0);
    for (uint64_t p = 0; p < n_pairs; p++) {
        a_off[p + 1] = a_off[p] + a_len[p];
        b_off[p + 1] = b_off[p] + b_len[p];
    }
    apa_batch* b = nullptr;
    static const uint32_t none[2] = {0u, 0u};  // an empty batch still says "packed input"
    int rc = batch_prepare(e, n_pairs, nullptr, a_off.data(), nullptr, b_off.data(), true, &b, a_planes ? a_planes : none, b_planes ? b_planes : none);
    if (rc == APA_OK) rc = batch_run(e, b, preset, trace, true);
    if (rc == APA_OK) rc = apa_batch_download(e, b, costs, cigar_pool, cigar_off, cigar_len);
    if (rc == APA_OK && stats) *stats = b->stats;
    apa_batch_free(e, b);
    return rc;
}

extern "C" int apa_batch_run_params(apa_engine* e, apa_batch* b, const apa_params* params, int trace) {
    RunParams rp;
    int rc = validate_params(params, &rp);
    if (rc != APA_OK) return rc;
    return batch_run(e, b, -1, trace, false, &rp);
}

extern "C" int apa_align_batch_params(apa_engine* e, const apa_params* params, int trace, uint64_t n_pairs, const uint8_t* a_all,
                                      const int64_t* a_off, const uint8_t* b_all, const int64_t* b_off, int64_t* costs,
                                      char** cigar_pool, int64_t* cigar_off, int64_t* cigar_len, apa_batch_stats* stats) {
    RunParams rp;
    int rc = validate_params(params, &rp);
    if (rc != APA_OK) return rc;
    apa_batch* b = nullptr;
    rc = apa_batch_upload(e, n_pairs, a_all, a_off, b_all, b_off, &b);
    if (rc == APA_OK) rc = batch_run(e, b, -1, trace, false, &rp);
    if (rc == APA_OK) rc = apa_batch_download(e, b, costs, cigar_pool, cigar_off, cigar_len);
    if (rc == APA_OK && stats) *stats = b->stats;
    apa_batch_free(e, b);
    return rc;
}

// ------------------------------------------------------------------------------------------------ band log (tests)
static int64_t band_log_impl(apa_engine* e, int preset, const apa_params* params, int trace, const uint8_t* a, uint64_t n, const uint8_t* b,
                             uint64_t m, int32_t* out, uint64_t cap);
extern "C" int64_t apa_debug_band_log(apa_engine* e, int preset, int trace, const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m,
                                      int32_t* out, uint64_t cap) {
    return band_log_impl(e, preset, nullptr, trace, a, n, b, m, out, cap);
}
extern "C" int64_t apa_debug_band_log_params(apa_engine* e, const apa_params* params, int trace, const uint8_t* a, uint64_t n,
                                             const uint8_t* b, uint64_t m, int32_t* out, uint64_t cap) {
    if (!params) return set_err(APA_ERR_BAD_INPUT, "null params");
    return band_log_impl(e, -1, params, trace, a, n, b, m, out, cap);
}
static int64_t band_log_impl(apa_engine* e, int preset, const apa_params* params, int trace, const uint8_t* a, uint64_t n, const uint8_t* b,
                             uint64_t m, int32_t* out, uint64_t cap) {
    int64_t a_off[2] = {0, (int64_t)n}, b_off[2] = {0, (int64_t)m};
    apa_batch* bt = nullptr;
    int rc = apa_batch_upload(e, 1, a, a_off, b, b_off, &bt);
    if (rc != APA_OK) return rc;
    const uint64_t rec_cap64 = 7ull * 64ull * (n / (uint64_t)(params ? std::max(1, params->block_width) : BLOCK_W) + 2);
    if (rec_cap64 > (1ull << 28)) {
        apa_batch_free(e, bt);
        return set_err(APA_ERR_TOO_LARGE, "band log: too many blocks for the debug log");
    }
    const uint32_t rec_cap = (uint32_t)rec_cap64;
    std::vector<int32_t> rec(rec_cap);
    uint32_t nrec = 0;
    cudaError_t ce = cudaMalloc(&bt->d_dbg, rec_cap * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&bt->d_dbg_n, 4);
    if (ce == cudaSuccess) ce = cudaMemset(bt->d_dbg_n, 0, 4);
    bt->dbg_cap = rec_cap;
    if (ce == cudaSuccess) rc = params ? apa_batch_run_params(e, bt, params, trace) : apa_batch_run(e, bt, preset, trace);
    if (ce == cudaSuccess && rc == APA_OK) ce = cudaMemcpy(&nrec, bt->d_dbg_n, 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && rc == APA_OK) ce = cudaMemcpy(rec.data(), bt->d_dbg, std::min(nrec, rec_cap) * 4, cudaMemcpyDeviceToHost);
    cudaFree(bt->d_dbg);
    cudaFree(bt->d_dbg_n);
    apa_batch_free(e, bt);
    if (ce != cudaSuccess) return set_err(APA_ERR_CUDA, cudaGetErrorString(ce));
    if (rc != APA_OK) return rc;
    if (nrec > rec_cap) return set_err(APA_ERR_TOO_LARGE, "band log truncated");
    // regroup into the oracle's layout: npass, then per pass: f_max, nblocks, nblocks x (j_s, j_e, fixed_s, fixed_e)
    std::vector<int32_t> o;
    o.push_back(0);
    int cur_pass = -1;
    size_t cnt_pos = 0;
    for (uint32_t r = 0; r + 7 <= nrec; r += 7) {
        if (rec[r] != cur_pass) {
            cur_pass = rec[r];
            o[0]++;
            o.push_back(rec[r + 1]);
            cnt_pos = o.size();
            o.push_back(0);
        }
        o[cnt_pos]++;
        for (int k = 3; k < 7; k++) o.push_back(rec[r + k]);
    }
    for (size_t t = 0; t < o.size() && t < cap; t++) out[t] = o[t];
    return (int64_t)o.size();
}

// ------------------------------------------------------------------------------------------------ pa_bitpacking::search
// Host side of search / trace: the pattern's match masks, the start column v0 (search.rs:58-66) and input validation.
struct SearchSetup {
    uint64_t nwords = 0, nhw = 0, padding = 0;
    std::vector<uint32_t> pmask, v0;  // per half-word: 4 mask words (A C T G) / (p, m)
};
static int search_setup(const uint8_t* pattern, uint64_t np, const uint8_t* text, uint64_t nt, float unmatched_cost, SearchSetup& S) {
    if (!(unmatched_cost >= 0.0f && unmatched_cost <= 1.0f)) return set_err(APA_ERR_BAD_INPUT, "unmatched_cost must be in [0, 1]");
    if (nt >= (1ull << 31) - 1024 || np >= (1ull << 24)) return set_err(APA_ERR_TOO_LARGE, "search: text < 2^31, pattern < 2^24");
    // ScatterProfile::build (profile.rs:28-66): per 32 rows of the pattern, the rows matching A / C / T / G.
    S.nwords = (np + 63) / 64;
    S.nhw = S.nwords * 2;
    S.padding = S.nwords * 64 - np;
    S.pmask.assign(S.nhw * 4, 0u);
    S.v0.assign(S.nhw * 2, 0u);
    for (uint64_t j = 0; j < np; j++) {
        uint32_t m4 = 0;  // bit i: matches base i (A0 C1 T2 G3)
        switch (pattern[j]) {
            case 'a': case 'A': m4 = 1; break;
            case 'c': case 'C': m4 = 2; break;
            case 't': case 'T': m4 = 4; break;
            case 'g': case 'G': m4 = 8; break;
            case 'n': case 'N': case '*': m4 = 15; break;
            case 'y': case 'Y': m4 = 6; break;  // C or T
            case 'r': case 'R': m4 = 9; break;  // A or G
            default: return set_err(APA_ERR_BAD_INPUT, "search: unknown pattern base (ACGT, N, *, Y, R)");
        }
        for (int i = 0; i < 4; i++)
            if (m4 >> i & 1) S.pmask[(j / 32) * 4 + i] |= 1u << (j % 32);
    }
    for (uint64_t j = np; j < S.nwords * 64; j++)
        for (int i = 0; i < 4; i++) S.pmask[(j / 32) * 4 + i] |= 1u << (j % 32);  // padding rows match everything
    for (uint64_t i = 0; i < nt; i++) {
        const uint8_t c = text[i] & 0xDF;  // acgtACGT only (profile.rs:31-38)
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return set_err(APA_ERR_BAD_INPUT, "search: text byte outside acgtACGT");
    }
    if (unmatched_cost > 0.0f) {  // search.rs:58-66
        for (uint64_t i = 0;; i++) {
            const uint64_t idx = (uint64_t)std::ceil((float)i / unmatched_cost);
            if (idx >= np) break;
            S.v0[(idx / 32) * 2] |= 1u << (idx % 32);
        }
    }
    return APA_OK;
}
// Device buffers of a search call come from the engine's cache (no cudaMalloc / cudaFree per call once it is warm).
struct SearchBufs {
    apa_engine* e;
    std::vector<void*> ptrs;
    ~SearchBufs() {
        for (void* p : ptrs) eng_release(e, p);
    }
    template <class T>
    cudaError_t get(T** out, size_t bytes) {
        void* p = nullptr;
        cudaError_t ce = eng_alloc(e, &p, std::max<size_t>(bytes, 16));
        if (ce == cudaSuccess) ptrs.push_back(p);
        *out = (T*)p;
        return ce;
    }
};
// The `out` vector of pa_bitpacking::search from the final column and the bottom deltas (search.rs:72-104): bottom row, then up
// the right column; the first `padding` values are skipped because the pattern was rounded up to a multiple of 64 rows.
static int search_assemble(const SearchSetup& S, uint64_t np, uint64_t nt, const std::vector<uint32_t>& vfin, const std::vector<int8_t>& hd,
                           int32_t* out) {
    auto word = [&](const std::vector<uint32_t>& vv, uint64_t w, int pm) -> uint64_t {
        return (uint64_t)vv[(2 * w) * 2 + pm] | ((uint64_t)vv[(2 * w + 1) * 2 + pm] << 32);
    };
    auto value = [&](const std::vector<uint32_t>& vv, uint64_t w) -> int {
        return __builtin_popcountll(word(vv, w, 0)) - __builtin_popcountll(word(vv, w, 1));
    };
    auto suffix = [&](const std::vector<uint32_t>& vv, uint64_t w, int j) -> int {
        const uint64_t mask = ~((1ull << (64 - j)) - 1ull);  // V::value_of_suffix, encoding.rs:33-38 (0 < j <= 64)
        return __builtin_popcountll(word(vv, w, 0) & mask) - __builtin_popcountll(word(vv, w, 1) & mask);
    };
    int b = 0;
    for (uint64_t w = 0; w < S.nwords; w++) b += value(S.v0, w);
    uint64_t n_out = 0, skipped = 0;
    out[n_out++] = b;
    for (uint64_t i = 0; i < nt; i++) {
        b += hd[i];
        if (skipped < S.padding)
            skipped++;
        else
            out[n_out++] = b;
    }
    for (uint64_t w = S.nwords; w-- > 0;) {
        for (int j = 1; j <= 64; j++) {
            const int val = b - suffix(vfin, w, j) + suffix(S.v0, w, j);
            if (skipped < S.padding)
                skipped++;
            else
                out[n_out++] = val;
        }
        b -= value(vfin, w);
        b += value(S.v0, w);
    }
    if (n_out != np + nt + 1) return set_err(APA_ERR_INTERNAL, "search: output length");
    return APA_OK;
}

extern "C" int apa_search(apa_engine* e, const uint8_t* pattern, uint64_t np, const uint8_t* text, uint64_t nt, float unmatched_cost,
                          int32_t* out) {
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    SearchSetup S;
    int rc = search_setup(pattern, np, text, nt, unmatched_cost, S);
    if (rc != APA_OK) return rc;
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<uint32_t> vfin(S.v0);
    std::vector<int8_t> hd(std::max<uint64_t>(nt, 1), 0);
    if (nt > 0 && S.nhw > 0) {
        // text segments, one warp each: long enough that the 2 * rows warm-up columns stay a small share, short enough to fill the GPU
        const int nhw = (int)S.nhw;
        const int warm = 2 * 32 * nhw;
        const long long min_seg = std::max<long long>(1024, 8ll * warm);
        const long long want_segs = (long long)e->sm_count * 16;
        long long seg_len = std::max<long long>(min_seg, ((long long)nt + want_segs - 1) / want_segs);
        seg_len = (seg_len + BLOCK_W - 1) / BLOCK_W * BLOCK_W;
        const long long n_segs = ((long long)nt + seg_len - 1) / seg_len;
        SearchBufs B{e, {}};
        uint8_t* d_text;
        uint4* d_pmask;
        uint2 *d_v0, *d_vwork, *d_vfin;
        int8_t* d_h;
        cudaStream_t st = e->stream;
        CUDA_TRY(B.get(&d_text, nt));
        CUDA_TRY(B.get(&d_pmask, S.nhw * 16));
        CUDA_TRY(B.get(&d_v0, S.nhw * 8));
        CUDA_TRY(B.get(&d_vwork, (size_t)n_segs * S.nhw * 8));
        CUDA_TRY(B.get(&d_vfin, S.nhw * 8));
        CUDA_TRY(B.get(&d_h, nt));
        CUDA_TRY(cudaMemcpyAsync(d_text, text, nt, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_pmask, S.pmask.data(), S.nhw * 16, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_v0, S.v0.data(), S.nhw * 8, cudaMemcpyHostToDevice, st));
        apa_search_kernel<false><<<(unsigned)((n_segs + 3) / 4), 128, 0, st>>>(d_text, (int)nt, d_pmask, nhw, d_v0, d_vwork, d_vfin, d_h, (int)seg_len,
                                                                               warm, nullptr);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(vfin.data(), d_vfin, S.nhw * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hd.data(), d_h, nt, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return search_assemble(S, np, nt, vfin, hd, out);
}

// SearchResult::trace(idx) (search.rs:135-230): the alignment that ends at out[idx]. The window of the text that can hold it
// (2 |pattern| columns, doubled while the cost found at the end position exceeds the target) is re-filled on the GPU with every
// column kept (apa_search_kernel<true>), and walked back on the GPU (apa_search_trace_kernel); the host formats the CIGAR text.
// cigar_out: NUL-terminated text (pa_types::Cigar::to_string: count omitted when 1), at most cigar_cap bytes; pos_out: start.i,
// start.j (where the walk stopped: poss[0] of the reference), end.i, end.j, cost.
extern "C" int apa_search_trace(apa_engine* e, const uint8_t* pattern, uint64_t np, const uint8_t* text, uint64_t nt, float unmatched_cost,
                                uint64_t idx, char* cigar_out, uint64_t cigar_cap, int32_t* pos_out) {
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    if (idx > np + nt) return set_err(APA_ERR_BAD_INPUT, "search trace: idx out of range");  // assert, search.rs:122
    std::vector<int32_t> out(np + nt + 1);
    int rc = apa_search(e, pattern, np, text, nt, unmatched_cost, out.data());
    if (rc != APA_OK) return rc;
    SearchSetup S;
    rc = search_setup(pattern, np, text, nt, unmatched_cost, S);
    if (rc != APA_OK) return rc;
    // idx_to_pos (search.rs:121-132) and the target cost (search.rs:137-140: the right column's values carry the unmatched rows)
    const int64_t pi = idx <= nt ? (int64_t)idx : (int64_t)nt, pj = idx <= nt ? (int64_t)np : (int64_t)(np - (idx - nt));
    int target = out[idx];
    if ((uint64_t)pi == nt) {  // V::value_from(&v0, pos.1): rows pj .. of the start column
        for (uint64_t j = (uint64_t)pj; j < S.nwords * 64; j++)
            target -= (int)((S.v0[(j / 32) * 2] >> (j % 32)) & 1u) - (int)((S.v0[(j / 32) * 2 + 1] >> (j % 32)) & 1u);
    }
    pos_out[2] = (int32_t)pi, pos_out[3] = (int32_t)pj, pos_out[4] = target;
    if (S.nhw == 0) {  // empty pattern: nothing to walk
        if (cigar_cap < 1) return set_err(APA_ERR_TOO_LARGE, "search trace: cigar buffer too small");
        cigar_out[0] = 0;
        pos_out[0] = (int32_t)pi, pos_out[1] = 0;
        return APA_OK;
    }
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const int nhw = (int)S.nhw;
    const uint64_t end = (uint64_t)pi;
    uint64_t width = 2 * np;
    std::vector<uint32_t> ones(S.nhw * 2, 0u);
    for (uint64_t h = 0; h < S.nhw; h++) ones[2 * h] = ~0u;
    for (;;) {
        const uint64_t start = end > width ? end - width : 0;
        const uint64_t ncol = end - start;
        if ((ncol + 1) * S.nhw * 8 > (8ull << 30)) return set_err(APA_ERR_TOO_LARGE, "search trace: window too large");
        SearchBufs B{e, {}};
        uint8_t* d_text;
        uint4* d_pmask;
        uint2 *d_v0, *d_vwork, *d_vfin, *d_fill;
        uint32_t* d_elems;
        int* d_out;
        const uint32_t elem_cap = (uint32_t)std::min<uint64_t>(ncol + np + 16, 1u << 30);
        CUDA_TRY(B.get(&d_text, nt + 1));
        CUDA_TRY(B.get(&d_pmask, S.nhw * 16));
        CUDA_TRY(B.get(&d_v0, S.nhw * 8));
        CUDA_TRY(B.get(&d_vwork, S.nhw * 8 * 4));
        CUDA_TRY(B.get(&d_vfin, S.nhw * 8));
        CUDA_TRY(B.get(&d_fill, (ncol + 1) * S.nhw * 8));
        CUDA_TRY(B.get(&d_elems, (size_t)elem_cap * 4));
        CUDA_TRY(B.get(&d_out, 8 * 4));
        if (nt) CUDA_TRY(cudaMemcpyAsync(d_text, text, nt, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_pmask, S.pmask.data(), S.nhw * 16, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_v0, start == 0 ? S.v0.data() : ones.data(), S.nhw * 8, cudaMemcpyHostToDevice, st));
        const int seg_len = (int)((std::max<uint64_t>(ncol, 1) + BLOCK_W - 1) / BLOCK_W * BLOCK_W);
        apa_search_kernel<true><<<1, 128, 0, st>>>(d_text + start, (int)ncol, d_pmask, nhw, d_v0, d_vwork, d_vfin, nullptr, seg_len, 0, d_fill);
        CUDA_TRY(cudaGetLastError());
        apa_search_trace_kernel<<<1, 32, 0, st>>>(d_text, d_pmask, nhw, d_fill, (int)start, (int)end, (int)pj, target, d_elems, elem_cap, d_out);
        CUDA_TRY(cudaGetLastError());
        int h_out[8] = {0};
        CUDA_TRY(cudaMemcpyAsync(h_out, d_out, 5 * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_out[0] == 1) {  // cost > target_cost: the window cut the path off
            if (start == 0) return set_err(APA_ERR_INTERNAL, "search trace: target cost not reached with the whole text");
            width *= 2;
            continue;
        }
        if (h_out[0] != 0) return set_err(APA_ERR_INTERNAL, "search trace: a reference panic path was reached (search.rs:176-226)");
        const uint32_t n_el = (uint32_t)h_out[1];
        if (n_el > elem_cap) return set_err(APA_ERR_INTERNAL, "search trace: element buffer");
        std::vector<uint32_t> el(n_el);
        if (n_el) CUDA_TRY(cudaMemcpy(el.data(), d_elems, (size_t)n_el * 4, cudaMemcpyDeviceToHost));
        std::string txt;
        static const char opc[4] = {'=', 'X', 'D', 'I'};
        for (uint32_t k = n_el; k-- > 0;) {  // cigar.reverse()
            const uint32_t cnt = el[k] & 0x3fffffffu;
            if (cnt != 1) txt += std::to_string(cnt);
            txt += opc[el[k] >> 30];
        }
        if (txt.size() + 1 > cigar_cap) return set_err(APA_ERR_TOO_LARGE, "search trace: cigar buffer too small");
        memcpy(cigar_out, txt.c_str(), txt.size() + 1);
        pos_out[0] = h_out[2], pos_out[1] = h_out[3];
        return APA_OK;
    }
}

// ------------------------------------------------------------------------------------------------ INT32 issue-rate probe
// Measures what the block DP's roofline is quoted against ("must be measured, do not assume"): the rate at which
// this GPU retires the integer instructions of the Myers step when every SM is full of warps that do nothing else. Eight
// independent dependent chains per thread, operands the compiler cannot fold. MODE 0 LOP3, 1 SHF (funnel shift), 2 IADD3,
// 3 IMAD (FMA pipe), 4 the block-DP mix per step: 8 LOP3 + 2 SHF on the ALU pipe next to 4 IMAD on the FMA pipe.
template <int MODE>
__global__ void __launch_bounds__(256) apa_peak_kernel(uint32_t* out, int iters, uint32_t k1, uint32_t k2) {
    uint32_t x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = threadIdx.x * 2654435761u + c * 40503u + blockIdx.x;
    asm volatile("" : "+r"(k1), "+r"(k2));  // operands in registers, opaque to constant folding
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t v = x[c];
            if (MODE == 0) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(k1), "r"(k2));
            } else if (MODE == 1) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(k1), "r"(k2));
            } else if (MODE == 2) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(v) : "r"(k1), "r"(k2));
            } else if (MODE == 3) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(k1), "r"(k2));
            } else {
                uint32_t w = v ^ k2;
#pragma unroll
                for (int r = 0; r < 2; r++) {  // two Myers-step mixes: (4 LOP3, 1 SHF, 2 IMAD) x 2
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(w), "r"(k2));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w) : "r"(k1), "r"(v));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(v) : "r"(w), "r"(k1));
                    asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(w) : "r"(v), "r"(k2));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(v) : "r"(w), "r"(k2));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w) : "r"(k1), "r"(v));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(w), "r"(k1));
                }
                v ^= w;
            }
            x[c] = v;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) acc ^= x[c];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;  // keeps the chains alive, (almost) never stores
}

// out[0..4]: lane-operations per second of LOP3 / SHF / IADD3 / IMAD alone and of the ALU-pipe instructions (LOP3 + SHF) inside
// the block-DP mix; out[5]: SM clock in Hz during the probe (cycles of SM 0 / elapsed time is not available from the host: the
// value is the device's current clock rate attribute); out[6]: the IMADs retired per second next to that mix.
extern "C" int apa_int32_peak(apa_engine* e, double* out) {
    if (!e || !out) return set_err(APA_ERR_NO_DEVICE, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    uint32_t* d_out = nullptr;
    const int grid = e->sm_count * 8, block = 256, iters = 2000;
    CUDA_TRY(cudaMalloc(&d_out, (size_t)grid * block * 4));
    struct Free {
        void* p;
        ~Free() { cudaFree(p); }
    } fr{d_out};
    cudaStream_t st = e->stream;
    auto run = [&](int mode, double& ms_out) -> cudaError_t {
        for (int rep = 0; rep < 2; rep++) {  // first repetition warms up
            cudaEventRecord(e->ev[0], st);
            switch (mode) {
                case 0: apa_peak_kernel<0><<<grid, block, 0, st>>>(d_out, iters, 0x9e3779b9u, 0x7f4a7c15u); break;
                case 1: apa_peak_kernel<1><<<grid, block, 0, st>>>(d_out, iters, 0x9e3779b9u, 7u); break;
                case 2: apa_peak_kernel<2><<<grid, block, 0, st>>>(d_out, iters, 0x9e3779b9u, 0x7f4a7c15u); break;
                case 3: apa_peak_kernel<3><<<grid, block, 0, st>>>(d_out, iters, 0x9e3779b9u, 0x7f4a7c15u); break;
                default: apa_peak_kernel<4><<<grid, block, 0, st>>>(d_out, iters, 0x9e3779b9u, 13u); break;
            }
            cudaEventRecord(e->ev[1], st);
            cudaError_t ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) return ce;
            float ms = 0;
            ce = cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
            if (ce != cudaSuccess) return ce;
            ms_out = ms;
        }
        return cudaGetLastError();
    };
    const double lanes = (double)grid * block * iters * 8.0;
    for (int mode = 0; mode < 4; mode++) {
        double ms = 0;
        CUDA_TRY(run(mode, ms));
        out[mode] = lanes * 16.0 / (ms * 1e-3);  // MODE 2: the two adds of a round fuse into one IADD3
    }
    double ms = 0;
    CUDA_TRY(run(4, ms));
    out[4] = lanes * 10.0 / (ms * 1e-3);
    out[6] = lanes * 4.0 / (ms * 1e-3);
    int khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device));
    out[5] = khz * 1e3;
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ K0 on its own (tests)
extern "C" int apa_pack_planes_device(apa_engine* e, const uint8_t* seq, uint64_t len, uint32_t* out, uint64_t out_cap_halfwords,
                                      uint64_t* n_halfwords) {
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    if (len >= (1ull << 31) - 1024) return set_err(APA_ERR_TOO_LARGE, "sequence length must be < 2^31");
    CUDA_TRY(cudaSetDevice(e->device));
    const uint64_t nhw = (((len + 63) / 64) * 2 + 2 + 15) & ~15ull;
    if (n_halfwords) *n_halfwords = nhw;
    if (out_cap_halfwords < nhw) return set_err(APA_ERR_TOO_LARGE, "apa_pack_planes_device: output too small");
    uint8_t* d_raw = nullptr;
    uint2* d_prof = nullptr;
    int64_t* d_off = nullptr;
    int* d_bad = nullptr;
    struct Free {
        void** p[4];
        ~Free() {
            for (void** q : p) cudaFree(*q);
        }
    } free_all{{(void**)&d_raw, (void**)&d_prof, (void**)&d_off, (void**)&d_bad}};
    // the sequence is placed at an odd offset: the kernel must not assume any alignment of the raw bases
    CUDA_TRY(cudaMalloc(&d_raw, len + 64));
    CUDA_TRY(cudaMalloc(&d_prof, nhw * 8));
    CUDA_TRY(cudaMalloc(&d_off, 6 * 8));
    CUDA_TRY(cudaMalloc(&d_bad, 4));
    const int64_t offs[6] = {3, 3 + (int64_t)len, 0, 0, 0, (int64_t)nhw};  // a_off[2] | b_off[2] (empty b) | plane offsets[2]
    CUDA_TRY(cudaMemcpy(d_off, offs, sizeof offs, cudaMemcpyHostToDevice));
    if (len) CUDA_TRY(cudaMemcpy(d_raw + 3, seq, len, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemsetAsync(d_prof, 0xff, nhw * 8, e->stream));  // (on the kernel's stream: e->stream does not wait for the legacy stream)
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4, e->stream));
    BatchDev bd{};
    bd.n_pairs = 1;
    bd.a_off = d_off;
    bd.b_off = d_off + 2;
    bd.ap_off = d_off + 4;
    bd.bp_off = d_off + 2;  // {0, 0}: b is empty and owns no half-words
    bd.aprof = d_prof;
    bd.bprof = d_prof;
    bd.raw_a = d_raw;
    bd.raw_b = d_raw;
    apa_pack_kernel<<<dim3(1, 4), 256, 0, e->stream>>>(bd, d_bad);
    CUDA_TRY(cudaGetLastError());
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(out, d_prof, nhw * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (bad) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT (the reference panics here: pa-bitpacking/src/profile.rs:113)");
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ block KAT entry
extern "C" int apa_block_compute(apa_engine* e, const uint8_t* a, uint64_t na, const uint8_t* b, uint64_t mb, uint8_t* h, uint64_t* v,
                                 int64_t* bottom_sum) {
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    for (uint64_t i = 0; i < na; i++)
        if (h[i] > 2) return set_err(APA_ERR_BAD_INPUT, "apa_block_compute: h holds 0 (0), 1 (+1) or 2 (-1) per column");
    const uint64_t nwords = (mb + 63) / 64, nhw = nwords * 2;
    if (na == 0 || nhw == 0) {  // nothing to sweep: the top edge is the bottom edge
        int64_t s = 0;
        for (uint64_t i = 0; i < na; i++) s += h[i] == 1 ? 1 : (h[i] == 2 ? -1 : 0);
        *bottom_sum = s;
        return APA_OK;
    }
    // host-side planes of a and b in the device layout (+2 padding half-words)
    const uint64_t nhw_a = ((na + 63) / 64) * 2;
    std::vector<uint2> bp(nhw + 2), apl(nhw_a + 2);
    if (apa_pack_planes_host(b, (int64_t)mb, 0, (int64_t)nhw + 2, (uint32_t*)bp.data()) |
        apa_pack_planes_host(a, (int64_t)na, 0, (int64_t)nhw_a + 2, (uint32_t*)apl.data()))
        return set_err(APA_ERR_BAD_INPUT, "byte outside ACGT");
    std::vector<uint2> vv(nhw);
    for (uint64_t w = 0; w < nwords; w++) {
        vv[2 * w] = make_uint2((uint32_t)v[2 * w], (uint32_t)v[2 * w + 1]);
        vv[2 * w + 1] = make_uint2((uint32_t)(v[2 * w] >> 32), (uint32_t)(v[2 * w + 1] >> 32));
    }
    uint2 *d_a = nullptr, *d_bp = nullptr, *d_v = nullptr;
    uint8_t* d_h = nullptr;
    int32_t* d_cum = nullptr;
    struct Free {  // the CUDA_TRYs below may return early
        void** p[5];
        ~Free() {
            for (void** q : p) cudaFree(*q);
        }
    } free_all{{(void**)&d_a, (void**)&d_bp, (void**)&d_v, (void**)&d_h, (void**)&d_cum}};
    cudaStream_t st = e->stream;
    CUDA_TRY(cudaMalloc(&d_a, (nhw_a + 2) * 8));
    CUDA_TRY(cudaMalloc(&d_bp, (nhw + 2) * 8));
    CUDA_TRY(cudaMalloc(&d_v, nhw * 8));
    CUDA_TRY(cudaMalloc(&d_cum, (nhw + 1) * 4));
    CUDA_TRY(cudaMalloc(&d_h, na));
    CUDA_TRY(cudaMemcpyAsync(d_a, apl.data(), (nhw_a + 2) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_bp, bp.data(), (nhw + 2) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_v, vv.data(), nhw * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_h, h, na, cudaMemcpyHostToDevice, st));
    apa_block_kernel<<<1, 32, 0, st>>>(d_a, (int)na, d_bp, (int)nhw, d_v, d_cum, d_h);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(vv.data(), d_v, nhw * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h, d_h, na, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (uint64_t w = 0; w < nwords; w++) {
        v[2 * w] = (uint64_t)vv[2 * w].x | ((uint64_t)vv[2 * w + 1].x << 32);
        v[2 * w + 1] = (uint64_t)vv[2 * w].y | ((uint64_t)vv[2 * w + 1].y << 32);
    }
    int64_t total = 0;
    for (uint64_t i = 0; i < na; i++) total += h[i] == 1 ? 1 : (h[i] == 2 ? -1 : 0);
    *bottom_sum = total;
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ process-wide engines
// One engine per device, created on first use and shared by the single-pair drop-in symbols and apa_align_batch_multi. An
// engine has one stream and one work queue: concurrent callers of the same device take turns (per-engine mutex).
struct EngineSlot {
    std::mutex mu;
    apa_engine* eng = nullptr;
};
static std::mutex g_slots_mu;
static std::unordered_map<int, EngineSlot*> g_slots;

static EngineSlot* engine_slot(int device) {  // nullptr on failure (apa_last_error() says why)
    EngineSlot* sl;
    {
        std::lock_guard<std::mutex> lk(g_slots_mu);
        auto it = g_slots.find(device);
        if (it == g_slots.end()) it = g_slots.emplace(device, new EngineSlot()).first;
        sl = it->second;
    }
    std::lock_guard<std::mutex> lk(sl->mu);
    if (!sl->eng && apa_engine_create(device, &sl->eng) != APA_OK) return nullptr;
    return sl;
}

// The process-wide engine of a device (the one apa_align_batch_multi and the single-pair symbols use), for callers that want
// their resident batches on the same engine - and so on the same scratch arena - as those calls. Not to be destroyed.
extern "C" int apa_shared_engine(int device, apa_engine** out) {
    *out = nullptr;
    EngineSlot* sl = engine_slot(device);
    if (!sl) return APA_ERR_NO_DEVICE;
    *out = sl->eng;
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU batch call
// Pairs are independent (astarpa2/src/lib.rs:50-53 builds a fresh aligner per call): the batch is cut into contiguous shards
// balanced by bases, one per device, each aligned by that device's engine on its own host thread - no data-path collective
// (SURVEY 8e). Results come back in input order; the CIGAR texts of all shards share one page-locked pool.
extern "C" int apa_align_batch_multi(const int* devices, int n_devices, int preset, int trace, uint64_t n_pairs, const uint8_t* a_all,
                                     const int64_t* a_off, const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool,
                                     int64_t* cigar_off, int64_t* cigar_len, apa_batch_stats* stats) {
    if (cigar_pool) *cigar_pool = nullptr;
    if (!devices || n_devices < 1) return set_err(APA_ERR_BAD_INPUT, "apa_align_batch_multi: no devices");
    for (int d = 0; d < n_devices; d++)
        for (int d2 = 0; d2 < d; d2++)
            if (devices[d] == devices[d2]) return set_err(APA_ERR_BAD_INPUT, "apa_align_batch_multi: device listed twice");
    // shard bounds: the prefix of pairs whose bases reach (d + 1) / n_devices of the total
    std::vector<uint64_t> lo(n_devices + 1, n_pairs);
    lo[0] = 0;
    {
        const long double total = (long double)(a_off[n_pairs] - a_off[0]) + (long double)(b_off[n_pairs] - b_off[0]);
        uint64_t p = 0;
        for (int d = 1; d < n_devices; d++) {
            const long double target = total * d / n_devices;
            while (p < n_pairs && (long double)(a_off[p] - a_off[0]) + (long double)(b_off[p] - b_off[0]) < target) p++;
            lo[d] = p;
        }
    }
    struct Shard {
        EngineSlot* slot = nullptr;
        apa_batch* b = nullptr;
        int rc = APA_OK;
        std::string err;
        uint64_t pool_base = 0;
    };
    std::vector<Shard> sh(n_devices);
    for (int d = 0; d < n_devices; d++) {
        sh[d].slot = engine_slot(devices[d]);
        if (!sh[d].slot) return APA_ERR_NO_DEVICE;  // message set by apa_engine_create
    }
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int d = 0; d < n_devices; d++) locks.emplace_back(sh[d].slot->mu);
    auto for_each_shard = [&](auto fn) {
        std::vector<std::thread> th;
        for (int d = 1; d < n_devices; d++) th.emplace_back([&, d]() { fn(d); if (sh[d].rc != APA_OK) sh[d].err = g_last_error; });
        fn(0);
        if (sh[0].rc != APA_OK) sh[0].err = g_last_error;
        for (auto& t : th) t.join();
    };
    for_each_shard([&](int d) {  // phase 1: upload + run
        Shard& s = sh[d];
        apa_engine* e = s.slot->eng;
        s.rc = batch_prepare(e, lo[d + 1] - lo[d], a_all, a_off + lo[d], b_all, b_off + lo[d], true, &s.b);
        if (s.rc == APA_OK) s.rc = batch_run(e, s.b, preset, trace, true);
    });
    int rc = APA_OK;
    for (int d = 0; d < n_devices && rc == APA_OK; d++)
        if (sh[d].rc != APA_OK) rc = set_err(sh[d].rc, "device " + std::to_string(devices[d]) + ": " + sh[d].err);
    char* pool = nullptr;
    if (rc == APA_OK) {
        uint64_t total_pool = 0;
        for (int d = 0; d < n_devices; d++) {
            sh[d].pool_base = total_pool;
            total_pool += sh[d].b->pool_used;
        }
        const bool want_cigars = trace && cigar_pool && cigar_off && cigar_len;
        if (want_cigars) {
            pool = (char*)pinned_pool_get(std::max<uint64_t>(total_pool, 1));
            if (!pool) rc = set_err(APA_ERR_TOO_LARGE, "host allocation of the CIGAR pool failed");
        }
        if (rc == APA_OK) {
            for_each_shard([&](int d) {  // phase 2: download, every shard into its slice of the outputs
                Shard& s = sh[d];
                const uint64_t p0 = lo[d], np = lo[d + 1] - lo[d];
                s.rc = batch_download(s.slot->eng, s.b, costs + p0, nullptr, want_cigars ? cigar_off + p0 : nullptr,
                                      want_cigars ? cigar_len + p0 : nullptr, want_cigars ? pool + s.pool_base : nullptr);
                if (s.rc == APA_OK && want_cigars)
                    for (uint64_t p = 0; p < np; p++) cigar_off[p0 + p] += (int64_t)s.pool_base;
                if (s.rc == APA_OK && stats) stats[d] = s.b->stats;
            });
            for (int d = 0; d < n_devices && rc == APA_OK; d++)
                if (sh[d].rc != APA_OK) rc = set_err(sh[d].rc, "device " + std::to_string(devices[d]) + ": " + sh[d].err);
        }
    }
    for (int d = 0; d < n_devices; d++) apa_batch_free(sh[d].slot->eng, sh[d].b);
    if (rc != APA_OK) {
        if (pool) apa_free(pool);
        return rc;
    }
    if (cigar_pool) *cigar_pool = pool;
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ drop-in symbols
// The reference's single-pair entry points (astarpa-c/src/lib.rs:8-101) on the device APA_DEVICE names (default 0).
static int default_device() {
    const char* ev = getenv("APA_DEVICE");
    return ev ? atoi(ev) : 0;
}

static uint64_t align_one(int preset, const apa_params* params, const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len,
                          uint8_t** cigar_ptr, uintptr_t* cigar_len) {
    EngineSlot* sl = engine_slot(default_device());
    if (!sl) {
        // Same contract as the reference, whose failures are Rust panics that abort the process
        // (astarpa-c/src/lib.rs has no error path). There is no CPU fallback.
        fprintf(stderr, "libastarpa_c (B200): %s\n", apa_last_error());
        abort();
    }
    std::lock_guard<std::mutex> lk(sl->mu);  // one stream per engine: serialise concurrent callers
    int64_t a_off[2] = {0, (int64_t)a_len}, b_off[2] = {0, (int64_t)b_len};
    int64_t cost = -1, coff = 0, clen = 0;
    char* pool = nullptr;
    int rc = params ? apa_align_batch_params(sl->eng, params, 1, 1, a, a_off, b, b_off, &cost, &pool, &coff, &clen, nullptr)
                    : apa_align_batch(sl->eng, preset, 1, 1, a, a_off, b, b_off, &cost, &pool, &coff, &clen, nullptr);
    if (rc != APA_OK) {
        fprintf(stderr, "libastarpa_c (B200): %s\n", apa_last_error());
        abort();
    }
    uint8_t* out = (uint8_t*)malloc((size_t)clen + 1);
    memcpy(out, pool + coff, (size_t)clen);
    out[clen] = 0;
    apa_free(pool);
    *cigar_ptr = out;
    *cigar_len = (uintptr_t)clen;
    return (uint64_t)cost;
}

extern "C" uint64_t astarpa2_simple(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                                    uintptr_t* cigar_len) {
    return align_one(APA_PRESET_SIMPLE, nullptr, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
extern "C" uint64_t astarpa2_full(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                                  uintptr_t* cigar_len) {
    return align_one(APA_PRESET_FULL, nullptr, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
// A*PA v1 entry points (astarpa-c/src/lib.rs:54-95). The v1 engine (astar_dt, an A* over single states) is out of scope
// (SURVEY 8f row 2); these symbols are served by the A*PA2 block engine bounded by the heuristic the caller names: GCSH with
// exact matches of length k, no local pruning (MatchConfig::new(k, r), pa-heuristic/src/matches.rs:404-410), pruning by match
// start. The cost is the edit distance either way and the CIGAR is an optimal alignment; its tie-breaks are A*PA2's, not v1's.
// Arguments the heuristic cannot honour are refused loudly, never ignored: inexact matches (r = 2, matches/inexact.rs) and
// Prune::Both (prune_end) are not built, so `astarpa` (= astarpa_gcsh(r = 2, k = 15, false), lib.rs:54-64) runs GCSH(r = 1,
// k = 15) and says so once on stderr, and astarpa_gcsh aborts with a message for r != 1 or prune_end, as a reference panic would.
static uint64_t align_gcsh(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uintptr_t k, uint8_t** cigar_ptr,
                           uintptr_t* cigar_len) {
    apa_params q;
    apa_params_preset(APA_PRESET_FULL, &q);
    q.k = (int32_t)k;
    q.p = 0;
    return align_one(-1, &q, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
extern "C" uint64_t astarpa(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                            uintptr_t* cigar_len) {
    static std::atomic<bool> said{false};
    if (!said.exchange(true))
        fprintf(stderr, "libastarpa_c (B200): astarpa() = A*PA v1 with inexact matches (r = 2, k = 15) is served by the A*PA2 engine with "
                        "GCSH(r = 1, k = 15): same cost, an optimal CIGAR with A*PA2's tie-breaks\n");
    return align_gcsh(a, a_len, b, b_len, 15, cigar_ptr, cigar_len);
}
extern "C" uint64_t astarpa_gcsh(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uintptr_t r, uintptr_t k,
                                 bool prune_end, uint8_t** cigar_ptr, uintptr_t* cigar_len) {
    if (r != 1 || prune_end || k < 4 || k > 16) {
        fprintf(stderr, "libastarpa_c (B200): astarpa_gcsh(r = %zu, k = %zu, prune_end = %d): only r = 1 (exact matches), k = 4..16, "
                        "prune_end = false are built\n", (size_t)r, (size_t)k, (int)prune_end);
        abort();
    }
    return align_gcsh(a, a_len, b, b_len, k, cigar_ptr, cigar_len);
}
extern "C" void astarpa_free_cigar(uint8_t* cigar) { free(cigar); }
