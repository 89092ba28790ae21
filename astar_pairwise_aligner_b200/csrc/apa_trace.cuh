// K3: on-device traceback for one pair (one warp). Replaces
//   Blocks::trace                 astarpa2/src/blocks/trace.rs:21-135
//   Blocks::parent                trace.rs:145-228
//   Blocks::dt_trace_block        trace.rs:231-416   (lanes = diagonals of the DT front)
//   Blocks::fill_with_blocks      astarpa2/src/blocks.rs:572-662  (block re-compute storing every column)
//   extend_left{,_simd}           trace.rs:443-500
//   Cigar::push_elem / to_string  pa-types (external; format pinned by astarpa-c/example.cpp:16)
// The sparse block store of the final pass (right-edge V column per 256-column block) is walked backwards;
// each block is first tried with the greedy diagonal-transition trace and otherwise re-filled.
#pragma once
#include "apa_align.cuh"

namespace APA_NS {

constexpr int DT_CACHE_ELEMS = (DT_MAX_G + 1) * (DT_MAX_G + 1);

// Shared-memory accesses of the DT fronts through an explicit 32-bit shared address held in a register: under the
// 48-register cap the compiler otherwise rebuilds the address of the warp's shared block (5 instructions) at every access.
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i32(uint32_t addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

struct CigarWriter {  // elements are pushed newest-first (the path is walked from the end); see emit_cigar_text.
    uint8_t* arena;
    uint32_t arena_size; // the elements grow down from here: PairCtx::cig_top (the arena end, or the bottom of block store B)
    uint32_t count;      // elements already stored in the arena
    uint32_t pend_op;    // pending (mergeable) element
    uint32_t pend_cnt;   // 0 = none
    uint32_t nbuf;       // finished elements held in registers: element k of the buffer sits in lane k's `buf`
    uint32_t buf;
};

// Write the buffered elements to the arena: element (count + k) goes to arena_end - 4 (count + k + 1), one coalesced store.
__device__ __forceinline__ void cig_drain(PairCtx& cx, CigarWriter& cw) {
    if (cw.nbuf == 0) return;
    const uint64_t need64 = ((uint64_t)cw.count + cw.nbuf) * 4u;  // 64-bit: must not wrap past the arena size
    const uint32_t need = (uint32_t)need64;
    if (need64 > cw.arena_size || cw.arena_size - need < cx.v_top) {
        cx.status = ST_OVERFLOW;
        cw.nbuf = 0;
        return;
    }
    const uint32_t lane = threadIdx.x & 31;
    if (lane < cw.nbuf) *(uint32_t*)(cw.arena + cw.arena_size - 4u * (cw.count + lane + 1u)) = cw.buf;
    cw.count += cw.nbuf;
    cw.nbuf = 0;
    cx.hi_bot = cw.arena_size - need;
}
__device__ __forceinline__ void cig_store(PairCtx& cx, CigarWriter& cw, uint32_t op, uint32_t cnt) {
    // cig_pack keeps 30 bits of count: pend_cnt is capped below 2^30 by cig_push, which starts a new element instead
    if ((threadIdx.x & 31) == cw.nbuf) cw.buf = cig_pack(op, cnt);
    cw.nbuf++;
    if (cw.nbuf == 32) cig_drain(cx, cw);
}
// Cigar::push_elem: merge with the previous element when the op is the same.
__device__ __forceinline__ void cig_push(PairCtx& cx, CigarWriter& cw, uint32_t op, uint32_t cnt) {
    if (cw.pend_cnt != 0 && cw.pend_op == op) {
        if ((uint64_t)cw.pend_cnt + cnt >= (1ull << 30)) cx.status = ST_TOO_LARGE;  // an element keeps 30 bits of count (cig_pack)
        cw.pend_cnt += cnt;
        return;
    }
    if (cw.pend_cnt != 0) cig_store(cx, cw, cw.pend_op, cw.pend_cnt);
    cw.pend_op = op;
    cw.pend_cnt = cnt;
}
__device__ __forceinline__ void cig_flush(PairCtx& cx, CigarWriter& cw) {
    if (cw.pend_cnt != 0) cig_store(cx, cw, cw.pend_op, cw.pend_cnt);
    cw.pend_cnt = 0;
    cig_drain(cx, cw);
}

// A column of the dense (re-filled) region or a stored sparse block, as seen by parent().
struct ColRef {
    BlkView v;       // for fill columns: js/je/top_val set, v points at the column's words, cum == nullptr
    bool is_fill;
};

// Block::index on a fill column: popcount walk done cooperatively.
__device__ __forceinline__ Cost fill_index(const ColRef& c, I j) {
    const int lane = threadIdx.x & 31;
    const int nhw = (c.v.je - c.v.js) >> 5;
    I jj = min(j, c.v.je);
    int off = jj - c.v.js;
    int hwj = off >> 5, bit = off & 31;
    int sum = 0;
    for (int hw = lane; hw < nhw && hw <= hwj; hw += 32) {
        uint2 pm = c.v.v[hw];
        if (hw < hwj)
            sum += __popc(pm.x) - __popc(pm.y);
        else if (bit) {
            uint32_t mask = (1u << bit) - 1u;
            sum += __popc(pm.x & mask) - __popc(pm.y & mask);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
    return c.v.top_val + sum + (j - jj);
}
__device__ __forceinline__ Cost col_index(const ColRef& c, I j) { return c.is_fill ? fill_index(c, j) : blk_index(c.v, j); }
__device__ __forceinline__ int col_get_diff(const ColRef& c, I j) { return blk_get_diff(c.v, j); }

struct TraceState {
    I ti, tj;
    Cost g;
    int top;       // topmost stored (sparse) block still on the stack
    bool fill;     // a dense region sits on top of block `top`
    I fis;         // its left column (== meta[top].col_e)
    int fcols;     // dense columns still on the stack: fis+1 .. fis+fcols
    I fjs, fje;    // rounded row range of the dense region
    Cost ftop0;    // value at (fis, fjs)
    uint2* fvals;  // fvals[(c-1) * fnhw + hw], c = 1..
};

__device__ __forceinline__ ColRef fill_col(const TraceState& ts, int c) {
    ColRef r;
    const int fnhw = (ts.fje - ts.fjs) >> 5;
    r.is_fill = true;
    r.v.js = ts.fjs;
    r.v.je = ts.fje;
    r.v.top_val = ts.ftop0 + c;
    r.v.bot_val = 0;
    r.v.v = ts.fvals + (size_t)(c - 1) * fnhw;
    r.v.cum = nullptr;
    r.v.ones = 0;
    return r;
}

// dt_trace_block (trace.rs:231-416). Returns true on success and updates ts.(ti,tj,g).
// Lanes = diagonals. The fronts (column reached per diagonal) of the current and previous level live in shared
// memory; the per-element record (extension length, parent diagonal) needed only for the final backtrack goes to
// `rec` in the arena: element (g,d) at g*g+g+d holds (ext << 2) | (parent_d + 1).
__device__ bool dev_dt_trace(PairCtx& cx, WarpSmem& sm, CigarWriter& cw, TraceState& ts, const BlkView& prev, I block_start,
                             int32_t* rec) {
    const int lane = threadIdx.x & 31;
    const I si = ts.ti, sj = ts.tj;
    const Cost g_st = ts.g;
    auto idx = [](int g, int d) { return g * g + g + d; };
    int8_t* chain = (int8_t*)sm.hrow;  // d of each level on the final path
    constexpr int DOFF = 48;
    // byte address (shared window) of front element d = 0 of level parity 0; parity 1 is 96 ints further
    const uint32_t dt0 = __shfl_sync(FULL, (uint32_t)__cvta_generic_to_shared(&sm.dt_i[0][DOFF]), 0);

    // reference closure extend_left_simd_and_check (trace.rs:313-329)
    auto reached = [&](I i, I j, Cost target) -> bool {
        if (i != block_start) return false;
        if (j < prev.js || j > prev.je) return false;  // Block::get -> None
        return blk_index(prev, j) == target;
    };
    int g = 0, d_lo = 0, d_hi = 0;
    int found_d = 0;
    bool found = false;
    {
        I i = si, j = sj;
        I ext = extend_left_packed(cx.aprof, cx.bprof, i, block_start, j);
        if (lane == 0) {
            sts_i32(dt0, i);
            rec[0] = (ext << 2) | 1;
        }
        __syncwarp();
        if (reached(i, j, g_st)) found = true;
    }
    while (!found) {
        const int ng = g + 1;
        const uint32_t cur = dt0 + (uint32_t)(g & 1) * 384u;  // front of level g: element d at cur + 4 d
        const uint32_t nxt = dt0 + (uint32_t)(ng & 1) * 384u;
        // expand + extend level ng, one diagonal per lane
        I min_fr = INT32_MAX, min_i = INT32_MAX;
        int succ_d = INT32_MAX;
        for (int base = d_lo - 1; base <= d_hi + 1; base += 32) {
            const int d = base + lane;
            I best = INT32_MAX;
            int pd = 0;
            bool in = d <= d_hi + 1;
            if (in) {
                if (d - 1 >= d_lo && d - 1 <= d_hi) {  // from d-1: (fr.i, -1) insertion
                    I y = lds_i32(cur + 4 * (d - 1));
                    if (y < best) best = y, pd = -1;
                }
                if (d >= d_lo && d <= d_hi) {  // from d: (fr.i - 1, 0) substitution
                    I y = lds_i32(cur + 4 * d) - 1;
                    if (y < best) best = y, pd = 0;
                }
                if (d + 1 >= d_lo && d + 1 <= d_hi) {  // from d+1: (fr.i - 1, +1) deletion
                    I y = lds_i32(cur + 4 * (d + 1)) - 1;
                    if (y < best) best = y, pd = 1;
                }
            }
            bool ok = false;
            if (in) {
                I i = best, ext = 0;
                if (best != INT32_MAX) {
                    I j = sj - (si - i) - d;
                    ext = extend_left_packed(cx.aprof, cx.bprof, i, block_start, j);
                    ok = reached(i, j, g_st - ng);
                    min_fr = min(min_fr, (I)(2u * (uint32_t)i - (uint32_t)d));
                    min_i = min(min_i, i);
                }
                sts_i32(nxt + 4 * d, i);
                rec[idx(ng, d)] = (ext << 2) | (pd + 1);
            }
            unsigned bal = __ballot_sync(FULL, ok);
            if (bal && succ_d == INT32_MAX) succ_d = base + (__ffs(bal) - 1);
        }
        __syncwarp();
        g = ng;
        d_lo -= 1;
        d_hi += 1;
        if (succ_d != INT32_MAX) {
            found = true;
            found_d = succ_d;
            break;
        }
        min_fr = __reduce_min_sync(FULL, min_fr);
        min_i = __reduce_min_sync(FULL, min_i);
        if (g == P_MAX_G(cx) / 2 && min_i > (block_start + si) / 2) return false;
        if (g == P_MAX_G(cx)) return false;
        // x-drop: shrink diagonals more than fr_drop behind (trace.rs:396-414)
        if (P_FR_DROP(cx) > 0) {
            const I thr = (I)((uint32_t)min_fr + (uint32_t)P_FR_DROP(cx));
            while (d_lo < d_hi) {
                I i = lds_i32(nxt + 4 * d_lo);
                if (i <= block_start || (I)(2u * (uint32_t)i - (uint32_t)d_lo) > thr)
                    d_lo++;
                else
                    break;
            }
            while (d_lo < d_hi) {
                I i = lds_i32(nxt + 4 * d_hi);
                if (i <= block_start || (I)(2u * (uint32_t)i - (uint32_t)d_hi) > thr)
                    d_hi--;
                else
                    break;
            }
        }
        if (d_lo > d_hi) return false;
    }
    // inner fn trace() (trace.rs:266-308): emit ops from st towards block_start.
    {
        int d = found_d;
        __syncwarp();
        if (lane == 0) {
            int dd = d;
            for (int l = g; l >= 0; l--) {
                chain[l] = (int8_t)dd;
                if (l > 0) dd += (rec[idx(l, dd)] & 3) - 1;
            }
        }
        __syncwarp();
        for (int l = 0; l <= g; l++) {
            int dl = chain[l];
            int e = rec[idx(l, dl)];
            if (l > 0) {
                int pd = (e & 3) - 1;
                cig_push(cx, cw, pd == -1 ? OP_INS : (pd == 0 ? OP_SUB : OP_DEL), 1);
            }
            I ext = e >> 2;
            if (ext > 0) cig_push(cx, cw, OP_MATCH, (uint32_t)ext);
        }
        __syncwarp();
        ts.ti = block_start;
        ts.tj = sj - (si - block_start) - d;
        ts.g = g_st - g;
    }
    return true;
}

// Blocks::trace (trace.rs:21-135). On entry the final pass's blocks are in meta[0..nblk].
// Produces the CIGAR elements (newest first) in the arena; returns false on error (cx.status set).
__device__ bool dev_trace(PairCtx& cx, WarpSmem& sm, CigarWriter& cw, Cost cost) {
    const int lane = threadIdx.x & 31;
    TraceState ts;
    ts.ti = cx.n;
    ts.tj = cx.m;
    ts.g = cost;
    ts.top = cx.nblk;
    ts.fill = false;
    ts.fcols = 0;
    ts.fvals = nullptr;
    ts.fis = ts.fjs = ts.fje = 0;
    ts.ftop0 = 0;
    // DT cache + dense region live above the V column store of the final pass.
    const uint32_t trace_base = cx.v_top;
    uint32_t cache_off = arena_alloc(cx, DT_CACHE_ELEMS * 4u);
    if (cx.status != ST_PENDING) return false;
    int32_t* cache = (int32_t*)(cx.arena + cache_off);
    const uint32_t fill_base = cx.v_top;
    (void)trace_base;

    while (!(ts.ti == 0 && ts.tj == 0)) {
        // Remove blocks to the right of `to` (trace.rs:45-47).
        if (ts.fill) {
            int keep = ts.ti - ts.fis;
            if (keep < ts.fcols) ts.fcols = keep;
            if (ts.fcols <= 0) {
                ts.fill = false;
                ts.fcols = 0;
            }
        }
        if (!ts.fill) {
            while (ts.top > 0 && cx.meta[ts.top].col_s >= ts.ti) ts.top--;
        }
        // DT trace first (trace.rs:50-65).
        if (P_DT_TRACE(cx) && ts.ti > 0 && !ts.fill) {
            const BlkMeta pm = cx.meta[ts.top - 1];
            if (pm.col_e < ts.ti - 1) {
                const BlkView prev = view_of(cx, pm);
                long long t_dt0 = APA_TIC();
                bool dt_ok = dev_dt_trace(cx, sm, cw, ts, prev, pm.col_e, cache);
                APA_TOC(cx.tphase[4], t_dt0);
                if (dt_ok) {
                    cx.dt_blocks++;
                    if (cx.status != ST_PENDING) return false;
                    continue;
                }
            }
        }
        // DP based traceback: re-fill the block when needed (trace.rs:69-125).
        if (ts.ti > 0 && !ts.fill) {
            const BlkMeta bm = cx.meta[ts.top];
            const BlkMeta pm = cx.meta[ts.top - 1];
            if (!(pm.col_e < ts.ti && ts.ti <= bm.col_e)) {
                cx.status = ST_ASSERT;
                return false;
            }
            if (pm.col_e < ts.ti - 1 || bm.col_e > ts.ti) {
                const BlkView prev = view_of(cx, pm);
                const I is = pm.col_e, ie = ts.ti;
                const I jr_s = bm.js, jr_e = ts.tj;
                ts.top -= 1;  // pop_last_block
                I height = min(jr_e - jr_s, (ie - is) * 5 / 4);
                stage_amask(sm, cx.aprof, is, ie - is, lane);
                for (;;) {
                    JRange r = jr_round_out(JRange{max(jr_e - height, pm.js), jr_e});
                    const int fnhw = (r.e - r.s) >> 5;
                    cx.v_top = fill_base;
                    uint32_t voff = arena_alloc(cx, (uint32_t)fnhw * 8u + (uint32_t)(fnhw + 1) * 4u);
                    const uint64_t fbytes = (uint64_t)(ie - is) * (uint64_t)fnhw * 8u;  // 64-bit: (cols x half-words) may exceed 2^32
                    if (fbytes > 0xF0000000ull) cx.status = ST_OVERFLOW;
                    uint32_t foff = (cx.status == ST_PENDING) ? arena_alloc(cx, (uint32_t)fbytes) : 0u;
                    if (cx.status != ST_PENDING) return false;
                    ts.fjs = r.s;
                    ts.fje = r.e;
                    ts.fis = is;
                    ts.fcols = ie - is;
                    ts.ftop0 = blk_index(prev, r.s);
                    ts.fvals = (uint2*)(cx.arena + foff);
                    block_dp<true>(sm, cx.bprof, prev, ie - is, r.s, r.e, (uint2*)(cx.arena + voff),
                                   (int32_t*)(cx.arena + voff + (size_t)fnhw * 8), ts.ftop0 + (ie - is), ts.fvals, cx.dpc);
                    cx.fill_blocks++;
                    ColRef lastc = fill_col(ts, ts.fcols);
                    if (fill_index(lastc, ts.tj) == ts.g) break;
                    if (r.s == 0) {  // "No trace found through block"
                        cx.status = ST_ASSERT;
                        return false;
                    }
                    height *= 2;
                }
                ts.fill = true;
            }
        }
        // parent() (trace.rs:145-228)
        {
            ColRef block, prevc;
            bool have_prev = true;
            if (ts.fill) {
                block = fill_col(ts, ts.fcols);
                if (ts.fcols >= 2) {
                    prevc = fill_col(ts, ts.fcols - 1);
                } else {
                    prevc.is_fill = false;
                    prevc.v = view_of(cx, cx.meta[ts.top]);
                }
            } else {
                block.is_fill = false;
                block.v = view_of(cx, cx.meta[ts.top]);
                if (ts.top >= 1) {
                    prevc.is_fill = false;
                    prevc.v = view_of(cx, cx.meta[ts.top - 1]);
                } else {
                    have_prev = false;
                }
            }
            // Greedy matching.
            I cnt = extend_left_packed(cx.aprof, cx.bprof, ts.ti, 0, ts.tj);
            if (cnt > 0) {
                cig_push(cx, cw, OP_MATCH, (uint32_t)cnt);
            } else {
                int vd = col_get_diff(block, ts.tj - 1);
                if (vd == 1) {
                    ts.g -= 1;
                    ts.tj -= 1;
                    cig_push(cx, cw, OP_INS, 1);
                } else {
                    if (!have_prev) {
                        cx.status = ST_ASSERT;
                        return false;
                    }
                    Cost hd = ts.tj < prevc.v.js ? 1 : ts.g - col_index(prevc, ts.tj);
                    if (hd == 1) {
                        ts.g -= 1;
                        ts.ti -= 1;
                        cig_push(cx, cw, OP_DEL, 1);
                    } else {
                        Cost dd;
                        if (ts.tj > prevc.v.je) {
                            dd = 1;
                        } else {
                            int pdv = col_get_diff(prevc, ts.tj - 1);
                            if (pdv == DIFF_NONE) {
                                cx.status = ST_ASSERT;
                                return false;
                            }
                            dd = pdv + hd;
                        }
                        if (dd == 1) {
                            ts.g -= 1;
                            ts.ti -= 1;
                            ts.tj -= 1;
                            cig_push(cx, cw, OP_SUB, 1);
                        } else {
                            cx.status = ST_ASSERT;  // "PARENT NOT FOUND IN TRACEBACK"
                            return false;
                        }
                    }
                }
            }
            if (cx.status != ST_PENDING) return false;
        }
    }
    if (ts.g != 0) {
        cx.status = ST_ASSERT;
        return false;
    }
    cig_flush(cx, cw);
    return cx.status == ST_PENDING;
}

// Cigar::reverse + to_string: the stored elements are newest-first, i.e. ascending addresses from
// arena_end - 4*count are already in forward order. Text: count omitted when 1, ops '=' 'X' 'D' 'I'.
__device__ __forceinline__ uint32_t ndigits(uint32_t x) {
    uint32_t n = 1;
    while (x >= 10) {
        x /= 10;
        n++;
    }
    return n;
}
__device__ long long emit_cigar_text(const CigarWriter& cw, char* pool, unsigned long long* pool_cursor, unsigned long long pool_cap,
                                     long long* out_len) {
    const int lane = threadIdx.x & 31;
    const uint32_t* elems = (const uint32_t*)(cw.arena + cw.arena_size - 4u * cw.count);
    __syncwarp();  // the elements were stored by lane 0 (cig_store); every lane reads them below
    // pass 1: total length
    unsigned long long total = 0;
    for (uint32_t k = lane; k < cw.count; k += 32) {
        uint32_t cnt = elems[k] & 0x3fffffffu;
        total += 1 + (cnt != 1 ? ndigits(cnt) : 0);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(pool_cursor, total + 1);
    off = __shfl_sync(FULL, off, 0);
    *out_len = (long long)total;
    if (off + total + 1 > pool_cap) return -1;
    // pass 2: write
    unsigned long long base = off;
    const char opc[4] = {'=', 'X', 'D', 'I'};
    for (uint32_t k0 = 0; k0 < cw.count; k0 += 32) {
        uint32_t k = k0 + lane;
        uint32_t len = 0, cnt = 0, op = 0;
        if (k < cw.count) {
            uint32_t e = elems[k];
            cnt = e & 0x3fffffffu;
            op = e >> 30;
            len = 1 + (cnt != 1 ? ndigits(cnt) : 0);
        }
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += y;
        }
        if (k < cw.count) {
            char* p = pool + base + (incl - len);
            uint32_t nd = len - 1;
            uint32_t x = cnt;
            for (uint32_t t = 0; t < nd; t++) {
                p[nd - 1 - t] = (char)('0' + x % 10);
                x /= 10;
            }
            p[nd] = opc[op];
        }
        base += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) pool[off + total] = 0;
    return (long long)off;
}

}  // namespace APA_NS
