// Shared device-side types and helpers of the B200 A*PA2 engine (sm_100a only).
//
// Everything here is written warp-per-pair: one warp owns one sequence pair. Scalar control state (ranges,
// costs, cursors) is kept warp-uniform — every lane computes the same value — so cooperative steps (the
// block-DP wavefront, popcount scans, ballots) can be called from anywhere without divergence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Phase timers (clock64 around the phases of a pair) cost registers; build with -DAPA_PHASE_TIMERS=1 to get
// apa_batch_stats.phase_cycles filled in (make TIMERS=1).
#ifndef APA_PHASE_TIMERS
#define APA_PHASE_TIMERS 0
#endif
#if APA_PHASE_TIMERS
#define APA_TIC() clock64()
#define APA_TOC(acc, t0) ((acc) += clock64() - (t0))
#else
#define APA_TIC() 0ll
#define APA_TOC(acc, t0) ((void)(t0))
#endif

// The device code is compiled twice: with APA_GENERAL=0 (namespace apa, apa_engine.cu) every AstarPa2 parameter is the
// compile-time constant of the two presets, which is what the hot path is tuned for; with APA_GENERAL=1 (namespace
// apa_gen, apa_general.cu) the same code reads them from RunParams at run time (other domains, heuristics, doubling
// types, block widths: astarpa2/src/params.rs, the configurations of astarpa2/src/tests.rs:19-119).
#ifndef APA_GENERAL
#define APA_GENERAL 0
#endif
#if APA_GENERAL
#define APA_NS apa_gen
#else
#define APA_NS apa
#endif

namespace APA_NS {

typedef int32_t I;
typedef int32_t Cost;

constexpr unsigned FULL = 0xffffffffu;
constexpr int WI = 64;   // reference word height (astarpa2/src/lib.rs:35). Ranges round to multiples of 64.
constexpr int HW = 32;   // our DP lane height: one 32-row half-word per lane (same recurrence, A.4 of SURVEY).
constexpr int BLOCK_W = 256;  // block_width of both presets (astarpa2/src/params.rs:84,111)

// Per-pair status codes written by the kernels.
enum Status : int32_t {
    ST_PENDING = 0,
    ST_DONE = 1,
    ST_OVERFLOW = 2,   // scratch arena too small: host re-runs the pair with a larger arena
    ST_BAD_INPUT = 3,  // byte outside ACGT
    ST_ASSERT = 4,     // a reference panic path was reached
    ST_TOO_LARGE = 5,  // a CIGAR run of 2^30 or more equal operations
};

struct JRange {
    I s, e;  // inclusive (astarpa2/src/ranges.rs:13-14)
};
__device__ __forceinline__ bool jr_empty(JRange r) { return r.s > r.e; }
__device__ __forceinline__ JRange jr_union(JRange a, JRange b) { return JRange{min(a.s, b.s), max(a.e, b.e)}; }
__device__ __forceinline__ JRange jr_inter(JRange a, JRange b) { return JRange{max(a.s, b.s), min(a.e, b.e)}; }
// Rust: s / 64 * 64 truncates toward zero; next_multiple_of rounds up (ranges.rs:71-76).
__device__ __forceinline__ I next_mult64(I x) {
    I r = x % WI;
    if (r < 0) r += WI;
    return r == 0 ? x : x + (WI - r);
}
__device__ __forceinline__ I trunc_mult64(I x) { return x / WI * WI; }
__device__ __forceinline__ JRange jr_round_out(JRange r) { return JRange{trunc_mult64(r.s), next_mult64(r.e)}; }
__device__ __forceinline__ JRange jr_round_in(JRange r) { return JRange{next_mult64(r.s), trunc_mult64(r.e)}; }
__device__ __forceinline__ I div_ceil_pos(I a, I b) {  // signed div_ceil, b > 0
    I q = a / b, r = a % b;
    return r > 0 ? q + 1 : q;
}

// Per-block metadata kept for the whole band-doubling search of one pair (astarpa2/src/block.rs:8-31).
struct BlkMeta {
    I orig_s, orig_e;  // original_j_range
    I js, je;          // rounded-out j_range (multiples of 64)
    I fs, fe;          // fixed_j_range (valid iff has_fixed)
    Cost top_val, bot_val;
    uint32_t v_off;     // byte offset of this block's V column inside the pair's arena (valid in the current pass)
    uint16_t has_fixed;
    uint16_t ones;      // column 0: all vertical deltas +1, nothing stored
    I col_s, col_e;     // i_range (left-exclusive)
    I j_h;              // Block.j_h (block.rs:30): row along which the pair's h row holds this block's horizontal deltas; J_H_NONE = None
    int32_t store_pass; // pass that wrote the V column at v_off: the column is still there during the NEXT pass only (two block stores)
    int32_t pad_[2];
};
static_assert(sizeof(BlkMeta) == 64, "BlkMeta layout");
constexpr I J_H_NONE = INT32_MIN;

// A read-only view of a stored right-edge column: V as (p,m) per 32-row half-word + running values.
struct BlkView {
    I js, je;            // rounded j_range
    Cost top_val, bot_val;
    const uint2* v;      // nhw entries
    const int32_t* cum;  // nhw+1 entries: value at row js + 32*hw
    int ones;
    __device__ __forceinline__ int nhw() const { return (je - js) >> 5; }
};

// Block::index (astarpa2/src/block.rs:69-122): value at row j >= js; rows past je extend with +1.
__device__ __forceinline__ Cost blk_index(const BlkView& b, I j) {
    if (j > b.je) return b.bot_val + (j - b.je);
    int off = j - b.js;
    int hw = off >> 5, bit = off & 31;
    if (b.ones) return b.top_val + off;
    Cost base = b.cum[hw];
    if (bit == 0) return base;
    uint2 pm = b.v[hw];
    uint32_t mask = (1u << bit) - 1u;
    return base + __popc(pm.x & mask) - __popc(pm.y & mask);
}
// Block::get_diff (block.rs:134-145): vertical delta from row j to j+1, or NONE outside the stored words.
constexpr int DIFF_NONE = 99;
__device__ __forceinline__ int blk_get_diff(const BlkView& b, I j) {
    if (j < b.js) return DIFF_NONE;
    int off = j - b.js;
    int hw = off >> 5;
    if (hw >= b.nhw()) return DIFF_NONE;
    if (b.ones) return 1;
    uint2 pm = b.v[hw];
    int bit = off & 31;
    return (int)((pm.x >> bit) & 1u) - (int)((pm.y >> bit) & 1u);
}

// RankTransform("ACGT") of an upper-case base: A0 C1 G2 T3 (pa-bitpacking/src/profile.rs:113).
__device__ __forceinline__ uint32_t rank_acgt(uint32_t c) {
    uint32_t x = (c >> 1) & 3u;  // A0 C1 T2 G3  (this is also QGrams::char_to_bits, qgrams.rs:30-32)
    return x ^ (x >> 1);         // -> A0 C1 G2 T3
}
__device__ __forceinline__ bool is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// K0 on the device: BitProfile::build (pa-bitpacking/src/profile.rs:112-133) for one sequence, by one warp (or by the warps
// of a grid, each taking every `stride`-th group of 32 half-words starting at `first`). raw: the bases as uploaded, any
// alignment; prof: nhw plane words (negated rank bits of A0 C1 G2 T3 per 32 bases, zero past the end of the sequence, padding
// half-words included). A group of 1 024 bases is fetched with 16-byte loads from the enclosing aligned words (ld.global.cg:
// the bytes may have been written by a copy engine while this kernel was already running, and L1 is not coherent with those
// writes) into `sbuf` (>= 1 040 bytes of this warp's shared memory, 16-byte aligned), then every lane turns the 32 bytes of
// its half-word into two plane words: 2-bit code (c >> 1) & 3 per byte, rank = code ^ (code >> 1), the four rank bits of a
// 32-bit word gathered by one multiply. Returns true (warp-uniform) if a byte outside ACGT was seen: the reference panics there
// (profile.rs:113).
__device__ __forceinline__ bool dev_pack_planes(const uint8_t* __restrict__ raw, I len, uint2* __restrict__ prof, int nhw, uint32_t* sbuf,
                                                int first = 0, int stride = 1) {
    const int lane = threadIdx.x & 31;
    const unsigned long long addr = (unsigned long long)raw;
    const uint32_t sh = (uint32_t)(addr & 15ull);
    const uint4* src = (const uint4*)(addr - sh);
    uint4* sb4 = (uint4*)sbuf;
    bool bad = false;
    for (int g = first; 32 * g < nhw; g += stride) {
        // bytes [1024 g, 1024 g + 1024) of the sequence = aligned 16-byte words 64 g .. 64 g + 64 (the last one only when sh > 0);
        // nothing past the end of the sequence is fetched (it may belong to a chunk that has not landed, or to no buffer at all)
        const long long last_word = ((long long)sh + len + 15) >> 4;  // exclusive
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int k = lane + 32 * r;
            const long long w = 64ll * g + k;
            if (k < 65 && w < last_word) sb4[k] = __ldcg(src + w);
        }
        __syncwarp();
        const int hw = 32 * g + lane;
        if (hw < nhw) {
            const I p0 = (I)hw * 32;
            const int nvalid = max(0, min(32, len - p0));
            uint32_t b0 = 0u, b1 = 0u, diff = 0u;
            if (nvalid > 0) {
                const uint32_t o = sh + 32u * (uint32_t)lane;
                const uint32_t* w32 = sbuf + (o >> 2);
                const uint32_t s8 = (o & 3u) * 8u;
                uint32_t lo = w32[0];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t hi = w32[j + 1];
                    const uint32_t x = __funnelshift_r(lo, hi, s8);  // bases 4 j .. 4 j + 3 of this half-word
                    lo = hi;
                    const uint32_t code = (x >> 1) & 0x03030303u;                  // A0 C1 T2 G3 per byte
                    const uint32_t rk = code ^ ((code >> 1) & 0x01010101u);        // A0 C1 G2 T3
                    b0 |= ((((rk & 0x01010101u) * 0x01020408u) >> 24) & 15u) << (4 * j);
                    b1 |= (((((rk >> 1) & 0x01010101u) * 0x01020408u) >> 24) & 15u) << (4 * j);
                    // the byte each code stands for, against the byte that is there (bytes past the end are not judged)
                    const uint32_t sel = (code & 3u) | ((code >> 4) & 0x30u) | ((code >> 8) & 0x300u) | ((code >> 12) & 0x3000u);
                    const int nb = min(4, nvalid - 4 * j);
                    const uint32_t bm = nb >= 4 ? ~0u : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
                    diff |= (__byte_perm(0x47544341u, 0u, sel) ^ x) & bm;
                }
            }
            const uint32_t vm = nvalid >= 32 ? ~0u : ((1u << nvalid) - 1u);
            prof[hw] = make_uint2(~b0 & vm, ~b1 & vm);
            bad |= diff != 0u;
        }
        __syncwarp();  // sbuf is overwritten by the next group
    }
    return __any_sync(FULL, bad);
}

// ---- 2-bit-plane packed sequences: prof[hw] = (negated rank bit0, negated rank bit1) of 32 consecutive bases.
// Every per-pair plane array carries two zero half-words of padding, so extract32 may read prof[hw + 1].
__device__ __forceinline__ uint2 extract32(const uint2* __restrict__ prof, I pos) {  // bits of bases [pos, pos + 32)
    int hw = pos >> 5, sh = pos & 31;
    uint2 lo = prof[hw], hi = prof[hw + 1];
    return make_uint2(__funnelshift_r(lo.x, hi.x, sh), __funnelshift_r(lo.y, hi.y, sh));
}
__device__ __forceinline__ uint2 extract32_end(const uint2* __restrict__ prof, I pos) {  // bit 31 = base pos - 1 (pos >= 1)
    if (pos >= 32) return extract32(prof, pos - 32);
    uint2 x = prof[0];
    int sh = 32 - pos;
    return make_uint2(x.x << sh, x.y << sh);
}
// extend_right (pa-heuristic/src/matches/prepruning.rs:25-32): advance (i, j) along the diagonal while a[i] == b[j],
// i < end_i, j < m; 32 bases per iteration.
__device__ __forceinline__ void extend_right_packed(const uint2* __restrict__ ap, const uint2* __restrict__ bp, I m, I& i, I j, I end_i) {
    for (;;) {
        int len = min(32, min(end_i - i, m - j));
        if (len <= 0) return;
        uint2 A = extract32(ap, i), B = extract32(bp, j);
        uint32_t mm = (A.x ^ B.x) | (A.y ^ B.y);
        if (len < 32) mm |= 0xffffffffu << len;
        int run = mm ? __ffs(mm) - 1 : 32;
        i += run;
        j += run;
        if (run < 32) return;
    }
}
// extend_left (astarpa2/src/blocks/trace.rs:443-451): step (i, j) back while a[i-1] == b[j-1], i > i0, j > 0.
__device__ __forceinline__ I extend_left_packed(const uint2* __restrict__ ap, const uint2* __restrict__ bp, I& i, I i0, I& j) {
    I cnt = 0;
    for (;;) {
        int len = min(32, min(i - i0, j));
        if (len <= 0) return cnt;
        uint2 A = extract32_end(ap, i), B = extract32_end(bp, j);
        uint32_t mm = (A.x ^ B.x) | (A.y ^ B.y);
        if (len < 32) mm |= 0x80000000u >> len;
        int run = mm ? __clz(mm) : 32;
        i -= run;
        j -= run;
        cnt += run;
        if (run < 32) return cnt;
    }
}

// CIGAR element: op in the top 2 bits, count in the low 30 (pa-types CigarElem; trace.rs:129-221).
enum CigOp : uint32_t { OP_MATCH = 0, OP_SUB = 1, OP_DEL = 2, OP_INS = 3 };
__device__ __forceinline__ uint32_t cig_pack(uint32_t op, uint32_t cnt) { return (op << 30) | cnt; }

}  // namespace APA_NS
