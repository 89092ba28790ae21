// K2: the gap-cost chaining seed heuristic (GCSH, r = 1 exact matches, k = 12, local pruning p = 14,
// prune-by-start) built and queried on the GPU, one warp per pair. Replaces the pa-heuristic slice used by
// astarpa2_full (astarpa2/src/params.rs:98-128):
//   Seeds (potential, seed_at, transform)                     pa-heuristic/src/seeds.rs:20-157   -> closed forms
//   QGrams + exact::hash_a / hash_to_smallvec                 matches/qgrams.rs:7-110, matches/exact.rs:15-69
//   MatchBuilder::push (transform filter, local pruning)      matches.rs:205-247
//   preserve_for_local_pruning, extend_right{,_simd}          matches/prepruning.rs:25-203  (lanes = DT diagonals)
//   MatchBuilder::sort/finish                                 matches.rs:249-332            (counting sort by seed)
//   MatchPruner::{new, prune_block}, ActiveRange              prune.rs:97-292
//   CSHI::{new, h, h_with_hint, distance, prune_block, update_contours}   heuristic/csh.rs:152-554
//   HintContours::{new, score, score_with_hint, update_layers} contour/hint_contours.rs:125-637
//
// Contours: in the A*PA2 call pattern (update_contours(Pos(0,0)) at the start of every pass, last_change =
// Layer::MAX, right_of = 0; domain.rs:365-370, csh.rs:521-547) HintContours::update_layers re-scores every layer
// from the lowest one ever touched by a pruned match up to the top, i.e. it leaves exactly the layering a fresh
// build over the still-active matches produces, and score_with_hint returns score() whatever the hint. So the GPU
// keeps a flat layer -> point-list structure, rebuilds it at the start of a pass when matches were pruned, and
// answers score(q) (highest layer holding a point >= q) with a 32-layers-per-probe search. The oracle implements
// the reference's literal data structure; the parity tests compare the resulting bands and CIGARs.
#pragma once
#include "apa_align.cuh"

// Build-kernel variants (A/B measurements on B200, profiles/README.md round 2); 1 = the shipped formulation.
#ifndef APA_LAYER_CACHE
#define APA_LAYER_CACHE 1  // contour build with the top 32 layers in registers
#endif
#ifndef APA_DT_V2
#define APA_DT_V2 1        // local pruning: per-lane nm entry loaded once, single-compare potential test
#endif
#ifndef APA_NM_LAZY
#define APA_NM_LAZY 0      // next_match_per_diag uninitialised behind a segment bitmap (less DRAM traffic, measured 2.2 ms slower: off)
#endif
#ifndef APA_TAB_DENSE
#define APA_TAB_DENSE 0    // seed table at load factor <= 0.63 instead of <= 0.5
#endif
#ifndef APA_BLOOM2
#define APA_BLOOM2 1       // blocked Bloom filter: two bits per key inside one 32-bit word, instead of one bit
#endif
#ifndef APA_BLOOM_LOG
#define APA_BLOOM_LOG 2    // filter bits per table slot = 2^APA_BLOOM_LOG
#endif
#ifndef APA_CAS4
#define APA_CAS4 0         // four seed insertions in flight per lane
#endif
#ifndef APA_PROBE2
#define APA_PROBE2 1       // two table probes in flight per lane in the window scan
#endif

namespace APA_NS {

constexpr int GCSH_K = 12;        // seed length (params.rs:103)
constexpr int GCSH_P = 14;        // local-pruning look-ahead in seeds (params.rs:105)
constexpr uint32_t HT_EMPTY = 0xffffffffu;

struct GcshH {
    static constexpr bool PRUNE = true;
#if APA_GENERAL
    int k_, p_;  // MatchConfig.length / local_pruning (pa-heuristic/src/matches.rs:388-423)
    __device__ __forceinline__ int K() const { return k_; }
    __device__ __forceinline__ int P() const { return p_; }
#else
    static __device__ __forceinline__ constexpr int K() { return GCSH_K; }
    static __device__ __forceinline__ constexpr int P() { return GCSH_P; }
#endif
    I n, m;
    I nseeds;
    I ttx, tty;  // transform(target)
    // matches sorted by (start.i, start.j)  == MatchPruner.by_start
    int M;
    const I* ms_i;
    const I* ms_j;
    I* px;             // transform(start).0
    I* py;             // transform(start).1
    uint8_t* active;   // MatchStatus::Active
    int* next;         // next point in the same layer's overflow list
    int* layer_head;   // [0 .. M+1] overflow list (third and later points of a layer); -1 = none
    int4* layer_pts;   // [0 .. M+1] first two points of each layer inline: (x0, y0, x1, y1); x = INT32_MIN when unused
    int nlayers;       // highest non-empty layer
    const uint32_t* base;  // [nseeds+1] first match of each seed in by_start order
    uint32_t* before_end;  // ActiveRange.before.end per seed
    uint32_t* after_start; // ActiveRange.after.start per seed (HT_EMPTY = not split yet)
    // Search start of score(), per call-site stream (never changes a result): the last answer, the coordinate of the
    // last query that bound it, and the layers-per-unit density used to extrapolate from one query to the next.
    // Streams: 0 band end (j_range), 1 fixed-range start, 2 fixed-range end, 3 plain (contour build, h0).
    int hint[4];
    int hq[4];
    int dens;          // nlayers * 256 / nseeds: a layer is one chained match, about one per (1 / match rate) seeds
    bool dirty;
    unsigned long long h_calls;
    unsigned long long probes;  // 32-layer probe rounds of score() (stats: score_probes / score_calls = rounds per query)
    long long t_h;

    // Seeds::potential (seeds.rs:79-81) for fixed-length seeds at 0, k, 2k, ...: number of seeds starting at >= i.
    __device__ __forceinline__ Cost pot(I i) const {
        I c = (i + K() - 1) / K();
        return c >= nseeds ? 0 : nseeds - c;
    }
    // RotateToFrontContour::contains on layer w (rotate_to_front.rs:32-44), without the rotation.
    __device__ __forceinline__ bool contains(int w, I qx, I qy) const {
        const int4 p = layer_pts[w];  // one 16-byte load answers most probes (~1.5 points per layer)
        if (qx <= p.x && qy <= p.y) return true;
        if (p.z == INT32_MIN) return false;
        if (qx <= p.z && qy <= p.w) return true;
        for (int idx = layer_head[w]; idx >= 0; idx = next[idx])
            if (qx <= px[idx] && qy <= py[idx]) return true;
        return false;
    }
    // HintContours::score (hint_contours.rs:258-272): highest layer containing a point >= q. Layers are monotone
    // (a point in layer w is dominated by one in layer w-1), so the warp probes 32 layers per step: first the 32
    // consecutive layers around the previous answer of the same call site (`slot`: the band code asks in three
    // streams - band end, fixed-range start, fixed-range end - each of which moves smoothly from block to block),
    // then, on a miss, strided probes away from that window (stride 1, 16, 256, ...) until the answer is bracketed,
    // then 32-ary refinement. The result does not depend on the hints.
    // (kept inlined at its call sites: a __noinline__ score() shrinks the pass kernel by 15 % but measured 5 % slower)
    __device__ int score(I qx, I qy, int slot) {
        const int lane = threadIdx.x & 31;
        if (nlayers == 0) return 0;
        int lo = 0;            // known: contained in layer lo (layer 0 holds (MAX, MAX))
        int hi = nlayers + 1;  // known: not contained in layer hi
        bool down;
        {
            // Along the chain, layer(q) ~ dens * (K - q_bind): the transformed coordinate that binds is y below the
            // main diagonal (band end) and x above it (band start). Extrapolate from the previous query of the stream.
            int pred = hint[slot];
            if (slot != 3) {
                const int qb = slot == 1 ? qx : qy;
                const int dq = max(-65536, min(65536, hq[slot] - qb));
                pred += (dq * dens) >> 8;
                hq[slot] = qb;
            }
            const int basew = max(1, min(pred - 15, nlayers - 31));
            probes++;
            const int w = basew + lane;
            const bool c = (w <= nlayers) && contains(w, qx, qy);
            const int cnt = __popc(__ballot_sync(FULL, c));
            const int last = min(basew + 31, nlayers);
            down = cnt == 0;
            if (down) {
                hi = basew;
            } else {
                lo = basew + cnt - 1;
                if (lo < last) hi = lo + 1;
            }
        }
        int cap = 1;
        while (hi - lo > 1) {
            const int stride = min((hi - lo + 30) >> 5, cap);  // ceil(candidates / 32), capped while galloping
            cap <<= 4;
            probes++;
            if (down) {  // probes hi - stride, hi - 2 stride, ...: the contained ones are the far (high-lane) end
                const int w = hi - (lane + 1) * stride;
                const bool c = (w <= lo) || contains(w, qx, qy);
                const unsigned bal = __ballot_sync(FULL, c);
                const int f = bal ? __ffs(bal) - 1 : 32;  // lanes < f: not contained
                const int nhi = hi - f * stride;
                if (f < 32) lo = max(lo, hi - (f + 1) * stride);
                hi = nhi;
            } else {  // probes lo + stride, lo + 2 stride, ...: the contained ones are the near (low-lane) end
                const int w = lo + (lane + 1) * stride;
                const bool c = (w < hi) && contains(w, qx, qy);
                const int cnt = __popc(__ballot_sync(FULL, c));
                const int nlo = lo + cnt * stride;
                if (cnt < 32) hi = min(hi, lo + (cnt + 1) * stride);
                lo = nlo;
            }
        }
        hint[slot] = lo;
        return lo;
    }
    // CSHI::h / h_with_hint (csh.rs:341-376): P(u) - layer(T(u)), or max(gap, potential) to the target in layer 0.
    __device__ Cost h(I i, I j, int slot = 3) {
        h_calls++;
        long long t0 = APA_TIC();
        Cost p = pot(i);
        int val = score(i - j - p, j - i - p, slot);
        Cost r;
        if (val == 0) {
            I d = (n - i) - (m - j);
            Cost gap = d < 0 ? -d : d;
            r = max(gap, p);  // potential_distance(pos, target) = P(pos) (seeds.rs:84-88: no seed covers i = n)
        } else {
            r = p - val;
        }
        APA_TOC(t_h, t0);
        return r;
    }
#if APA_LAYER_CACHE
    // HintContours::new over the active arrows (hint_contours.rs:213-255) == state after update_layers.
    // Matches are taken from the last to the first; each asks score(end) and joins layer score + 1. Along a chain almost every
    // query is answered by one of the top few layers, so the warp keeps the TOP 32 LAYERS IN REGISTERS (lane l: the two inline
    // points of layer nlayers - l and whether it has an overflow list): a query is a compare + ballot with no memory round
    // trip, an insertion updates the owning lane and writes through to layer_pts / layer_head so that the table in memory is
    // always what the plain algorithm would have left. Queries that no cached layer contains fall back to score() on memory.
    __device__ void build_layers() {
        const int lane = threadIdx.x & 31;
        for (int w = lane; w <= nlayers + 1 && w <= M + 1; w += 32) {
            layer_head[w] = -1;
            layer_pts[w] = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
        }
        __syncwarp();
        nlayers = 0;
        hint[3] = 0;
        int4 cp = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);  // cached layer nlayers - lane (valid iff that is >= 1)
        bool covf = false;                                                 // ... has an overflow list
        for (int base = M - 1; base >= 0; base -= 32) {
            // the next 32 matches (indices base, base - 1, ...): one coalesced load each of active / px / py
            const int my = base - lane;
            I mx = 0, my_y = 0;
            bool mok = false;
            if (my >= 0 && active[my]) {
                mx = px[my];
                my_y = py[my];
                mok = (mx + 1 <= ttx) && (my_y + 1 <= tty);  // transform(end) = transform(start) + (1, 1): P(end.i) = P(start.i) - 1
            }
            unsigned todo = __ballot_sync(FULL, mok);
            while (todo) {
                const int t = __ffs(todo) - 1;
                todo &= todo - 1;
                const int idx = base - t;
                const I sx = __shfl_sync(FULL, mx, t), sy = __shfl_sync(FULL, my_y, t);
                const I ex = sx + 1, ey = sy + 1;
                // score(end): the highest layer containing a point >= (ex, ey)
                const int wl = nlayers - lane;
                bool c = false;
                if (wl >= 1) {
                    c = (ex <= cp.x && ey <= cp.y) || (cp.z != INT32_MIN && ex <= cp.z && ey <= cp.w);
                    if (!c && covf)
                        for (int q = layer_head[wl]; q >= 0; q = next[q])
                            if (ex <= px[q] && ey <= py[q]) {
                                c = true;
                                break;
                            }
                }
                const unsigned bal = __ballot_sync(FULL, c);
                int v;
                if (bal)
                    v = nlayers - (__ffs(bal) - 1) + 1;
                else if (nlayers <= 32)
                    v = 1;  // every layer >= 1 is cached and none contains the point: layer 0
                else
                    v = score(ex, ey, 3) + 1;  // below the cached window: search the table in memory
                if (v > nlayers) {  // a new top layer: every cached layer moves one lane up
                    cp.x = __shfl_up_sync(FULL, cp.x, 1);
                    cp.y = __shfl_up_sync(FULL, cp.y, 1);
                    cp.z = __shfl_up_sync(FULL, cp.z, 1);
                    cp.w = __shfl_up_sync(FULL, cp.w, 1);
                    covf = __shfl_up_sync(FULL, (int)covf, 1) != 0;
                    if (lane == 0) {
                        cp = make_int4(sx, sy, INT32_MIN, INT32_MIN);
                        covf = false;
                        layer_pts[v] = cp;
                        layer_head[v] = -1;
                    }
                    nlayers = v;
                } else if (nlayers - v < 32) {  // joins a cached layer: its lane updates registers and memory
                    if (lane == nlayers - v) {
                        if (cp.z == INT32_MIN) {
                            cp.z = sx;
                            cp.w = sy;
                            layer_pts[v] = cp;
                        } else {
                            next[idx] = layer_head[v];
                            layer_head[v] = idx;
                            covf = true;
                        }
                    }
                } else if (lane == 0) {  // joins a layer below the window: memory only (every layer <= nlayers has a first point)
                    int4 p = layer_pts[v];
                    if (p.z == INT32_MIN) {
                        p.z = sx;
                        p.w = sy;
                        layer_pts[v] = p;
                    } else {
                        next[idx] = layer_head[v];
                        layer_head[v] = idx;
                    }
                }
                hint[3] = v;
                __syncwarp();
            }
        }
        hint[0] = hint[1] = hint[2] = hint[3];
        hq[0] = hq[1] = hq[2] = hq[3] = 0;
        dens = min(256, (int)(((long long)nlayers << 8) / max(1, nseeds)));
        dirty = false;
    }
#else
    // HintContours::new over the active arrows (hint_contours.rs:213-255) == state after update_layers.
    __device__ void build_layers() {
        const int lane = threadIdx.x & 31;
        for (int w = lane; w <= nlayers + 1 && w <= M + 1; w += 32) {
            layer_head[w] = -1;
            layer_pts[w] = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
        }
        __syncwarp();
        nlayers = 0;
        hint[3] = 0;
        for (int idx = M - 1; idx >= 0; idx--) {
            if (!active[idx]) continue;
            I ex = px[idx] + 1, ey = py[idx] + 1;  // transform(end): P(end.i) = P(start.i) - 1
            if (!(ex <= ttx && ey <= tty)) continue;
            int v = score(ex, ey, 3) + 1;
            if (lane == 0) {
                int4 p = (v > nlayers) ? make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN) : layer_pts[v];
                if (v > nlayers) layer_head[v] = -1;
                if (p.x == INT32_MIN) {
                    p.x = px[idx];
                    p.y = py[idx];
                    layer_pts[v] = p;
                } else if (p.z == INT32_MIN) {
                    p.z = px[idx];
                    p.w = py[idx];
                    layer_pts[v] = p;
                } else {
                    next[idx] = layer_head[v];
                    layer_head[v] = idx;
                }
            }
            if (v > nlayers) nlayers = v;
            hint[3] = v;
            __syncwarp();
        }
        hint[0] = hint[1] = hint[2] = hint[3];
        hq[0] = hq[1] = hq[2] = hq[3] = 0;
        dens = min(256, (int)(((long long)nlayers << 8) / max(1, nseeds)));
        dirty = false;
    }
#endif
    __device__ void update_contours() {
        if (dirty) build_layers();
    }
    // MatchPruner::prune_block (prune.rs:245-292): i_range = is..ie, j_range = js..je used as inclusive bounds.
    // The seeds of the block (<= 22 for a 256-column block) are independent of each other: one lane per seed.
    __device__ void prune_block(I is, I ie, I js, I je) {
        const int lane = threadIdx.x & 31;
        const I s0 = (is + K()) / K();  // first seed with col >= is + 1
        bool changed = false;
        for (I sb = s0; sb < nseeds && sb * K() <= ie; sb += 32) {
            const I s = sb + lane;
            if (s < nseeds && s * K() <= ie) {
                const uint32_t b_start = base[s], a_end = base[s + 1];
                uint32_t b_end = before_end[s];
                uint32_t a_start = after_start[s];
                if (a_start == HT_EMPTY) {
                    a_start = b_end;
                    while (a_start >= b_start + 1 && ms_j[a_start - 1] > je) {
                        b_end -= 1;
                        a_start -= 1;
                    }
                }
                while (b_end > b_start && ms_j[b_end - 1] >= js) {
                    active[b_end - 1] = 0;
                    b_end -= 1;
                    changed = true;
                }
                while (a_start < a_end && ms_j[a_start] <= je) {
                    active[a_start] = 0;
                    a_start += 1;
                    changed = true;
                }
                before_end[s] = b_end;
                after_start[s] = a_start;
            }
        }
        if (__any_sync(FULL, changed)) dirty = true;
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------------ local pruning
// Warp-uniform values the build keeps in registers. The compiler otherwise re-derives arena pointers from their
// allocation arithmetic at every use (10+ instructions each under the 48-register cap); a value that went through a
// shuffle cannot be rematerialised, so it is kept (or spilled, which costs one load).
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { return __shfl_sync(FULL, v, 0); }
template <class T>
__device__ __forceinline__ T* pin_ptr(T* p) {
    unsigned long long v = (unsigned long long)p;
    unsigned lo = __shfl_sync(FULL, (unsigned)v, 0), hi = __shfl_sync(FULL, (unsigned)(v >> 32), 0);
    return (T*)(((unsigned long long)hi << 32) | lo);
}

// next_match_per_diag (matches.rs:147-148): diagonal -> start.i of the left-most kept match, default MAX. One entry per
// diagonal the transform filter lets through (2 ns + 2 (p + 2) + 1 of them), of which a pair touches a few hundred: the array is
// NOT initialised. A bitmap (one bit per segment of 32 entries, 64 bytes at n = 100 k) says which segments hold values; the first
// store into a segment fills it with MAX (one coalesced store), loads from an unmarked segment answer MAX without looking.
struct NmpdView {
    I* vb;           // biased: vb[d] for dmin <= d <= dmax
    uint32_t* segs;  // bit s: entries [32 s, 32 s + 32) (relative to dmin) are initialised
    I dmin, dmax;
#if !APA_NM_LAZY
    __device__ __forceinline__ I get(I d) const { return (d < dmin || d > dmax) ? INT32_MAX : vb[d]; }
    __device__ __forceinline__ void set(I d, I si) {
        if (d >= dmin && d <= dmax && (threadIdx.x & 31) == 0) vb[d] = si;
        __syncwarp();
    }
#else
    __device__ __forceinline__ I get(I d) const {
        if (d < dmin || d > dmax) return INT32_MAX;
        const uint32_t sgm = (uint32_t)(d - dmin) >> 5;
        const uint32_t w = segs[sgm >> 5];  // both loads are issued together; the value is ignored for an unmarked segment
        const I v = vb[d];
        return ((w >> (sgm & 31)) & 1u) ? v : INT32_MAX;
    }
    __device__ __forceinline__ void set(I d, I si) {  // warp-uniform arguments, called by the whole warp
        if (d < dmin || d > dmax) return;
        const int lane = threadIdx.x & 31;
        const uint32_t sgm = (uint32_t)(d - dmin) >> 5;
        const uint32_t w = segs[sgm >> 5];
        if (!((w >> (sgm & 31)) & 1u)) {
            const I e = dmin + (I)(sgm << 5) + lane;
            if (e <= dmax) vb[e] = INT32_MAX;
            __syncwarp();
            if (lane == 0) segs[sgm >> 5] = w | (1u << (sgm & 31));
        }
        if (lane == 0) vb[d] = si;
        __syncwarp();
    }
#endif
};

struct PruneWin {};
#if APA_DT_V2
// preserve_for_local_pruning (prepruning.rs:95-203) for an exact match of seed `seed` starting at (seed * k, sj).
// Warp-uniform result. Lanes 0 .. 2 pd hold the DT front, lane <-> diagonal (ei - ej) + (lane - pd): a lane's diagonal never
// changes, so its next_match_per_diag entry is loaded once, before the loop. Potentials in closed form: P(i) = ns - ceil(i / k)
// for i <= ns k, so "g + P(fr) >= P(start)" (prepruning.rs:158-168) is the single compare fr <= (seed + g) k.
__device__ bool dev_preserve_for_local_pruning(const GcshH& H, PruneWin&, const uint2* __restrict__ ap, const uint2* __restrict__ bp, I seed,
                                               I sj, const NmpdView& nm) {
    const int lane = threadIdx.x & 31;
    const int GK = H.K();
    const I si = seed * GK;
    const I ei = si + GK, ej = sj + GK;
    const I last = min(seed + H.P() - 1, H.nseeds - 1);
    const I end_i = (last + 1) * GK;
    const int pd = last + 1 - seed;  // start_pot - P(end_i), <= GCSH_P
    const I dd = ei - ej + (lane - pd);  // this lane's diagonal
    const I nmv = (lane <= 2 * pd) ? nm.get(dd) : INT32_MAX;
    // g = 0
    I f0 = ei;
    extend_right_packed(ap, bp, H.m, f0, ej, end_i);
    if (f0 >= end_i) return true;
    if (__shfl_sync(FULL, nmv, pd) <= f0) return true;
    I fr = (lane == pd) ? f0 : INT32_MIN;
    int lo = pd, hi = pd + 1;  // d_range
    I dead_below = si + GK;    // (seed + g) k at g = 1
    for (Cost g = 1; g < pd; g++, dead_below += GK) {
        // expand: next[d] = max(fr[d+1], fr[d] + 1, fr[d-1] + 1) over sources inside d_range (lanes outside hold MIN)
        const I up = __shfl_down_sync(FULL, fr, 1);  // fr[d+1]
        const I dn = __shfl_up_sync(FULL, fr, 1);    // fr[d-1]
        const I here = (lane >= lo && lane < hi) ? fr + 1 : INT32_MIN;
        const I from_up = (lane + 1 >= lo && lane + 1 < hi) ? up : INT32_MIN;
        const I from_dn = (lane - 1 >= lo && lane - 1 < hi) ? dn + 1 : INT32_MIN;
        fr = max(here, max(from_up, from_dn));
        lo -= 1;
        hi += 1;
        // check & shrink (from both ends only, like the reference)
        bool in = lane >= lo && lane < hi;
        const unsigned alive = __ballot_sync(FULL, in && fr > dead_below);
        if (alive == 0) return false;
        lo = __ffs(alive) - 1;
        hi = 32 - __clz(alive);
        // extend
        in = lane >= lo && lane < hi;
        bool ok = false;
        if (in) {
            const I old_i = fr;
            extend_right_packed(ap, bp, H.m, fr, fr - dd, end_i);
            ok = (fr >= end_i) || (old_i <= nmv && nmv <= fr);
        }
        if (__any_sync(FULL, ok)) return true;
    }
    return false;
}

#else
// preserve_for_local_pruning (prepruning.rs:95-203) for an exact match of seed `seed` starting at (seed * k, sj).
// Warp-uniform result. Potentials in closed form: P(seed * k) = ns - seed.
__device__ bool dev_preserve_for_local_pruning(const GcshH& H, PruneWin&, const uint2* __restrict__ ap, const uint2* __restrict__ bp, I seed,
                                               I sj, const NmpdView& nm) {
    const int lane = threadIdx.x & 31;
    const I si = seed * H.K();
    const I ei = si + H.K(), ej = sj + H.K();
    const Cost start_pot = H.nseeds - seed;
    const I last = min(seed + H.P() - 1, H.nseeds - 1);
    const I end_i = (last + 1) * H.K();
    const int pd = last + 1 - seed;  // start_pot - P(end_i), <= GCSH_P
    // g = 0
    I f0 = ei;
    extend_right_packed(ap, bp, H.m, f0, ej, end_i);
    if (f0 >= end_i) return true;
    if (nm.get(ei - ej) <= f0) return true;
    // lanes 0 .. 2*pd hold the front; lane d <-> diagonal e + (d - pd)
    I fr = (lane == pd) ? f0 : INT32_MIN;
    int lo = pd, hi = pd + 1;  // d_range
    const I dd = ei - ej + (lane - pd);  // this lane's diagonal
    for (Cost g = 1; g < pd; g++) {
        // expand: next[d] = max(fr[d+1], fr[d] + 1, fr[d-1] + 1) over sources inside d_range
        I up = __shfl_down_sync(FULL, fr, 1);  // fr[d+1]
        I dn = __shfl_up_sync(FULL, fr, 1);    // fr[d-1]
        I nx = INT32_MIN;
        if (lane + 1 >= lo && lane + 1 < hi) nx = max(nx, up);
        if (lane >= lo && lane < hi) nx = max(nx, fr + 1);
        if (lane - 1 >= lo && lane - 1 < hi) nx = max(nx, dn + 1);
        fr = nx;
        lo -= 1;
        hi += 1;
        // check & shrink
        bool in = lane >= lo && lane < hi;
        bool dead = in && (g + H.pot(fr) >= start_pot);
        unsigned alive = __ballot_sync(FULL, in && !dead);
        if (alive == 0) return false;
        lo = __ffs(alive) - 1;
        hi = 32 - __clz(alive);
        // extend
        in = lane >= lo && lane < hi;
        bool ok = false;
        if (in) {
            I j = fr - dd;
            I old_i = fr;
            extend_right_packed(ap, bp, H.m, fr, j, end_i);
            I nmv = nm.get(dd);
            ok = (fr >= end_i) || (old_i <= nmv && nmv <= fr);
        }
        if (__any_sync(FULL, ok)) return true;
    }
    return false;
}

#endif

constexpr uint32_t KMER_MUL = 0x9E3779B1u;
constexpr uint32_t STAGE_MULTI = 0x40000000u;  // staged hit whose k-mer occurs in several seeds of a

// CSHI::new (csh.rs:199-308): find matches, filter, sort, build the pruner state and the contours.
// All scratch comes from the pair's arena; structures that die after the precomputation are placed last so the
// block store can reuse their space. Returns false on arena overflow (cx.status set).
__device__ bool gcsh_build(PairCtx& cx, WarpSmem& sm, GcshH& H) {
    PruneWin& win = *(PruneWin*)sm.dt_i;
    const int lane = threadIdx.x & 31;
    const I n = cx.n, m = cx.m;
    H.n = n;
    H.m = m;
    const int GK = H.K(), GP = H.P();
    H.nseeds = n >= GK ? (n - GK) / GK + 1 : 0;  // fixed_length_seeds, qgrams.rs:99-109
    H.ttx = n - m;  // transform(target): P(n) = 0
    H.tty = m - n;
    H.h_calls = 0;
    H.probes = 0;
    H.t_h = 0;
    H.hint[0] = H.hint[1] = H.hint[2] = H.hint[3] = 0;
    H.hq[0] = H.hq[1] = H.hq[2] = H.hq[3] = 0;
    H.dens = 0;
    H.dirty = false;
    H.nlayers = 0;
    const I ns = H.nseeds;

    // capacity for matches scales with the arena (re-run with a larger arena on overflow)
    const uint32_t avail = cx.hi_bot - cx.v_top;
    int mcap = (int)min((uint32_t)(1u << 30), avail / 192u);
    if (mcap < 64) {
        cx.status = ST_OVERFLOW;
        return false;
    }
    // ---- live structures
    uint32_t off_base = arena_alloc(cx, (uint32_t)(ns + 2) * 4u);
    uint32_t off_bend = arena_alloc(cx, (uint32_t)(ns + 1) * 4u);
    uint32_t off_astart = arena_alloc(cx, (uint32_t)(ns + 1) * 4u);
    uint32_t off_msi = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_msj = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_px = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_py = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_next = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_lh = arena_alloc(cx, (uint32_t)(mcap + 2) * 4u);
    uint32_t off_lp = arena_alloc(cx, (uint32_t)(mcap + 2) * 16u);
    uint32_t off_act = arena_alloc(cx, (uint32_t)mcap);
    const uint32_t live_end = cx.v_top;
    // ---- dead after the precomputation
    // Open-addressing table of the seeds: load factor 0.31 .. 0.63 (16 384 slots = 128 KB at n = 100 k). k-mer presence filter in
    // front of it: a blocked Bloom filter, two bits per key inside one 32-bit word, 4 bits of filter per table slot (8 KB at
    // n = 100 k: the filters of all resident warps, ~45 MB, stay in the 126 MB L2; about 5 % of unrelated windows pass).
    int log_t = 5;
#if APA_TAB_DENSE
    while ((5ll << log_t) < 8ll * ns) log_t++;  // 2^log_t >= 1.6 ns: load factor 0.31 .. 0.63
#else
    while ((1 << log_t) < 2 * ns) log_t++;      // 2^log_t >= 2 ns: load factor 0.25 .. 0.5
#endif
    const uint32_t tsize = 1u << log_t;
    const int log_bw = log_t + APA_BLOOM_LOG - 5;  // filter words: 2^APA_BLOOM_LOG bits per table slot
    const uint32_t bm_words = 1u << log_bw;
    uint32_t off_tab = arena_alloc(cx, tsize * 8u);
    uint32_t off_bm = arena_alloc(cx, bm_words * 4u);
    const I dmin = (n - m) - ns - (GP + 2), dmax = (n - m) + ns + (GP + 2);
    const uint32_t nm_entries = (uint32_t)(dmax - dmin + 1);
    const uint32_t nm_segw = (nm_entries + 1023u) >> 10;  // bitmap words: one bit per 32 entries
    uint32_t off_nm = arena_alloc(cx, ((nm_entries + 31u) & ~31u) * 4u);
    uint32_t off_nmseg = arena_alloc(cx, nm_segw * 4u);
    uint32_t off_arr = arena_alloc(cx, (uint32_t)mcap * 16u);
    uint32_t off_stage = arena_alloc(cx, 32u * 32u * 8u);
    if (cx.status != ST_PENDING) return false;

    uint32_t* cnt = pin_ptr((uint32_t*)(cx.arena + off_base));  // counts, then exclusive prefix
    uint32_t* before_end = (uint32_t*)(cx.arena + off_bend);
    uint32_t* after_start = (uint32_t*)(cx.arena + off_astart);
    I* ms_i = (I*)(cx.arena + off_msi);
    I* ms_j = (I*)(cx.arena + off_msj);
    uint2* tab = pin_ptr((uint2*)(cx.arena + off_tab));
    uint32_t* bm = pin_ptr((uint32_t*)(cx.arena + off_bm));
    I* nm_v = (I*)(cx.arena + off_nm);
    uint32_t* nm_seg = (uint32_t*)(cx.arena + off_nmseg);
    NmpdView nm{pin_ptr(nm_v - dmin), pin_ptr(nm_seg), dmin, dmax};
    int4* arr = pin_ptr((int4*)(cx.arena + off_arr));        // kept matches in arrival order: (seed, j, rank within seed, -)
    uint2* stage = pin_ptr((uint2*)(cx.arena + off_stage));  // per lane: up to 32 staged hits (j, first seed | STAGE_MULTI)
    const uint2* ap = pin_ptr(cx.aprof);
    const uint2* bp = pin_ptr(cx.bprof);

    for (uint32_t t = lane; t < tsize; t += 32) tab[t] = make_uint2(0u, HT_EMPTY);
    for (uint32_t t = lane; t < bm_words; t += 32) bm[t] = 0u;
#if APA_NM_LAZY
    for (uint32_t t = lane; t < nm_segw; t += 32) nm_seg[t] = 0u;
#else
    for (uint32_t t = lane; t < nm_entries; t += 32) nm_v[t] = INT32_MAX;
#endif
    for (I t = lane; t < ns + 2; t += 32) cnt[t] = 0u;
    __syncwarp();

    // ---- hash the seeds of a (hash_to_smallvec, exact.rs:48-55). Key: bit t = rank bit0 of char t, bit k+t = rank bit1.
    const uint32_t kmask = (1u << GK) - 1u;
    // filter word and the two bits of a key: word from the top bits of the hash, bit positions from the next 5 + 5
    auto bloom_word = [&](uint32_t hsh) -> uint32_t { return hsh >> (32 - log_bw); };
#if APA_BLOOM2
    auto bloom_bits = [&](uint32_t hsh) -> uint32_t { return (1u << ((hsh >> 5) & 31u)) | (1u << (hsh & 31u)); };
#else
    auto bloom_bits = [&](uint32_t hsh) -> uint32_t { return 1u << ((hsh >> (32 - log_bw - 5)) & 31u); };
#endif
    // Four seeds per lane and round: the four compare-and-swaps are in flight together (each is a DRAM-latency round trip into a
    // table no cache holds); a lane whose slot was taken walks on alone. Only this warp touches the table, the CAS settles
    // collisions between its own lanes.
#if APA_CAS4
    // Four seeds per lane and round, their compare-and-swaps in flight together (each is a DRAM-latency round trip into a table
    // no cache holds). Only the four results stay in registers: a lane whose slot was taken recomputes the key and walks on
    // alone. Only this warp touches the table; the CAS settles collisions between its own lanes.
    const unsigned long long EMPTY64 = (unsigned long long)HT_EMPTY << 32;
    auto seed_key = [&](I sd) -> uint32_t {
        const uint2 w = extract32(ap, sd * GK);  // planes are stored negated
        return (~w.x & kmask) | ((~w.y & kmask) << GK);
    };
    for (I s0 = 0; s0 < ns; s0 += 128) {
        unsigned long long got[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const I sd = s0 + 32 * u + lane;
            got[u] = EMPTY64;
            if (sd < ns) {
                const uint32_t key = seed_key(sd);
                const uint32_t hsh = key * KMER_MUL;
                atomicOr(&bm[bloom_word(hsh)], bloom_bits(hsh));
                got[u] = atomicCAS((unsigned long long*)&tab[hsh >> (32 - log_t)], EMPTY64, ((unsigned long long)(uint32_t)sd << 32) | key);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (got[u] != EMPTY64) {
                const I sd = s0 + 32 * u + lane;
                const uint32_t key = seed_key(sd);
                uint32_t slot = (key * KMER_MUL) >> (32 - log_t);
                do {
                    slot = (slot + 1) & (tsize - 1);
                } while (atomicCAS((unsigned long long*)&tab[slot], EMPTY64, ((unsigned long long)(uint32_t)sd << 32) | key) != EMPTY64);
            }
    }
    __syncwarp();

#else
    const unsigned long long EMPTY64 = (unsigned long long)HT_EMPTY << 32;
    for (I s0 = 0; s0 < ns; s0 += 32) {
        I sd = s0 + lane;
        if (sd < ns) {
            const uint2 w = extract32(ap, sd * GK);  // planes are stored negated
            const uint32_t key = (~w.x & kmask) | ((~w.y & kmask) << GK);
            const uint32_t hsh = key * KMER_MUL;
            atomicOr(&bm[bloom_word(hsh)], bloom_bits(hsh));
            uint32_t slot = hsh >> (32 - log_t);
            unsigned long long want = ((unsigned long long)(uint32_t)sd << 32) | key;  // uint2{key, seed}
            for (;;) {
                unsigned long long old = atomicCAS((unsigned long long*)&tab[slot], EMPTY64, want);
                if (old == EMPTY64) break;
                slot = (slot + 1) & (tsize - 1);
            }
        }
    }
    __syncwarp();

#endif
    // smallest seed > `after` whose k-mer is `key` (INT32_MAX if none); n_out = number of seeds with that k-mer.
    // probe_from continues from a slot whose entry `e` the caller has already loaded.
    auto probe_from = [&](uint32_t key, uint32_t slot, uint2 e, I after, int& n_out) -> I {
        I best = INT32_MAX;
        int c = 0;
        while (e.y != HT_EMPTY) {
            if (e.x == key) {
                c++;
                if ((I)e.y > after && (I)e.y < best) best = (I)e.y;
            }
            slot = (slot + 1) & (tsize - 1);
            e = tab[slot];
        }
        n_out = c;
        return best;
    };
    auto probe = [&](uint32_t key, I after, int& n_out) -> I {
        const uint32_t slot = (key * KMER_MUL) >> (32 - log_t);
        return probe_from(key, slot, tab[slot], after, n_out);
    };

    // ---- all windows of b, right to left (b_qgrams_rev, qgrams.rs:81-97), 1024 per round: lane l owns the 32 windows
    // j = jhi, jhi - 1, ... with jhi = jtop - 32 l. (1) every lane tests its windows against the presence filter (one
    // 4-byte load each, L2-resident), (2) the survivors (~1 in 9 for unrelated windows) are looked up in the hash table and
    // the hits staged per lane, (3) the hits are pushed in the reference's arrival order - j descending, seeds ascending -
    // through MatchBuilder::push (matches.rs:205-247), one hit at a time with the lanes on the diagonals of the pruning front.
    int M = 0;
    for (I jtop = m - GK; jtop >= 0; jtop -= 1024) {
        const I jhi = jtop - 32 * lane;
        const I wbase = max(jhi - 31, 0);  // bit s of `surv` <-> window j = wbase + s
        uint32_t surv = 0u;
        uint32_t w0x = 0u, w0y = 0u, w1x = 0u, w1y = 0u;
        if (jhi >= 0) {
            const uint2 lo = extract32(bp, wbase), hi = extract32(bp, wbase + 32);
            w0x = ~lo.x, w0y = ~lo.y, w1x = ~hi.x, w1y = ~hi.y;
#pragma unroll 8
            for (int sft = 0; sft < 32; sft++) {
                const uint32_t key = (__funnelshift_r(w0x, w1x, sft) & kmask) | ((__funnelshift_r(w0y, w1y, sft) & kmask) << GK);
                const uint32_t hsh = key * KMER_MUL;
                const uint32_t bits = bloom_bits(hsh);
                surv |= ((bm[bloom_word(hsh)] & bits) == bits ? 1u : 0u) << sft;
            }
            const int nwin = jhi - wbase + 1;  // windows above jhi belong to the previous lane
            if (nwin < 32) surv &= (1u << nwin) - 1u;
        }
        int nh = 0;
        uint2 st0 = make_uint2(0u, 0u), st1 = make_uint2(0u, 0u);
        auto stage_hit = [&](int sft, I seed0, int c) {  // the first two hits of a lane stay in registers (1.7 hits per lane and round on average)
            const uint2 rec = make_uint2((uint32_t)(wbase + sft), (uint32_t)seed0 | (c > 1 ? STAGE_MULTI : 0u));
            if (nh == 0)
                st0 = rec;
            else if (nh == 1)
                st1 = rec;
            else
                stage[lane * 32 + nh] = rec;
            nh++;
        };
        auto window_key = [&](int sft) -> uint32_t {
            return (__funnelshift_r(w0x, w1x, sft) & kmask) | ((__funnelshift_r(w0y, w1y, sft) & kmask) << GK);
        };
#if APA_PROBE2
        // two survivors per lane and iteration: both first table entries are loaded before either chain is followed
        while (__any_sync(FULL, surv != 0u)) {
            if (surv) {
                const int sa = 31 - __clz(surv);  // highest j first
                surv &= ~(1u << sa);
                const int sb = surv ? 31 - __clz(surv) : -1;
                if (sb >= 0) surv &= ~(1u << sb);
                const uint32_t ka = window_key(sa), kb = window_key(sb < 0 ? 0 : sb);
                const uint32_t slot_a = (ka * KMER_MUL) >> (32 - log_t), slot_b = (kb * KMER_MUL) >> (32 - log_t);
                const uint2 ea = tab[slot_a];
                uint2 eb = make_uint2(0u, HT_EMPTY);
                if (sb >= 0) eb = tab[slot_b];
                int c;
                I seed0 = probe_from(ka, slot_a, ea, -1, c);
                if (c > 0) stage_hit(sa, seed0, c);
                if (sb >= 0) {
                    seed0 = probe_from(kb, slot_b, eb, -1, c);
                    if (c > 0) stage_hit(sb, seed0, c);
                }
            }
        }
#else
        while (__any_sync(FULL, surv != 0u)) {
            if (surv) {
                const int sft = 31 - __clz(surv);  // highest j first
                surv &= ~(1u << sft);
                int c;
                const I seed0 = probe(window_key(sft), -1, c);
                if (c > 0) stage_hit(sft, seed0, c);
            }
        }
#endif
        __syncwarp();
        unsigned has = __ballot_sync(FULL, nh > 0);
        while (has) {
            const int l = __ffs(has) - 1;
            has &= has - 1;
            const int cntl = __shfl_sync(FULL, nh, l);
            for (int hh = 0; hh < cntl; hh++) {
                uint2 rec;
                if (hh < 2) {
                    rec.x = __shfl_sync(FULL, hh == 0 ? st0.x : st1.x, l);
                    rec.y = __shfl_sync(FULL, hh == 0 ? st0.y : st1.y, l);
                } else {
                    rec = stage[l * 32 + hh];
                }
                const I jj = (I)rec.x;
                I seed = (I)(rec.y & ~STAGE_MULTI);
                const bool multi = (rec.y & STAGE_MULTI) != 0u;
                for (;;) {
                    // MatchBuilder::push (matches.rs:205-247)
                    const I si = seed * GK;
                    const Cost p = ns - seed;  // P(si)
                    const bool pass_t = (si - jj - p <= H.ttx) && (jj - si - p <= H.tty);
                    if (pass_t && (GP == 0 || dev_preserve_for_local_pruning(H, win, ap, bp, seed, jj, nm))) {  // local_pruning = 0: keep all (matches.rs:213)
                        if (M >= mcap) {
                            cx.status = ST_OVERFLOW;
                            return false;
                        }
                        nm.set(si - jj, si);
                        if (lane == 0) {
                            const uint32_t rank = cnt[seed];
                            cnt[seed] = rank + 1u;
                            arr[M] = make_int4(seed, jj, (int)rank, 0);
                        }
                        M++;
                        __syncwarp();
                    }
                    if (!multi) break;
                    const uint2 w = extract32(bp, jj);
                    const uint32_t kk = (~w.x & kmask) | ((~w.y & kmask) << GK);
                    int dummy;
                    seed = probe(kk, seed, dummy);
                    if (seed == INT32_MAX) break;
                }
            }
        }
    }
    __syncwarp();
    // ---- sort by (start.i, start.j): counting sort by seed; arrival order inside a seed is j descending
    {
        uint32_t run = 0;
        for (I s0 = 0; s0 < ns + 1; s0 += 32) {
            I s = s0 + lane;
            uint32_t v = (s < ns) ? cnt[s] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += y;
            }
            if (s < ns + 1) cnt[s] = run + incl - v;  // exclusive prefix; cnt[ns] = M
            run += __shfl_sync(FULL, incl, 31);
        }
        __syncwarp();
        for (int t = lane; t < M; t += 32) {
            const int4 a4 = arr[t];
            const I s = a4.x;
            uint32_t c_s = cnt[s + 1] - cnt[s];
            uint32_t pos = cnt[s] + (c_s - 1u - (uint32_t)a4.z);
            ms_i[pos] = s * GK;
            ms_j[pos] = a4.y;
        }
        __syncwarp();
    }
    H.M = M;
    H.ms_i = ms_i;
    H.ms_j = ms_j;
    H.px = (I*)(cx.arena + off_px);
    H.py = (I*)(cx.arena + off_py);
    H.next = (int*)(cx.arena + off_next);
    H.layer_head = (int*)(cx.arena + off_lh);
    H.layer_pts = (int4*)(cx.arena + off_lp);
    H.active = cx.arena + off_act;
    H.base = cnt;
    H.before_end = before_end;
    H.after_start = after_start;
    for (int t = lane; t < M; t += 32) {
        I i = ms_i[t], j = ms_j[t];
        Cost p = H.pot(i);
        H.px[t] = i - j - p;
        H.py[t] = j - i - p;
        H.active[t] = 1;
    }
    for (I s = lane; s < ns; s += 32) {  // MatchPruner::new active ranges (prune.rs:171-189)
        before_end[s] = cnt[s + 1];
        after_start[s] = HT_EMPTY;
    }
    __syncwarp();
    // the hash table, next_match_per_diag and the arrival arrays are dead now: the block store starts here
    cx.v_top = live_end;
    cx.v_base = live_end;
    H.nlayers = 0;
    if (lane == 0) {
        H.layer_head[0] = -1;
        H.layer_pts[0] = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
    }
    __syncwarp();
    H.build_layers();
    return true;
}

}  // namespace APA_NS
