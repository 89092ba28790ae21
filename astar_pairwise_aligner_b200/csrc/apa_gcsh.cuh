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

namespace apa {

constexpr int GCSH_K = 12;        // seed length (params.rs:103)
constexpr int GCSH_P = 14;        // local-pruning look-ahead in seeds (params.rs:105)
constexpr uint32_t HT_EMPTY = 0xffffffffu;

struct GcshH {
    static constexpr bool PRUNE = true;
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
    int hint;
    bool dirty;
    unsigned long long h_calls;
    long long t_h;

    // Seeds::potential (seeds.rs:79-81) for fixed-length seeds at 0, k, 2k, ...: number of seeds starting at >= i.
    __device__ __forceinline__ Cost pot(I i) const {
        I c = (i + GCSH_K - 1) / GCSH_K;
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
    // (a point in layer w is dominated by one in layer w-1), so 32 layers are probed per step.
    __device__ int score(I qx, I qy) {
        const int lane = threadIdx.x & 31;
        if (nlayers == 0) return 0;
        int lo = 0;            // known: contained in layer lo (layer 0 holds (MAX, MAX))
        int hi = nlayers + 1;  // known: not contained in layer hi
        int basew = max(1, min(hint - 15, nlayers - 31));
        for (;;) {
            int w = basew + lane;
            bool c = (w < hi) && (w > lo) && contains(w, qx, qy);
            unsigned bal = __ballot_sync(FULL, c);
            // windows are clipped to (lo, hi): lanes outside vote 0
            int first = max(basew, lo + 1);       // first probed layer
            int last = min(basew + 31, hi - 1);   // last probed layer
            if (last < first) break;
            int cntc = __popc(bal);
            if (cntc == 0) {
                hi = first;
            } else {
                lo = first + cntc - 1;
                if (lo < last) hi = lo + 1;
            }
            if (hi - lo <= 1) break;
            // next window: centred bisection of the remaining interval
            int mid = lo + (hi - lo) / 2;
            basew = max(lo + 1, mid - 15);
        }
        hint = lo;
        return lo;
    }
    // CSHI::h / h_with_hint (csh.rs:341-376): P(u) - layer(T(u)), or max(gap, potential) to the target in layer 0.
    __device__ Cost h(I i, I j) {
        h_calls++;
        long long t0 = APA_TIC();
        Cost p = pot(i);
        int val = score(i - j - p, j - i - p);
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
    // HintContours::new over the active arrows (hint_contours.rs:213-255) == state after update_layers.
    __device__ void build_layers() {
        const int lane = threadIdx.x & 31;
        for (int w = lane; w <= nlayers + 1 && w <= M + 1; w += 32) {
            layer_head[w] = -1;
            layer_pts[w] = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
        }
        __syncwarp();
        nlayers = 0;
        hint = 0;
        for (int idx = M - 1; idx >= 0; idx--) {
            if (!active[idx]) continue;
            I ex = px[idx] + 1, ey = py[idx] + 1;  // transform(end): P(end.i) = P(start.i) - 1
            if (!(ex <= ttx && ey <= tty)) continue;
            int v = score(ex, ey) + 1;
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
            hint = v;
            __syncwarp();
        }
        dirty = false;
    }
    __device__ void update_contours() {
        if (dirty) build_layers();
    }
    // MatchPruner::prune_block (prune.rs:245-292): i_range = is..ie, j_range = js..je used as inclusive bounds.
    __device__ void prune_block(I is, I ie, I js, I je) {
        const int lane = threadIdx.x & 31;
        I s0 = (is + 1 + GCSH_K - 1) / GCSH_K;  // first seed with col >= is + 1
        for (I s = s0; s < nseeds && s * GCSH_K <= ie; s++) {
            uint32_t b_start = base[s], b_end = before_end[s];
            uint32_t a_start = after_start[s];
            const uint32_t a_end = base[s + 1];
            if (a_start == HT_EMPTY) {
                a_start = b_end;
                while (a_start >= b_start + 1 && ms_j[a_start - 1] > je) {
                    b_end -= 1;
                    a_start -= 1;
                }
            }
            bool changed = false;
            while (b_end > b_start && ms_j[b_end - 1] >= js) {
                if (lane == 0) active[b_end - 1] = 0;
                b_end -= 1;
                changed = true;
            }
            while (a_start < a_end && ms_j[a_start] <= je) {
                if (lane == 0) active[a_start] = 0;
                a_start += 1;
                changed = true;
            }
            if (lane == 0) {
                before_end[s] = b_end;
                after_start[s] = a_start;
            }
            if (changed) dirty = true;
        }
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------------ local pruning
struct NmpdView {  // next_match_per_diag (matches.rs:147-148): diagonal -> start.i of the left-most kept match, default MAX
    I* v;
    I dmin, dmax;
    __device__ __forceinline__ I get(I d) const { return (d < dmin || d > dmax) ? INT32_MAX : v[d - dmin]; }
};

// preserve_for_local_pruning (prepruning.rs:95-203) for an exact match starting at (si, sj). Warp-uniform result.
__device__ bool dev_preserve_for_local_pruning(const GcshH& H, const uint2* __restrict__ ap, const uint2* __restrict__ bp, I si, I sj,
                                               const NmpdView& nm) {
    const int lane = threadIdx.x & 31;
    const I ei = si + GCSH_K, ej = sj + GCSH_K;
    const Cost start_pot = H.pot(si);
    const I seed_idx = si / GCSH_K;
    const I last = min(seed_idx + GCSH_P - 1, H.nseeds - 1);
    const I end_i = (last + 1) * GCSH_K;
    const Cost end_pot = H.pot(end_i);
    const int pd = start_pot - end_pot;  // <= GCSH_P
    // g = 0
    I f0 = ei;
    extend_right_packed(ap, bp, H.m, f0, ej, end_i);
    if (f0 >= end_i) return true;
    if (nm.get(ei - ej) <= f0) return true;
    // lanes 0 .. 2*pd hold the front; lane d <-> diagonal e + (d - pd)
    I fr = (lane == pd) ? f0 : INT32_MIN;
    int lo = pd, hi = pd + 1;  // d_range
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
            I dd = ei - ej + (lane - pd);
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

__device__ __forceinline__ uint32_t kmer_hash(uint32_t key, int log_t) { return (key * 0x9E3779B1u) >> (32 - log_t); }

// CSHI::new (csh.rs:199-308): find matches, filter, sort, build the pruner state and the contours.
// All scratch comes from the pair's arena; structures that die after the precomputation are placed last so the
// block store can reuse their space. Returns false on arena overflow (cx.status set).
__device__ bool gcsh_build(PairCtx& cx, GcshH& H) {
    const int lane = threadIdx.x & 31;
    const I n = cx.n, m = cx.m;
    H.n = n;
    H.m = m;
    H.nseeds = n >= GCSH_K ? (n - GCSH_K) / GCSH_K + 1 : 0;  // fixed_length_seeds, qgrams.rs:99-109
    H.ttx = n - m;  // transform(target): P(n) = 0
    H.tty = m - n;
    H.h_calls = 0;
    H.t_h = 0;
    H.hint = 0;
    H.dirty = false;
    H.nlayers = 0;
    const I ns = H.nseeds;

    // capacity for matches scales with the arena (re-run with a larger arena on overflow)
    const uint32_t avail = cx.hi_bot - cx.v_top;
    int mcap = (int)min((uint32_t)(1u << 30), avail / 180u);
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
    int log_t = 5;
    while ((1 << log_t) < 2 * ns) log_t++;
    const uint32_t tsize = 1u << log_t;
    uint32_t off_tab = arena_alloc(cx, tsize * 8u);
    const I dmin = (n - m) - ns - (GCSH_P + 2), dmax = (n - m) + ns + (GCSH_P + 2);
    uint32_t off_nm = arena_alloc(cx, (uint32_t)(dmax - dmin + 1) * 4u);
    uint32_t off_arr_s = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_arr_j = arena_alloc(cx, (uint32_t)mcap * 4u);
    uint32_t off_arr_r = arena_alloc(cx, (uint32_t)mcap * 4u);
    if (cx.status != ST_PENDING) return false;

    uint32_t* cnt = (uint32_t*)(cx.arena + off_base);  // counts, then exclusive prefix
    uint32_t* before_end = (uint32_t*)(cx.arena + off_bend);
    uint32_t* after_start = (uint32_t*)(cx.arena + off_astart);
    I* ms_i = (I*)(cx.arena + off_msi);
    I* ms_j = (I*)(cx.arena + off_msj);
    uint2* tab = (uint2*)(cx.arena + off_tab);
    NmpdView nm{(I*)(cx.arena + off_nm), dmin, dmax};
    I* arr_s = (I*)(cx.arena + off_arr_s);
    I* arr_j = (I*)(cx.arena + off_arr_j);
    uint32_t* arr_r = (uint32_t*)(cx.arena + off_arr_r);

    for (uint32_t t = lane; t < tsize; t += 32) tab[t] = make_uint2(0u, HT_EMPTY);
    for (I t = lane; t <= dmax - dmin; t += 32) nm.v[t] = INT32_MAX;
    for (I t = lane; t < ns + 2; t += 32) cnt[t] = 0u;
    __syncwarp();

    // ---- hash the seeds of a (hash_to_smallvec, exact.rs:48-55). Key: bit t = rank bit0 of char t, bit k+t = rank bit1.
    for (I s0 = 0; s0 < ns; s0 += 32) {
        I s = s0 + lane;
        if (s < ns) {
            const uint2 w = extract32(cx.aprof, s * GCSH_K);  // planes are stored negated
            const uint32_t km = (1u << GCSH_K) - 1u;
            const uint32_t key = (~w.x & km) | ((~w.y & km) << GCSH_K);
            uint32_t slot = kmer_hash(key, log_t);
            unsigned long long want = ((unsigned long long)(uint32_t)s << 32) | key;  // uint2{key, seed}
            for (;;) {
                unsigned long long old = atomicCAS((unsigned long long*)&tab[slot], ((unsigned long long)HT_EMPTY << 32), want);
                if (old == ((unsigned long long)HT_EMPTY << 32)) break;
                slot = (slot + 1) & (tsize - 1);
            }
        }
    }
    __syncwarp();

    // ---- scan all windows of b right to left (b_qgrams_rev, qgrams.rs:81-97), push matches in arrival order
    int M = 0;
    // next larger seed with this key after `after` (or the smallest when after < 0); returns count via cnt_out
    auto probe = [&](uint32_t key, I after, int& cnt_out) -> I {
        uint32_t slot = kmer_hash(key, log_t);
        I best = INT32_MAX;
        int c = 0;
        for (;;) {
            uint2 e = tab[slot];
            if (e.y == HT_EMPTY) break;
            if (e.x == key) {
                c++;
                if ((I)e.y > after && (I)e.y < best) best = (I)e.y;
            }
            slot = (slot + 1) & (tsize - 1);
        }
        cnt_out = c;
        return best;
    };
    const uint32_t kmask = (1u << GCSH_K) - 1u;
    for (I jt = m - GCSH_K; jt >= 0; jt -= 32) {
        const I j = jt - lane;
        uint32_t key = 0;
        int c = 0;
        I seed0 = INT32_MAX;
        if (j >= 0) {
            const uint2 w = extract32(cx.bprof, j);  // planes are stored negated
            key = (~w.x & kmask) | ((~w.y & kmask) << GCSH_K);
            seed0 = probe(key, -1, c);
        }
        unsigned bal = __ballot_sync(FULL, c > 0);
        while (bal) {
            const int l = __ffs(bal) - 1;
            bal &= bal - 1;
            const I jj = jt - l;
            const uint32_t kk = __shfl_sync(FULL, key, l);
            const int cc = __shfl_sync(FULL, c, l);
            I seed = __shfl_sync(FULL, seed0, l);
            for (int t = 0; t < cc; t++) {
                const I si = seed * GCSH_K;
                // MatchBuilder::push (matches.rs:205-247)
                const Cost p = H.pot(si);
                const bool pass_t = (si - jj - p <= H.ttx) && (jj - si - p <= H.tty);
                if (pass_t && dev_preserve_for_local_pruning(H, cx.aprof, cx.bprof, si, jj, nm)) {
                    if (M >= mcap) {
                        cx.status = ST_OVERFLOW;
                        return false;
                    }
                    const I d = si - jj;
                    if (d >= dmin && d <= dmax) {
                        if (lane == 0) nm.v[d - dmin] = si;
                    }
                    if (lane == 0) {
                        arr_s[M] = seed;
                        arr_j[M] = jj;
                        arr_r[M] = cnt[seed];
                        cnt[seed] += 1;
                    }
                    M++;
                    __syncwarp();
                }
                if (t + 1 < cc) {
                    int dummy;
                    seed = probe(kk, seed, dummy);
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
            I s = arr_s[t];
            uint32_t c_s = cnt[s + 1] - cnt[s];
            uint32_t pos = cnt[s] + (c_s - 1u - arr_r[t]);
            ms_i[pos] = s * GCSH_K;
            ms_j[pos] = arr_j[t];
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

}  // namespace apa
