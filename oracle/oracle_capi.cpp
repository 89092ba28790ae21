// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).
// C-ABI surface of the CPU restatement, loaded with ctypes by tests/, __graft_entry__.smoke() and by
// bench.py's cpu_baseline / --impl reference legs. Nothing in astarpa_pairwise_aligner_b200/ links this.
//
// Parity status: the Rust reference cannot be compiled in this image (no cargo/rustc; un-vendored deps:
// pa-types, pa-generate, bio, triple_accel). This restatement is pinned against the reference's own
// golden vectors / known-answer tests (tests/test_oracle_golden.py lists each with its file:line):
// cost == Levenshtein on pa-test's 8 pairs and astarpa's regression pairs, example.c cost 2,
// example.cpp CIGAR text format, the pa-bitpacking block KAT, qgram KATs. Exact A*PA2 CIGAR strings and
// band shapes are NOT pinned by any reference test ("parity unpinned" for those).
#include <malloc.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "astarpa2.hpp"
#include "search.hpp"

using namespace oracle;

static AstarPa2Params preset_params(int preset) {
    // 0 = astarpa2_simple, 1 = astarpa2_full (astarpa2/src/lib.rs:44-53)
    // Further configurations mirror astarpa2/src/tests.rs:19-119 where expressible with this slice.
    switch (preset) {
        case 0: return AstarPa2Params::simple();
        case 1: return AstarPa2Params::full();
        case 2: {  // tests.rs:81-89 nw_prune: GCSH k=15 exact, prune start, bw=256, BlockParams::default()
            AstarPa2Params q = AstarPa2Params::full();
            q.k = 15;
            q.p = 0;
            q.doubling_start_gap = true;  // DoublingType::band_doubling() = BandDoubling{Gap, 2.0}
            q.front = BlockParams{};
            return q;
        }
        case 3: {  // tests.rs:91-104 dt_trace: as 2 with dt_trace
            AstarPa2Params q = AstarPa2Params::full();
            q.k = 15;
            q.p = 0;
            q.doubling_start_gap = true;
            q.front = BlockParams{};
            q.front.dt_trace = true;
            return q;
        }
        case 4: {  // tests.rs:58-66 band_doubling_edlib: GapCost, bw=64
            AstarPa2Params q = AstarPa2Params::simple();
            q.doubling_start_gap = true;
            q.block_width = 64;
            q.front = BlockParams{};
            q.prune = true;  // ..nw() has prune: true; a no-op for GapCost (heuristic.rs:149-155)
            return q;
        }
        case 5: {  // tests.rs:106-119 incremental_doubling: GapCost, bw=64, dt_trace + incremental
            AstarPa2Params q = AstarPa2Params::simple();
            q.doubling_start_gap = true;
            q.block_width = 64;
            q.front = BlockParams{};
            q.front.dt_trace = true;
            q.front.incremental_doubling = true;
            q.prune = true;
            return q;
        }
        case 6: {  // tests.rs:48-56 band_doubling_dijkstra: NoCost, bw=64
            AstarPa2Params q = AstarPa2Params::simple();
            q.heuristic = HeuristicKind::None;
            q.doubling_start_gap = true;
            q.block_width = 64;
            q.front = BlockParams{};
            q.prune = true;
            return q;
        }
        case 7: {  // tests.rs:24-32 band_doubling_gapgap: Domain::GapGap bw=64
            AstarPa2Params q = AstarPa2Params::simple();
            q.domain = DomainKind::GapGap;
            q.doubling_start_gap = true;
            q.block_width = 64;
            q.front = BlockParams{};
            q.prune = true;
            return q;
        }
        case 8: {  // tests.rs:34-46 dt_trace_gapgap: Domain::GapGap bw=256 dt_trace
            AstarPa2Params q = AstarPa2Params::simple();
            q.domain = DomainKind::GapGap;
            q.doubling_start_gap = true;
            q.block_width = 256;
            q.front = BlockParams{};
            q.front.dt_trace = true;
            q.prune = true;
            return q;
        }
        case 9: {  // tests.rs:19-22 full(): Domain::Full, no doubling, bw=1
            AstarPa2Params q = AstarPa2Params::simple();
            q.domain = DomainKind::Full;
            q.doubling = false;
            q.block_width = 1;
            q.front = BlockParams{};
            q.prune = true;
            return q;
        }
        // Further configurations the parameter space allows (astarpa2/src/params.rs, band.rs), used to pin the general GPU kernel.
        case 10: {  // Domain::GapStart, bw = 64, BandDoubling{Gap, 2}
            AstarPa2Params q = AstarPa2Params::simple();
            q.domain = DomainKind::GapStart;
            q.doubling_start_gap = true;
            q.block_width = 64;
            q.front = BlockParams{};
            q.front.dt_trace = true;
            return q;
        }
        case 11: {  // GapCost with LinearSearch{start: Gap, delta: 48}, bw = 32
            AstarPa2Params q = AstarPa2Params::simple();
            q.doubling_start_gap = true;
            q.linear = true;
            q.delta = 48;
            q.block_width = 32;
            return q;
        }
        case 12: {  // GCSH k = 12, p = 14 like astarpa2_full, but dense h (sparse_h = false), bw = 32, start Zero, factor 1.5, fr_drop 20
            AstarPa2Params q = AstarPa2Params::full();
            q.sparse_h = false;
            q.block_width = 32;
            q.doubling_start_zero = true;
            q.factor = 1.5f;
            q.front.fr_drop = 20;
            return q;
        }
        case 13: {  // AstarPa2Params::nw() (params.rs:46-68): Domain::Full, no doubling, bw = 256, no dt_trace
            AstarPa2Params q = AstarPa2Params::simple();
            q.domain = DomainKind::Full;
            q.doubling = false;
            q.block_width = 256;
            q.front = BlockParams{true, true, false, false, false, 40, 20};
            q.sparse_h = false;
            q.prune = false;
            return q;
        }
        case 14: {  // Dijkstra (NoCost) with bw = 1 and dt_trace, GCSH-free pruning flag on
            AstarPa2Params q = AstarPa2Params::simple();
            q.heuristic = HeuristicKind::None;
            q.doubling_start_gap = true;
            q.block_width = 1;
            q.front = BlockParams{};
            q.front.dt_trace = true;
            q.prune = true;
            return q;
        }
        case 15: {  // GCSH k = 8, p = 3, prune off, bw = 100 (not a multiple of 32), LinearSearch{H0, 200}
            AstarPa2Params q = AstarPa2Params::full();
            q.k = 8;
            q.p = 3;
            q.prune = false;
            q.block_width = 100;
            q.linear = true;
            q.delta = 200;
            return q;
        }
        case 16:    // the heuristic of the legacy C entry points (astarpa-c/src/lib.rs:54-95): GCSH(MatchConfig::new(k, r = 1)) =
        case 17:    // exact matches of length k, no local pruning (matches.rs:404-410), Prune::Start - on the A*PA2 block engine
        case 18: {  // with the knobs of full(). k = 8 / 12 / 15.
            AstarPa2Params q = AstarPa2Params::full();
            q.k = preset == 16 ? 8 : (preset == 17 ? 12 : 15);
            q.p = 0;
            return q;
        }
    }
    throw RefPanic("unknown preset");
}

struct OracleStats {  // mirrored by tests/oracle_lib.py
    int64_t f_max_tries, num_blocks, computed_lanes, computed_cells, h_calls, num_matches, h0;
    int64_t dt_trace_tries, dt_trace_success, fill_tries, fill_success;
};

static void fill_stats(OracleStats* st, const AstarPa2Stats& s) {
    if (!st) return;
    st->f_max_tries = (int64_t)s.f_max_tries;
    st->num_blocks = (int64_t)s.block_stats.num_blocks;
    st->computed_lanes = (int64_t)s.block_stats.computed_lanes;
    st->computed_cells = (int64_t)s.block_stats.computed_cells;
    st->h_calls = (int64_t)s.h_calls;
    st->num_matches = (int64_t)s.num_matches;
    st->h0 = s.h0;
    st->dt_trace_tries = (int64_t)s.trace_stats.dt_trace_tries;
    st->dt_trace_success = (int64_t)s.trace_stats.dt_trace_success;
    st->fill_tries = (int64_t)s.trace_stats.fill_tries;
    st->fill_success = (int64_t)s.trace_stats.fill_success;
}

extern "C" {

// Returns cost (>= 0), or -1 when the restated reference would panic (message in err, if given).
// *cigar_out is malloc'd NUL-terminated text (free with oracle_free) when trace != 0.
int64_t oracle_align(int preset, int trace, const uint8_t* a, size_t n, const uint8_t* b, size_t m, char** cigar_out,
                     size_t* cigar_len, OracleStats* stats, int self_check, char* err, size_t err_cap) {
    if (cigar_out) *cigar_out = nullptr;
    if (cigar_len) *cigar_len = 0;
    try {
        AlignResult r = cost_or_align(a, n, b, m, preset_params(preset), trace != 0, nullptr, self_check != 0);
        fill_stats(stats, r.stats);
        if (r.has_cigar && cigar_out) {
            std::string s = r.cigar.to_string();
            char* p = (char*)malloc(s.size() + 1);
            memcpy(p, s.c_str(), s.size() + 1);
            *cigar_out = p;
            if (cigar_len) *cigar_len = s.size();
        }
        return r.cost;
    } catch (const std::exception& e) {
        if (err && err_cap) snprintf(err, err_cap, "%s", e.what());
        return -1;
    }
}

void oracle_free(void* p) { free(p); }

// Per-pass band log: for each pass: f_max, nblocks, then nblocks x (j_s, j_e, fixed_s, fixed_e).
// Returns the number of int32 written (or needed if > cap), -1 on panic.
int64_t oracle_align_log(int preset, int trace, const uint8_t* a, size_t n, const uint8_t* b, size_t m, int32_t* out,
                         size_t cap) {
    try {
        std::vector<PassLog> log;
        cost_or_align(a, n, b, m, preset_params(preset), trace != 0, &log, false);
        size_t w = 0;
        auto put = [&](int32_t x) {
            if (w < cap) out[w] = x;
            w++;
        };
        put((int32_t)log.size());
        for (auto& pl : log) {
            put(pl.f_max);
            put((int32_t)pl.fixed_ranges.size());
            for (size_t t = 0; t < pl.fixed_ranges.size(); t++) {
                put(pl.j_ranges[t].s);
                put(pl.j_ranges[t].e);
                put(pl.fixed_ranges[t].s);
                put(pl.fixed_ranges[t].e);
            }
        }
        return (int64_t)w;
    } catch (const std::exception&) {
        return -1;
    }
}

int64_t oracle_levenshtein(const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
    return levenshtein_bitvector(a, n, b, m);
}
int64_t oracle_levenshtein_dp(const uint8_t* a, size_t n, const uint8_t* b, size_t m) { return levenshtein_dp(a, n, b, m); }

// Parse CIGAR text ("=I4=X=") and verify it against a, b (Cigar::verify). Returns cost or -1.
int64_t oracle_cigar_verify(const char* cigar, size_t len, const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
    Cigar c;
    size_t t = 0;
    while (t < len) {
        int64_t cnt = 0;
        bool has = false;
        while (t < len && cigar[t] >= '0' && cigar[t] <= '9') {
            cnt = cnt * 10 + (cigar[t] - '0');
            has = true;
            t++;
        }
        if (t >= len) return -1;
        if (!has) cnt = 1;
        if (has && cnt <= 1) return -1;  // to_string never prints a count of 0 or 1
        CigarOp op;
        switch (cigar[t]) {
            case '=': op = OpMatch; break;
            case 'X': op = OpSub; break;
            case 'I': op = OpIns; break;
            case 'D': op = OpDel; break;
            default: return -1;
        }
        t++;
        if (!c.ops.empty() && c.ops.back().op == op) return -1;  // push_elem would have merged
        c.ops.push_back(CigarElem{op, (I)cnt});
    }
    return c.verify(a, n, b, m);
}

// pa-bitpacking rectangle (simd::compute semantics): a[na] x b words; h_in/h_out as (p,m) bit pairs packed
// one byte per column (bit0=p, bit1=m); v as interleaved u64 pairs. Returns sum of bottom deltas.
int64_t oracle_bp_compute(const uint8_t* a, size_t na, const uint8_t* b, size_t mb, uint8_t* h, uint64_t* v) {
    std::vector<Bits> pa, pb;
    bitprofile_build(a, na, b, mb, pa, pb);
    std::vector<H> hh(na);
    for (size_t i = 0; i < na; i++) hh[i] = H{(uint64_t)(h[i] & 1), (uint64_t)((h[i] >> 1) & 1)};
    std::vector<V> vv(pb.size());
    for (size_t j = 0; j < pb.size(); j++) vv[j] = V{v[2 * j], v[2 * j + 1]};
    Cost s = bp_compute(pa.data(), na, pb.data(), pb.size(), hh.data(), vv.data());
    for (size_t i = 0; i < na; i++) h[i] = (uint8_t)(hh[i].p | (hh[i].m << 1));
    for (size_t j = 0; j < pb.size(); j++) {
        v[2 * j] = vv[j].p;
        v[2 * j + 1] = vv[j].m;
    }
    return s;
}

// SearchResult::trace(idx) (search.rs:135-230): CIGAR text into cigar_out (NUL-terminated, cap bytes), pos_out = {start.i, start.j,
// end.i, end.j, cost}. Returns strlen, -1 on a reference panic, -2 when cap is too small.
int64_t oracle_search_trace(const uint8_t* pattern, size_t np, const uint8_t* text, size_t nt, float unmatched_cost, uint64_t idx,
                            char* cigar_out, size_t cap, int32_t* pos_out) {
    try {
        SearchTrace tr = search_trace(pattern, np, text, nt, unmatched_cost, (size_t)idx);
        if (tr.cigar.size() + 1 > cap) return -2;
        memcpy(cigar_out, tr.cigar.c_str(), tr.cigar.size() + 1);
        pos_out[0] = tr.start.i, pos_out[1] = tr.start.j, pos_out[2] = tr.end.i, pos_out[3] = tr.end.j, pos_out[4] = tr.cost;
        return (int64_t)tr.cigar.size();
    } catch (const RefPanic&) {
        return -1;
    }
}

// The reference's micro-benchmark of the block kernel (pa-bitpacking/benches/nw/main.rs:139-159): `reps` evaluations of the
// a[na] x b[mb] rectangle with all-(+1) input deltas, profile built once. Returns the last bottom-delta sum (keeps the loop live).
int64_t oracle_bp_compute_bench(const uint8_t* a, size_t na, const uint8_t* b, size_t mb, int reps) {
    std::vector<Bits> pa, pb;
    bitprofile_build(a, na, b, mb, pa, pb);
    std::vector<H> hh(na);
    std::vector<V> vv(pb.size());
    Cost s = 0;
    for (int r = 0; r < reps; r++) {
        for (size_t i = 0; i < na; i++) hh[i] = H{1, 0};
        for (size_t j = 0; j < pb.size(); j++) vv[j] = V{~0ull, 0ull};
        s += bp_compute(pa.data(), na, pb.data(), pb.size(), hh.data(), vv.data());
        asm volatile("" ::"r"(hh.data()), "r"(vv.data()) : "memory");
    }
    return s;
}

uint64_t oracle_to_qgram(const uint8_t* s, int k) { return QGrams::to_qgram(s, k); }

// pa_bitpacking::search (search.rs:46-118): out must hold np + nt + 1 values. Returns the count, or -1 on a reference panic.
int64_t oracle_search(const uint8_t* pattern, size_t np, const uint8_t* text, size_t nt, float unmatched_cost, int32_t* out) {
    try {
        std::vector<Cost> o = search(pattern, np, text, nt, unmatched_cost);
        for (size_t i = 0; i < o.size(); i++) out[i] = o[i];
        return (int64_t)o.size();
    } catch (const RefPanic&) {
        return -1;
    }
}

// GCSH introspection: number of matches kept after transform filter + local pruning, and h(0,0).
int64_t oracle_gcsh_info(const uint8_t* a, size_t n, const uint8_t* b, size_t m, int k, int p, int64_t* h0,
                         int32_t* match_starts, size_t cap) {
    try {
        GcshI h(a, n, b, m, MatchConfig{(I)k, 1, (size_t)p});
        if (h0) *h0 = h.h0;
        size_t w = 0;
        for (auto& mt : h.matches.by_start) {
            if (2 * w + 1 < cap) {
                match_starts[2 * w] = mt.start.i;
                match_starts[2 * w + 1] = mt.start.j;
            }
            w++;
        }
        return (int64_t)w;
    } catch (const std::exception&) {
        return -1;
    }
}

// Multi-threaded batch (pairs are independent: astarpa2/src/lib.rs:50-53 builds a fresh aligner per call).
// Sequences are concatenated; offsets have n_pairs+1 entries. costs[i] = -1 on panic.
// cigar_lens (optional) receives strlen of each CIGAR; the text itself is discarded (timing leg).
// Returns wall seconds.
double oracle_align_batch(int preset, int trace, size_t n_pairs, const uint8_t* a_all, const int64_t* a_off,
                          const uint8_t* b_all, const int64_t* b_off, int n_threads, int64_t* costs,
                          int64_t* cigar_lens, int64_t* computed_cells, uint64_t* cigar_hash) {
    // Keep large vectors on the per-thread heaps instead of mmap/munmap per pair (page-fault and mmap-lock storms with
    // many threads would otherwise dominate; the Rust reference would use a pooling allocator in such a setting).
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 256);
    std::atomic<size_t> next{0};
    AstarPa2Params params = preset_params(preset);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_pairs) break;
            try {
                AlignResult r = cost_or_align(a_all + a_off[i], (size_t)(a_off[i + 1] - a_off[i]), b_all + b_off[i],
                                              (size_t)(b_off[i + 1] - b_off[i]), params, trace != 0);
                costs[i] = r.cost;
                if (computed_cells) computed_cells[i] = (int64_t)r.stats.block_stats.computed_cells;
                if (r.has_cigar) {
                    std::string s = r.cigar.to_string();
                    if (cigar_lens) cigar_lens[i] = (int64_t)s.size();
                    if (cigar_hash) {
                        uint64_t hsh = 1469598103934665603ull;  // FNV-1a
                        for (unsigned char c : s) hsh = (hsh ^ c) * 1099511628211ull;
                        cigar_hash[i] = hsh;
                    }
                } else if (cigar_lens) {
                    cigar_lens[i] = 0;
                }
            } catch (const std::exception&) {
                costs[i] = -1;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
