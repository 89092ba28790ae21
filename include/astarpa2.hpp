// C++ host mirror of the reference's Rust API for the A*PA2 path, header-only, over the C-ABI of libastarpa_c.so
// (include/astarpa_b200.h). The image has no Rust toolchain, so this is the compiled-language host layer a C++
// caller (and tools/pa_bin.cpp) uses instead of the Rust crates:
//
//   astarpa2::AstarPa2::simple(trace) / ::full(trace)   AstarPa2Params::{simple,full}().make_aligner(trace)
//                                                        astarpa2/src/params.rs:70-132
//   AstarPa2::align(a, b) -> (Cost, optional<Cigar>)    pa_types::Aligner::align as implemented at
//                                                        astarpa2/src/lib.rs:210-215 (CIGAR iff trace)
//   AstarPa2::cost(a, b)                                 AstarPa2::cost, astarpa2/src/lib.rs:177-179
//   astarpa2::astarpa2_simple / astarpa2_full            astarpa2/src/lib.rs:44-53
//   astarpa2::AstarPa2Params{...}.make_aligner(trace)    AstarPa2Params::make_aligner (params.rs:132-226): other domains,
//                                                        heuristics, doubling types, block widths (general kernel)
//   AstarPa2::align_with_stats(a, b)                     AstarPa2StatsAligner::align_with_stats, astarpa2/src/lib.rs:200-208
//   astarpa2::search(pattern, text, unmatched_cost)     pa_bitpacking::search(..).out, pa-bitpacking/src/search.rs:46-118
//   AstarPa2::align_batch(pairs)                         no reference counterpart: one call carries a whole batch
//                                                        to the GPU (SURVEY 8b "needed extension")
//   Cigar::{to_string, parse, verify}                    pa_types::Cigar (external crate; text format pinned by
//                                                        astarpa-c/example.cpp:16: "=I4=X=", count omitted when 1)
//
// Errors: the reference panics (invalid bases, internal assertions); here every failure throws astarpa2::Error
// carrying apa_last_error(). There is no CPU fallback: without a usable B200 the constructor throws.
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "astarpa_b200.h"

namespace astarpa2 {

using Cost = int32_t;               // pa_types::Cost
using Seq = std::string_view;       // pa_types::Seq = &[u8]; bytes over ACGT

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

enum class CigarOp : uint8_t { Match = '=', Sub = 'X', Del = 'D', Ins = 'I' };
struct CigarElem {
    CigarOp op;
    uint32_t cnt;
    bool operator==(const CigarElem& o) const { return op == o.op && cnt == o.cnt; }
};

// Run-length CIGAR. I consumes a base of b, D a base of a (astarpa2/src/blocks/trace.rs:176-204).
struct Cigar {
    std::vector<CigarElem> ops;

    std::string to_string() const {
        std::string s;
        for (const auto& e : ops) {
            if (e.cnt != 1) s += std::to_string(e.cnt);
            s += (char)e.op;
        }
        return s;
    }
    static Cigar parse(std::string_view text) {
        Cigar c;
        uint64_t cnt = 0;
        bool have = false;
        for (char ch : text) {
            if (ch >= '0' && ch <= '9') {
                cnt = cnt * 10 + (uint64_t)(ch - '0');
                have = true;
                if (cnt > 0x3fffffffu) throw Error(APA_ERR_BAD_INPUT, "CIGAR count too large");
            } else if (ch == '=' || ch == 'X' || ch == 'D' || ch == 'I') {
                if (have && cnt == 0) throw Error(APA_ERR_BAD_INPUT, "CIGAR element with count 0");
                c.ops.push_back(CigarElem{(CigarOp)ch, have ? (uint32_t)cnt : 1u});
                cnt = 0;
                have = false;
            } else {
                throw Error(APA_ERR_BAD_INPUT, std::string("unexpected character in CIGAR: ") + ch);
            }
        }
        if (have) throw Error(APA_ERR_BAD_INPUT, "CIGAR ends with a count");
        return c;
    }
    // Cigar::verify: walks the path over (a, b); returns its unit cost, or -1 if it is not a valid alignment of a and b.
    int64_t verify(Seq a, Seq b) const {
        size_t i = 0, j = 0;
        int64_t cost = 0;
        for (const auto& e : ops) {
            for (uint32_t t = 0; t < e.cnt; t++) {
                switch (e.op) {
                    case CigarOp::Match:
                        if (i >= a.size() || j >= b.size() || a[i] != b[j]) return -1;
                        i++, j++;
                        break;
                    case CigarOp::Sub:
                        if (i >= a.size() || j >= b.size() || a[i] == b[j]) return -1;
                        i++, j++, cost++;
                        break;
                    case CigarOp::Del:
                        if (i >= a.size()) return -1;
                        i++, cost++;
                        break;
                    case CigarOp::Ins:
                        if (j >= b.size()) return -1;
                        j++, cost++;
                        break;
                }
            }
        }
        return (i == a.size() && j == b.size()) ? cost : -1;
    }
};

using Alignment = std::pair<Cost, std::optional<Cigar>>;

struct BatchResult {
    std::vector<Cost> costs;
    std::vector<std::string> cigars;  // CIGAR text per pair; empty vector when trace == false
    apa_batch_stats stats{};
    std::vector<apa_pair_stats> pair_stats;  // filled by align_batch_with_stats only
};

class AstarPa2;
// Flat parameters (astarpa2/src/params.rs:8-42) = the C-ABI struct; simple() / full() / nw() as in params.rs:46-128.
struct AstarPa2Params : apa_params {
    static AstarPa2Params preset(int which) {
        AstarPa2Params q;
        int rc = apa_params_preset(which, &q);
        if (rc != APA_OK) throw Error(rc, std::string("apa_params_preset: ") + apa_last_error());
        return q;
    }
    static AstarPa2Params simple() { return preset(APA_PRESET_SIMPLE); }
    static AstarPa2Params full() { return preset(APA_PRESET_FULL); }
    static AstarPa2Params nw() {  // params.rs:46-68
        AstarPa2Params q = simple();
        q.domain = APA_DOMAIN_FULL;
        q.heuristic = APA_HEURISTIC_NONE;
        q.doubling = APA_DOUBLING_NONE;
        q.dt_trace = q.sparse_h = q.prune = 0;
        return q;
    }
    // The serde JSON form of the reference / pa-bench (params.rs:7-42); throws Error naming a field this engine does not serve.
    static AstarPa2Params from_json(const std::string& json) {
        AstarPa2Params q;
        char err[512] = {0};
        int rc = apa_params_from_json(json.c_str(), &q, err, sizeof err);
        if (rc != APA_OK) throw Error(rc, err);
        return q;
    }
    inline AstarPa2 make_aligner(bool trace, int device = 0) const;
};

class AstarPa2 {
  public:
    enum Preset { Simple = APA_PRESET_SIMPLE, Full = APA_PRESET_FULL };

    AstarPa2(Preset preset, bool trace, int device = 0) : preset_(preset), trace_(trace) {
        int rc = apa_engine_create(device, &engine_);
        if (rc != APA_OK) throw Error(rc, std::string("apa_engine_create: ") + apa_last_error());
    }
    // Explicit parameters: served by the general kernel (apa_align_batch_params).
    AstarPa2(const apa_params& params, bool trace, int device = 0) : preset_(Full), trace_(trace), params_(params), has_params_(true) {
        int rc = apa_engine_create(device, &engine_);
        if (rc != APA_OK) throw Error(rc, std::string("apa_engine_create: ") + apa_last_error());
    }
    static AstarPa2 simple(bool trace = true, int device = 0) { return AstarPa2(Simple, trace, device); }
    static AstarPa2 full(bool trace = true, int device = 0) { return AstarPa2(Full, trace, device); }
    AstarPa2(const AstarPa2&) = delete;
    AstarPa2& operator=(const AstarPa2&) = delete;
    AstarPa2(AstarPa2&& o) noexcept : engine_(o.engine_), preset_(o.preset_), trace_(o.trace_), params_(o.params_), has_params_(o.has_params_) {
        o.engine_ = nullptr;
    }
    ~AstarPa2() {
        if (engine_) apa_engine_destroy(engine_);
    }

    bool trace() const { return trace_; }
    Preset preset() const { return preset_; }

    // Aligner::align (astarpa2/src/lib.rs:210-215)
    Alignment align(Seq a, Seq b) {
        const std::pair<Seq, Seq> one[1] = {{a, b}};
        BatchResult r = align_batch(one, 1, trace_);
        if (!trace_) return {r.costs[0], std::nullopt};
        return {r.costs[0], Cigar::parse(r.cigars[0])};
    }
    // AstarPa2::cost (astarpa2/src/lib.rs:177-179)
    Cost cost(Seq a, Seq b) {
        const std::pair<Seq, Seq> one[1] = {{a, b}};
        return align_batch(one, 1, false).costs[0];
    }
    BatchResult align_batch(const std::vector<std::pair<Seq, Seq>>& pairs) { return align_batch(pairs.data(), pairs.size(), trace_); }

    // AstarPa2StatsAligner::align_with_stats (astarpa2/src/lib.rs:200-208): the alignment plus the per-pair counters.
    std::pair<Alignment, apa_pair_stats> align_with_stats(Seq a, Seq b) {
        BatchResult r = align_batch_with_stats({{a, b}});
        Alignment al{r.costs[0], std::nullopt};
        if (trace_) al.second = Cigar::parse(r.cigars[0]);
        return {std::move(al), r.pair_stats[0]};
    }
    // Through an HBM-resident batch (upload / run / download), which is where the per-pair counters live.
    BatchResult align_batch_with_stats(const std::vector<std::pair<Seq, Seq>>& pairs) {
        const size_t n = pairs.size();
        std::string a_all, b_all;
        std::vector<int64_t> a_off(n + 1, 0), b_off(n + 1, 0);
        for (size_t p = 0; p < n; p++) {
            a_all.append(pairs[p].first);
            b_all.append(pairs[p].second);
            a_off[p + 1] = (int64_t)a_all.size();
            b_off[p + 1] = (int64_t)b_all.size();
        }
        apa_batch* bt = nullptr;
        auto check = [&](int rc, const char* what) {
            if (rc == APA_OK) return;
            std::string msg = std::string(what) + ": " + apa_last_error();
            if (bt) apa_batch_free(engine_, bt);
            throw Error(rc, msg);
        };
        check(apa_batch_upload(engine_, n, (const uint8_t*)a_all.data(), a_off.data(), (const uint8_t*)b_all.data(), b_off.data(), &bt),
              "apa_batch_upload");
        check(has_params_ ? apa_batch_run_params(engine_, bt, &params_, trace_ ? 1 : 0) : apa_batch_run(engine_, bt, (int)preset_, trace_ ? 1 : 0),
              "apa_batch_run");
        BatchResult r;
        std::vector<int64_t> costs(n ? n : 1), off(n ? n : 1), len(n ? n : 1);
        char* pool = nullptr;
        check(trace_ ? apa_batch_download(engine_, bt, costs.data(), &pool, off.data(), len.data())
                     : apa_batch_download(engine_, bt, costs.data(), nullptr, nullptr, nullptr),
              "apa_batch_download");
        r.costs.assign(costs.begin(), costs.begin() + n);
        if (trace_ && pool) {
            r.cigars.resize(n);
            for (size_t p = 0; p < n; p++) r.cigars[p].assign(pool + off[p], (size_t)len[p]);
        }
        apa_free(pool);
        r.pair_stats.resize(n ? n : 1);
        check(apa_batch_download_pair_stats(engine_, bt, r.pair_stats.data()), "apa_batch_download_pair_stats");
        r.pair_stats.resize(n);
        check(apa_batch_get_stats(bt, &r.stats), "apa_batch_get_stats");
        apa_batch_free(engine_, bt);
        return r;
    }

    // Concatenated form (what the C-ABI takes): pair p is a_all[a_off[p] .. a_off[p+1]) vs b_all[b_off[p] .. b_off[p+1]).
    BatchResult align_batch_concat(const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all, const int64_t* b_off, size_t n,
                                   bool trace) {
        BatchResult r;
        std::vector<int64_t> costs(n ? n : 1), off(n ? n : 1), len(n ? n : 1);
        char* pool = nullptr;
        int rc = has_params_ ? apa_align_batch_params(engine_, &params_, trace ? 1 : 0, n, a_all, a_off, b_all, b_off, costs.data(), &pool,
                                                      off.data(), len.data(), &r.stats)
                             : apa_align_batch(engine_, (int)preset_, trace ? 1 : 0, n, a_all, a_off, b_all, b_off, costs.data(), &pool,
                                               off.data(), len.data(), &r.stats);
        if (rc != APA_OK) throw Error(rc, std::string("apa_align_batch: ") + apa_last_error());
        r.costs.resize(n);
        for (size_t p = 0; p < n; p++) r.costs[p] = (Cost)costs[p];
        if (trace && pool) {
            r.cigars.resize(n);
            for (size_t p = 0; p < n; p++) r.cigars[p].assign(pool + off[p], (size_t)len[p]);
        }
        apa_free(pool);
        return r;
    }

  private:
    BatchResult align_batch(const std::pair<Seq, Seq>* pairs, size_t n, bool trace) {
        std::string a_all, b_all;
        std::vector<int64_t> a_off(n + 1, 0), b_off(n + 1, 0);
        size_t ta = 0, tb = 0;
        for (size_t p = 0; p < n; p++) ta += pairs[p].first.size(), tb += pairs[p].second.size();
        a_all.reserve(ta);
        b_all.reserve(tb);
        for (size_t p = 0; p < n; p++) {
            a_all.append(pairs[p].first);
            b_all.append(pairs[p].second);
            a_off[p + 1] = (int64_t)a_all.size();
            b_off[p + 1] = (int64_t)b_all.size();
        }
        return align_batch_concat((const uint8_t*)a_all.data(), a_off.data(), (const uint8_t*)b_all.data(), b_off.data(), n, trace);
    }
    apa_engine* engine_ = nullptr;
    Preset preset_;
    bool trace_;
    apa_params params_{};
    bool has_params_ = false;
};
inline AstarPa2 AstarPa2Params::make_aligner(bool trace, int device) const { return AstarPa2(static_cast<const apa_params&>(*this), trace, device); }

// pa_bitpacking::search(pattern, text, unmatched_cost).out (pa-bitpacking/src/search.rs:46-118): the costs along the bottom row
// and up the right column of the semi-global DP; |pattern| + |text| + 1 values. Pattern may hold N, *, Y, R.
inline std::vector<Cost> search(Seq pattern, Seq text, float unmatched_cost = 0.0f, int device = 0) {
    apa_engine* e = nullptr;
    int rc = apa_engine_create(device, &e);
    if (rc != APA_OK) throw Error(rc, std::string("apa_engine_create: ") + apa_last_error());
    std::vector<Cost> out(pattern.size() + text.size() + 1);
    rc = apa_search(e, (const uint8_t*)pattern.data(), pattern.size(), (const uint8_t*)text.data(), text.size(), unmatched_cost, out.data());
    std::string msg = rc == APA_OK ? "" : std::string("apa_search: ") + apa_last_error();
    apa_engine_destroy(e);
    if (rc != APA_OK) throw Error(rc, msg);
    return out;
}

// SearchResult of pa_bitpacking::search with its trace() (pa-bitpacking/src/search.rs:5-16,135-230): `out` as above; trace(idx)
// is the alignment of the pattern that ends at out[idx] - CIGAR plus the position where the walk stopped and where it ended
// (i = text column, j = pattern row).
struct SearchTrace {
    Cigar cigar;
    std::pair<int32_t, int32_t> start, end;
    Cost cost;
};
class SearchResult {
  public:
    std::vector<Cost> out;
    SearchResult(Seq pattern, Seq text, float unmatched_cost = 0.0f, int device = 0)
        : pattern_(pattern), text_(text), unmatched_cost_(unmatched_cost), device_(device) {
        out = search(pattern, text, unmatched_cost, device);
    }
    SearchTrace trace(size_t idx) const {
        apa_engine* e = nullptr;
        int rc = apa_engine_create(device_, &e);
        if (rc != APA_OK) throw Error(rc, std::string("apa_engine_create: ") + apa_last_error());
        std::string text(2 * (pattern_.size() + text_.size()) + 16, '\0');
        int32_t pos[5] = {0, 0, 0, 0, 0};
        rc = apa_search_trace(e, (const uint8_t*)pattern_.data(), pattern_.size(), (const uint8_t*)text_.data(), text_.size(), unmatched_cost_,
                              idx, text.data(), text.size(), pos);
        std::string msg = rc == APA_OK ? "" : std::string("apa_search_trace: ") + apa_last_error();
        apa_engine_destroy(e);
        if (rc != APA_OK) throw Error(rc, msg);
        text.resize(strlen(text.c_str()));
        return SearchTrace{Cigar::parse(text), {pos[0], pos[1]}, {pos[2], pos[3]}, (Cost)pos[4]};
    }

  private:
    std::string pattern_, text_;
    float unmatched_cost_;
    int device_;
};

// One call over several GPUs of the box (apa_align_batch_multi): contiguous shards of the batch, one per device, results in
// input order. The engines are the process-wide ones of the library.
inline BatchResult align_batch_multi(const std::vector<int>& devices, AstarPa2::Preset preset, bool trace,
                                     const std::vector<std::pair<Seq, Seq>>& pairs) {
    const size_t n = pairs.size();
    std::string a_all, b_all;
    std::vector<int64_t> a_off(n + 1, 0), b_off(n + 1, 0);
    for (size_t p = 0; p < n; p++) {
        a_all.append(pairs[p].first);
        b_all.append(pairs[p].second);
        a_off[p + 1] = (int64_t)a_all.size();
        b_off[p + 1] = (int64_t)b_all.size();
    }
    BatchResult r;
    std::vector<int64_t> costs(n ? n : 1), off(n ? n : 1), len(n ? n : 1);
    char* pool = nullptr;
    int rc = apa_align_batch_multi(devices.data(), (int)devices.size(), (int)preset, trace ? 1 : 0, n, (const uint8_t*)a_all.data(), a_off.data(),
                                   (const uint8_t*)b_all.data(), b_off.data(), costs.data(), &pool, off.data(), len.data(), nullptr);
    if (rc != APA_OK) throw Error(rc, std::string("apa_align_batch_multi: ") + apa_last_error());
    r.costs.assign(costs.begin(), costs.begin() + n);
    if (trace && pool) {
        r.cigars.resize(n);
        for (size_t p = 0; p < n; p++) r.cigars[p].assign(pool + off[p], (size_t)len[p]);
    }
    apa_free(pool);
    return r;
}

// astarpa2::astarpa2_simple / astarpa2_full (astarpa2/src/lib.rs:44-53): a fresh aligner per call, with trace.
inline std::pair<Cost, Cigar> astarpa2_simple(Seq a, Seq b) {
    auto r = AstarPa2::simple(true).align(a, b);
    return {r.first, std::move(*r.second)};
}
inline std::pair<Cost, Cigar> astarpa2_full(Seq a, Seq b) {
    auto r = AstarPa2::full(true).align(a, b);
    return {r.first, std::move(*r.second)};
}

}  // namespace astarpa2
