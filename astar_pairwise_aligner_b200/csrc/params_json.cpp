// AstarPa2Params from the serde JSON the reference and pa-bench exchange (astarpa2/src/params.rs:7-42, #[serde(deny_unknown_fields)]):
//   {"name": "...", "domain": "Full" | "GapStart" | "GapGap" | {"Astar": null},
//    "heuristic": {"type": "None" | "Zero" | "Gap" | "GCSH", "r": 2, "k": 15, "p": 0, "prune": "Start", ...},   (pa-heuristic/src/cli.rs:46-98)
//    "doubling": "None" | {"BandDoubling": {"start": "Zero" | "Gap" | "H0", "factor": 2.0}} | {"LinearSearch": {"start": .., "delta": 1.0}},
//    "block_width": 256, "front": {"sparse": true, "simd": .., "no_ilp": .., "incremental_doubling": .., "dt_trace": .., "max_g": 40, "fr_drop": 10},
//    "sparse_h": true, "prune": true, "viz": false}
// Missing fields take serde's defaults (r = 2, k = 15, p = 0, prune = Start; the BlockParams flags false / 0). Values this engine
// does not serve (other heuristics, r = 2 with GCSH, Prune::End / Both, LocalDoubling, viz, variable seed lengths) are refused with
// APA_ERR_BAD_INPUT and a message naming the field - never ignored. Host code only.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/astarpa_b200.h"

namespace {

struct JVal {
    enum Kind { Null, Bool, Num, Str, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<std::pair<std::string, std::unique_ptr<JVal>>> obj;
    const JVal* get(const char* key) const {
        for (auto& kv : obj)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
};

struct Parser {
    const char* p;
    std::string err;
    void ws() {
        while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') p++;
    }
    bool fail(const std::string& m) {
        if (err.empty()) err = m;
        return false;
    }
    bool string(std::string& out) {
        if (*p != '"') return fail("expected a string");
        p++;
        while (*p && *p != '"') {
            if (*p == '\\') {
                p++;
                if (!*p) return fail("unterminated string");
                out += (*p == 'n' ? '\n' : (*p == 't' ? '\t' : *p));
                p++;
            } else {
                out += *p++;
            }
        }
        if (*p != '"') return fail("unterminated string");
        p++;
        return true;
    }
    bool value(JVal& v) {
        ws();
        if (*p == '{') {
            v.kind = JVal::Obj;
            p++;
            ws();
            if (*p == '}') {
                p++;
                return true;
            }
            for (;;) {
                ws();
                std::string key;
                if (!string(key)) return false;
                ws();
                if (*p != ':') return fail("expected ':' after \"" + key + "\"");
                p++;
                auto child = std::make_unique<JVal>();
                if (!value(*child)) return false;
                v.obj.emplace_back(key, std::move(child));
                ws();
                if (*p == ',') {
                    p++;
                    continue;
                }
                if (*p == '}') {
                    p++;
                    return true;
                }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '"') {
            v.kind = JVal::Str;
            return string(v.str);
        }
        if (!strncmp(p, "true", 4)) {
            v.kind = JVal::Bool, v.b = true, p += 4;
            return true;
        }
        if (!strncmp(p, "false", 5)) {
            v.kind = JVal::Bool, v.b = false, p += 5;
            return true;
        }
        if (!strncmp(p, "null", 4)) {
            v.kind = JVal::Null, p += 4;
            return true;
        }
        if (*p == '-' || (*p >= '0' && *p <= '9')) {
            char* end = nullptr;
            v.kind = JVal::Num;
            v.num = strtod(p, &end);
            if (end == p) return fail("bad number");
            p = end;
            return true;
        }
        return fail(*p == '[' ? "arrays do not occur in AstarPa2Params" : "unexpected character");
    }
};

}  // namespace

extern "C" int apa_params_from_json(const char* json, apa_params* out, char* err, uint64_t err_cap) {
    std::string msg;
    auto bad = [&](const std::string& m) {
        msg = "AstarPa2Params JSON: " + m;
        if (err && err_cap) {
            strncpy(err, msg.c_str(), (size_t)err_cap - 1);
            err[err_cap - 1] = 0;
        }
        return APA_ERR_BAD_INPUT;
    };
    if (!json || !out) return bad("null argument");
    Parser ps{json, {}};
    JVal root;
    if (!ps.value(root)) return bad(ps.err);
    ps.ws();
    if (*ps.p) return bad("trailing characters");
    if (root.kind != JVal::Obj) return bad("expected an object");
    auto known = [&](const JVal& o, std::initializer_list<const char*> keys, const char* where) -> std::string {
        for (auto& kv : o.obj) {
            bool ok = false;
            for (const char* k : keys) ok |= kv.first == k;
            if (!ok) return std::string("unknown field `") + kv.first + "` in " + where;  // deny_unknown_fields
        }
        return "";
    };
    auto num = [&](const JVal* v, double dflt, bool& ok) -> double {
        if (!v) return dflt;
        if (v->kind != JVal::Num) ok = false;
        return v->num;
    };
    auto boolean = [&](const JVal* v, bool dflt, bool& ok) -> bool {
        if (!v) return dflt;
        if (v->kind != JVal::Bool) ok = false;
        return v->b;
    };
    std::string e = known(root, {"name", "domain", "heuristic", "doubling", "block_width", "front", "sparse_h", "prune", "viz"}, "AstarPa2Params");
    if (!e.empty()) return bad(e);
    apa_params q{};
    bool ok = true;
    // domain
    const JVal* d = root.get("domain");
    if (!d) return bad("missing field `domain`");
    std::string dom = d->kind == JVal::Str ? d->str : (d->kind == JVal::Obj && d->obj.size() == 1 ? d->obj[0].first : "");
    if (dom == "Full") q.domain = APA_DOMAIN_FULL;
    else if (dom == "GapStart") q.domain = APA_DOMAIN_GAP_START;
    else if (dom == "GapGap") q.domain = APA_DOMAIN_GAP_GAP;
    else if (dom == "Astar") q.domain = APA_DOMAIN_ASTAR;
    else return bad("unknown `domain`");
    // heuristic (pa-heuristic/src/cli.rs:46-98)
    const JVal* h = root.get("heuristic");
    if (!h || h->kind != JVal::Obj) return bad("missing field `heuristic`");
    e = known(*h, {"type", "r", "k", "p", "prune", "kmin", "kmax", "max_matches", "skip_prune"}, "heuristic");
    if (!e.empty()) return bad(e);
    const JVal* ht = h->get("type");
    if (!ht || ht->kind != JVal::Str) return bad("missing field `heuristic.type`");
    if (ht->str == "None" || ht->str == "Zero") q.heuristic = APA_HEURISTIC_NONE;
    else if (ht->str == "Gap") q.heuristic = APA_HEURISTIC_GAP;
    else if (ht->str == "GCSH") q.heuristic = APA_HEURISTIC_GCSH;
    else return bad("heuristic.type `" + ht->str + "` is not built (None, Zero, Gap, GCSH are)");
    q.r = (int32_t)num(h->get("r"), 2, ok);
    q.k = (int32_t)num(h->get("k"), 15, ok);
    q.p = (int32_t)num(h->get("p"), 0, ok);
    std::string hprune = "Start";
    if (const JVal* hp = h->get("prune")) {
        if (hp->kind != JVal::Str) return bad("heuristic.prune must be a string");
        hprune = hp->str;
    }
    for (const char* k : {"kmin", "kmax", "max_matches", "skip_prune"})
        if (const JVal* v = h->get(k))
            if (v->kind != JVal::Null) return bad(std::string("heuristic.") + k + " is not built (fixed-length seeds only)");
    const bool gcsh = q.domain == APA_DOMAIN_ASTAR && q.heuristic == APA_HEURISTIC_GCSH;
    if (gcsh && q.r != 1) return bad("heuristic.r = 2 (inexact matches) is not built; r must be 1");
    if (gcsh && hprune != "Start" && hprune != "None") return bad("heuristic.prune `" + hprune + "` is not built (Start, None are)");
    if (!gcsh) q.r = 1, q.k = 12, q.p = 14;  // unused outside GCSH: the values apa_params_preset leaves there
    // doubling (astarpa2/src/band.rs:26-45)
    const JVal* db = root.get("doubling");
    if (!db) return bad("missing field `doubling`");
    q.factor = 2.0f, q.delta = 1, q.doubling_start = APA_START_H0;
    if (db->kind == JVal::Str) {
        if (db->str != "None") return bad("doubling `" + db->str + "` is not built");
        q.doubling = APA_DOUBLING_NONE;
    } else if (db->kind == JVal::Obj && db->obj.size() == 1 && db->obj[0].second->kind == JVal::Obj) {
        const std::string& kind = db->obj[0].first;
        const JVal& body = *db->obj[0].second;
        if (kind == "BandDoubling") q.doubling = APA_DOUBLING_BAND;
        else if (kind == "LinearSearch") q.doubling = APA_DOUBLING_LINEAR;
        else return bad("doubling `" + kind + "` is not built");
        e = known(body, {"start", "factor", "delta"}, "doubling");
        if (!e.empty()) return bad(e);
        const JVal* st = body.get("start");
        if (!st || st->kind != JVal::Str) return bad("missing field `doubling.start`");
        if (st->str == "Zero") q.doubling_start = APA_START_ZERO;
        else if (st->str == "Gap") q.doubling_start = APA_START_GAP;
        else if (st->str == "H0") q.doubling_start = APA_START_H0;
        else return bad("unknown doubling.start");
        if (q.doubling == APA_DOUBLING_BAND) {
            if (!body.get("factor")) return bad("missing field `doubling.factor`");
            q.factor = (float)num(body.get("factor"), 2.0, ok);
        } else {
            if (!body.get("delta")) return bad("missing field `doubling.delta`");
            const double dl = num(body.get("delta"), 1.0, ok);
            if (dl != std::floor(dl)) return bad("doubling.delta must be a whole number here");
            q.delta = (int32_t)dl;
        }
    } else {
        return bad("malformed `doubling`");
    }
    if (!root.get("block_width")) return bad("missing field `block_width`");
    q.block_width = (int32_t)num(root.get("block_width"), 256, ok);
    // front: BlockParams (astarpa2/src/blocks.rs:31-60)
    const JVal* f = root.get("front");
    if (!f || f->kind != JVal::Obj) return bad("missing field `front`");
    e = known(*f, {"sparse", "simd", "no_ilp", "incremental_doubling", "dt_trace", "max_g", "fr_drop"}, "front");
    if (!e.empty()) return bad(e);
    if (!f->get("sparse")) return bad("missing field `front.sparse`");
    q.sparse = boolean(f->get("sparse"), true, ok);
    boolean(f->get("simd"), false, ok);    // scheduling hints of the CPU kernel: no meaning here
    boolean(f->get("no_ilp"), false, ok);
    q.incremental_doubling = boolean(f->get("incremental_doubling"), false, ok);
    q.dt_trace = boolean(f->get("dt_trace"), false, ok);
    q.max_g = (int32_t)num(f->get("max_g"), 0, ok);
    q.fr_drop = (int32_t)num(f->get("fr_drop"), 0, ok);
    if (!q.dt_trace && (q.max_g < 1 || q.max_g > 40)) q.max_g = 40;  // only read by DT-trace
    q.sparse_h = boolean(root.get("sparse_h"), false, ok);
    q.prune = boolean(root.get("prune"), false, ok) && !(gcsh && hprune == "None");
    if (boolean(root.get("viz"), false, ok)) return bad("viz is not built");
    if (!ok) return bad("a field has the wrong type");
    *out = q;
    return APA_OK;
}
