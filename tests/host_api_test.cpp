// C++ host mirror (include/astarpa2.hpp) test program, built and run by tests/test_pa_bin.py.
//   host_api_test cigar            CPU only: Cigar::{parse,to_string,verify} (text format pinned by astarpa-c/example.cpp:16)
//   host_api_test align            needs a B200: the pa_types::Aligner-shaped calls on the reference's example pair
#include <cstdio>
#include <cstring>

#include "astarpa2.hpp"

static int fails = 0;
#define CHECK(x)                                              \
    do {                                                      \
        if (!(x)) {                                           \
            printf("FAIL line %d: %s\n", __LINE__, #x);       \
            fails++;                                          \
        }                                                     \
    } while (0)

int main(int argc, char** argv) {
    using namespace astarpa2;
    const char* mode = argc > 1 ? argv[1] : "cigar";
    const Seq a = "ACTCGCT", b = "AACTCGTT";  // astarpa-c/example.c:8-9
    if (!strcmp(mode, "cigar")) {
        Cigar c = Cigar::parse("=I4=X=");  // astarpa-c/example.cpp:16
        CHECK(c.ops.size() == 5);
        CHECK((c.ops[1] == CigarElem{CigarOp::Ins, 1}));
        CHECK((c.ops[2] == CigarElem{CigarOp::Match, 4}));
        CHECK(c.to_string() == "=I4=X=");
        CHECK(c.verify(a, b) == 2);
        CHECK(c.verify(b, a) == -1);
        CHECK(Cigar::parse("7=").verify(a, a) == 0);
        CHECK(Cigar::parse("").verify("", "") == 0);
        CHECK(Cigar::parse("3D").verify("ACG", "") == 3);
        CHECK(Cigar::parse("12I").to_string() == "12I");
        CHECK(Cigar::parse("6=X").verify(a, a) == -1);  // X on equal bases is not a valid substitution
        bool threw = false;
        try {
            Cigar::parse("3M");
        } catch (const Error&) {
            threw = true;
        }
        CHECK(threw);
        threw = false;
        try {
            Cigar::parse("3=4");
        } catch (const Error&) {
            threw = true;
        }
        CHECK(threw);
    } else if (!strcmp(mode, "nodevice")) {
        bool threw = false;
        try {
            AstarPa2::full(true);
        } catch (const Error& e) {
            threw = e.code == APA_ERR_NO_DEVICE;
            printf("%s\n", e.what());
        }
        CHECK(threw);
    } else {
        for (int preset = 0; preset < 2; preset++) {
            AstarPa2 al(preset ? AstarPa2::Full : AstarPa2::Simple, true);
            auto [cost, cigar] = al.align(a, b);
            CHECK(cost == 2);  // astarpa-c/example.c:23-29
            CHECK(cigar.has_value() && cigar->verify(a, b) == 2);
            CHECK(al.cost(a, b) == 2);
            std::vector<std::pair<Seq, Seq>> pairs = {{a, b}, {a, a}, {"", b}, {b, a}};
            BatchResult r = al.align_batch(pairs);
            CHECK(r.costs.size() == 4 && r.costs[0] == 2 && r.costs[1] == 0 && r.costs[2] == 8 && r.costs[3] == 2);
            CHECK(r.cigars[1] == "7=" && r.cigars[2] == "8I");
            for (size_t p = 0; p < pairs.size(); p++) CHECK(Cigar::parse(r.cigars[p]).verify(pairs[p].first, pairs[p].second) == r.costs[p]);
        }
        auto [c1, g1] = astarpa2_simple(a, b);
        auto [c2, g2] = astarpa2_full(a, b);
        CHECK(c1 == 2 && c2 == 2 && g1.verify(a, b) == 2 && g2.verify(a, b) == 2);
        // one call over the GPUs of the box (apa_align_batch_multi), and SearchResult::trace (search.rs:135-230)
        {
            std::vector<std::pair<Seq, Seq>> pairs = {{a, b}, {a, a}, {"", b}, {b, a}};
            BatchResult r = align_batch_multi({0}, AstarPa2::Full, true, pairs);
            CHECK(r.costs.size() == 4 && r.costs[0] == 2 && r.costs[1] == 0 && r.costs[2] == 8 && r.costs[3] == 2);
            for (size_t p = 0; p < pairs.size(); p++) CHECK(Cigar::parse(r.cigars[p]).verify(pairs[p].first, pairs[p].second) == r.costs[p]);
            SearchResult sr("AC", "CTTACTTA", 0.0f);  // the reference's doc-test, search.rs:29-32
            CHECK((sr.out == std::vector<Cost>{0, 0, 1, 2, 1, 0, 1, 2, 1, 0, 0}));
            SearchTrace tr = sr.trace(5);
            CHECK(tr.cigar.to_string() == "2=" && tr.start == std::make_pair(3, 0) && tr.end == std::make_pair(5, 2) && tr.cost == 0);
        }
        // explicit parameters (general kernel) and the stats surface (align_with_stats, lib.rs:200-208)
        {
            AstarPa2Params q = AstarPa2Params::nw();  // params.rs:46-68: the full n*m rectangle in one pass
            AstarPa2 nw = q.make_aligner(true);
            auto [al, st] = nw.align_with_stats(a, b);
            CHECK(al.first == 2 && al.second.has_value() && al.second->verify(a, b) == 2);
            CHECK(st.f_max_tries == 1 && st.h0 == 0 && st.num_matches == 0);
            AstarPa2Params g = AstarPa2Params::simple();
            g.domain = APA_DOMAIN_GAP_GAP;  // Edlib-like band (tests.rs:24-32)
            g.block_width = 64;
            g.doubling_start = APA_START_GAP;
            CHECK(g.make_aligner(false).cost(a, b) == 2);
            AstarPa2 full = AstarPa2::full(true);
            auto [al2, st2] = full.align_with_stats(a, b);
            CHECK(al2.first == 2 && st2.f_max_tries >= 1 && st2.h0 <= 2);
            bool threw2 = false;
            try {
                AstarPa2Params bad = AstarPa2Params::full();
                bad.r = 2;  // inexact matches are not built
                bad.make_aligner(true).align(a, b);
            } catch (const Error& e) {
                threw2 = e.code == APA_ERR_BAD_INPUT;
            }
            CHECK(threw2);
        }
        {   // pa_bitpacking::search doc-test (pa-bitpacking/src/search.rs:29-32)
            std::vector<Cost> out = search("AC", "CTTACTTA", 0.0f);
            const std::vector<Cost> want = {0, 0, 1, 2, 1, 0, 1, 2, 1, 0, 0};
            CHECK(out == want);
        }
        AstarPa2 cost_only = AstarPa2::full(false);
        auto r = cost_only.align(a, b);
        CHECK(r.first == 2 && !r.second.has_value());  // trace = false: Aligner::align returns no CIGAR (lib.rs:74-77)
        bool threw = false;
        try {
            cost_only.cost("ACGN", "ACGT");
        } catch (const Error& e) {
            threw = e.code == APA_ERR_BAD_INPUT;
        }
        CHECK(threw);
    }
    printf(fails ? "host_api_test %s: %d FAILED\n" : "host_api_test %s: ok\n", mode, fails);
    return fails ? 1 : 0;
}
