// pa-bin equivalent for the B200 A*PA2 path (reference: pa-bin/src/main.rs:9-37, pa-bin/src/lib.rs:16-131).
//
//   pa-bin [OPTIONS] <--input <INPUT> | --length <LENGTH>>
//     -i, --input <INPUT>        a .seq, .txt or FASTA file (or a directory of them) with sequence pairs
//     -o, --output <OUTPUT>      write a .csv of `{cost},{cigar}` lines
//         --aligner <ALIGNER>    astarpa | astarpa2-simple | astarpa2-full   [default: astarpa2-full]
//     -n, --length <LENGTH>      generated input: target length             (pa_generate::DatasetGenerator flags)
//     -e, --error-rate <RATE>    generated input: error rate                 [default: 0.05]
//         --seed <SEED>  --cnt <CNT>  --error-model <uniform|noisy-insert|noisy-delete|symmetric-repeat>
//   Not in the reference: --device D, --batch-bases B (bases of a+b per GPU batch), --cost-only (AstarPa2::cost),
//   --dry-run (parse / generate only and print `n,m,fnv1a(a),fnv1a(b)` per pair; touches no GPU).
//
// The reference aligns one pair per Aligner::align call (main.rs:26); one pair per call cannot feed a GPU, so pairs
// are collected into batches and each batch is one apa_align_batch call. Output order = input order.
// `--aligner astarpa` (A*PA v1) is served by the A*PA2-full engine: same optimal cost, a valid CIGAR.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/apa_generate.h"
#include "../../include/astarpa2.hpp"
#include "pa_input.hpp"

namespace {

struct Args {
    std::string input, output, aligner = "astarpa2-full", error_model = "uniform", params_file;
    bool have_length = false, have_seed = false, cost_only = false, dry_run = false;
    uint64_t length = 1000, cnt = 1, seed = 0, batch_bases = 2000000000ull;
    double error_rate = 0.05;
    int device = 0;
};

[[noreturn]] void usage(int code) {
    fprintf(code ? stderr : stdout,
            "Globally align pairs of sequences using A*PA2 on a B200\n\n"
            "Usage: pa-bin [OPTIONS] <--input <INPUT>|--length <LENGTH>>\n\n"
            "Options:\n"
            "  -i, --input <INPUT>      A .seq, .txt, or Fasta file with sequence pairs to align\n"
            "  -o, --output <OUTPUT>    Write a .csv of `{cost},{cigar}` lines\n"
            "      --aligner <ALIGNER>  The aligner to use [default: astarpa2-full] [possible values: astarpa,\n"
            "                           astarpa2-simple, astarpa2-full]\n"
            "      --params <FILE>      AstarPa2Params as the reference's serde JSON (a pa-bench job's `params`); overrides --aligner\n"
            "      --device <D>         CUDA device index [default: 0]\n"
            "      --batch-bases <B>    Bases (a + b) per GPU batch [default: 2000000000]\n"
            "      --cost-only          Compute costs only (no traceback); the csv then holds `{cost},`\n"
            "      --dry-run            Parse/generate the pairs and print `n,m,fnv(a),fnv(b)`; no GPU work\n"
            "  -h, --help               Print help\n\n"
            "Generated input:\n"
            "  -n, --length <LENGTH>          Target length of each generated sequence [default: 1000]\n"
            "  -e, --error-rate <ERROR_RATE>  Error rate between sequences [default: 0.05]\n"
            "      --seed <SEED>              RNG seed (pair p uses seed + p) [default: random]\n"
            "      --cnt <CNT>                Number of pairs to generate [default: 1]\n"
            "      --error-model <MODEL>      uniform | noisy-insert | noisy-delete | symmetric-repeat [default: uniform]\n");
    exit(code);
}

Args parse(int argc, char** argv) {
    Args a;
    auto need = [&](int& k) -> const char* {
        if (k + 1 >= argc) {
            fprintf(stderr, "error: %s needs a value\n", argv[k]);
            usage(2);
        }
        return argv[++k];
    };
    for (int k = 1; k < argc; k++) {
        std::string s = argv[k];
        std::string val;
        size_t eq = s.find('=');
        bool has_eq = s.rfind("--", 0) == 0 && eq != std::string::npos;
        if (has_eq) {
            val = s.substr(eq + 1);
            s = s.substr(0, eq);
        }
        auto value = [&]() -> std::string { return has_eq ? val : std::string(need(k)); };
        if (s == "-h" || s == "--help") usage(0);
        else if (s == "-i" || s == "--input") a.input = value();
        else if (s == "-o" || s == "--output") a.output = value();
        else if (s == "--aligner") a.aligner = value();
        else if (s == "--params") a.params_file = value();
        else if (s == "-n" || s == "--length") a.length = strtoull(value().c_str(), nullptr, 10), a.have_length = true;
        else if (s == "-e" || s == "--error-rate") a.error_rate = atof(value().c_str());
        else if (s == "--seed") a.seed = strtoull(value().c_str(), nullptr, 10), a.have_seed = true;
        else if (s == "--cnt") a.cnt = strtoull(value().c_str(), nullptr, 10);
        else if (s == "--error-model") a.error_model = value();
        else if (s == "--device") a.device = atoi(value().c_str());
        else if (s == "--batch-bases") a.batch_bases = strtoull(value().c_str(), nullptr, 10);
        else if (s == "--cost-only") a.cost_only = true;
        else if (s == "--dry-run") a.dry_run = true;
        else {
            fprintf(stderr, "error: unexpected argument '%s'\n", argv[k]);
            usage(2);
        }
    }
    // clap group "input_type": exactly one of --input / --length (pa-bin/src/lib.rs:44-48)
    if (a.input.empty() == !a.have_length) {
        fprintf(stderr, "error: exactly one of --input <INPUT> and --length <LENGTH> is required\n");
        usage(2);
    }
    if (a.aligner != "astarpa" && a.aligner != "astarpa2-simple" && a.aligner != "astarpa2-full") {
        fprintf(stderr, "error: invalid value '%s' for '--aligner' [possible values: astarpa, astarpa2-simple, astarpa2-full]\n",
                a.aligner.c_str());
        exit(2);
    }
    return a;
}

int model_id(const std::string& m) {
    if (m == "uniform") return 0;
    if (m == "noisy-insert") return 1;
    if (m == "noisy-delete") return 2;
    if (m == "symmetric-repeat") return 3;
    fprintf(stderr, "error: invalid value '%s' for '--error-model'\n", m.c_str());
    exit(2);
}

uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h;
}

struct Runner {
    const Args& args;
    explicit Runner(const Args& a) : args(a) {}
    astarpa2::AstarPa2* aligner = nullptr;
    FILE* out = nullptr;
    std::string a_all, b_all;
    std::vector<int64_t> a_off{0}, b_off{0};
    uint64_t done = 0;

    void flush() {
        const size_t n = a_off.size() - 1;
        if (n == 0) return;
        astarpa2::BatchResult r = aligner->align_batch_concat((const uint8_t*)a_all.data(), a_off.data(), (const uint8_t*)b_all.data(),
                                                              b_off.data(), n, !args.cost_only);
        for (size_t p = 0; p < n; p++) {
            if (out) fprintf(out, "%d,%s\n", (int)r.costs[p], args.cost_only ? "" : r.cigars[p].c_str());  // main.rs:31-33
        }
        done += n;
        fprintf(stderr, "Done: %3llu\r", (unsigned long long)done);
        a_all.clear();
        b_all.clear();
        a_off.assign(1, 0);
        b_off.assign(1, 0);
    }
    bool push(std::string&& a, std::string&& b) {
        if (args.dry_run) {
            printf("%zu,%zu,%016llx,%016llx\n", a.size(), b.size(), (unsigned long long)fnv1a(a), (unsigned long long)fnv1a(b));
            done++;
            return true;
        }
        if (a_off.size() > 1 && a_all.size() + b_all.size() + a.size() + b.size() > args.batch_bases) flush();
        a_all += a;
        b_all += b;
        a_off.push_back((int64_t)a_all.size());
        b_off.push_back((int64_t)b_all.size());
        return true;
    }
};

}  // namespace

int main(int argc, char** argv) {
    Args args = parse(argc, argv);
    try {
        Runner run(args);
        std::unique_ptr<astarpa2::AstarPa2> al;
        if (!args.dry_run) {
            auto preset = args.aligner == "astarpa2-simple" ? astarpa2::AstarPa2::Simple : astarpa2::AstarPa2::Full;
            if (!args.params_file.empty()) {  // AstarPa2Params::make_aligner (params.rs:132-226) from the serde JSON form
                FILE* pf = fopen(args.params_file.c_str(), "r");
                if (!pf) {
                    fprintf(stderr, "error: cannot read %s\n", args.params_file.c_str());
                    return 1;
                }
                std::string json;
                char buf[4096];
                size_t got;
                while ((got = fread(buf, 1, sizeof buf, pf)) > 0) json.append(buf, got);
                fclose(pf);
                al.reset(new astarpa2::AstarPa2(astarpa2::AstarPa2Params::from_json(json), !args.cost_only, args.device));
            } else {
                al.reset(new astarpa2::AstarPa2(preset, !args.cost_only, args.device));  // AlignerType::build, lib.rs:25-33
            }
            run.aligner = al.get();
            if (!args.output.empty()) {
                run.out = fopen(args.output.c_str(), "w");
                if (!run.out) {
                    fprintf(stderr, "error: cannot create %s\n", args.output.c_str());
                    return 1;
                }
            }
            fprintf(stderr, "Done: %3d\r", 0);
        }
        if (!args.input.empty()) {
            pa_input::process_input(args.input, [&](std::string&& a, std::string&& b) { return run.push(std::move(a), std::move(b)); });
        } else {
            // Generated input (lib.rs:111-126). The reference draws from ChaCha8 seeded with --seed (a random seed in
            // 0..1000 is chosen and printed when absent); the stream of that external crate is not reproduced: pair p
            // comes from this library's generator with seed + p.
            uint64_t seed = args.seed;
            if (!args.have_seed) {
                seed = std::random_device{}() % 1000;
                fprintf(stderr, "Seed: %llu\n", (unsigned long long)seed);
            }
            const int model = model_id(args.error_model);
            for (uint64_t p = 0; p < args.cnt; p++) {
                const uint64_t cap = 3 * args.length + 64;
                std::string a(args.length ? args.length : 1, '\0'), b(cap, '\0');
                int64_t bl = apa_generate_pair(args.length, args.error_rate, model, seed + p, (uint8_t*)a.data(), (uint8_t*)b.data(), cap);
                if (bl < 0) throw astarpa2::Error(APA_ERR_TOO_LARGE, "generator buffer too small");
                a.resize(args.length);
                b.resize((size_t)bl);
                if (!run.push(std::move(a), std::move(b))) break;
            }
        }
        if (!args.dry_run) {
            run.flush();
            fprintf(stderr, "\n");
            if (run.out) fclose(run.out);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "\npa-bin: %s\n", e.what());
        return 1;
    }
    return 0;
}
