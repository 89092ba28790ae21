// Input side of the pa-bin equivalent: sequence-pair files in the reference's three formats
// (pa-bin/src/lib.rs:66-131). Pure host code, no GPU.
//   .seq            consecutive line pairs, first line prefixed '>' and second '<' (lib.rs:85-88 asserts both)
//   .txt            consecutive line pairs, no prefix
//   .fna .fa .fasta consecutive FASTA records form a pair; sequence lines of a record are concatenated
//                   (bio::io::fasta::Reader semantics: line terminators stripped, nothing else rewritten)
// A trailing unpaired line / record is ignored, as itertools::tuples() does (lib.rs:84,96-98).
// A directory argument means every file in it (lib.rs:70-78); here in sorted name order for reproducible output.
#pragma once
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace pa_input {

struct InputError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using PairFn = std::function<bool(std::string&& a, std::string&& b)>;  // return false to stop (ControlFlow::Break)

inline std::string extension_of(const std::string& path) {
    size_t slash = path.find_last_of('/');
    size_t dot = path.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot + 1 == path.size()) return "";
    return path.substr(dot + 1);
}

inline bool read_line(std::istream& in, std::string& line) {  // BufRead::lines(): strips "\n" and a preceding "\r"
    if (!std::getline(in, line)) return false;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return true;
}

// .seq / .txt (lib.rs:80-93). Returns false when the callback asked to stop.
inline bool read_line_pairs(std::istream& in, bool seq_format, const std::string& name, const PairFn& fn) {
    std::string a, b;
    uint64_t lineno = 0;
    while (read_line(in, a)) {
        lineno++;
        if (!read_line(in, b)) break;  // unpaired last line
        lineno++;
        if (seq_format) {
            if (a.empty() || a[0] != '>') throw InputError(name + ":" + std::to_string(lineno - 1) + ": expected a line starting with '>'");
            if (b.empty() || b[0] != '<') throw InputError(name + ":" + std::to_string(lineno) + ": expected a line starting with '<'");
            a.erase(0, 1);
            b.erase(0, 1);
        }
        if (!fn(std::move(a), std::move(b))) return false;
        a.clear();
        b.clear();
    }
    return true;
}

// FASTA records, paired consecutively (lib.rs:95-106).
inline bool read_fasta_pairs(std::istream& in, const std::string& name, const PairFn& fn) {
    std::string line, cur, first;
    bool in_record = false, have_first = false;
    uint64_t lineno = 0;
    auto finish = [&]() -> bool {  // a record ended
        if (!have_first) {
            first = std::move(cur);
            have_first = true;
            cur.clear();
            return true;
        }
        have_first = false;
        bool go = fn(std::move(first), std::move(cur));
        first.clear();
        cur.clear();
        return go;
    };
    while (read_line(in, line)) {
        lineno++;
        if (!line.empty() && line[0] == '>') {
            if (in_record && !finish()) return false;
            in_record = true;
            continue;
        }
        if (!in_record) {
            if (line.empty()) continue;
            throw InputError(name + ":" + std::to_string(lineno) + ": expected '>' at the start of a FASTA record");
        }
        cur += line;
    }
    if (in_record && !finish()) return false;
    return true;
}

inline bool is_dir(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
inline bool is_file(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

inline bool read_file_pairs(const std::string& path, const PairFn& fn) {
    const std::string ext = extension_of(path);
    if (ext.empty()) throw InputError(path + ": unknown file extension");
    const bool lines = ext == "seq" || ext == "txt";
    const bool fasta = ext == "fna" || ext == "fa" || ext == "fasta";
    if (!lines && !fasta) throw InputError(path + ": unknown file extension \"" + ext + "\". Must be in {seq,txt,fna,fa,fasta}.");
    std::ifstream in(path, std::ios::binary);
    if (!in) throw InputError(path + ": cannot open");
    return lines ? read_line_pairs(in, ext == "seq", path, fn) : read_fasta_pairs(in, path, fn);
}

// Cli::process_input_pairs for --input (lib.rs:66-110).
inline void process_input(const std::string& input, const PairFn& fn) {
    std::vector<std::string> files;
    if (is_file(input)) {
        files.push_back(input);
    } else if (is_dir(input)) {
        DIR* d = opendir(input.c_str());
        if (!d) throw InputError(input + " is not a file or directory");
        while (dirent* e = readdir(d)) {
            std::string nm = e->d_name;
            if (nm == "." || nm == "..") continue;
            std::string p = input + (input.back() == '/' ? "" : "/") + nm;
            if (is_file(p)) files.push_back(p);
        }
        closedir(d);
        std::sort(files.begin(), files.end());
    } else {
        throw InputError(input + " is not a file or directory");
    }
    for (const auto& f : files)
        if (!read_file_pairs(f, fn)) return;
}

}  // namespace pa_input
