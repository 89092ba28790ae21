"""Extracts the reference's own golden vectors for the A*PA2 path into tests/golden/*.json.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources (file:line under /root/reference):
  pa-test/src/lib.rs:9-18          8 hard-coded pairs (cost must equal Levenshtein, CIGAR must verify)
  astarpa/src/tests.rs:134-169     5 regression pairs of past heuristic bugs
  astarpa-c/example.c:9-10,23-29   ACTCGCT / AACTCGTT => cost 2 through astarpa2_simple and astarpa2_full
  astarpa-c/example.cpp:16         CIGAR text "=I4=X=" (format: count omitted when 1)
  pa-heuristic/src/matches/qgrams.rs:117-144  q-gram known answers
The expected costs are computed here with a textbook O(nm) DP (independent of oracle/ and of the product).
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lev(a, b):
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j - 1] + (a[i - 1] != b[j - 1]), prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[-1]


def main():
    out = {"pairs": []}
    src = open(os.path.join(REF, "pa-test/src/lib.rs")).read()
    body = src[src.index("fn test_sequences"):src.index("const FIXED")]
    seqs = re.findall(r'b"([ACGT]+)"', body)
    assert len(seqs) == 16
    for k in range(0, 16, 2):
        out["pairs"].append({"src": "pa-test/src/lib.rs:9-18", "a": seqs[k], "b": seqs[k + 1]})
    src = open(os.path.join(REF, "astarpa/src/tests.rs")).read()
    body = src[src.index("mod edge_cases"):]
    seqs = re.findall(r'"([ACGT]+)"\s*\.as_bytes\(\)', body.replace("\n", " "))
    assert len(seqs) == 10, len(seqs)
    for k in range(0, 10, 2):
        out["pairs"].append({"src": "astarpa/src/tests.rs:134-169", "a": seqs[k], "b": seqs[k + 1]})
    out["pairs"].append({"src": "astarpa-c/example.c:9-10", "a": "ACTCGCT", "b": "AACTCGTT", "expect_cost": 2})
    for p in out["pairs"]:
        c = lev(p["a"], p["b"])
        assert p.get("expect_cost", c) == c
        p["cost"] = c
    out["cigar_format"] = {"src": "astarpa-c/example.cpp:16", "a": "ACTCGCT", "b": "AACTCGTT", "cigar": "=I4=X=", "cost": 2}
    out["qgram"] = {"src": "pa-heuristic/src/matches/qgrams.rs:117-124",
                    "char_to_bits": {"A": 0, "C": 1, "G": 3, "T": 2},
                    "to_qgram": {"ACGT": 0b00011110, "TGCA": 0b10110100}}
    json.dump(out, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)
    print("wrote", len(out["pairs"]), "pairs")


if __name__ == "__main__":
    main()
