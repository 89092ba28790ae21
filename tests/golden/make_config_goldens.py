"""Regression fixtures for the 16 restated A*PA2 configurations (oracle/oracle_capi.cpp preset_params):

    python tests/golden/make_config_goldens.py        # writes tests/golden/config_vectors.json

For a few seeded pairs per configuration: the optimal cost (cross-checked here against an independent Levenshtein), the CIGAR
text the oracle's restatement of the reference produces, the number of f_max tries and a digest of the per-pass band log. These
are NOT reference outputs (the Rust reference cannot be built here; exact CIGARs and bands are unpinned by the
reference's own tests): they freeze the oracle so that the oracle and the CUDA path cannot drift together unnoticed. The CPU
suite checks the oracle against them, the GPU suite checks the CUDA path against them.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [(300, 0.10, 0, 101), (1500, 0.05, 1, 102), (1500, 0.20, 2, 103), (2500, 0.08, 3, 104)]


def digest(band_log):
    return hashlib.sha256(json.dumps(band_log).encode()).hexdigest()[:16]


def main():
    import astar_pairwise_aligner_b200 as A  # generator only (host code of the library)
    import oracle_lib as O
    out = {"cases": [list(c) for c in CASES], "configs": {}}
    for preset in range(16):
        rows = []
        for n, e, model, seed in CASES:
            a, b = A.generate_pair(n, e, model, seed)
            cost, cigar, st = O.align(a, b, preset, True)
            assert cost == O.levenshtein(a, b) and O.cigar_verify(cigar, a, b) == cost
            rows.append({"cost": cost, "cigar": cigar, "f_max_tries": st["f_max_tries"], "band_log": digest(O.band_log(a, b, preset, True))})
        out["configs"][str(preset)] = rows
    with open(os.path.join(HERE, "config_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["configs"]), "configurations x", len(CASES), "pairs")


if __name__ == "__main__":
    main()
