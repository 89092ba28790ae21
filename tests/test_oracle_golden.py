"""CPU tests: pin the oracle (CPU restatement of the reference path) against the reference's own golden
vectors and known-answer tests, then against an independent ground truth on a random grid
shaped like pa-test's (pa-test/src/lib.rs:24-63)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
PRESETS = list(range(16))  # 0 simple, 1 full, 2..9: configurations of astarpa2/src/tests.rs:19-119, 10..15: further
# points of the parameter space (other domain / LinearSearch / dense h / nw() / bw = 1 / k = 8), see oracle/oracle_capi.cpp


@pytest.mark.parametrize("idx", range(len(GOLD["pairs"])))
def test_golden_pairs_cost_and_cigar(oracle, idx):
    p = GOLD["pairs"][idx]
    a, b = p["a"].encode(), p["b"].encode()
    assert oracle.levenshtein(a, b) == p["cost"]
    assert oracle.levenshtein_dp(a, b) == p["cost"]
    for preset in PRESETS:
        cost, cigar, _ = oracle.align(a, b, preset, True, self_check=True)
        assert cost == p["cost"], (preset, p["src"])
        assert oracle.cigar_verify(cigar, a, b) == cost  # pa-test/src/lib.rs:98
        cost2, none, _ = oracle.align(a, b, preset, False)
        assert cost2 == cost and none is None


def test_example_c_costs(oracle):
    # astarpa-c/example.c:23-29: cost 2 through astarpa2_simple and astarpa2_full
    a, b = b"ACTCGCT", b"AACTCGTT"
    assert oracle.align(a, b, oracle.PRESET_SIMPLE)[0] == 2
    assert oracle.align(a, b, oracle.PRESET_FULL)[0] == 2


def test_cigar_text_format(oracle):
    # astarpa-c/example.cpp:16: "=I4=X=" is a valid optimal CIGAR text for this pair: count omitted when 1.
    g = GOLD["cigar_format"]
    assert oracle.cigar_verify(g["cigar"], g["a"].encode(), g["b"].encode()) == g["cost"]
    # malformed texts are rejected by the checker
    assert oracle.cigar_verify("1=I4=X=", g["a"].encode(), g["b"].encode()) == -1
    assert oracle.cigar_verify("==I3=X=", g["a"].encode(), g["b"].encode()) == -1
    cost, cigar, _ = oracle.align(g["a"].encode(), g["b"].encode(), oracle.PRESET_FULL)
    assert cost == 2 and "1" not in cigar.replace("10", "").replace("11", "")


def test_qgram_kat(oracle):
    # pa-heuristic/src/matches/qgrams.rs:117-124
    L = oracle.lib()
    for s, q in GOLD["qgram"]["to_qgram"].items():
        assert L.oracle_to_qgram(s.encode(), len(s)) == q
    for c, v in GOLD["qgram"]["char_to_bits"].items():
        assert L.oracle_to_qgram(c.encode(), 1) == v


@pytest.mark.parametrize("h", [64, 128, 256, 512])
def test_block_kernel_kat(oracle, apa, h):
    # pa-bitpacking/benches/nw/main.rs:142-149: 256 columns x h rows, all-(+1) input deltas:
    # sum of bottom horizontal deltas == levenshtein(a, b) - |b|.
    a, _ = apa.generate_pair(256, 0.0, 0, 31415)
    b, _ = apa.generate_pair(h, 0.0, 0, 27182)
    hbuf = np.ones(256, dtype=np.uint8)
    v = np.zeros(2 * (h // 64), dtype=np.uint64)
    v[0::2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    s = oracle.lib().oracle_bp_compute(a, len(a), b, len(b), hbuf.ctypes.data, v.ctypes.data)
    assert s == oracle.levenshtein(a, b) - len(b)


def _grid(ns, es, models, seed0):
    for n in ns:
        for e in es:
            for model in models:
                yield n, e, model, seed0 + 1000 * n + model


NS = [0, 1, 2, 3, 7, 10, 17, 20, 50, 100, 190, 254, 255, 256, 257, 258, 300, 500, 511, 512, 513, 515]
ES = [0.0, 0.01, 0.05, 0.1, 0.2, 0.3, 0.5, 0.7, 1.0]


@pytest.mark.parametrize("preset", PRESETS)
def test_random_grid_against_ground_truth(oracle, apa, preset):
    # pa-test/src/lib.rs:65-99: cost == Levenshtein (triple_accel there, full-matrix bit-vector here) and the
    # CIGAR verifies. self_check = the reference's cfg!(test) incremental-vs-scratch assertion (blocks.rs:471-543).
    for n, e, model, seed in _grid(NS, ES[preset % 3::3], range(4), 31415):
        a, b = apa.generate_pair(n, e, model, seed)
        lev = oracle.levenshtein(a, b)
        cost, cigar, _ = oracle.align(a, b, preset, True, self_check=True)
        assert cost == lev, (preset, n, e, model)
        assert oracle.cigar_verify(cigar, a, b) == lev, (preset, n, e, model)


@pytest.mark.parametrize("n,e", [(3000, 0.05), (10000, 0.05), (10000, 0.15), (30000, 0.08)])
def test_presets_larger(oracle, apa, n, e):
    a, b = apa.generate_pair(n, e, 0, 31415 + n)
    lev = oracle.levenshtein(a, b)
    for preset in (0, 1):
        cost, cigar, st = oracle.align(a, b, preset, True, self_check=True)
        assert cost == lev and oracle.cigar_verify(cigar, a, b) == lev
        assert oracle.align(a, b, preset, False)[0] == lev


def test_ground_truth_cross_check(oracle, apa):
    for n, e, model, seed in _grid([0, 1, 5, 63, 64, 65, 130, 400], [0.0, 0.1, 0.5], range(4), 99):
        a, b = apa.generate_pair(n, e, model, seed)
        assert oracle.levenshtein(a, b) == oracle.levenshtein_dp(a, b)


def test_bad_input_panics(oracle):
    # BitProfile::build panics on bytes outside ACGT (pa-bitpacking/src/profile.rs:113)
    with pytest.raises(oracle.OraclePanic):
        oracle.align(b"ACGN", b"ACGT", 0)


def _config_goldens():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_vectors.json")))


@pytest.mark.parametrize("preset", PRESETS)
def test_oracle_matches_committed_config_vectors(oracle, apa, preset):
    # tests/golden/config_vectors.json (tests/golden/make_config_goldens.py): frozen oracle outputs per configuration - cost,
    # CIGAR text, number of f_max tries, digest of the band log - so that oracle and CUDA path cannot drift together.
    import hashlib
    import json
    g = _config_goldens()
    for (n, e, model, seed), want in zip(g["cases"], g["configs"][str(preset)]):
        a, b = apa.generate_pair(n, e, model, seed)
        cost, cigar, st = oracle.align(a, b, preset, True)
        assert (cost, cigar, st["f_max_tries"]) == (want["cost"], want["cigar"], want["f_max_tries"]), (preset, n)
        digest = hashlib.sha256(json.dumps(oracle.band_log(a, b, preset, True)).encode()).hexdigest()[:16]
        assert digest == want["band_log"], (preset, n)


@pytest.mark.parametrize("preset", [p for p in PRESETS if p >= 2])
def test_configurations_larger(oracle, apa, preset):
    # the non-preset configurations on longer pairs: optimal cost (independent Levenshtein), valid CIGAR, cost-only == traced
    big = preset in (9, 13)  # Domain::Full computes the whole rectangle
    for n, e, model in [(3000, 0.05, 0), (3000, 0.2, 1)] if big else [(6000, 0.05, 0), (6000, 0.15, 2), (12000, 0.08, 3)]:
        a, b = apa.generate_pair(n, e, model, 4000 + n + preset)
        lev = oracle.levenshtein(a, b)
        cost, cigar, _ = oracle.align(a, b, preset, True, self_check=True)
        assert cost == lev and oracle.cigar_verify(cigar, a, b) == lev, (preset, n, e)
        assert oracle.align(a, b, preset, False)[0] == lev
