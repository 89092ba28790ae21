"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI library, against the
oracle on the same seeded inputs — bit-exact cost AND CIGAR text (integer work, no tolerance)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))

PRESETS = [0, 1]  # astarpa2_simple, astarpa2_full


def _check_pairs(apa, oracle, pairs, preset, trace=True):
    costs, cigars = apa.AstarPa2(preset, trace).align_batch(pairs)
    for k, (a, b) in enumerate(pairs):
        oc, ocg, _ = oracle.align(a, b, preset, trace)
        assert int(costs[k]) == oc, (preset, k, len(a), len(b), int(costs[k]), oc)
        if trace:
            if cigars[k] != ocg:
                gl = oracle.parse_band_log(_gpu_band_log(apa, a, b, preset))
                ol = oracle.band_log(a, b, preset, True)
                first = next((i for i, (x, y) in enumerate(zip(gl, ol)) if x != y), None)
                raise AssertionError(f"CIGAR differs: preset {preset} pair {k} n={len(a)} m={len(b)} cost {oc}; "
                                     f"band logs equal: {gl == ol}; first differing pass {first}")
        else:
            assert cigars is None


def _gpu_band_log(apa, a, b, preset, trace=True):
    eng = apa._engine(0)
    L = apa.load_library()
    L.apa_debug_band_log.restype = C.c_int64
    L.apa_debug_band_log.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p,
                                     C.c_uint64]
    cap = 16 + 8 * (len(a) // 64 + 4) * 40
    buf = np.zeros(cap, dtype=np.int32)
    w = L.apa_debug_band_log(eng._h, preset, int(trace), a, len(a), b, len(b), buf.ctypes.data, cap)
    assert 0 <= w <= cap, w
    return buf[:w]


@pytest.mark.parametrize("h", [64, 128, 256, 512, 1024, 2112])
def test_block_kernel_kat_gpu(apa, oracle, engine, h):
    # pa-bitpacking/benches/nw/main.rs:142-149 on the GPU kernel, plus equality of every output with the oracle.
    for na in (256, 100, 700):
        a, _ = apa.generate_pair(na, 0.0, 0, 31415 + h)
        b, _ = apa.generate_pair(h, 0.0, 0, 27182 + na)
        s, hout, vout = engine.block_compute(a, b)
        assert s == oracle.levenshtein(a, b) - len(b)
        hb = np.ones(na, dtype=np.uint8)
        v = np.zeros(2 * (h // 64), dtype=np.uint64)
        v[0::2] = np.uint64(0xFFFFFFFFFFFFFFFF)
        so = oracle.lib().oracle_bp_compute(a, len(a), b, len(b), hb.ctypes.data, v.ctypes.data)
        assert so == s
        assert (hb == hout).all()
        assert (v == vout).all()


def test_block_kernel_hmode_update_gpu(apa, oracle, engine):
    # HMode::Update (astarpa2/src/blocks.rs:665-748): arbitrary deltas along the top edge in, the bottom edge out, arbitrary left
    # column - what incremental doubling feeds the kernel. Random +1 / 0 / -1 top deltas and random consistent left columns
    # against the oracle's bp_compute, single- and multi-chunk heights, partial last slab.
    rng = np.random.default_rng(3)
    for na, h in [(1, 64), (37, 64), (256, 128), (300, 1024), (700, 2112), (256, 4096)]:
        a, _ = apa.generate_pair(na, 0.0, 0, 11 + na)
        b, _ = apa.generate_pair(h, 0.0, 0, 13 + h)
        hin = rng.integers(0, 3, size=na).astype(np.uint8)
        p = rng.integers(0, 1 << 62, size=h // 64, dtype=np.uint64) * np.uint64(3)
        m = rng.integers(0, 1 << 62, size=h // 64, dtype=np.uint64) & ~p
        v = np.zeros(2 * (h // 64), dtype=np.uint64)
        v[0::2], v[1::2] = p, m
        s, hout, vout = engine.block_compute(a, b, v=v.copy(), h=hin)
        hb, vo = hin.copy(), v.copy()
        so = oracle.lib().oracle_bp_compute(a, len(a), b, len(b), hb.ctypes.data, vo.ctypes.data)
        assert (hb == hout).all() and (vo == vout).all() and so == s, (na, h)


@pytest.mark.parametrize("preset", PRESETS)
def test_golden_pairs_gpu(apa, oracle, preset):
    pairs = [(p["a"].encode(), p["b"].encode()) for p in GOLD["pairs"]]
    costs, cigars = apa.AstarPa2(preset, True).align_batch(pairs)
    for p, c, cg in zip(GOLD["pairs"], costs, cigars):
        assert int(c) == p["cost"], p["src"]
        assert oracle.cigar_verify(cg, p["a"].encode(), p["b"].encode()) == p["cost"]
    _check_pairs(apa, oracle, pairs, preset)


NS = [0, 1, 2, 3, 7, 10, 17, 20, 50, 100, 190, 254, 255, 256, 257, 258, 300, 500, 511, 512, 513, 515, 1000, 2049]
ES = [0.0, 0.01, 0.05, 0.1, 0.2, 0.3, 0.5, 0.7, 1.0]


@pytest.mark.parametrize("preset", PRESETS)
@pytest.mark.parametrize("model", range(4))
def test_random_grid_gpu(apa, oracle, preset, model):
    # pa-test/src/lib.rs:24-63 grid (n list, e list, 4 error models), all pairs in one GPU batch.
    pairs = [apa.generate_pair(n, e, model, 31415 + 1000 * n + model) for n in NS for e in ES]
    _check_pairs(apa, oracle, pairs, preset)
    _check_pairs(apa, oracle, pairs[::7], preset, trace=False)


@pytest.mark.parametrize("preset", PRESETS)
@pytest.mark.parametrize("n,e", [(10000, 0.05), (10000, 0.15), (30000, 0.08), (100000, 0.05), (100000, 0.15)])
def test_long_pairs_gpu(apa, oracle, preset, n, e):
    pairs = [apa.generate_pair(n, e, 0, 31415 + s) for s in range(3)]
    _check_pairs(apa, oracle, pairs, preset)


@pytest.mark.parametrize("preset", PRESETS)
def test_batch_mixed_lengths_gpu(apa, oracle, preset):
    rng = np.random.default_rng(7)
    pairs = [apa.generate_pair(int(rng.integers(0, 6000)), float(rng.choice([0.01, 0.05, 0.1, 0.2])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(300)]
    _check_pairs(apa, oracle, pairs, preset)


@pytest.mark.parametrize("preset", PRESETS)
def test_band_log_matches_oracle(apa, oracle, preset):
    for n, e in [(3000, 0.1), (20000, 0.05), (20000, 0.15)]:
        a, b = apa.generate_pair(n, e, 0, 5)
        assert oracle.parse_band_log(_gpu_band_log(apa, a, b, preset)) == oracle.band_log(a, b, preset, True)


def test_drop_in_symbols_gpu(apa):
    # astarpa-c/example.c:8-33 through ctypes: cost 2 through every entry point, CIGAR released by astarpa_free_cigar.
    L = apa.load_library()
    a, b = b"ACTCGCT", b"AACTCGTT"
    for fn in (L.astarpa2_simple, L.astarpa2_full, L.astarpa):
        cig = C.c_void_p()
        ln = C.c_size_t()
        cost = fn(a, len(a), b, len(b), C.byref(cig), C.byref(ln))
        assert cost == 2
        text = C.string_at(cig.value)
        assert len(text) == ln.value
        L.astarpa_free_cigar(cig)


LEGACY_WORKER = r'''
import ctypes as C, sys
sys.path.insert(0, {root!r})
import astar_pairwise_aligner_b200 as A
L = A.load_library()
cig, ln = C.c_void_p(), C.c_size_t()
L.astarpa_gcsh(b"ACGTACGTACGTAAC", 15, b"ACGTACGTACGTAAC", 15, {r}, {k}, {pe}, C.byref(cig), C.byref(ln))
print("RETURNED")
'''


def test_legacy_gcsh_symbols_gpu(apa, oracle, tmp_path):
    # astarpa-c/src/lib.rs:54-95: astarpa_gcsh(r, k, prune_end) and astarpa() (= r 2, k 15). r = 1: the heuristic the caller
    # names - GCSH with exact matches of length k, no local pruning, pruning by start - bounds the A*PA2 engine; cost and CIGAR
    # against the oracle configured the same way (configurations 16 / 17 / 18 = k 8 / 12 / 15). Arguments that are not built
    # (r = 2, prune_end) are refused with a message and an abort (the reference's failure mode is a panic), never ignored.
    import subprocess
    import sys
    L = apa.load_library()
    pairs = [apa.generate_pair(n, e, m, 77 + n) for n, e, m in [(0, 0.0, 0), (40, 0.1, 1), (900, 0.05, 0), (5000, 0.1, 2), (20000, 0.05, 3)]]
    pairs.append((b"ACTCGCT", b"AACTCGTT"))  # astarpa-c/example.c
    for k, cfg in ((8, 16), (12, 17), (15, 18)):
        for a, b in pairs:
            cig, ln = C.c_void_p(), C.c_size_t()
            cost = L.astarpa_gcsh(a, len(a), b, len(b), 1, k, False, C.byref(cig), C.byref(ln))
            text = C.string_at(cig.value).decode()
            assert len(text) == ln.value
            L.astarpa_free_cigar(cig)
            oc, ocg, _ = oracle.align(a, b, cfg, True)
            assert (cost, text) == (oc, ocg), (k, len(a))
    for a, b in pairs:  # astarpa(): k = 15, served with exact matches (announced once on stderr)
        cig, ln = C.c_void_p(), C.c_size_t()
        cost = L.astarpa(a, len(a), b, len(b), C.byref(cig), C.byref(ln))
        text = C.string_at(cig.value).decode()
        L.astarpa_free_cigar(cig)
        oc, ocg, _ = oracle.align(a, b, 18, True)
        assert (cost, text) == (oc, ocg)
    root = os.path.dirname(HERE)
    for r, k, pe in ((2, 15, False), (1, 12, True), (1, 40, False)):
        script = tmp_path / f"legacy_{r}_{k}_{int(pe)}.py"
        script.write_text(LEGACY_WORKER.format(root=root, r=r, k=k, pe=pe))
        out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
        assert out.returncode != 0 and "RETURNED" not in out.stdout and "only r = 1" in out.stderr, (r, k, pe, out.stderr[-500:])


def test_drop_in_c_program_gpu(apa, tmp_path):
    # A C caller compiled against include/astarpa.h and linked to libastarpa_c.so, like astarpa-c/example.c.
    import subprocess
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "dropin_example")
    libdir = os.path.dirname(apa.lib_path())
    subprocess.check_call(["gcc", "-O1", os.path.join(HERE, "dropin_example.c"), "-I", os.path.join(root, "include"), "-L", libdir,
                           "-lastarpa_c", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(" ok") == 4


@pytest.mark.parametrize("preset", PRESETS)
def test_hard_shapes_gpu(apa, oracle, preset):
    """Shapes the uniform generator does not produce: low-complexity / tandem-repeat sequences (many identical k-mers:
    duplicate seeds per window, crowded diagonals), very unequal lengths, unrelated sequences, high divergence."""
    rng = np.random.default_rng(11)

    def rnd(n):
        return bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))

    pairs = [
        (b"A" * 600, b"A" * 613),
        (b"ACGT" * 200, b"ACGT" * 190 + b"AC"),
        (b"ACGTTGCAAGTC" * 60, b"ACGTTGCAAGTC" * 61),          # period = k = 12: every seed identical
        (b"AACCGGTT" * 90 + rnd(300), rnd(250) + b"AACCGGTT" * 95),
        (rnd(3000), rnd(40)),
        (rnd(25), rnd(2500)),
        (rnd(4000), rnd(4100)),                                  # unrelated: distance ~ 0.53 n
        (b"", rnd(300)),
        (rnd(300), b""),
        (b"ACGTACGTACG", b"ACGTACGTACG"),                        # n < k
    ]
    pairs += [apa.generate_pair(n, e, model, 4000 + n + model) for n, e, model in
              [(6000, 0.3, 0), (6000, 0.5, 1), (9000, 0.3, 2), (12000, 0.2, 3), (20000, 0.25, 0)]]
    _check_pairs(apa, oracle, pairs, preset)


def test_arena_overflow_retry_gpu(apa, oracle, monkeypatch):
    # A deliberately tiny scratch arena: pairs overflow (ST_OVERFLOW) and are re-run with 4x arenas until they fit.
    monkeypatch.setenv("APA_ARENA_BYTES", "65536")
    pairs = [apa.generate_pair(n, 0.08, 0, 900 + n) for n in (200, 5000, 20000, 40000)]
    for preset in PRESETS:
        _check_pairs(apa, oracle, pairs, preset)
    eng = apa._engine(0)
    bt = eng.upload(*apa._concat(pairs))
    bt.run(1, True)
    assert bt.stats()["retries"] > 0
    bt.free()


@pytest.mark.parametrize("preset", PRESETS)
def test_million_bp_pair_gpu(apa, oracle, preset):
    # BASELINE configs[3]/[4] sizes (n = 1 000 000): cost against the oracle, CIGAR verified against the pair.
    a, b = apa.generate_pair(1000000, 0.05 if preset == 0 else 0.15, 0, 77)
    cost, cigar = apa.AstarPa2(preset, True).align(a, b)
    assert oracle.cigar_verify(cigar, a, b) == cost
    oc, ocg, _ = oracle.align(a, b, preset, True)
    assert cost == oc and cigar == ocg


def test_ten_million_bp_pair_gpu(apa, oracle):
    # BASELINE configs[4]: one ONT-like pair, n = 10 000 000, e = 5 % (one warp walks the whole pair; replicas only
    # across GPUs). Cost and CIGAR against the oracle; the CIGAR is also replayed over the pair.
    a, b = apa.generate_pair(10000000, 0.05, 0, 4242)
    cost, cigar = apa.AstarPa2(1, True).align(a, b)
    assert oracle.cigar_verify(cigar, a, b) == cost
    oc, ocg, _ = oracle.align(a, b, 1, True)
    assert cost == oc and cigar == ocg


def test_config2_cost_only_batch_gpu(apa, oracle):
    # BASELINE configs[1] shape (n = 10 000, e = 5 %, cost only) on a slice of the batch: costs against the oracle for
    # both presets, and cost-only == cost of the traced run.
    pairs = [apa.generate_pair(10000, 0.05, 0, 31415 + s) for s in range(64)]
    for preset in PRESETS:
        costs, cigars = apa.AstarPa2(preset, False).align_batch(pairs)
        assert cigars is None
        traced, _ = apa.AstarPa2(preset, True).align_batch(pairs)
        assert (costs == traced).all()
        for k in range(0, 64, 7):
            assert int(costs[k]) == oracle.align(pairs[k][0], pairs[k][1], preset, False)[0]


@pytest.mark.parametrize("preset", PRESETS)
def test_fused_kernel_path_gpu(apa, oracle, preset, monkeypatch):
    # Batches whose per-pair arenas do not fit in HBM run the fused single-kernel path (per-warp arenas, apa_align_kernel_*);
    # APA_SPLIT=0 forces it. Same bit-exact contract; every register variant.
    monkeypatch.setenv("APA_SPLIT", "0")
    rng = np.random.default_rng(23)
    pairs = [apa.generate_pair(int(rng.integers(0, 9000)), float(rng.choice([0.02, 0.05, 0.15])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(96)]
    for regs in ("64", "48", "40"):
        monkeypatch.setenv("APA_REGS", regs)
        _check_pairs(apa, oracle, pairs, preset)


def test_bad_input_gpu(apa):
    with pytest.raises(apa.AstarPaError):
        apa.AstarPa2(0, True).align_batch([(b"ACGT", b"ACGT"), (b"ACGN", b"ACGT")])


def test_full_size_properties_gpu(apa, oracle):
    """BASELINE config sizes (n = 100k) through size-independent properties: identical inputs cost 0 with an
    all-match CIGAR; cost is symmetric under swapping a and b; a CIGAR verifies against its pair."""
    a, b = apa.generate_pair(100000, 0.05, 0, 424242)
    al = apa.AstarPa2(0, True)
    c0, cg0 = al.align(a, a)
    assert c0 == 0 and cg0 == "100000="
    c1, cg1 = al.align(a, b)
    c2, cg2 = al.align(b, a)
    assert c1 == c2
    assert oracle.cigar_verify(cg1, a, b) == c1 and oracle.cigar_verify(cg2, b, a) == c2


# ---------------------------------------------------------------------------------------------- general parameters
# the other Domains / DoublingTypes / block widths / heuristics of AstarPa2Params, served by the general
# kernel (apa_general.cu). Same bit-exact contract as the presets: cost, CIGAR text and the per-pass band log against the
# oracle configured identically (tests/params_matrix.py mirrors oracle/oracle_capi.cpp preset_params).
from params_matrix import GENERAL_PRESETS, params_for  # noqa: E402


def _gpu_band_log_params(apa, a, b, params, trace=True):
    eng = apa._engine(0)
    L = apa.load_library()
    cap = 16 + 8 * (len(a) + 4) * 40
    buf = np.zeros(cap, dtype=np.int32)
    w = L.apa_debug_band_log_params(eng._h, C.byref(params), int(trace), a, len(a), b, len(b), buf.ctypes.data, cap)
    assert 0 <= w <= cap, (w, L.apa_last_error())
    return buf[:w]


def _check_pairs_params(apa, oracle, pairs, preset, trace=True):
    params = params_for(apa, preset)
    costs, cigars = apa.AstarPa2(params, trace).align_batch(pairs)
    for k, (a, b) in enumerate(pairs):
        oc, ocg, _ = oracle.align(a, b, preset, trace)
        assert int(costs[k]) == oc, (preset, k, len(a), len(b), int(costs[k]), oc)
        if trace and cigars[k] != ocg:
            gl = oracle.parse_band_log(_gpu_band_log_params(apa, a, b, params))
            ol = oracle.band_log(a, b, preset, True)
            first = next((i for i, (x, y) in enumerate(zip(gl, ol)) if x != y), None)
            raise AssertionError(f"CIGAR differs: config {preset} pair {k} n={len(a)} m={len(b)} cost {oc}; "
                                 f"band logs equal: {gl == ol}; first differing pass {first}")


@pytest.mark.parametrize("preset", GENERAL_PRESETS)
def test_general_params_grid_gpu(apa, oracle, preset):
    # the pa-test grid (pa-test/src/lib.rs:24-63), thinned per configuration; golden pairs first
    pairs = [(p["a"].encode(), p["b"].encode()) for p in GOLD["pairs"]]
    pairs += [apa.generate_pair(n, e, model, 31415 + 1000 * n + model)
              for n in NS for e in ES[preset % 3::3] for model in ((preset + n) % 4,)]
    _check_pairs_params(apa, oracle, pairs, preset)
    _check_pairs_params(apa, oracle, pairs[::5], preset, trace=False)


@pytest.mark.parametrize("preset", GENERAL_PRESETS)
def test_general_params_band_log_gpu(apa, oracle, preset):
    big = preset in (9, 13)  # Domain::Full computes n * m cells
    for n, e in [(700, 0.1), (3000, 0.08)] if big else [(700, 0.1), (5000, 0.08), (12000, 0.05)]:
        a, b = apa.generate_pair(n, e, 0, 7 + preset)
        params = params_for(apa, preset)
        assert oracle.parse_band_log(_gpu_band_log_params(apa, a, b, params)) == oracle.band_log(a, b, preset, True), (preset, n)
        cost, cigar = apa.AstarPa2(params, True).align(a, b)
        oc, ocg, _ = oracle.align(a, b, preset, True)
        assert (cost, cigar) == (oc, ocg), (preset, n)


def test_general_kernel_equals_preset_kernels_gpu(apa):
    # the presets through the general kernel give what the tuned kernels give
    rng = np.random.default_rng(11)
    pairs = [apa.generate_pair(int(rng.integers(0, 20000)), float(rng.choice([0.02, 0.05, 0.15])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(200)]
    for preset, params in ((0, apa.AstarPa2Params.simple()), (1, apa.AstarPa2Params.full())):
        c0, g0 = apa.AstarPa2(preset, True).align_batch(pairs)
        c1, g1 = apa.AstarPa2(params, True).align_batch(pairs)
        assert (c0 == c1).all() and g0 == g1


def test_general_params_rejects_unsupported_gpu(apa):
    P = apa.AstarPa2Params
    for bad in (P.full().replace(r=2), P.full().replace(k=20), P.simple().replace(block_width=512), P.simple().replace(sparse=0),
                P.simple().replace(doubling="none"), P.simple().replace(max_g=100)):
        with pytest.raises(apa.AstarPaError):
            apa.AstarPa2(bad, True).align(b"ACGT", b"ACGT")


@pytest.mark.parametrize("preset", GENERAL_PRESETS)
def test_pair_stats_match_oracle_gpu(apa, oracle, preset):
    # SURVEY 8f row 4, the stats surface (align_with_stats, astarpa2/src/lib.rs:200-208): the counters that do not depend
    # on incremental-vs-recompute accounting must equal the oracle's AstarPa2Stats / TraceStats, pair by pair.
    pairs = [apa.generate_pair(n, e, m, 991 + n) for n, e, m in [(0, 0.0, 0), (50, 0.2, 1), (700, 0.1, 2), (3000, 0.05, 0), (3000, 0.2, 3)]]
    for aligner in ([apa.AstarPa2(preset, True)] if preset < 2 else []) + [apa.AstarPa2(params_for(apa, preset), True)]:
        costs, cigars, stats = aligner.align_batch_with_stats(pairs)
        for (a, b), c, cg, st in zip(pairs, costs, cigars, stats):
            oc, ocg, ost = oracle.align(a, b, preset, True)
            assert (int(c), cg) == (oc, ocg)
            par = params_for(apa, preset)
            gcsh = par.domain == 3 and par.heuristic == 2  # h_calls / num_matches are GCSH counters (0 otherwise)
            keys = ("f_max_tries", "h0", "dt_trace_success", "fill_tries") + (("num_matches", "h_calls") if gcsh else ())
            got = {k: st[k] for k in keys}
            want = {k: ost[k] for k in got}
            assert got == want, (preset, len(a), got, want)
            assert st["computed_cells"] >= ost["computed_cells"] >= 0


# ---------------------------------------------------------------------------------------------- intra-pair parallelism
@pytest.mark.parametrize("coop", ["1", "4", "8"])
@pytest.mark.parametrize("preset", PRESETS)
def test_coop_pass_kernel_gpu(apa, oracle, preset, coop, monkeypatch):
    # apa_coop.cuh: the warps of a CTA share the chunks of one pair's band. Small batches pick it automatically (8 warps per
    # pair); APA_COOP forces 1 / 4 / 8 so that every variant sees the same shapes: tall bands (astarpa2_simple at high
    # divergence: tens of chunks per block), single-chunk bands, empty and tiny pairs.
    monkeypatch.setenv("APA_COOP", coop)
    pairs = [apa.generate_pair(n, e, model, 4242 + n + model) for n, e, model in
             [(0, 0.0, 0), (1, 1.0, 0), (40, 0.3, 1), (700, 0.5, 2), (5000, 0.3, 0), (20000, 0.25, 3), (60000, 0.1, 0), (60000, 0.02, 1),
              (100000, 0.05, 0)]]
    _check_pairs(apa, oracle, pairs, preset)
    _check_pairs(apa, oracle, pairs[3:7], preset, trace=False)


@pytest.mark.parametrize("preset", PRESETS)
def test_waves_when_arenas_do_not_fit_gpu(apa, oracle, preset, monkeypatch):
    # A work list whose per-pair arenas exceed the memory budget runs the phase-split path in waves (BASELINE configs[3] does
    # this for real); the budget is forced down so that 40 pairs take several waves.
    rng = np.random.default_rng(5)
    pairs = [apa.generate_pair(int(rng.integers(2000, 30000)), float(rng.choice([0.05, 0.15])), 0, int(rng.integers(1 << 40)))
             for _ in range(40)]
    monkeypatch.setenv("APA_BUDGET_BYTES", str(12 << 20))
    _check_pairs(apa, oracle, pairs, preset)


# ---------------------------------------------------------------------------------------------- pa_bitpacking::search
def test_search_gpu(apa, oracle):
    # SURVEY 8f row 4, second half: the semi-global pattern search (pa-bitpacking/src/search.rs:46-118) on the block-DP
    # kernel. The reference's doc-test vector, then random patterns with wildcards against the oracle: single- and multi-chunk
    # patterns (up to 3 000 rows), single- and multi-slab texts, every unmatched_cost class, empty inputs.
    assert apa.search(b"AC", b"CTTACTTA", 0.0).tolist() == [0, 0, 1, 2, 1, 0, 1, 2, 1, 0, 0]
    import random
    rng = random.Random(3)
    for n_p, n_t in [(1, 0), (1, 1), (5, 70), (63, 200), (64, 256), (65, 257), (130, 1000), (1000, 3000), (1025, 600), (3000, 5000),
                     (20, 100000), (0, 10)]:
        for u in (0.0, 1.0, 0.5, 0.3):
            p = bytes(rng.choice(b"ACGTACGTACGTNYR*acgt") for _ in range(n_p))
            t = bytes(rng.choice(b"ACGTacgt") for _ in range(n_t))
            assert apa.search(p, t, u).tolist() == oracle.search(p, t, u), (n_p, n_t, u)
    # a planted occurrence is found with cost 0 at its end column
    t = bytes(rng.choice(b"ACGT") for _ in range(4000))
    p = t[1234:1300]
    out = apa.search(p, t, 0.0)
    assert out[1300] == 0 and out[:len(t) + 1].min() == 0
    for bad in ((b"AX", b"ACGT", 0.0), (b"AC", b"ACGN", 0.0), (b"AC", b"ACGT", 2.0)):
        with pytest.raises(apa.AstarPaError):
            apa.search(*bad)
    # many text segments (one warp each, 2 * rows warm-up columns): short and multi-chunk patterns on a long text
    for n_p, n_t, u in [(33, 3000000, 0.0), (200, 2000000, 0.3), (1500, 1200000, 0.0)]:
        p = bytes(rng.choice(b"ACGTACGTACGTNYR") for _ in range(n_p))
        t = bytes(rng.choice(b"ACGT") for _ in range(n_t))
        assert apa.search(p, t, u).tolist() == oracle.search(p, t, u), (n_p, n_t, u)


def test_search_trace_gpu(apa, oracle):
    # SearchResult::trace (pa-bitpacking/src/search.rs:135-230) on the GPU against the oracle's literal restatement: planted
    # occurrences with errors, ends along the bottom row and up the right column, windows that have to be doubled, patterns whose
    # length is and is not a multiple of 64, every unmatched_cost class. Where the reference would panic, the call fails.
    import random
    rng = random.Random(5)
    assert oracle.search_trace(b"AC", b"CTTACTTA", 0.0, 5)[0] == "2="
    checked = 0
    for n_p, n_t in [(2, 8), (20, 300), (64, 1000), (65, 1000), (128, 5000), (200, 700), (1000, 9000), (2100, 20000)]:
        for u in (0.0, 0.5, 1.0):
            t = bytearray(rng.choice(b"ACGT") for _ in range(n_t))
            p = bytearray(rng.choice(b"ACGT") for _ in range(n_p))
            if n_t > 2 * n_p:  # plant the pattern with a few edits
                at = rng.randrange(0, n_t - n_p)
                t[at:at + n_p] = p
                for _ in range(max(1, n_p // 25)):
                    t[at + rng.randrange(n_p)] = rng.choice(b"ACGT")
            if n_p >= 20:
                p[3] = ord("N")
            p, t = bytes(p), bytes(t)
            out = oracle.search(p, t, u)
            bottom = out[:n_t + 1]
            idxs = {bottom.index(min(bottom[1:])) if n_t else 0, n_t, n_t // 2, n_t + n_p, n_t + n_p // 2, 1, rng.randrange(len(out))}
            for idx in sorted(idxs):
                try:
                    want = oracle.search_trace(p, t, u, idx)
                except oracle.OraclePanic:
                    with pytest.raises(apa.AstarPaError):
                        apa.search_trace(p, t, u, idx)
                    continue
                assert apa.search_trace(p, t, u, idx) == want, (n_p, n_t, u, idx)
                checked += 1
    assert checked > 100
    with pytest.raises(apa.AstarPaError):
        apa.search_trace(b"ACGT", b"ACGTACGT", 0.0, 100)


@pytest.mark.parametrize("preset", GENERAL_PRESETS)
def test_committed_config_vectors_gpu(apa, oracle, preset):
    # the CUDA path against the committed fixtures (tests/golden/config_vectors.json), not only against the live oracle:
    # cost, CIGAR text, f_max tries and the band-log digest of 4 seeded pairs per configuration.
    import hashlib
    g = json.load(open(os.path.join(HERE, "golden", "config_vectors.json")))
    params = params_for(apa, preset)
    pairs = [apa.generate_pair(n, e, model, seed) for n, e, model, seed in g["cases"]]
    aligners = [apa.AstarPa2(params, True)] + ([apa.AstarPa2(preset, True)] if preset < 2 else [])
    for al in aligners:
        costs, cigars, stats = al.align_batch_with_stats(pairs)
        for k, want in enumerate(g["configs"][str(preset)]):
            assert (int(costs[k]), cigars[k], stats[k]["f_max_tries"]) == (want["cost"], want["cigar"], want["f_max_tries"]), (preset, k)
    for (a, b), want in zip(pairs, g["configs"][str(preset)]):
        log = oracle.parse_band_log(_gpu_band_log_params(apa, a, b, params))
        assert hashlib.sha256(json.dumps(log).encode()).hexdigest()[:16] == want["band_log"], preset


def test_incremental_doubling_cells_gpu(apa, oracle):
    # Incremental doubling and block reuse as WORK SAVERS (astarpa2/src/blocks.rs:190-197,342-469): on multi-pass pairs the cells
    # the GPU evaluates (64 * lanes * cols of every computed range) stay within 10 % of the reference's BlockStats count, one
    # warp per pair and with a whole CTA per pair; results are those of the oracle as everywhere else.
    pairs = [apa.generate_pair(n, e, m, 70 + n) for n, e, m in [(20000, 0.15, 0), (100000, 0.15, 0), (20000, 0.3, 0), (40000, 0.3, 3),
                                                                (60000, 0.2, 0), (10000, 0.4, 0)]]
    for preset in (1, 2, 3):
        aligners = [apa.AstarPa2(params_for(apa, preset), True)] + ([apa.AstarPa2(1, True)] if preset == 1 else [])
        for aligner in aligners:
            costs, cigars, stats = aligner.align_batch_with_stats(pairs)
            for (a, b), c, cg, st in zip(pairs, costs, cigars, stats):
                oc, ocg, ost = oracle.align(a, b, preset, True)
                assert (int(c), cg) == (oc, ocg)
                tries = ost["f_max_tries"]
                assert st["f_max_tries"] == tries >= 3  # 3 .. 7 passes each (the oracle says)
                # The first pass of a pair writes no h row here (single-pass pairs pay nothing), so the second pass recomputes
                # what the first had fixed: with the band doubling from pass to pass that is 2^-(tries - 1) of the reference's
                # total on top - within 10 % from five passes on.
                limit = 1.1 if tries >= 5 else 1.0 + 2.0 ** -(tries - 2)
                assert ost["computed_cells"] <= st["computed_cells"] <= limit * ost["computed_cells"], \
                    (preset, len(a), tries, st["computed_cells"], ost["computed_cells"])
