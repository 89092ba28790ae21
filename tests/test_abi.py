"""CPU tests of the drop-in boundary: libastarpa_c.so loads without a GPU and exports every symbol the
headers declare (include/astarpa.h mirrors astarpa-c/astarpa.h:15-65). No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(astarpa\w*|apa_\w+)\s*\(", src)) - {"apa_batch_stats"})


def test_exports_all_declared_symbols(apa):
    L = ctypes.CDLL(apa.lib_path())
    names = _declared("astarpa.h") + _declared("astarpa_b200.h")
    assert {"astarpa2_simple", "astarpa2_full", "astarpa", "astarpa_gcsh", "astarpa_free_cigar"} <= set(names)
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback(apa):
    """Without a device the engine refuses to start (the product path never routes through oracle/)."""
    import astar_pairwise_aligner_b200 as A
    if A.load_library().apa_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(A.AstarPaError):
        A.Engine(0)


def test_no_cpu_fallback_shared_engine_and_multi(apa):
    """The process-wide engine and the multi-GPU entry fail the same way without a device."""
    import numpy as np
    import astar_pairwise_aligner_b200 as A
    if A.load_library().apa_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(A.AstarPaError):
        A.Engine.shared(0)
    with pytest.raises(A.AstarPaError):
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)
        off = np.array([0, 4], dtype=np.int64)
        A.align_batch_multi([0], seq, off, seq, off)


def test_library_holds_sm100a_kernels(apa):
    """The shipped library carries sm_100a machine code for every kernel DESIGN.md names (static check, no GPU)."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([tool, "-lelf", apa.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in elfs and not re.search(r"sm_(?!100a)\d+", elfs), elfs
    res = subprocess.run([tool, "-res-usage", apa.lib_path()], capture_output=True, text=True).stdout
    for k in ("apa_phase_build_kernel", "apa_phase_pass_kernel", "apa_phase_cont_kernel", "apa_phase_trace_kernel",
              "apa_phase_pass_coop_kernel", "apa_align_kernel_r64", "apa_general_kernel", "apa_pack_kernel", "apa_block_kernel",
              "apa_search_kernel", "apa_search_trace_kernel", "apa_peak_kernel"):
        assert k in res, k


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "astar_pairwise_aligner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt and "oracle_lib" not in txt and "liboracle" not in txt, f


def test_generator_is_deterministic(apa):
    a1, b1 = apa.generate_pair(1000, 0.05, 0, 31415)
    a2, b2 = apa.generate_pair(1000, 0.05, 0, 31415)
    assert (a1, b1) == (a2, b2) and len(a1) == 1000 and set(a1 + b1) <= set(b"ACGT")
    aa, ao, bb, bo = apa.generate_batch(3, 1000, 0.05, 0, 31415, threads=2)
    assert aa[:1000].tobytes() == a1 and bb[bo[0]:bo[1]].tobytes() == b1


def test_host_packer_matches_layout(apa):
    """apa_pack_planes_host (K0, BitProfile::build layout, pa-bitpacking/src/profile.rs:112-133): negated rank-bit planes per
    32 bases, zero padding, ACGT validation — AVX2 / AVX-512 bodies against a scalar restatement, for ragged lengths and
    segment starts (the engine packs long sequences in segments)."""
    import numpy as np
    L = apa.load_library()
    L.apa_pack_planes_host.restype = ctypes.c_int
    L.apa_pack_planes_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    rng = np.random.default_rng(5)
    rank = {65: 0, 67: 1, 71: 2, 84: 3}
    for n in [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4097]:
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).astype(np.uint8)
        nhw = ((n + 63) // 64) * 2 + 2
        ref = np.zeros(2 * nhw, dtype=np.uint32)
        for i, c in enumerate(seq):
            r = rank[int(c)]
            ref[2 * (i // 32)] |= np.uint32(((r & 1) ^ 1) << (i % 32))
            ref[2 * (i // 32) + 1] |= np.uint32(((r >> 1) ^ 1) << (i % 32))
        buf = np.ascontiguousarray(np.concatenate([seq, np.zeros(64, np.uint8)]))
        for h0 in (0, 1, 2, 3):
            if h0 >= nhw:
                continue
            out = ref.copy()
            out[2 * h0:] = 0xDEADBEEF
            assert L.apa_pack_planes_host(buf.ctypes.data, n, h0, nhw, out.ctypes.data) == 0
            assert (out == ref).all(), (n, h0)
        if n > 40:
            buf[37] = ord("N")
            assert L.apa_pack_planes_host(buf.ctypes.data, n, 0, nhw, ref.ctypes.data) == 1


def test_params_presets_and_struct_layout(apa):
    """apa_params / apa_pair_stats as the C header declares them (17 int32/float fields = 68 bytes, 8 int64 = 64 bytes), and the
    presets filled by apa_params_preset equal AstarPa2Params::simple() / ::full() (astarpa2/src/params.rs:70-128). Host only."""
    assert ctypes.sizeof(apa.AstarPa2Params) == 68 and ctypes.sizeof(apa.PairStats) == 64
    s, f = apa.AstarPa2Params.simple(), apa.AstarPa2Params.full()
    for q in (s, f):
        assert (q.domain, q.doubling, q.doubling_start, q.block_width) == (3, 1, 2, 256)  # Astar, BandDoubling{H0, 2.0}
        assert (q.factor, q.sparse, q.dt_trace, q.max_g, q.fr_drop, q.sparse_h) == (2.0, 1, 1, 40, 10, 1)
    assert (s.heuristic, s.prune, s.incremental_doubling) == (1, 0, 0)  # GapCost, no pruning (params.rs:70-96)
    assert (f.heuristic, f.k, f.r, f.p, f.prune, f.incremental_doubling) == (2, 12, 1, 14, 1, 1)  # GCSH (params.rs:98-128)
    nw = apa.AstarPa2Params.nw()
    assert (nw.domain, nw.doubling, nw.dt_trace) == (0, 0, 0)  # params.rs:46-68
    g = f.replace(domain="gap_gap", block_width=64, doubling_start="gap")
    assert (g.domain, g.block_width, g.doubling_start, g.k) == (2, 64, 1, 12) and f.domain == 3  # replace() copies


FULL_JSON = ('{"name":"full","domain":{"Astar":null},"heuristic":{"type":"GCSH","r":1,"k":12,"p":14,"prune":"Start","kmin":null,'
             '"kmax":null,"max_matches":null,"skip_prune":null},"doubling":{"BandDoubling":{"start":"H0","factor":2.0}},"block_width":256,'
             '"front":{"sparse":true,"simd":true,"no_ilp":false,"incremental_doubling":true,"dt_trace":true,"max_g":40,"fr_drop":10},'
             '"sparse_h":true,"prune":true,"viz":false}')
SIMPLE_JSON = ('{"name":"simple","domain":{"Astar":null},"heuristic":{"type":"Gap","r":2,"k":15,"p":0,"prune":"Start","kmin":null,'
               '"kmax":null,"max_matches":null,"skip_prune":null},"doubling":{"BandDoubling":{"start":"H0","factor":2.0}},"block_width":256,'
               '"front":{"sparse":true,"simd":true,"no_ilp":false,"incremental_doubling":false,"dt_trace":true,"max_g":40,"fr_drop":10},'
               '"sparse_h":true,"prune":false,"viz":false}')


def test_params_from_json(apa):
    """apa_params_from_json: the serde form of AstarPa2Params (astarpa2/src/params.rs:7-42, what pa-bench job files hold). The two
    documents above are serde_json of AstarPa2Params::full() / ::simple() (params.rs:70-128) written out field by field."""
    P = apa.AstarPa2Params

    def fields(q):
        return {k: getattr(q, k) for k, _ in q._fields_}

    assert fields(P.from_json(FULL_JSON)) == fields(P.full())
    want = fields(P.simple())
    got = fields(P.from_json(SIMPLE_JSON))
    assert got == want, {k: (got[k], want[k]) for k in got if got[k] != want[k]}
    # serde defaults for absent fields; the other domains / doubling types
    q = P.from_json('{"domain":"GapGap","heuristic":{"type":"None"},"doubling":{"LinearSearch":{"start":"Gap","delta":48.0}},'
                    '"block_width":64,"front":{"sparse":true}}')
    assert (q.domain, q.heuristic, q.doubling, q.doubling_start, q.delta, q.block_width, q.dt_trace, q.sparse_h, q.prune) == \
           (2, 0, 2, 1, 48, 64, 0, 0, 0)
    q = P.from_json('{"domain":"Full","heuristic":{"type":"Zero"},"doubling":"None","block_width":1,"front":{"sparse":true,"dt_trace":false}}')
    assert (q.domain, q.doubling, q.block_width) == (0, 0, 1)
    # Prune::None inside the heuristic switches pruning off
    assert P.from_json(FULL_JSON.replace('"prune":"Start"', '"prune":"None"')).prune == 0
    # refused, never ignored: inexact matches, other heuristics, Prune::Both, LocalDoubling, viz, unknown fields, malformed text
    for bad in (FULL_JSON.replace('"r":1', '"r":2'), FULL_JSON.replace('"GCSH"', '"SH"'), FULL_JSON.replace('"prune":"Start"', '"prune":"Both"'),
                FULL_JSON.replace('{"BandDoubling":{"start":"H0","factor":2.0}}', '"LocalDoubling"'), FULL_JSON.replace('"viz":false', '"viz":true'),
                FULL_JSON.replace('"sparse_h":true', '"sparse_hh":true'), FULL_JSON.replace('"kmin":null', '"kmin":10'), FULL_JSON[:-1],
                '{"domain":"Full"}', FULL_JSON.replace('"delta"', '"x"').replace('2.0}}', '2.0,"start_increment":3}}')):
        with pytest.raises(apa.AstarPaError):
            P.from_json(bad)
