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
