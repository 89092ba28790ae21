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
