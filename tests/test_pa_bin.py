"""pa-bin equivalent and the C++ host mirror (reference: pa-bin/src/lib.rs:49-131, main.rs:9-37).
CPU tests cover the parsers (.seq / .txt / FASTA / directory), the CLI contract and Cigar; GPU tests run the whole tool."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.fixture(scope="session")
def pa_bin(apa):
    p = os.path.join(os.path.dirname(apa.lib_path()), "pa-bin")
    if not os.path.exists(p):
        import __graft_entry__
        __graft_entry__.build()
    assert os.path.exists(p)
    return p


@pytest.fixture(scope="session")
def host_test_exe(apa, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "host_api_test")
    libdir = os.path.dirname(apa.lib_path())
    subprocess.check_call(["g++", "-O1", "-std=c++17", os.path.join(HERE, "host_api_test.cpp"), "-I", os.path.join(ROOT, "include"),
                           "-L", libdir, "-lastarpa_c", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def _dry(pa_bin, *args):
    out = subprocess.run([pa_bin, *args, "--dry-run"], capture_output=True, text=True, timeout=60)
    return out.returncode, out.stdout.split(), out.stderr


PAIRS = [(b"ACGTACGT", b"ACGTTCGT"), (b"A", b"CCCC"), (b"", b"ACG"), (b"GATTACA" * 30, b"GATTACA" * 29 + b"GAT")]


def _expect(pairs):
    return ["%d,%d,%s,%s" % (len(a), len(b), _fnv(a), _fnv(b)) for a, b in pairs]


def test_seq_and_txt_formats(pa_bin, tmp_path):
    seq = tmp_path / "in.seq"
    seq.write_bytes(b"".join(b">" + a + b"\n<" + b + b"\n" for a, b in PAIRS) + b">ACGT\n")  # trailing unpaired line ignored
    rc, lines, err = _dry(pa_bin, "-i", str(seq))
    assert rc == 0 and lines == _expect(PAIRS), err
    txt = tmp_path / "in.txt"
    txt.write_bytes(b"".join(a + b"\r\n" + b + b"\r\n" for a, b in PAIRS))  # CRLF: lines() strips both
    rc, lines, err = _dry(pa_bin, "--input=" + str(txt))
    assert rc == 0 and lines == _expect(PAIRS), err


def test_seq_format_requires_prefixes(pa_bin, tmp_path):
    bad = tmp_path / "bad.seq"
    bad.write_bytes(b">ACGT\nACGT\n")  # the reference asserts '<' (pa-bin/src/lib.rs:87)
    rc, _, err = _dry(pa_bin, "-i", str(bad))
    assert rc != 0 and "'<'" in err


def test_fasta_format(pa_bin, tmp_path):
    fa = tmp_path / "in.fasta"
    recs = []
    for k, (a, b) in enumerate(PAIRS):
        for nm, s in (("a", a), ("b", b)):
            body = b"\n".join(s[i:i + 60] for i in range(0, len(s), 60))  # multi-line records
            recs.append(b">%s%d some description\n" % (nm.encode(), k) + body + b"\n")
    fa.write_bytes(b"".join(recs) + b">orphan\nACGT\n")  # unpaired last record ignored
    rc, lines, err = _dry(pa_bin, "-i", str(fa))
    assert rc == 0 and lines == _expect(PAIRS), err
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"ACGT\n>x\nACGT\n")
    assert _dry(pa_bin, "-i", str(bad))[0] != 0


def test_directory_input_and_unknown_extension(pa_bin, tmp_path):
    d = tmp_path / "dir"
    d.mkdir()
    (d / "b.txt").write_bytes(PAIRS[1][0] + b"\n" + PAIRS[1][1] + b"\n")
    (d / "a.seq").write_bytes(b">" + PAIRS[0][0] + b"\n<" + PAIRS[0][1] + b"\n")
    rc, lines, err = _dry(pa_bin, "-i", str(d))
    assert rc == 0 and lines == _expect(PAIRS[:2]), err  # sorted by file name
    (d / "c.xyz").write_bytes(b"ACGT\nACGT\n")
    rc, _, err = _dry(pa_bin, "-i", str(d))
    assert rc != 0 and "extension" in err


def test_cli_contract(pa_bin, apa):
    # clap group input_type: exactly one of --input / --length (pa-bin/src/lib.rs:44-48)
    assert subprocess.run([pa_bin], capture_output=True).returncode == 2
    assert subprocess.run([pa_bin, "-i", "x.seq", "-n", "10"], capture_output=True).returncode == 2
    assert subprocess.run([pa_bin, "-n", "10", "--aligner", "edlib"], capture_output=True).returncode == 2
    assert subprocess.run([pa_bin, "--help"], capture_output=True).returncode == 0
    # generated input is deterministic in (--seed, pair index) and matches the library generator
    rc, lines, _ = _dry(pa_bin, "-n", "300", "-e", "0.1", "--seed", "7", "--cnt", "3")
    assert rc == 0 and lines == _expect([apa.generate_pair(300, 0.1, 0, 7 + p) for p in range(3)])
    rc, lines, _ = _dry(pa_bin, "-n", "300", "-e", "0.1", "--seed", "7", "--error-model", "noisy-delete")
    assert rc == 0 and lines == _expect([apa.generate_pair(300, 0.1, 2, 7)])


def test_pa_bin_fails_loudly_without_gpu(pa_bin, apa):
    if apa.load_library().apa_device_count() > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([pa_bin, "-n", "100", "--seed", "1"], capture_output=True, text=True)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


def test_host_cigar_cpp(host_test_exe, apa):
    out = subprocess.run([host_test_exe, "cigar"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    if apa.load_library().apa_device_count() == 0:
        out = subprocess.run([host_test_exe, "nodevice"], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_host_api_cpp_gpu(host_test_exe):
    out = subprocess.run([host_test_exe, "align"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("aligner,preset", [("astarpa2-simple", 0), ("astarpa2-full", 1)])
def test_pa_bin_end_to_end_gpu(pa_bin, apa, oracle, tmp_path, aligner, preset):
    # BASELINE configs[0]: n = 1000, e = 5 % through the CLI; plus file input. CSV lines must equal the oracle's (cost, CIGAR).
    pairs = [apa.generate_pair(1000, 0.05, 0, 31415 + p) for p in range(5)]
    out = tmp_path / "gen.csv"
    r = subprocess.run([pa_bin, "-n", "1000", "-e", "0.05", "--seed", "31415", "--cnt", "5", "--aligner", aligner, "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    want = ["%d,%s" % oracle.align(a, b, preset, True)[:2] for a, b in pairs]
    assert out.read_text().split() == want
    seq = tmp_path / "in.seq"
    seq.write_bytes(b"".join(b">" + a + b"\n<" + b + b"\n" for a, b in pairs))
    out2 = tmp_path / "file.csv"
    r = subprocess.run([pa_bin, "-i", str(seq), "--aligner", aligner, "-o", str(out2), "--batch-bases", "4500"],  # several batches
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert out2.read_text().split() == want
    out3 = tmp_path / "cost.csv"
    r = subprocess.run([pa_bin, "-i", str(seq), "--aligner", aligner, "-o", str(out3), "--cost-only"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and out3.read_text().split() == [w.split(",")[0] + "," for w in want]


@pytest.mark.gpu
def test_pa_bin_params_json_gpu(pa_bin, apa, oracle, tmp_path):
    # --params: AstarPa2Params in the reference's serde JSON form (a pa-bench job's `params`, astarpa2/src/params.rs:7-42).
    # The reference's test configuration `dt_trace` (astarpa2/src/tests.rs:91-104: GCSH k = 15 exact, BandDoubling from Gap,
    # BlockParams::default() with dt_trace) = oracle configuration 3; unsupported values are refused before any alignment.
    pairs = [apa.generate_pair(1500, 0.08, m, 500 + m) for m in range(4)]
    seq = tmp_path / "in.seq"
    seq.write_bytes(b"".join(b">" + a + b"\n<" + b + b"\n" for a, b in pairs))
    js = tmp_path / "params.json"
    js.write_text('{"name":"dt_trace","domain":{"Astar":null},"heuristic":{"type":"GCSH","r":1,"k":15,"p":0,"prune":"Start"},'
                  '"doubling":{"BandDoubling":{"start":"Gap","factor":2.0}},"block_width":256,'
                  '"front":{"sparse":true,"simd":true,"no_ilp":false,"incremental_doubling":true,"dt_trace":true,"max_g":40,"fr_drop":20},'
                  '"sparse_h":true,"prune":true}')
    out = tmp_path / "out.csv"
    r = subprocess.run([pa_bin, "-i", str(seq), "--params", str(js), "-o", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert out.read_text().split() == ["%d,%s" % oracle.align(a, b, 3, True)[:2] for a, b in pairs]
    js.write_text(js.read_text().replace('"r":1', '"r":2'))
    r = subprocess.run([pa_bin, "-i", str(seq), "--params", str(js), "-o", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "r must be 1" in r.stderr
