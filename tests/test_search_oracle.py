"""pa_bitpacking::search: the oracle restatement against the reference's own doc-test vector and against an
independent dynamic program of the documented semantics (pa-bitpacking/src/search.rs:16-45)."""
import math
import random

import pytest


def _match(pc, tc):
    pc, tc = pc.upper(), tc.upper()
    if pc in "N*":
        return True
    if pc == "Y":
        return tc in "CT"
    if pc == "R":
        return tc in "AG"
    return pc == tc


def brute_search(p: bytes, t: bytes, u: float):
    """Semi-global DP: free start anywhere in the text (top row 0); unmatched pattern prefix / suffix rows cost 1 where the
    reference sets a bit of v0 (every ceil(i / u)-th row, search.rs:58-66). Output: bottom row, then up the right column with
    the unmatched suffix added."""
    n_p, n_t = len(p), len(t)
    ones = set()
    if u > 0:
        i = 0
        while True:
            idx = math.ceil(i / u)
            if idx >= n_p:
                break
            ones.add(idx)
            i += 1
    col0 = [0]
    for j in range(n_p):
        col0.append(col0[-1] + (1 if j in ones else 0))
    D = [[0] * (n_t + 1) for _ in range(n_p + 1)]
    for j in range(n_p + 1):
        D[j][0] = col0[j]
    for i in range(1, n_t + 1):
        for j in range(1, n_p + 1):
            D[j][i] = min(D[j - 1][i] + 1, D[j][i - 1] + 1, D[j - 1][i - 1] + (0 if _match(chr(p[j - 1]), chr(t[i - 1])) else 1))
    return D[n_p][:] + [D[j][n_t] + col0[n_p] - col0[j] for j in range(n_p - 1, -1, -1)]


def test_search_reference_doctest(oracle):
    # pa-bitpacking/src/search.rs:29-32
    assert oracle.search(b"AC", b"CTTACTTA", 0.0) == [0, 0, 1, 2, 1, 0, 1, 2, 1, 0, 0]


def test_search_against_independent_dp(oracle):
    rng = random.Random(1)
    for _ in range(400):
        n_p = rng.choice([1, 2, 5, 63, 64, 65, 100, 128, 130])
        n_t = rng.choice([0, 1, 3, 10, 70, 200])
        p = bytes(rng.choice(b"ACGTNYR*acgt") for _ in range(n_p))
        t = bytes(rng.choice(b"ACGTacgt") for _ in range(n_t))
        u = rng.choice([0.0, 1.0, 0.5, 0.25])
        assert oracle.search(p, t, u) == brute_search(p, t, u), (n_p, n_t, u)


def test_search_panics_like_the_reference(oracle):
    with pytest.raises(oracle.OraclePanic):
        oracle.search(b"AC", b"ACGN", 0.0)  # text must be acgtACGT (profile.rs:37)
    with pytest.raises(oracle.OraclePanic):
        oracle.search(b"AX", b"ACGT", 0.0)
    with pytest.raises(oracle.OraclePanic):
        oracle.search(b"AC", b"ACGT", 1.5)


def test_search_trace_oracle_properties(oracle):
    # SearchResult::trace (search.rs:135-230) restated: the doc-test's occurrence, and for random planted occurrences the CIGAR
    # replayed over text[start.i:end.i] x pattern[start.j:end.j] costs exactly out[idx] (unmatched_cost = 0: free start in the top
    # row and in the left column), uses every base of both slices, and '=' / 'X' agree with the bases.
    import random
    import re
    assert oracle.search_trace(b"AC", b"CTTACTTA", 0.0, 5) == ("2=", (3, 0), (5, 2), 0)
    rng = random.Random(11)
    for n_p, n_t in [(12, 200), (64, 900), (100, 3000), (700, 5000)]:
        t = bytearray(rng.choice(b"ACGT") for _ in range(n_t))
        p = bytes(rng.choice(b"ACGT") for _ in range(n_p))
        at = rng.randrange(0, n_t - n_p)
        t[at:at + n_p] = p
        for _ in range(n_p // 20):
            t[at + rng.randrange(n_p)] = rng.choice(b"ACGT")
        t = bytes(t)
        out = oracle.search(p, t, 0.0)
        for idx in (out.index(min(out[1:n_t + 1]), 1), n_t // 3, n_t, n_t + n_p // 2):
            cigar, (si, sj), (ei, ej), cost = oracle.search_trace(p, t, 0.0, idx)
            i, j, c = si, sj, 0
            for cnt, op in re.findall(r"(\d*)([=XID])", cigar):
                for _ in range(int(cnt or 1)):
                    if op in "=X":
                        assert (t[i] == p[j]) == (op == "="), (idx, i, j)
                        i, j, c = i + 1, j + 1, c + (op == "X")
                    elif op == "D":
                        i, c = i + 1, c + 1
                    else:
                        j, c = j + 1, c + 1
            assert (i, j) == (ei, ej) and c == cost == out[idx] and (si == 0 or sj == 0), (n_p, idx, cigar[:40])
