"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's A*PA2 path (see oracle/*.hpp headers). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

PRESET_SIMPLE, PRESET_FULL = 0, 1


class OracleStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "f_max_tries", "num_blocks", "computed_lanes", "computed_cells", "h_calls", "num_matches", "h0",
        "dt_trace_tries", "dt_trace_success", "fill_tries", "fill_success")]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        u8p = C.c_char_p
        L.oracle_align.restype = C.c_int64
        L.oracle_align.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t), C.POINTER(OracleStats), C.c_int, C.c_char_p, C.c_size_t]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_align_log.restype = C.c_int64
        L.oracle_align_log.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, C.c_void_p, C.c_size_t]
        for f in (L.oracle_levenshtein, L.oracle_levenshtein_dp):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]
        L.oracle_cigar_verify.restype = C.c_int64
        L.oracle_cigar_verify.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, u8p, C.c_size_t]
        L.oracle_bp_compute.restype = C.c_int64
        L.oracle_bp_compute.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.oracle_bp_compute_bench.restype = C.c_int64
        L.oracle_bp_compute_bench.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int]
        L.oracle_to_qgram.restype = C.c_uint64
        L.oracle_to_qgram.argtypes = [u8p, C.c_int]
        L.oracle_gcsh_info.restype = C.c_int64
        L.oracle_gcsh_info.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                       C.c_void_p, C.c_size_t]
        L.oracle_align_batch.restype = C.c_double
        L.oracle_align_batch.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_search.restype = C.c_int64
        L.oracle_search.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_float, C.c_void_p]
        L.oracle_search_trace.restype = C.c_int64
        L.oracle_search_trace.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_float, C.c_uint64, C.c_char_p, C.c_size_t, C.c_void_p]
        L.oracle_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


class OraclePanic(RuntimeError):
    pass


def align(a: bytes, b: bytes, preset=PRESET_FULL, trace=True, self_check=False):
    """Returns (cost, cigar_text or None, stats dict). Raises OraclePanic where the reference would panic."""
    L = lib()
    cig = C.c_void_p()
    clen = C.c_size_t()
    st = OracleStats()
    err = C.create_string_buffer(256)
    cost = L.oracle_align(preset, int(trace), a, len(a), b, len(b), C.byref(cig), C.byref(clen), C.byref(st),
                          int(self_check), err, 256)
    if cost < 0:
        raise OraclePanic(err.value.decode())
    text = None
    if cig.value:
        text = C.string_at(cig.value, clen.value).decode()
        L.oracle_free(cig)
    return cost, text, {k: getattr(st, k) for k, _ in OracleStats._fields_}


def band_log(a: bytes, b: bytes, preset=PRESET_FULL, trace=True):
    L = lib()
    cap = 16 + 8 * (len(a) // (64 if preset < 2 else 1) + 4) * 40  # configurations >= 2 may use block_width 1
    buf = np.zeros(cap, dtype=np.int32)
    w = L.oracle_align_log(preset, int(trace), a, len(a), b, len(b), buf.ctypes.data, cap)
    if w < 0:
        raise OraclePanic("panic in band log")
    assert w <= cap
    return parse_band_log(buf[:w])


def parse_band_log(buf):
    out = []
    pos = 1
    for _ in range(int(buf[0])):
        f_max, nb = int(buf[pos]), int(buf[pos + 1])
        pos += 2
        rows = buf[pos:pos + 4 * nb].reshape(nb, 4).tolist()
        pos += 4 * nb
        out.append((f_max, rows))
    return out


def levenshtein(a: bytes, b: bytes) -> int:
    return lib().oracle_levenshtein(a, len(a), b, len(b))


def levenshtein_dp(a: bytes, b: bytes) -> int:
    return lib().oracle_levenshtein_dp(a, len(a), b, len(b))


def cigar_verify(cigar: str, a: bytes, b: bytes) -> int:
    c = cigar.encode()
    return lib().oracle_cigar_verify(c, len(c), a, len(a), b, len(b))


def align_batch(a_all, a_off, b_all, b_off, preset=PRESET_FULL, trace=True, threads=None):
    """Multi-threaded timing leg. Returns (seconds, costs, cigar_lens, computed_cells, cigar_hash)."""
    L = lib()
    n = len(a_off) - 1
    threads = threads or L.oracle_hardware_threads()
    costs = np.zeros(n, dtype=np.int64)
    clens = np.zeros(n, dtype=np.int64)
    cells = np.zeros(n, dtype=np.int64)
    chash = np.zeros(n, dtype=np.uint64)
    a_all = np.ascontiguousarray(a_all, dtype=np.uint8)
    b_all = np.ascontiguousarray(b_all, dtype=np.uint8)
    a_off = np.ascontiguousarray(a_off, dtype=np.int64)
    b_off = np.ascontiguousarray(b_off, dtype=np.int64)
    sec = L.oracle_align_batch(preset, int(trace), n, a_all.ctypes.data, a_off.ctypes.data, b_all.ctypes.data,
                               b_off.ctypes.data, threads, costs.ctypes.data, clens.ctypes.data, cells.ctypes.data,
                               chash.ctypes.data)
    return sec, costs, clens, cells, chash


def search(pattern: bytes, text: bytes, unmatched_cost: float = 0.0):
    """pa_bitpacking::search(...).out (pa-bitpacking/src/search.rs:46-118) as a list of ints."""
    out = np.zeros(len(pattern) + len(text) + 1, dtype=np.int32)
    w = lib().oracle_search(pattern, len(pattern), text, len(text), unmatched_cost, out.ctypes.data)
    if w < 0:
        raise OraclePanic("panic in search")
    return out[:w].tolist()


def search_trace(pattern: bytes, text: bytes, unmatched_cost: float, idx: int):
    """SearchResult::trace(idx) (pa-bitpacking/src/search.rs:135-230): (cigar text, (start_i, start_j), (end_i, end_j), cost)."""
    cap = 2 * (len(pattern) + len(text)) + 16
    buf = C.create_string_buffer(cap)
    pos = np.zeros(5, dtype=np.int32)
    w = lib().oracle_search_trace(pattern, len(pattern), text, len(text), unmatched_cost, idx, buf, cap, pos.ctypes.data)
    if w < 0:
        raise OraclePanic("panic in search trace")
    return buf.value.decode(), (int(pos[0]), int(pos[1])), (int(pos[2]), int(pos[3])), int(pos[4])
