"""B200-native A*PA2 hot path behind the reference's interface.

Host-side mirror (Python, ctypes over the C-ABI shared library ``libastarpa_c.so``) of the Rust entry points
of the reference for this path:

* ``astarpa2_simple(a, b)`` / ``astarpa2_full(a, b)``  -> ``(cost, cigar)``   (astarpa2/src/lib.rs:44-53)
* ``AstarPa2(preset, trace).align(a, b)`` / ``.cost(a, b)``                    (astarpa2/src/lib.rs:177-215,
  the ``pa_types::Aligner`` trait: ``align(&mut self, a, b) -> (Cost, Option<Cigar>)``)
* ``AstarPa2.align_batch(pairs)`` — the batch extension a GPU needs (include/astarpa_b200.h).

All compute happens in hand-written sm_100a CUDA kernels inside the shared library. There is no CPU fallback:
if the library is missing or no B200-class device is usable, calls raise ``AstarPaError``.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["AstarPa2", "AstarPa2Params", "AstarPaError", "Engine", "astarpa2_simple", "astarpa2_full", "generate_pair",
           "generate_batch", "search", "PRESET_SIMPLE", "PRESET_FULL", "lib_path", "load_library", "align_batch_multi", "free_pool",
           "cigar_digests", "gen_library"]

PRESET_SIMPLE, PRESET_FULL = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AstarPaError(RuntimeError):
    pass


class BatchStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("computed_cells", C.c_uint64),
                ("dp_word_steps", C.c_uint64), ("passes", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("retries", C.c_uint64), ("fill_blocks", C.c_uint64), ("dt_blocks", C.c_uint64),
                ("phase_cycles", C.c_uint64 * 8), ("phase_ms", C.c_double * 3), ("score_calls", C.c_uint64), ("score_probes", C.c_uint64),
                ("pass_warps_per_pair", C.c_uint32), ("upload_mode", C.c_uint32), ("upload_chunks", C.c_uint32), ("waves", C.c_uint32), ("dp_issue_steps", C.c_uint64), ("upload_chunks_raw", C.c_uint32), ("overlapped", C.c_uint32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_cycles"] = list(d["phase_cycles"])
        d["phase_ms"] = list(d["phase_ms"])
        return d


class PairStats(C.Structure):
    """``apa_pair_stats``: per-pair counters (AstarPa2Stats / TraceStats of the reference, see include/astarpa_b200.h)."""
    _fields_ = [(k, C.c_int64) for k in ("f_max_tries", "h0", "num_matches", "h_calls", "computed_cells", "dt_trace_success",
                                         "fill_tries", "reserved")]


class AstarPa2Params(C.Structure):
    """``apa_params`` (include/astarpa_b200.h): the flat AstarPa2Params of the reference (astarpa2/src/params.rs:8-42).

    ``AstarPa2Params.simple()`` / ``.full()`` are the presets; ``.nw()`` is the full n*m alignment (params.rs:46-68);
    other configurations are built by setting fields, e.g. the reference's test matrix (astarpa2/src/tests.rs:19-119)."""
    DOMAIN = {"full": 0, "gap_start": 1, "gap_gap": 2, "astar": 3}
    HEURISTIC = {"none": 0, "gap": 1, "gcsh": 2}
    DOUBLING = {"none": 0, "band_doubling": 1, "linear_search": 2}
    START = {"zero": 0, "gap": 1, "h0": 2}
    _fields_ = [("domain", C.c_int32), ("heuristic", C.c_int32), ("k", C.c_int32), ("r", C.c_int32), ("p", C.c_int32),
                ("doubling", C.c_int32), ("doubling_start", C.c_int32), ("factor", C.c_float), ("delta", C.c_int32),
                ("block_width", C.c_int32), ("sparse", C.c_int32), ("incremental_doubling", C.c_int32), ("dt_trace", C.c_int32),
                ("max_g", C.c_int32), ("fr_drop", C.c_int32), ("sparse_h", C.c_int32), ("prune", C.c_int32)]

    @classmethod
    def _preset(cls, preset):
        q = cls()
        _check(load_library().apa_params_preset(preset, C.byref(q)))
        return q

    @classmethod
    def simple(cls):
        return cls._preset(PRESET_SIMPLE)

    @classmethod
    def full(cls):
        return cls._preset(PRESET_FULL)

    @classmethod
    def nw(cls):
        """AstarPa2Params::nw() (params.rs:46-68): Domain::Full, no doubling, no dt_trace."""
        q = cls._preset(PRESET_SIMPLE)
        q.domain, q.heuristic, q.doubling, q.dt_trace, q.sparse_h, q.prune = 0, 0, 0, 0, 0, 0
        return q

    @classmethod
    def from_json(cls, text):
        """AstarPa2Params from the reference's serde JSON (apa_params_from_json); raises AstarPaError naming an unsupported field."""
        L = load_library()
        L.apa_params_from_json.argtypes = [C.c_char_p, C.POINTER(cls), C.c_char_p, C.c_uint64]
        q = cls()
        err = C.create_string_buffer(512)
        rc = L.apa_params_from_json(text.encode() if isinstance(text, str) else text, C.byref(q), err, 512)
        if rc != 0:
            raise AstarPaError(err.value.decode())
        return q

    def replace(self, **kw):
        """Copy with fields replaced; domain / heuristic / doubling / doubling_start also accept their names."""
        q = type(self).from_buffer_copy(self)
        names = {"domain": self.DOMAIN, "heuristic": self.HEURISTIC, "doubling": self.DOUBLING, "doubling_start": self.START}
        for k, v in kw.items():
            if isinstance(v, str):
                v = names[k][v]
            setattr(q, k, v)
        return q


def lib_path():
    # APA_LIB: an alternative build of the same library (A/B measurements of kernel variants, see csrc/Makefile `ab`)
    return os.environ.get("APA_LIB") or os.path.join(_HERE, "libastarpa_c.so")


def load_library():
    """Load libastarpa_c.so (built in-tree by __graft_entry__.build()). Fails loudly when absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise AstarPaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.apa_last_error.restype = C.c_char_p
    L.apa_device_count.restype = C.c_int
    L.apa_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.apa_engine_destroy.argtypes = [C.c_void_p]
    L.apa_batch_upload.argtypes = [C.c_void_p, C.c_uint64, vp, vp, vp, vp, C.POINTER(C.c_void_p)]
    L.apa_batch_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.apa_batch_download.argtypes = [C.c_void_p, C.c_void_p, vp, C.POINTER(C.c_void_p), vp, vp]
    L.apa_batch_get_stats.argtypes = [C.c_void_p, C.POINTER(BatchStats)]
    L.apa_batch_free.argtypes = [C.c_void_p, C.c_void_p]
    L.apa_align_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(C.c_void_p), vp, vp,
                                  C.POINTER(BatchStats)]
    pp = C.POINTER(AstarPa2Params)
    L.apa_params_preset.argtypes = [C.c_int, pp]
    L.apa_batch_run_params.argtypes = [C.c_void_p, C.c_void_p, pp, C.c_int]
    L.apa_align_batch_params.argtypes = [C.c_void_p, pp, C.c_int, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(C.c_void_p), vp, vp,
                                         C.POINTER(BatchStats)]
    L.apa_debug_band_log_params.restype = C.c_int64
    L.apa_debug_band_log_params.argtypes = [C.c_void_p, pp, C.c_int, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64]
    L.apa_batch_download_pair_stats.argtypes = [C.c_void_p, C.c_void_p, vp]
    L.apa_search.argtypes = [C.c_void_p, vp, C.c_uint64, vp, C.c_uint64, C.c_float, vp]
    L.apa_search_trace.argtypes = [C.c_void_p, vp, C.c_uint64, vp, C.c_uint64, C.c_float, C.c_uint64, C.c_char_p, C.c_uint64, vp]
    L.apa_free.argtypes = [C.c_void_p]
    L.apa_pinned_alloc.restype = C.c_void_p
    L.apa_pinned_alloc.argtypes = [C.c_uint64]
    L.apa_pinned_free.argtypes = [C.c_void_p]
    L.apa_align_batch_multi.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(C.c_void_p), vp, vp, vp]
    L.apa_packed_layout.argtypes = [C.c_uint64, vp, vp]
    L.apa_pack_sequences.argtypes = [C.c_uint64, vp, vp, vp, vp, C.c_int]
    L.apa_align_batch_packed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(C.c_void_p), vp, vp,
                                         C.POINTER(BatchStats)]
    L.apa_int32_peak.argtypes = [C.c_void_p, vp]
    L.apa_pack_planes_device.argtypes = [C.c_void_p, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.apa_block_compute.argtypes = [C.c_void_p, vp, C.c_uint64, vp, C.c_uint64, vp, vp, C.POINTER(C.c_int64)]
    for name in ("astarpa2_simple", "astarpa2_full", "astarpa"):
        f = getattr(L, name)
        f.restype = C.c_uint64
        f.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.astarpa_gcsh.restype = C.c_uint64
    L.astarpa_gcsh.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_bool,
                               C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.astarpa_free_cigar.argtypes = [C.c_void_p]
    _LIB = L
    return L


def bind_to_device_numa(device):
    """apa_bind_host_thread_to_device: keep this thread (and what it spawns / first touches) on the GPU's NUMA node.
    Returns the number of CPUs bound to (0 = unchanged)."""
    L = load_library()
    L.apa_bind_host_thread_to_device.argtypes = [C.c_int]
    return int(L.apa_bind_host_thread_to_device(int(device)))


def _check(rc):
    if rc != 0:
        raise AstarPaError(f"libastarpa_c error {rc}: {load_library().apa_last_error().decode()}")


# ----------------------------------------------------------------------------------------------- generator
_GEN = None


def gen_library():
    """libapa_generate.so: the synthetic-input generator and text digests (host code for tests / bench.py, include/apa_generate.h).
    Deliberately a library of its own: the product library holds the aligner only."""
    global _GEN
    if _GEN is None:
        path = os.path.join(_HERE, "libapa_generate.so")
        if not os.path.exists(path):
            raise AstarPaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        G = C.CDLL(path)
        vp = C.c_void_p
        G.apa_generate_pair.restype = C.c_int64
        G.apa_generate_pair.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_uint64, vp, vp, C.c_uint64]
        G.apa_generate_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_uint64, vp, vp, C.c_uint64, vp, C.c_int]
        G.apa_fnv1a_batch.argtypes = [vp, vp, vp, C.c_uint64, vp]
        G.apa_fnv1a_batch.restype = None
        _GEN = G
    return _GEN


def cigar_digests(pool, off, ln):
    """FNV-1a (64 bit) of every CIGAR text of a downloaded pool (pool: c_void_p; off / ln: int64 arrays) -> uint64 array."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    ln = np.ascontiguousarray(ln, dtype=np.int64)
    out = np.zeros(len(off), dtype=np.uint64)
    if len(off):
        gen_library().apa_fnv1a_batch(pool, off.ctypes.data, ln.ctypes.data, len(off), out.ctypes.data)
    return out


def generate_pair(n, e, model=0, seed=31415):
    """Synthetic pair (stands in for pa_generate::generate_model, pa-test/src/lib.rs:60)."""
    L = gen_library()
    a = np.empty(max(n, 1), dtype=np.uint8)
    cap = 3 * n + 64
    b = np.empty(cap, dtype=np.uint8)
    bl = L.apa_generate_pair(n, float(e), model, seed, a.ctypes.data, b.ctypes.data, cap)
    if bl < 0:
        raise AstarPaError("generator buffer too small")
    return a[:n].tobytes(), b[:bl].tobytes()


def generate_batch(n_pairs, n, e, model=0, seed0=31415, threads=None):
    """Returns (a_all, a_off, b_all, b_off) numpy arrays; pair p uses seed seed0 + p."""
    L = gen_library()
    threads = threads or (os.cpu_count() or 1)
    stride = int(n * (1 + e) + n * e + 64)
    a_all = np.empty(max(n_pairs * n, 1), dtype=np.uint8)
    b_buf = np.empty(max(n_pairs * stride, 1), dtype=np.uint8)
    b_len = np.zeros(max(n_pairs, 1), dtype=np.int64)
    rc = L.apa_generate_batch(n_pairs, n, float(e), model, seed0, a_all.ctypes.data, b_buf.ctypes.data, stride,
                              b_len.ctypes.data, threads)
    if rc != 0:
        raise AstarPaError("generator stride too small")
    a_off = np.arange(n_pairs + 1, dtype=np.int64) * n
    b_off = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(b_len[:n_pairs], out=b_off[1:])
    b_all = np.empty(max(int(b_off[-1]), 1), dtype=np.uint8)
    for p in range(n_pairs):
        b_all[b_off[p]:b_off[p + 1]] = b_buf[p * stride:p * stride + b_len[p]]
    return a_all[:n_pairs * n], a_off, b_all[:int(b_off[-1])], b_off


def pack_sequences(seq_all, seq_off, threads=None, pinned=True):
    """apa_pack_sequences: the 2-bit plane form of a set of sequences (the engine's own layout, see include/astarpa_b200.h).
    Returns (planes uint32 array [2 per half-word], lengths int64 array); planes live in page-locked memory when pinned."""
    L = load_library()
    seq_all = np.ascontiguousarray(seq_all, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    n = len(seq_off) - 1
    lens = np.ascontiguousarray(seq_off[1:] - seq_off[:-1], dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    _check(L.apa_packed_layout(n, lens.ctypes.data, off.ctypes.data))
    words = max(int(off[-1]) * 2, 1)
    planes = pinned_copy(np.zeros(words, dtype=np.uint32)) if pinned else np.zeros(words, dtype=np.uint32)
    _check(L.apa_pack_sequences(n, seq_all.ctypes.data if seq_all.size else None, seq_off.ctypes.data, planes.ctypes.data, off.ctypes.data,
                                threads or (os.cpu_count() or 1)))
    return planes, lens


def pinned_copy(arr):
    """Copy a numpy array into page-locked host memory; returns a numpy view (keep it alive; never freed)."""
    L = load_library()
    arr = np.ascontiguousarray(arr)
    p = L.apa_pinned_alloc(arr.nbytes)
    if not p:
        raise AstarPaError("pinned allocation failed: " + L.apa_last_error().decode())
    buf = (C.c_uint8 * max(arr.nbytes, 1)).from_address(p)
    out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
    out[...] = arr
    return out


# ----------------------------------------------------------------------------------------------- engine
class Engine:
    """One per GPU: owns the stream, the scratch arenas and the device work queue."""

    def __init__(self, device=0):
        L = load_library()
        self._L = L
        self._h = None
        h = C.c_void_p()
        _check(L.apa_engine_create(device, C.byref(h)))
        self._h = h

    @classmethod
    def shared(cls, device=0):
        """The process-wide engine of `device` (apa_shared_engine): the one align_batch_multi and the drop-in symbols use."""
        L = load_library()
        L.apa_shared_engine.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        self = cls.__new__(cls)
        self._L = L
        self._h = None
        self._owned = False
        h = C.c_void_p()
        _check(L.apa_shared_engine(device, C.byref(h)))
        self._h = h
        return self

    def close(self):
        if self._h and getattr(self, "_owned", True):
            self._L.apa_engine_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, a_all, a_off, b_all, b_off):
        return Batch(self, a_all, a_off, b_all, b_off)

    def align_batch_raw(self, a_all, a_off, b_all, b_off, preset=PRESET_FULL, trace=True):
        """apa_align_batch: host buffers in, host buffers out (the bases stream to HBM while the kernel runs).
        Returns (costs, pool pointer or None, cigar_off, cigar_len, stats dict); release the pool with free_pool()."""
        a_all = np.ascontiguousarray(a_all, dtype=np.uint8)
        b_all = np.ascontiguousarray(b_all, dtype=np.uint8)
        a_off = np.ascontiguousarray(a_off, dtype=np.int64)
        b_off = np.ascontiguousarray(b_off, dtype=np.int64)
        n = len(a_off) - 1
        costs = np.zeros(max(n, 1), dtype=np.int64)
        off = np.zeros(max(n, 1), dtype=np.int64)
        ln = np.zeros(max(n, 1), dtype=np.int64)
        pool = C.c_void_p()
        st = BatchStats()
        ap = a_all.ctypes.data if a_all.size else None
        bp = b_all.ctypes.data if b_all.size else None
        if isinstance(preset, AstarPa2Params):  # explicit parameters: the general kernel
            _check(self._L.apa_align_batch_params(self._h, C.byref(preset), int(trace), n, ap, a_off.ctypes.data, bp, b_off.ctypes.data,
                                                  costs.ctypes.data, C.byref(pool), off.ctypes.data, ln.ctypes.data, C.byref(st)))
        else:
            _check(self._L.apa_align_batch(self._h, preset, int(trace), n, ap, a_off.ctypes.data, bp, b_off.ctypes.data,
                                           costs.ctypes.data, C.byref(pool), off.ctypes.data, ln.ctypes.data, C.byref(st)))
        return costs[:n], pool, off[:n], ln[:n], st.as_dict()

    def align_batch_packed(self, a_planes, a_len, b_planes, b_len, preset=PRESET_FULL, trace=True):
        """apa_align_batch_packed: as align_batch_raw, on sequences already packed by pack_sequences (2-bit planes)."""
        n = len(a_len)
        a_len = np.ascontiguousarray(a_len, dtype=np.int64)
        b_len = np.ascontiguousarray(b_len, dtype=np.int64)
        costs = np.zeros(max(n, 1), dtype=np.int64)
        off = np.zeros(max(n, 1), dtype=np.int64)
        ln = np.zeros(max(n, 1), dtype=np.int64)
        pool = C.c_void_p()
        st = BatchStats()
        _check(self._L.apa_align_batch_packed(self._h, preset, int(trace), n, a_planes.ctypes.data, a_len.ctypes.data, b_planes.ctypes.data,
                                              b_len.ctypes.data, costs.ctypes.data, C.byref(pool), off.ctypes.data, ln.ctypes.data, C.byref(st)))
        return costs[:n], pool, off[:n], ln[:n], st.as_dict()

    def free_pool(self, pool):
        if pool and pool.value:
            self._L.apa_free(pool)

    def int32_peak(self):
        """apa_int32_peak: measured integer issue rates of this GPU (lane-operations per second), for the block-DP roofline."""
        out = np.zeros(7, dtype=np.float64)
        _check(self._L.apa_int32_peak(self._h, out.ctypes.data))
        d = dict(zip(["lop3", "shf", "iadd3", "imad", "dp_mix_alu", "clock_hz", "dp_mix_imad"], [float(x) for x in out]))
        # the ALU pipe's ceiling for the block DP: the better of LOP3 alone and the LOP3 + SHF share of the DP mix
        d["alu_lane_ops_per_s"] = max(d["lop3"], d["dp_mix_alu"])
        return d

    def pack_planes(self, seq: bytes):
        """K0 on the device (apa_pack_planes_device): the 2-bit planes of one sequence as a uint32 array, 2 per half-word."""
        nhw = ((((len(seq) + 63) // 64) * 2 + 2 + 15) // 16) * 16
        out = np.zeros(2 * nhw, dtype=np.uint32)
        got = C.c_uint64()
        sb = np.frombuffer(seq, dtype=np.uint8) if seq else np.zeros(1, np.uint8)
        _check(self._L.apa_pack_planes_device(self._h, sb.ctypes.data, len(seq), out.ctypes.data, nhw, C.byref(got)))
        assert got.value == nhw
        return out

    def search(self, pattern: bytes, text: bytes, unmatched_cost: float = 0.0):
        """pa_bitpacking::search(pattern, text, unmatched_cost).out (pa-bitpacking/src/search.rs:46-118) as a numpy int32 array:
        the costs along the bottom row and up the right column of the semi-global DP, |pattern| + |text| + 1 values."""
        out = np.zeros(len(pattern) + len(text) + 1, dtype=np.int32)
        pb = np.frombuffer(pattern, dtype=np.uint8) if pattern else np.zeros(1, np.uint8)
        tb = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
        _check(self._L.apa_search(self._h, pb.ctypes.data, len(pattern), tb.ctypes.data, len(text), float(unmatched_cost), out.ctypes.data))
        return out

    def search_trace(self, pattern: bytes, text: bytes, unmatched_cost: float, idx: int):
        """SearchResult::trace(idx) (pa-bitpacking/src/search.rs:135-230): (cigar text, (start_i, start_j), (end_i, end_j), cost)."""
        cap = 2 * (len(pattern) + len(text)) + 16
        buf = C.create_string_buffer(cap)
        pos = np.zeros(5, dtype=np.int32)
        pb = np.frombuffer(pattern, dtype=np.uint8) if pattern else np.zeros(1, np.uint8)
        tb = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(1, np.uint8)
        _check(self._L.apa_search_trace(self._h, pb.ctypes.data, len(pattern), tb.ctypes.data, len(text), float(unmatched_cost), int(idx),
                                        buf, cap, pos.ctypes.data))
        return buf.value.decode(), (int(pos[0]), int(pos[1])), (int(pos[2]), int(pos[3])), int(pos[4])

    def block_compute(self, a: bytes, b: bytes, v=None, h=None):
        """pa_bitpacking::simd::compute on the GPU. h: top-edge deltas, one byte per column (0, 1 = +1, 2 = -1; default all +1);
        v: left-edge (p, m) u64 pairs per 64-row word (default all +1). Returns (bottom_sum, h_out, v_out)."""
        na, mb = len(a), len(b)
        nwords = (mb + 63) // 64
        h = np.ones(max(na, 1), dtype=np.uint8) if h is None else np.ascontiguousarray(h, dtype=np.uint8).copy()
        vv = np.zeros(2 * max(nwords, 1), dtype=np.uint64)
        if v is None:
            vv[0::2] = np.uint64(0xFFFFFFFFFFFFFFFF)
        else:
            vv[:2 * nwords] = v
        s = C.c_int64()
        ab = np.frombuffer(a, dtype=np.uint8) if na else np.zeros(1, np.uint8)
        bb = np.frombuffer(b, dtype=np.uint8) if mb else np.zeros(1, np.uint8)
        _check(self._L.apa_block_compute(self._h, ab.ctypes.data, na, bb.ctypes.data, mb, h.ctypes.data, vv.ctypes.data,
                                         C.byref(s)))
        return s.value, h[:na], vv[:2 * nwords]


class Batch:
    """A batch of pairs resident in HBM (apa_batch)."""

    def __init__(self, eng, a_all, a_off, b_all, b_off):
        self._eng = eng
        self._L = eng._L
        self._h = None
        a_all = np.ascontiguousarray(a_all, dtype=np.uint8)
        b_all = np.ascontiguousarray(b_all, dtype=np.uint8)
        a_off = np.ascontiguousarray(a_off, dtype=np.int64)
        b_off = np.ascontiguousarray(b_off, dtype=np.int64)
        self.n_pairs = len(a_off) - 1
        h = C.c_void_p()
        ap = a_all.ctypes.data if a_all.size else None
        bp = b_all.ctypes.data if b_all.size else None
        _check(self._L.apa_batch_upload(eng._h, self.n_pairs, ap, a_off.ctypes.data, bp, b_off.ctypes.data, C.byref(h)))
        self._h = h

    def run(self, preset=PRESET_FULL, trace=True):
        if isinstance(preset, AstarPa2Params):
            _check(self._L.apa_batch_run_params(self._eng._h, self._h, C.byref(preset), int(trace)))
        else:
            _check(self._L.apa_batch_run(self._eng._h, self._h, preset, int(trace)))
        return self

    def download(self, cigars=True):
        n = self.n_pairs
        costs = np.zeros(max(n, 1), dtype=np.int64)
        if not cigars:
            _check(self._L.apa_batch_download(self._eng._h, self._h, costs.ctypes.data, None, None, None))
            return costs[:n], None
        off = np.zeros(max(n, 1), dtype=np.int64)
        ln = np.zeros(max(n, 1), dtype=np.int64)
        pool = C.c_void_p()
        _check(self._L.apa_batch_download(self._eng._h, self._h, costs.ctypes.data, C.byref(pool), off.ctypes.data,
                                          ln.ctypes.data))
        out = None
        if pool.value:
            out = [C.string_at(pool.value + int(off[p]), int(ln[p])).decode() for p in range(n)]
            self._L.apa_free(pool)
        return costs[:n], out

    def download_raw(self):
        """As download(), without building Python strings: (costs, pool pointer, off, len); free pool with free_pool."""
        n = self.n_pairs
        costs = np.zeros(max(n, 1), dtype=np.int64)
        off = np.zeros(max(n, 1), dtype=np.int64)
        ln = np.zeros(max(n, 1), dtype=np.int64)
        pool = C.c_void_p()
        _check(self._L.apa_batch_download(self._eng._h, self._h, costs.ctypes.data, C.byref(pool), off.ctypes.data,
                                          ln.ctypes.data))
        return costs[:n], pool, off[:n], ln[:n]

    def free_pool(self, pool):
        if pool and pool.value:
            self._L.apa_free(pool)

    def pair_stats(self):
        """Per-pair counters of the last run: list of dicts (apa_pair_stats)."""
        arr = (PairStats * max(self.n_pairs, 1))()
        _check(self._L.apa_batch_download_pair_stats(self._eng._h, self._h, C.cast(arr, C.c_void_p)))
        keys = [k for k, _ in PairStats._fields_ if k != "reserved"]
        return [{k: getattr(arr[p], k) for k in keys} for p in range(self.n_pairs)]

    def stats(self):
        st = BatchStats()
        _check(self._L.apa_batch_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def free(self):
        if self._h:
            self._L.apa_batch_free(self._eng._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _concat(pairs):
    a_off = np.zeros(len(pairs) + 1, dtype=np.int64)
    b_off = np.zeros(len(pairs) + 1, dtype=np.int64)
    for p, (a, b) in enumerate(pairs):
        a_off[p + 1] = a_off[p] + len(a)
        b_off[p + 1] = b_off[p] + len(b)
    a_all = np.frombuffer(b"".join(a for a, _ in pairs), dtype=np.uint8)
    b_all = np.frombuffer(b"".join(b for _, b in pairs), dtype=np.uint8)
    return a_all, a_off, b_all, b_off


def align_batch_multi(devices, a_all, a_off, b_all, b_off, preset=PRESET_FULL, trace=True):
    """apa_align_batch_multi: one call, the GPUs of `devices` (contiguous shards balanced by bases, no collective).
    Returns (costs, pool pointer or None, cigar_off, cigar_len, list of per-device stats dicts); free the pool with apa_free
    (Engine.free_pool of any engine, or free_pool below)."""
    L = load_library()
    a_all = np.ascontiguousarray(a_all, dtype=np.uint8)
    b_all = np.ascontiguousarray(b_all, dtype=np.uint8)
    a_off = np.ascontiguousarray(a_off, dtype=np.int64)
    b_off = np.ascontiguousarray(b_off, dtype=np.int64)
    n = len(a_off) - 1
    devs = (C.c_int * len(devices))(*devices)
    costs = np.zeros(max(n, 1), dtype=np.int64)
    off = np.zeros(max(n, 1), dtype=np.int64)
    ln = np.zeros(max(n, 1), dtype=np.int64)
    pool = C.c_void_p()
    st = (BatchStats * len(devices))()
    _check(L.apa_align_batch_multi(C.cast(devs, C.c_void_p), len(devices), preset, int(trace), n, a_all.ctypes.data if a_all.size else None,
                                   a_off.ctypes.data, b_all.ctypes.data if b_all.size else None, b_off.ctypes.data, costs.ctypes.data,
                                   C.byref(pool), off.ctypes.data, ln.ctypes.data, C.cast(st, C.c_void_p)))
    return costs[:n], pool, off[:n], ln[:n], [x.as_dict() for x in st]


def free_pool(pool):
    if pool and pool.value:
        load_library().apa_free(pool)


_DEFAULT_ENGINES = {}


def _engine(device=0):
    if device not in _DEFAULT_ENGINES:
        _DEFAULT_ENGINES[device] = Engine(device)
    return _DEFAULT_ENGINES[device]


class AstarPa2:
    """Mirror of ``AstarPa2Params::{simple,full}().make_aligner(trace)`` (astarpa2/src/params.rs:70-132), or of
    ``params.make_aligner(trace)`` for an explicit ``AstarPa2Params`` (served by the general kernel)."""

    def __init__(self, preset="full", trace=True, device=0):
        self.preset = preset if isinstance(preset, AstarPa2Params) else {"simple": PRESET_SIMPLE, "full": PRESET_FULL, 0: 0, 1: 1}[preset]
        self.trace = trace
        self.device = device

    def align_batch(self, pairs):
        """pairs: list of (a, b) byte strings over ACGT. Returns (costs ndarray, list of CIGAR strings or None)."""
        eng = _engine(self.device)
        costs, pool, off, ln, _ = eng.align_batch_raw(*_concat(pairs), self.preset, self.trace)
        cigars = None
        if self.trace and pool.value:
            cigars = [C.string_at(pool.value + int(off[p]), int(ln[p])).decode() for p in range(len(pairs))]
        eng.free_pool(pool)
        return costs, cigars

    def align(self, a: bytes, b: bytes):
        """Aligner::align (astarpa2/src/lib.rs:210-215): (cost, cigar or None)."""
        costs, cigs = self.align_batch([(a, b)])
        return int(costs[0]), (cigs[0] if cigs is not None else None)

    def align_with_stats(self, a: bytes, b: bytes):
        """AstarPa2StatsAligner::align_with_stats (astarpa2/src/lib.rs:200-208): ((cost, cigar or None), stats dict)."""
        out = self.align_batch_with_stats([(a, b)])
        return (int(out[0][0]), out[1][0] if out[1] is not None else None), out[2][0]

    def align_batch_with_stats(self, pairs):
        """(costs, cigars or None, per-pair stats dicts) through an HBM-resident batch."""
        batch = _engine(self.device).upload(*_concat(pairs))
        try:
            batch.run(self.preset, self.trace)
            costs, cigars = batch.download(cigars=self.trace)
            return costs, cigars, batch.pair_stats()
        finally:
            batch.free()

    def cost(self, a: bytes, b: bytes):
        """AstarPa2::cost (astarpa2/src/lib.rs:177-179)."""
        saved, self.trace = self.trace, False
        try:
            return self.align(a, b)[0]
        finally:
            self.trace = saved


def search(pattern: bytes, text: bytes, unmatched_cost: float = 0.0, device=0):
    """pa_bitpacking::search (pa-bitpacking/src/search.rs:46): see Engine.search."""
    return _engine(device).search(pattern, text, unmatched_cost)


def search_trace(pattern: bytes, text: bytes, unmatched_cost: float, idx: int, device=0):
    """SearchResult::trace (pa-bitpacking/src/search.rs:135-230): see Engine.search_trace."""
    return _engine(device).search_trace(pattern, text, unmatched_cost, idx)


def astarpa2_simple(a: bytes, b: bytes):
    """astarpa2::astarpa2_simple (astarpa2/src/lib.rs:44-47)."""
    return AstarPa2("simple", True).align(a, b)


def astarpa2_full(a: bytes, b: bytes):
    """astarpa2::astarpa2_full (astarpa2/src/lib.rs:50-53)."""
    return AstarPa2("full", True).align(a, b)
