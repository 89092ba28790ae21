"""GPU parity tests at the shapes the headline numbers are measured on (run with -m gpu on a B200).

The bench line of BASELINE configs[2] comes from `apa_phase_{build,pass,trace}_kernel` with ONE warp per pair (batches of more
than 1 184 pairs), fed by a multi-chunk streaming upload; configs[3] from the phase-split path in waves. The grid tests of
test_gpu_parity.py are small batches (the cooperative pass kernel, one upload chunk), so these tests run the big shapes
against the oracle: every cost and every CIGAR (by FNV-1a digest of the text) of the batch, on every upload path, with the
path taken asserted from apa_batch_stats. Reference for the test shape: pa-test/src/lib.rs:65-99 (cost == ground truth and a
CIGAR that verifies, for every generated pair)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_batch(oracle, a_all, a_off, b_all, b_off, preset, trace=True):
    _, costs, _, _, digests = oracle.align_batch(a_all, a_off, b_all, b_off, preset, trace)
    assert (costs >= 0).all(), "oracle panic"
    return costs, digests


def _check_against(apa, costs, pool, off, ln, want_costs, want_digests, what):
    assert (costs == want_costs).all(), (what, np.flatnonzero(costs != want_costs)[:5])
    got = apa.cigar_digests(pool, off, ln)
    assert (got == want_digests).all(), (what, np.flatnonzero(got != want_digests)[:5])


@pytest.mark.parametrize("preset", [1, 0])
def test_headline_shape_one_warp_per_pair_gpu(apa, oracle, engine, preset, monkeypatch):
    # 2 000 pairs of n = 100 000 at e = 5 % (400 MB of bases): more pairs than the cooperative kernel takes (1 184), more
    # bases than one upload chunk. All costs and CIGARs against the oracle, through (1) the resident path bench.py times as
    # `value`, (2) apa_align_batch from pageable memory (host-packed planes streamed under the running kernel), (3)
    # apa_align_batch from page-locked memory: raw bases only (device-side K0 for every pair), and the default - raw chunks by
    # DMA, host-packed chunks by the host threads - which is the `e2e` path of bench.py.
    n_pairs = 2000
    a_all, a_off, b_all, b_off = apa.generate_batch(n_pairs, 100000, 0.05, 0, 31415)
    want_costs, want_digests = _oracle_batch(oracle, a_all, a_off, b_all, b_off, preset)
    # independent ground truth for a sample (full-matrix Levenshtein)
    for p in range(0, n_pairs, 250):
        a, b = a_all[a_off[p]:a_off[p + 1]].tobytes(), b_all[b_off[p]:b_off[p + 1]].tobytes()
        assert oracle.levenshtein(a, b) == want_costs[p]

    batch = engine.upload(a_all, a_off, b_all, b_off)
    batch.run(preset, True)
    st = batch.stats()
    assert st["pass_warps_per_pair"] == 1 and st["kernel_launches"] == 4 and st["retries"] == 0 and st["waves"] == 1, st  # build, first pass, continuation, trace
    costs, pool, off, ln = batch.download_raw()
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "resident")
    # a CIGAR of the batch replayed over its pair
    p = 1234
    text = C.string_at(pool.value + int(off[p]), int(ln[p])).decode()
    assert oracle.cigar_verify(text, a_all[a_off[p]:a_off[p + 1]].tobytes(), b_all[b_off[p]:b_off[p + 1]].tobytes()) == want_costs[p]
    batch.free_pool(pool)
    batch.free()

    costs, pool, off, ln, st = engine.align_batch_raw(a_all, a_off, b_all, b_off, preset, True)
    assert st["upload_mode"] == 2 and st["upload_chunks"] >= 8 and st["pass_warps_per_pair"] == 1, st
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "pageable, streamed")
    engine.free_pool(pool)

    a_pin, b_pin = apa.pinned_copy(a_all), apa.pinned_copy(b_all)
    monkeypatch.setenv("APA_RAW", "1")  # (an engine with >= 12 host threads to itself would host-pack a batch this large)
    costs, pool, off, ln, st = engine.align_batch_raw(a_pin, a_off, b_pin, b_off, preset, True)
    assert st["upload_mode"] == 4 and st["upload_chunks"] >= 8 and st["pass_warps_per_pair"] == 1, st
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "pinned, raw streamed")
    engine.free_pool(pool)
    # the default for page-locked inputs: streamed by both producers (raw chunks by DMA, host-packed chunks)
    monkeypatch.delenv("APA_RAW")
    costs, pool, off, ln, st = engine.align_batch_raw(a_pin, a_off, b_pin, b_off, preset, True)
    assert st["upload_mode"] == 5 and st["upload_chunks"] >= 8 and 1 <= st["upload_chunks_raw"] <= st["upload_chunks"], st
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "pinned, both producers")
    engine.free_pool(pool)
    # cost-only run of the same batch
    c2, pool2, _, _, _ = engine.align_batch_raw(a_pin, a_off, b_pin, b_off, preset, False)
    assert (c2 == want_costs).all() and not pool2.value
    L = apa.load_library()
    L.apa_pinned_free(a_pin.ctypes.data)
    L.apa_pinned_free(b_pin.ctypes.data)


def test_config3_waves_gpu(apa, oracle, engine, monkeypatch):
    # BASELINE configs[3] (n = 1 000 000, e = 15 %, astarpa2_full, cost + CIGAR) on 50 pairs through the wave path the full
    # config takes (per-pair arenas of the whole batch do not fit in HBM: here the budget is capped so that the 50 pairs need
    # three waves; 1 000 pairs need waves on their own), eight warps per pair.
    n_pairs = 50
    a_all, a_off, b_all, b_off = apa.generate_batch(n_pairs, 1000000, 0.15, 0, 31415)
    want_costs, want_digests = _oracle_batch(oracle, a_all, a_off, b_all, b_off, 1)
    batch = engine.upload(a_all, a_off, b_all, b_off)
    batch.run(1, True)
    st0 = batch.stats()
    costs, pool, off, ln = batch.download_raw()
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "all at once")
    batch.free_pool(pool)
    batch.free()
    monkeypatch.setenv("APA_BUDGET_BYTES", str(5 << 30))
    monkeypatch.setenv("APA_RAW", "1")  # (an engine with >= 12 host threads to itself would host-pack a batch this large)
    a_pin, b_pin = apa.pinned_copy(a_all), apa.pinned_copy(b_all)
    costs, pool, off, ln, st = engine.align_batch_raw(a_pin, a_off, b_pin, b_off, 1, True)
    assert st["waves"] >= 2 and st["pass_warps_per_pair"] == 8 and st["upload_mode"] == 3, (st, st0)
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "waves")
    p = 7
    text = C.string_at(pool.value + int(off[p]), int(ln[p])).decode()
    assert oracle.cigar_verify(text, a_all[a_off[p]:a_off[p + 1]].tobytes(), b_all[b_off[p]:b_off[p + 1]].tobytes()) == want_costs[p]
    engine.free_pool(pool)
    L = apa.load_library()
    L.apa_pinned_free(a_pin.ctypes.data)
    L.apa_pinned_free(b_pin.ctypes.data)


def test_device_pack_matches_host_packer_gpu(apa, engine):
    # K0 (BitProfile::build, pa-bitpacking/src/profile.rs:112-133) on the device against the host packer, ragged lengths,
    # unaligned raw bases; a byte outside ACGT is refused like the reference's panic (profile.rs:113).
    L = apa.load_library()
    L.apa_pack_planes_host.restype = C.c_int
    L.apa_pack_planes_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    rng = np.random.default_rng(9)
    for n in [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 1000, 1023, 1024, 1025, 4097, 32768, 100003, 1000001]:
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).astype(np.uint8)
        got = engine.pack_planes(seq.tobytes())
        nhw = len(got) // 2
        want = np.zeros(2 * nhw, dtype=np.uint32)
        buf = np.ascontiguousarray(np.concatenate([seq, np.zeros(64, np.uint8)]))
        assert L.apa_pack_planes_host(buf.ctypes.data, n, 0, nhw, want.ctypes.data) == 0
        assert (got == want).all(), n
    for bad in (b"ACGN", b"A" * 1000 + b"a" + b"C" * 50, b"ACG\x00"):
        with pytest.raises(apa.AstarPaError):
            engine.pack_planes(bad)


@pytest.mark.parametrize("preset", [0, 1])
def test_raw_upload_paths_equal_host_packed_gpu(apa, oracle, engine, preset, monkeypatch):
    # Same batch through every upload path: resident from pinned memory (raw DMA + apa_pack_kernel), resident from pageable
    # memory (host packer), apa_align_batch from pinned memory (raw, K0 inside the kernel that opens the pair: phase-split,
    # cooperative, fused and general kernels), apa_align_batch from pageable memory. Ragged lengths incl. empty sequences.
    rng = np.random.default_rng(31)
    pairs = [apa.generate_pair(int(rng.integers(0, 5000)), float(rng.choice([0.0, 0.05, 0.2])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(150)]
    pairs += [(b"", b""), (b"ACGT", b""), (b"", b"TTGCA")]
    a_all, a_off, b_all, b_off = apa._concat(pairs)
    want = [oracle.align(a, b, preset, True)[:2] for a, b in pairs]
    a_pin, b_pin = apa.pinned_copy(a_all), apa.pinned_copy(b_all)

    def check(costs, pool, off, ln, what):
        for k, (oc, ocg) in enumerate(want):
            assert int(costs[k]) == oc, (what, k)
            assert C.string_at(pool.value + int(off[k]), int(ln[k])).decode() == ocg, (what, k)

    for what, aa, bb in (("resident pinned", a_pin, b_pin), ("resident pageable", a_all, b_all)):
        batch = engine.upload(aa, a_off, bb, b_off)
        batch.run(preset, True)
        costs, pool, off, ln = batch.download_raw()
        check(costs, pool, off, ln, what)
        batch.free_pool(pool)
        batch.free()
    for env in ({}, {"APA_SPLIT": "0"}, {"APA_COOP": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for what, aa, bb, mode in (("e2e pinned", a_pin, b_pin, 3), ("e2e pageable", a_all, b_all, 1)):
            costs, pool, off, ln, st = engine.align_batch_raw(aa, a_off, bb, b_off, preset, True)
            assert st["upload_mode"] == mode, (what, st)
            check(costs, pool, off, ln, (what, env))
            engine.free_pool(pool)
        for k in env:
            monkeypatch.delenv(k)
    # bad input through the in-kernel K0
    bad_a = apa.pinned_copy(np.frombuffer(b"ACGTNACGT", dtype=np.uint8))
    bad_b = apa.pinned_copy(np.frombuffer(b"ACGTACGT", dtype=np.uint8))
    with pytest.raises(apa.AstarPaError):
        engine.align_batch_raw(bad_a, np.array([0, 9]), bad_b, np.array([0, 8]), preset, True)


def test_align_batch_multi_gpu(apa, oracle):
    # apa_align_batch_multi: contiguous shards over the devices of the box (all of them, and device 0 alone), results in
    # input order with one CIGAR pool; equal to the single-engine call pair by pair.
    L = apa.load_library()
    ndev = L.apa_device_count()
    rng = np.random.default_rng(13)
    pairs = [apa.generate_pair(int(rng.integers(0, 30000)), float(rng.choice([0.02, 0.05, 0.15])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(400)]
    args = apa._concat(pairs)
    for preset in (0, 1):
        ref_costs, ref_cigars = apa.AstarPa2(preset, True).align_batch(pairs)
        for devices in ([0], list(range(ndev))):
            costs, pool, off, ln, stats = apa.align_batch_multi(devices, *args, preset, True)
            assert (costs == ref_costs).all() and len(stats) == len(devices)
            for k in range(len(pairs)):
                assert C.string_at(pool.value + int(off[k]), int(ln[k])).decode() == ref_cigars[k], (preset, devices, k)
            apa.free_pool(pool)
            c2, pool2, _, _, _ = apa.align_batch_multi(devices, *args, preset, False)
            assert (c2 == ref_costs).all() and not pool2.value
        for k in range(0, len(pairs), 57):
            assert int(ref_costs[k]) == oracle.align(pairs[k][0], pairs[k][1], preset, False)[0]
    # empty batch, more devices than pairs
    costs, pool, off, ln, _ = apa.align_batch_multi(list(range(ndev)), *apa._concat(pairs[:1]), 1, True)
    assert int(costs[0]) == int(ref_costs[0])
    apa.free_pool(pool)
    costs, pool, off, ln, _ = apa.align_batch_multi([0], *apa._concat([]), 1, True)
    assert len(costs) == 0
    with pytest.raises(apa.AstarPaError):
        apa.align_batch_multi([0, 0], *args, 1, True)


SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch.distributed as dist
import astar_pairwise_aligner_b200 as A
from astar_pairwise_aligner_b200.sharding import align_batch_sharded, gpu_align_fn
import oracle_lib as O

dist.init_process_group("gloo")
rank = dist.get_rank()
ndev = A.load_library().apa_device_count()
rng = np.random.default_rng(17)
pairs = [A.generate_pair(int(rng.integers(0, 20000)), 0.06, k % 4, 500 + k) for k in range(120)]
args = A._concat(pairs)
for preset in (0, 1):
    res = align_batch_sharded(*args, preset, True, gpu_align_fn(rank % ndev), dist)
    if rank == 0:
        for k, (a, b) in enumerate(pairs):
            oc, ocg, _ = O.align(a, b, preset, True)
            assert int(res[0][k]) == oc and res[1][k] == ocg, (preset, k)
if rank == 0:
    print("GPU_SHARD_OK", len(pairs))
dist.destroy_process_group()
'''


def test_sharded_two_ranks_gpu(apa, tmp_path):
    # astar_pairwise_aligner_b200/sharding.py on the GPU: two ranks (gloo rendezvous; rank r drives device r mod #devices
    # through apa_align_batch_multi) align their shards, rank 0 gathers and checks every pair against the oracle.
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER.format(root=root))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, er) in zip(procs, outs):
        assert p.returncode == 0, er[-3000:]
    assert "GPU_SHARD_OK 120" in outs[0][0]


def test_streamed_upload_with_arena_retries_gpu(apa, oracle, engine, monkeypatch):
    # A streamed batch (several upload chunks, both producers) whose pairs overflow a deliberately small arena: the retried pairs
    # find their planes in HBM (packed on the host or by the kernel that first opened them) and must not be packed again from
    # raw bases that never travelled.
    n_pairs = 700
    a_all, a_off, b_all, b_off = apa.generate_batch(n_pairs, 100000, 0.05, 0, 99)
    want_costs, want_digests = _oracle_batch(oracle, a_all, a_off, b_all, b_off, 1)
    a_pin, b_pin = apa.pinned_copy(a_all), apa.pinned_copy(b_all)
    monkeypatch.setenv("APA_ARENA_BYTES", "500000")
    costs, pool, off, ln, st = engine.align_batch_raw(a_pin, a_off, b_pin, b_off, 1, True)
    assert st["upload_mode"] == 5 and st["retries"] > 0, st
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "streamed + retries")
    engine.free_pool(pool)
    L = apa.load_library()
    L.apa_pinned_free(a_pin.ctypes.data)
    L.apa_pinned_free(b_pin.ctypes.data)


@pytest.mark.parametrize("preset", [0, 1])
def test_packed_input_gpu(apa, oracle, engine, preset):
    # apa_align_batch_packed: sequences packed to 2-bit planes by the caller (apa_pack_sequences, the engine's own layout), small
    # batch (plain copy) and a streamed one; equal to the ASCII entry pair by pair, and to the oracle.
    rng = np.random.default_rng(41)
    pairs = [apa.generate_pair(int(rng.integers(0, 4000)), float(rng.choice([0.0, 0.05, 0.2])), int(rng.integers(0, 4)),
                               int(rng.integers(1 << 40))) for _ in range(120)] + [(b"", b""), (b"ACGT", b""), (b"", b"TTGCA")]
    a_all, a_off, b_all, b_off = apa._concat(pairs)
    ap, al = apa.pack_sequences(a_all, a_off)
    bp, bl = apa.pack_sequences(b_all, b_off)
    costs, pool, off, ln, st = engine.align_batch_packed(ap, al, bp, bl, preset, True)
    assert st["upload_mode"] == 6, st
    for k, (a, b) in enumerate(pairs):
        oc, ocg, _ = oracle.align(a, b, preset, True)
        assert int(costs[k]) == oc and C.string_at(pool.value + int(off[k]), int(ln[k])).decode() == ocg, k
    engine.free_pool(pool)
    with pytest.raises(apa.AstarPaError):
        apa.pack_sequences(np.frombuffer(b"ACGTNACG", dtype=np.uint8), np.array([0, 8]))
    # streamed: 700 pairs of n = 100 000 (140 MB of bases = 4 chunks)
    a_all, a_off, b_all, b_off = apa.generate_batch(700, 100000, 0.05, 0, 777)
    want_costs, want_digests = _oracle_batch(oracle, a_all, a_off, b_all, b_off, preset)
    ap, al = apa.pack_sequences(a_all, a_off)
    bp, bl = apa.pack_sequences(b_all, b_off)
    costs, pool, off, ln, st = engine.align_batch_packed(ap, al, bp, bl, preset, True)
    assert st["upload_mode"] == 7 and st["upload_chunks"] >= 3, st
    _check_against(apa, costs, pool, off, ln, want_costs, want_digests, "packed input, streamed")
    engine.free_pool(pool)
