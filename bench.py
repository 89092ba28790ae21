#!/usr/bin/env python
"""bench.py — the A*PA2 hot path on B200: GCUPS (effective DP cells/s) on synthetic n=100k, e=5% pairs.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path over one batch of synthetic pairs (BASELINE.json configs[2]:
n = 100 000, e = 5 %, with CIGAR traceback). --scaling weak (default): --pairs pairs per GPU; --scaling strong: --pairs
pairs in total, cut into contiguous shards balanced by bases (astar_pairwise_aligner_b200/sharding.py), one per rank.
Every run checks the GPU results against the oracle on the pairs of the CPU-baseline sample (cost and CIGAR text digest)
and the e2e results against the resident ones on every pair; a mismatch fails the run (`parity_checked_pairs`).
`value`   : effective GCUPS = sum |a||b| / device time, inputs already resident in HBM (CUDA events on the
            engine's stream around the kernels of each step).
`e2e`     : the same metric through the public C-ABI batch call with HOST (pinned) buffers: H2D of the sequences,
            kernels, D2H of costs + CIGAR text, every step.
`roofline`: dominant kernel = the longest of the three phase kernels (build / pass / trace), each timed with CUDA
            events on the engine's stream. HBM: algorithmic bytes of that kernel (block DP: 48 B per 64-row x
            256-col lane-block) / its launch duration vs the measured copy peak; the path is integer and
            latency bound, so the INT32-pipe fraction of the block DP is reported next to it.
Inputs (2 GB at the default size) are larger than the 126 MB L2, so no explicit flush between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GCUPS (effective DP cells/s, sum |a||b| / time) on n=100k e=5% pairs"
UNIT = "GCUPS"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default=os.environ.get("APA_BENCH_PRESET", "full"), choices=["simple", "full"])
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("APA_BENCH_PAIRS", "10000")), help="pairs per GPU")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--e", type=float, default=0.05)
    ap.add_argument("--no-trace", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end iterations (0 = --steps)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --pairs per GPU; strong: --pairs in total, sharded over the ranks")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


KERNELS = ["apa_phase_build_kernel", "apa_phase_pass_kernel", "apa_phase_trace_kernel"]
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch from the `ncu --set full` captures of the headline
# shape (astarpa2_full, n=100k, e=5 %, cost+CIGAR), in MB PER PAIR; scaled by the pairs of a launch. Source files are
# named next to each figure. None for shapes that were not captured.
NCU_TRAFFIC_MB_PER_PAIR = {  # profiles/r2c_launches.csv (10 000 pairs per launch; round 1: 5.447 / 0.784 / 0.271)
    "apa_phase_build_kernel": 4.831, "apa_phase_pass_kernel": 0.863, "apa_phase_trace_kernel": 0.279}


def traffic_per_launch(args, kernel):
    """Measured DRAM traffic of `kernel` per launch in GB (from the committed ncu capture), or None."""
    if args.preset == "full" and args.n == 100000 and abs(args.e - 0.05) < 1e-9 and not args.no_trace:
        mb = NCU_TRAFFIC_MB_PER_PAIR.get(kernel)
        return None if mb is None else mb * args.pairs / 1e3
    return None


def make_batch(A, args, rank):
    seed0 = 31415 + rank * args.pairs  # 31415: the reference's fixed seed (pa-test/src/lib.rs:51)
    return A.generate_batch(args.pairs, args.n, args.e, 0, seed0)


def cpu_quota():
    """CPUs this container may actually use (cgroup v2 cpu.max), or None when unlimited/unknown."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if q == "max" else float(q) / float(per)
    except Exception:
        return None


def cpu_leg(args, a_all, a_off, b_all, b_off, preset_id, trace, sample):
    """Oracle (CPU port of the reference path) on a bounded sample, all host threads. Returns dict."""
    import oracle_lib as O
    threads = O.lib().oracle_hardware_threads()
    quota = cpu_quota()
    if quota:  # more threads than the cgroup CPU quota only adds contention
        threads = max(1, min(threads, int(round(quota))))
    if sample <= 0:
        # calibrate on `threads` pairs, then size the sample for ~4 s of wall time
        k = min(threads, len(a_off) - 1)
        sec, *_ = O.align_batch(a_all[:a_off[k]], a_off[:k + 1], b_all[:b_off[k]], b_off[:k + 1], preset_id, trace, threads)
        per_pair = max(sec, 1e-4) / 1.0  # k pairs ran concurrently on k threads: wall ~ one pair
        sample = int(max(threads, min(len(a_off) - 1, 4.0 / per_pair * threads)))
    sample = min(sample, len(a_off) - 1)
    sec, costs, clens, cells, chash = O.align_batch(a_all[:a_off[sample]], a_off[:sample + 1], b_all[:b_off[sample]],
                                                    b_off[:sample + 1], preset_id, trace, threads)
    assert (costs >= 0).all(), "oracle panic in the CPU baseline"
    eff = float(np.sum((a_off[1:sample + 1] - a_off[:sample]).astype(np.float64) * (b_off[1:sample + 1] - b_off[:sample])))
    cpu_leg.sample_results = (costs.copy(), chash.copy())  # for the in-run parity check of the GPU results
    return {"value": eff / sec / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample} pairs of the same workload (n={args.n}, e={args.e}, preset {args.preset}, "
                      f"{'with' if trace else 'no'} CIGAR), {sec:.2f} s wall on {threads} threads",
            "computed_gcups": float(cells.sum()) / sec / 1e9, "bp_per_s": float(a_off[sample]) / sec, "seconds": sec,
            "pairs": sample, "pairs_per_s": sample / sec, "hardware_threads": O.lib().oracle_hardware_threads(),
            "cgroup_cpu_quota": quota}


def cpu_micro():
    """The reference's own micro-benchmark (pa-bitpacking/benches/nw/main.rs:139-159): the block kernel
    on a 256-column x h-row rectangle with +1 input deltas, one thread, restated reference layout (oracle port). GCUPS."""
    import oracle_lib as O
    L = O.lib()
    import astar_pairwise_aligner_b200 as A
    out = {}
    for h in (64, 128, 256, 512):
        a, _ = A.generate_pair(256, 0.0, 0, 31415)
        b, _ = A.generate_pair(h, 0.0, 0, 31416)
        hb = np.ones(256, dtype=np.uint8)
        v = np.zeros(2 * (h // 64), dtype=np.uint64)
        reps = 20000
        t0 = time.perf_counter()
        L.oracle_bp_compute_bench(a, len(a), b, len(b), reps)
        dt = time.perf_counter() - t0
        out[f"256x{h}"] = 256.0 * h * reps / dt / 1e9
    return {"gcups_one_thread": out, "what": "oracle_bp_compute (restated simd::compute, 4 x u64 lanes x 2) on 256 columns x h rows, +1 deltas",
            "source": "pa-bitpacking/benches/nw/main.rs:139-159"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    trace = not args.no_trace
    preset_id = {"simple": 0, "full": 1}[args.preset]
    # which BASELINE.json config this shape is (the default run is configs[2], the one the metric is quoted on)
    shape = (args.n, round(args.e, 3), trace, args.preset)
    cfg = {(100000, 0.05, True, "full"): "BASELINE configs[2]", (10000, 0.05, False, "full"): "BASELINE configs[1]",
           (1000000, 0.15, True, "full"): "BASELINE configs[3]", (10000000, 0.05, True, "full"): "BASELINE configs[4]"}.get(shape, "other shape")
    per = "pairs/GPU" if args.scaling == "weak" else f"pairs in total over {world} GPU(s)"
    workload = (f"{cfg}: {args.pairs} {per}, n={args.n}, e={args.e:g}, uniform errors, astarpa2_{args.preset}, "
                f"{'cost+CIGAR' if trace else 'cost only'}; inputs {2 * args.pairs * args.n / 1e9:.2f} GB/GPU > L2 (no flush needed)")
    config = {"workload": workload, "pairs_per_gpu": args.pairs, "n": args.n, "e": args.e, "preset": args.preset, "trace": trace,
              "sharding": f"independent pairs, {world} rank(s), no data-path collective", "l2": "inputs larger than L2"}

    import astar_pairwise_aligner_b200 as A

    if args.impl == "reference":
        # The reference's own CPU implementation of the path, restated (oracle port): the Rust crate cannot be
        # built in this image (no cargo/rustc). Rank 0 only.
        if rank != 0:
            return
        a_all, a_off, b_all, b_off = make_batch(A, args, 0)
        first = cpu_leg(args, a_all, a_off, b_all, b_off, preset_id, trace, args.cpu_sample)
        sample = first["pairs"]
        secs = []
        for it in range(args.warmup + args.steps):
            r = cpu_leg(args, a_all, a_off, b_all, b_off, preset_id, trace, sample)
            if it >= args.warmup:
                secs.append(r["seconds"])
            last = r
        eff = float(np.sum((a_off[1:sample + 1] - a_off[:sample]).astype(np.float64) * (b_off[1:sample + 1] - b_off[:sample])))
        val = eff * len(secs) / sum(secs) / 1e9
        last["value"] = val
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
                          "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                          "cpu_baseline": last, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    if args.scaling == "strong":
        # one batch of --pairs pairs in total (the same on every rank: seeded), this rank aligns its contiguous shard
        from astar_pairwise_aligner_b200.sharding import shard_bounds, slice_batch
        full = make_batch(A, args, 0)
        s0, s1 = shard_bounds(full[1], full[3], world)[rank]
        a_all, a_off, b_all, b_off = slice_batch(*full, s0, s1)
        del full
    else:
        a_all, a_off, b_all, b_off = make_batch(A, args, rank)
    n_local = len(a_off) - 1
    # one process per GPU: this rank's threads and page-locked buffers stay on the GPU's NUMA node (APA_BENCH_NUMA=0 turns it off)
    cpus_before = os.sched_getaffinity(0)
    numa_cpus = A.bind_to_device_numa(local_rank) if os.environ.get("APA_BENCH_NUMA", "1") != "0" else 0
    eng = A.Engine.shared(local_rank)  # the engine apa_align_batch_multi runs on: resident and end-to-end steps share one scratch arena
    eff_cells = float(np.sum((a_off[1:] - a_off[:-1]).astype(np.float64) * (b_off[1:] - b_off[:-1])))
    total_bp = float(a_off[-1])
    a_pin, b_pin = A.pinned_copy(a_all), A.pinned_copy(b_all)

    # ---- resident-input timing (value)
    batch = eng.upload(a_pin, a_off, b_pin, b_off)
    for _ in range(args.warmup):
        batch.run(preset_id, trace)
    barrier()
    sampler = ClockSampler(local_rank)
    kernel_ms, launches, st = 0.0, 0, None
    phase_ms = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.run(preset_id, trace)  # synchronises its stream; kernel_ms is CUDA-event time on that stream
        st = batch.stats()
        kernel_ms += st["kernel_ms"]
        phase_ms += np.array(st["phase_ms"])  # CUDA events around each phase kernel, on the engine's stream
        launches += st["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    overlapped = bool(st.get("overlapped"))
    if overlapped:
        # The timed steps ran the three phase kernels overlapped at their tails. Their individual durations (roofline, kernel
        # shares - what an ncu launch list shows, where launches are serialised) come from a few steps run back to back.
        os.environ["APA_OVERLAP"] = "0"
        phase_ms = np.zeros(3)
        n_serial = min(args.steps, 3)
        serial_ms = 0.0
        for _ in range(n_serial):
            batch.run(preset_id, trace)
            phase_ms += np.array(batch.stats()["phase_ms"])
            serial_ms += batch.stats()["kernel_ms"]
        del os.environ["APA_OVERLAP"]
        phase_ms *= args.steps / n_serial
        serial_ms /= n_serial
    else:
        serial_ms = kernel_ms / args.steps
    costs, pool, off, ln = batch.download_raw()
    digests = A.cigar_digests(pool, off, ln) if trace else None
    d2h_bytes = batch.stats()["d2h_bytes"]
    batch.free_pool(pool)
    batch.free()  # the end-to-end steps bring their own copy of the batch
    ms_step = kernel_ms / args.steps

    # ---- end-to-end through the public multi-GPU batch call with host (page-locked) buffers (e2e): H2D of this step's bases,
    # kernels, D2H of costs + CIGAR texts inside the timed region, every step; this rank drives its own GPU
    e2e_ms = []
    h2d_bytes = 0
    e2e_steps = args.e2e_steps or args.steps
    s2 = None
    for it in range(args.warmup + e2e_steps):
        barrier()
        t1 = time.perf_counter()
        c2, pool2, off2, ln2, s2 = A.align_batch_multi([local_rank], a_pin, a_off, b_pin, b_off, preset_id, trace)
        dt = (time.perf_counter() - t1) * 1e3
        s2 = s2[0]
        h2d_bytes, d2h_bytes = s2["h2d_bytes"], s2["d2h_bytes"]
        if it >= args.warmup:
            e2e_ms.append(dt)
        if it == 0:  # the e2e path gives what the resident path gives, on every pair
            assert (c2 == costs).all(), "e2e costs differ from the resident run"
            assert not trace or (A.cigar_digests(pool2, off2, ln2) == digests).all(), "e2e CIGARs differ from the resident run"
        A.free_pool(pool2)
    clocks = sampler.stop()
    e2e_step = float(np.mean(e2e_ms)) if e2e_ms else float('nan')

    # ---- the same end to end with the bases already in 2-bit planes on the host (apa_align_batch_packed): what a pipeline that
    # keeps its reads packed sees; 4x fewer bytes through host memory and PCIe. Packing (apa_pack_sequences) is outside the timed region.
    a_pl, a_len = A.pack_sequences(a_all, a_off)
    b_pl, b_len = A.pack_sequences(b_all, b_off)
    pk_ms, pk_h2d = [], 0
    for it in range(2 + min(e2e_steps, 5)):
        barrier()
        t1 = time.perf_counter()
        c3, pool3, off3, ln3, s3 = eng.align_batch_packed(a_pl, a_len, b_pl, b_len, preset_id, trace)
        dt = (time.perf_counter() - t1) * 1e3
        pk_h2d = s3["h2d_bytes"]
        if it >= 2:
            pk_ms.append(dt)
        if it == 0:
            assert (c3 == costs).all() and (not trace or (A.cigar_digests(pool3, off3, ln3) == digests).all()), "packed-input results differ"
        eng.free_pool(pool3)
    pk_step = float(np.mean(pk_ms))

    # ---- reduce over ranks: max time, sum of work
    agg = np.array([ms_step, e2e_step, *(phase_ms / args.steps), pk_step], dtype=np.float64)
    work = np.array([eff_cells, float(st["computed_cells"]), total_bp, float(st["dp_word_steps"]), float(a_off[-1] + b_off[-1]),
                     float(st["dp_issue_steps"]), float(n_local)], dtype=np.float64)
    if dist is not None:
        import torch
        t_agg = torch.tensor(agg, device="cuda")
        t_work = torch.tensor(work, device="cuda")
        dist.all_reduce(t_agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_work, op=dist.ReduceOp.SUM)
        agg, work = t_agg.cpu().numpy(), t_work.cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_step, e2e_step = float(agg[0]), float(agg[1])
    k_ms = [float(x) for x in agg[2:5]]
    pk_step = float(agg[5])
    eff_all, comp_all, bp_all, wsteps_all, bases_all, isteps_all, pairs_all = work
    value = eff_all / (ms_step / 1e3) / 1e9
    hbm_peak, peak_src, sm_max = peaks()
    comp_gpu = comp_all / world
    sm_mhz = clocks.get("sm_mhz") or sm_max
    # Dominant kernel = the longest of the three phase kernels (CUDA events around each launch, live in this run).
    # Algorithmic HBM bytes per launch:
    #   pass kernel  (K1 block DP): 48 B per (64 rows x 256 cols) lane-block = computed_cells * 48 / 16384
    #   build kernel (K2 GCSH)    : 0.7 B per base pair of input (2-bit planes of a and b, seed table, matches)
    #   trace kernel (K3)         : 2 B per base of a (DT records + V columns re-read) + the CIGAR text written
    alg = {KERNELS[0]: 0.7 * bases_all / world / 2, KERNELS[1]: comp_gpu * 48.0 / 16384.0,
           KERNELS[2]: 2.0 * bp_all / world + (d2h_bytes if trace else 0)}
    if sum(k_ms) <= 0:  # fused single-kernel path (arenas of the whole batch did not fit in HBM)
        dom, dom_ms = "apa_align_kernel", ms_step
        alg[dom] = sum(alg.values())
    else:
        dom = KERNELS[int(np.argmax(k_ms))]
        dom_ms = max(k_ms)
    achieved = alg[dom] / (dom_ms / 1e3) / 1e9
    # INT32 view of the block DP: ALU-pipe instructions per 32-row word step (csrc/apa_blockdp.cuh: 8 LOP3 + 2 SHF; the 4 IMAD
    # run on the FMA pipe and are not counted) against the ALU-pipe issue rate MEASURED on this GPU by apa_int32_peak
    # (a dependent LOP3/SHF chain per warp, all SMs full) - not an assumed figure.
    int_peak_meas = eng.int32_peak()
    int_ops = (wsteps_all / world) * 10.0
    int_peak = int_peak_meas["alu_lane_ops_per_s"]
    pass_ms = k_ms[1] if k_ms[1] > 0 else ms_step
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32 bit-vectors (i32 costs)",
        "data": "synthetic", "config": config,
        "computed_gcups": comp_all / (ms_step / 1e3) / 1e9, "aligned_bp_per_s": bp_all / (ms_step / 1e3),
        "pairs_per_s": pairs_all / (ms_step / 1e3), "wall_ms_per_step": wall_ms / args.steps,
        "passes_per_pair": st["passes"] / max(n_local, 1), "retries": st["retries"],
        "h_queries_per_pair": st["score_calls"] / max(n_local, 1),
        "contour_probe_rounds_per_query": (st["score_probes"] / st["score_calls"]) if st["score_calls"] else None,
        # block DP: useful 32-row lane-steps / lane-steps issued (32 lanes x steps of every chunk sweep, ramps included)
        "lane_utilisation": (wsteps_all / isteps_all) if isteps_all else None,
        "host": {"numa_bound_cpus_rank0": numa_cpus, "cpus_visible": len(os.sched_getaffinity(0))},
        "path": {"pass_warps_per_pair": st["pass_warps_per_pair"], "waves": st["waves"], "e2e_upload_mode": s2["upload_mode"] if s2 else None,
                 "e2e_upload_chunks": s2["upload_chunks"] if s2 else None},
        "kernel_overlap": {"overlapped_in_timed_steps": overlapped, "ms_per_step_back_to_back": serial_ms,
                           "note": "per-kernel ms below are from back-to-back launches (APA_OVERLAP=0); the timed steps launch the three "
                                   "phase kernels together and let each fill the SM slots the previous one's tail leaves empty"},
        "kernels": [{"name": k, "ms_per_launch": t, "share_of_step": t / serial_ms if serial_ms else None,
                     "algorithmic_gb_per_launch": alg[k] / 1e9} for k, t in zip(KERNELS, k_ms)],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic_per_launch(args, dom), "traffic_unit": "GB per launch (dram read + write)",
                     "peak_source": peak_src, "kernel": dom, "kernel_ms_per_launch": dom_ms,
                     "note": "integer/latency-bound path: ~0.003 algorithmic B/cell in the block DP; the INT32-pipe view of the "
                             "block DP (apa_phase_pass_kernel) is in the int32_* fields",
                     "int32_ops_per_s": int_ops / (pass_ms / 1e3), "int32_peak_ops_per_s": int_peak,
                     "int32_frac": int_ops / (pass_ms / 1e3) / int_peak,
                     "int32_peak_source": "measured in this run (apa_int32_peak: LOP3 + SHF chains, ALU pipe)",
                     "int32_peak_detail": int_peak_meas},
        "e2e": {"value": eff_all / (e2e_step / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_step, "steps": len(e2e_ms), "api": "apa_align_batch_multi (host page-locked buffers in, host buffers out)"},
        "e2e_packed_input": {"value": eff_all / (pk_step / 1e3) / 1e9, "unit": UNIT, "ms_per_step": pk_step, "h2d_bytes_per_step": int(pk_h2d),
                             "api": "apa_align_batch_packed (2-bit planes in page-locked host memory in, host buffers out; packing not timed)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if any(st["phase_cycles"]):  # only with a TIMERS=1 build
        out["phase_cycles"] = dict(zip(["heuristic_build", "block_dp", "passes_total", "trace_total", "dt_trace", "cigar_text", "h_queries",
                                        "prune_update"], [int(x) for x in st["phase_cycles"]]))
    if world == 1:
        os.sched_setaffinity(0, cpus_before)  # the CPU baseline gets every CPU this process may use
        out["cpu_baseline"] = cpu_leg(args, a_all, a_off, b_all, b_off, preset_id, trace, args.cpu_sample)
        # in-run parity: the GPU results of this very batch against the oracle's, on every pair of the CPU sample
        o_costs, o_digests = cpu_leg.sample_results
        k = len(o_costs)
        assert (costs[:k] == o_costs).all(), f"GPU costs differ from the oracle on pairs {np.flatnonzero(costs[:k] != o_costs)[:5]}"
        if trace:
            assert (digests[:k] == o_digests).all(), f"GPU CIGARs differ from the oracle on pairs {np.flatnonzero(digests[:k] != o_digests)[:5]}"
        out["parity_checked_pairs"] = int(k)
        out["parity"] = "GPU cost" + (" + CIGAR text (FNV-1a)" if trace else "") + f" == oracle on the first {k} pairs of the batch; e2e == resident on all {n_local}"
        out["cpu_baseline"]["micro"] = cpu_micro()
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
