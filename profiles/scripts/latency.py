"""Single-call latency of the drop-in symbol astarpa2_full (one pair per call, astarpa-c/astarpa.h:27-32) against the CPU port
(oracle) on one thread: python profiles/scripts/latency.py   (run on the GPU box). Median of 15 calls after 3 warm-up calls."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import astar_pairwise_aligner_b200 as A  # noqa: E402
import oracle_lib as O  # noqa: E402

L = A.load_library()
rows = []
for n in (1000, 10000, 100000, 1000000):
    a, b = A.generate_pair(n, 0.05, 0, 31415)
    gpu, cpu = [], []
    for it in range(18 if n < 1000000 else 6):
        cig, ln = C.c_void_p(), C.c_size_t()
        t0 = time.perf_counter()
        cost = L.astarpa2_full(a, len(a), b, len(b), C.byref(cig), C.byref(ln))
        dt = time.perf_counter() - t0
        L.astarpa_free_cigar(cig)
        if it >= 3:
            gpu.append(dt)
    for it in range(8 if n < 1000000 else 3):
        t0 = time.perf_counter()
        oc, _, _ = O.align(a, b, 1, True)
        dt = time.perf_counter() - t0
        if it >= 1:
            cpu.append(dt)
    assert oc == cost
    gpu.sort(), cpu.sort()
    rows.append({"n": n, "e": 0.05, "gpu_ms_per_call": 1e3 * gpu[len(gpu) // 2], "cpu_port_ms_per_call_1_thread": 1e3 * cpu[len(cpu) // 2]})
print(json.dumps({"what": "astarpa2_full, one pair per call through the C-ABI (host buffers in, CIGAR text out)", "rows": rows}))
