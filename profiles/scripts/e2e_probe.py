import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import astar_pairwise_aligner_b200 as A
args = A.generate_batch(10000, 100000, 0.05, 0, 31415)
a_pin, b_pin = A.pinned_copy(args[0]), A.pinned_copy(args[2])
eng = A.Engine(0)
for mode in ((None, "1", "0") if not os.environ.get("PROBE_BUILD_CTAS") else (None,)):
    if mode is None: os.environ.pop("APA_RAW", None)
    else: os.environ["APA_RAW"] = mode
    for thr in ((None, "8", "4") if mode is None and not os.environ.get("PROBE_BUILD_CTAS") else (None,)):
        if thr is None: os.environ.pop("APA_PACK_THREADS", None)
        else: os.environ["APA_PACK_THREADS"] = thr
        ts = []
        for it in range(4):
            t0 = time.perf_counter()
            c, pool, off, ln, st = eng.align_batch_raw(a_pin, args[1], b_pin, args[3], 1, True)
            ts.append((time.perf_counter() - t0) * 1e3)
            eng.free_pool(pool)
        print("APA_RAW", mode, "threads", thr, "e2e ms", [round(x, 1) for x in ts], "mode", st["upload_mode"], "raw chunks", st["upload_chunks_raw"], "of", st["upload_chunks"], "h2d MB", st["h2d_bytes"] >> 20, flush=True)
