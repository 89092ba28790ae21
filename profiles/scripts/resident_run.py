"""One resident batch, run AB_REPS + 1 times (default: BASELINE configs[2] shape; AB_PAIRS / AB_N / AB_E / AB_PRESET override).
Prints the per-phase CUDA-event times and a digest of all costs and CIGAR texts. Used by ab_phase.py (APA_LIB = the variant
library) and as the target of the ncu captures (profiles/scripts/r2_capture.sh)."""

import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import astar_pairwise_aligner_b200 as A
pairs, n, e = int(os.environ.get("AB_PAIRS", "10000")), int(os.environ.get("AB_N", "100000")), float(os.environ.get("AB_E", "0.05"))
preset = int(os.environ.get("AB_PRESET", "1"))
reps = int(os.environ.get("AB_REPS", "4"))
args = A.generate_batch(pairs, n, e, 0, 31415)
eng = A.Engine(0)
b = eng.upload(*args)
b.run(preset, True)
ph, tot = [], []
for _ in range(reps):
    b.run(preset, True)
    st = b.stats()
    ph.append(st["phase_ms"]); tot.append(st["kernel_ms"])
costs, pool, off, ln = b.download_raw()
dig = A.cigar_digests(pool, off, ln)
h = int(np.bitwise_xor.reduce(dig * np.uint64(0x9E3779B97F4A7C15) + costs.astype(np.uint64)))
print(json.dumps({"phase_ms": [float(x) for x in np.median(np.array(ph), axis=0)], "kernel_ms": float(np.median(tot)), "digest": h,
                  "retries": st["retries"]}))
