#!/bin/bash
mkdir -p gpurun_out
APA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 2 --e2e-steps 3 --cpu-sample 8 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
grep -E "apa_align_batch|batch_run" gpurun_out/bench_dbg.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dbg.json'))
print(d['ms_per_step'], d['e2e'])
PY
nproc; python -c "import os; print(os.cpu_count(), len(os.sched_getaffinity(0)))"
