#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14 | tee gpurun_out/pytest_gpu.log
APA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 64 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
grep batch_run gpurun_out/bench_full.err | sed -n 5,6p
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print(d['ms_per_step'], [ (k['name'],round(k['ms_per_launch'],2)) for k in d['kernels']], 'e2e', d['e2e']['ms_per_step'], 'hq', d['h_queries_per_pair'], d['contour_probe_rounds_per_query'])
PY
