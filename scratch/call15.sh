#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), 'GCUPS', round(d['value']), 'computed', round(d['computed_gcups']), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']], 'cpu', round(d['cpu_baseline']['value']), 'passes', d['passes_per_pair'], 'retries', d['retries'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
APA_DEBUG_TIMING=1 timeout 500 python bench.py --n 1000000 --e 0.15 --pairs 1000 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 16 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; summ gpurun_out/bench_cfg4.json
tail -n 5 gpurun_out/bench_cfg4.err
timeout 300 python bench.py --n 10000000 --e 0.05 --pairs 1 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; summ gpurun_out/bench_cfg5.json
