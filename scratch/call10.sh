#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), 'GCUPS', round(d['value']), 'computed', round(d['computed_gcups']), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']], 'e2e', d['e2e'].get('ms_per_step'), 'cpu', round(d['cpu_baseline']['value']), 'passes', d['passes_per_pair'], 'retries', d['retries'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --e2e-steps 2 --cpu-sample 64 > gpurun_out/bench_full_c.json 2> gpurun_out/bench_full_c.err; summ gpurun_out/bench_full_c.json
timeout 500 python bench.py --n 1000000 --e 0.15 --pairs 1000 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 16 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; summ gpurun_out/bench_cfg4.json
timeout 300 python bench.py --n 10000000 --e 0.05 --pairs 1 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; summ gpurun_out/bench_cfg5.json
for f in gpurun_out/bench_cfg4.err gpurun_out/bench_cfg5.err gpurun_out/bench_full_c.err; do tail -n 3 $f; done
