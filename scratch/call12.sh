#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "coop or waves" 2>&1 | tail -30 | tee gpurun_out/pytest_coop.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not coop and not waves" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), 'GCUPS', round(d['value']), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']], 'cpu', round(d['cpu_baseline']['value']), 'passes', d['passes_per_pair'], 'retries', d['retries'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
timeout 300 python bench.py --preset simple --pairs 2000 --steps 3 --warmup 2 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_simple_auto.json 2> gpurun_out/bench_simple_auto.err; summ gpurun_out/bench_simple_auto.json
APA_COOP=4 timeout 300 python bench.py --preset simple --pairs 2000 --steps 3 --warmup 2 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_simple_c4.json 2> gpurun_out/bench_simple_c4.err; summ gpurun_out/bench_simple_c4.json
APA_COOP=8 timeout 300 python bench.py --preset simple --pairs 500 --steps 3 --warmup 2 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_simple500_c8.json 2> gpurun_out/bench_simple500_c8.err; summ gpurun_out/bench_simple500_c8.json
APA_COOP=1 timeout 300 python bench.py --preset simple --pairs 500 --steps 3 --warmup 2 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_simple500_c1.json 2> gpurun_out/bench_simple500_c1.err; summ gpurun_out/bench_simple500_c1.json
tail -n 3 gpurun_out/bench_simple_c4.err
