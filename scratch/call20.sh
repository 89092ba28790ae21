#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
for r in 56 64; do
  APA_BUILD_REGS=$r APA_PASS_REGS=$r APA_TRACE_REGS=$r timeout 600 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_regs$r.json 2> gpurun_out/bench_regs$r.err; summ gpurun_out/bench_regs$r.json
done
