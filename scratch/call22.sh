#!/bin/bash
mkdir -p gpurun_out
APA_PARTS=2 timeout 300 python -m pytest tests -m gpu -x -q -k "mixed_lengths or full_size or config2" 2>&1 | tail -3
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
for p in 1 2 3 4; do
  APA_PARTS=$p timeout 300 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_parts$p.json 2> gpurun_out/bench_parts$p.err; summ gpurun_out/bench_parts$p.json
done
