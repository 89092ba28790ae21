#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d['kernels']])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
L=$PWD/astar_pairwise_aligner_b200
for v in c ab1; do
  lib=$L/libapa_$v.so; [ $v = c ] && lib=$L/libastarpa_c.so
  APA_LIB=$lib timeout 600 python bench.py --steps 4 --warmup 3 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_full_$v.json 2> gpurun_out/bench_full_$v.err; summ gpurun_out/bench_full_$v.json
  APA_LIB=$lib timeout 300 python bench.py --preset simple --pairs 2000 --steps 3 --warmup 3 --e2e-steps 0 --cpu-sample 8 > gpurun_out/bench_simple_$v.json 2> gpurun_out/bench_simple_$v.err; summ gpurun_out/bench_simple_$v.json
done
