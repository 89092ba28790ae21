#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
APA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 64 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
grep batch_run gpurun_out/bench_full.err | sed -n 5,6p
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print(d['ms_per_step'], [ (k['name'],round(k['ms_per_launch'],2)) for k in d['kernels']], 'e2e', d['e2e']['ms_per_step'])
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:apa_phase -c 3 -o gpurun_out/phase_full2 -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
