#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print(d['ms_per_step'], [ (k['name'],round(k['ms_per_launch'],2)) for k in d['kernels']], 'e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:apa_phase -c 3 -o gpurun_out/phase_full3 -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu2.log 2>&1
timeout 300 python bench.py --preset simple --pairs 2000 --steps 3 --warmup 3 --cpu-sample 16 > gpurun_out/bench_simple.json 2> gpurun_out/bench_simple.err
timeout 300 python bench.py --n 10000 --no-trace --pairs 10000 --steps 5 --warmup 3 --cpu-sample 64 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
ls -la gpurun_out
