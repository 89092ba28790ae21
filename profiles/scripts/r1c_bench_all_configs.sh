#!/bin/bash
# Run on a B200 box from the repo root (under gpurun): the commands behind the profiles/r1c_* artefacts.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), 'GCUPS', round(d['value']), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d.get('kernels',[])], 'e2e', d['e2e'].get('ms_per_step'), 'cpu', round(d['cpu_baseline']['value']))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; summ gpurun_out/bench_full.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; summ gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:apa_phase -c 3 --csv --log-file gpurun_out/dram_bytes.csv \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:apa_phase -c 3 -o gpurun_out/phase_full_r1c -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu2.log 2>&1
timeout 300 python bench.py --preset simple --pairs 2000 --steps 3 --warmup 3 --cpu-sample 16 > gpurun_out/bench_simple.json 2> gpurun_out/bench_simple.err; summ gpurun_out/bench_simple.json
timeout 300 python bench.py --preset simple --pairs 500 --steps 3 --warmup 3 --cpu-sample 16 > gpurun_out/bench_simple500.json 2> gpurun_out/bench_simple500.err; summ gpurun_out/bench_simple500.json
timeout 300 python bench.py --n 10000 --no-trace --pairs 10000 --steps 5 --warmup 3 --cpu-sample 64 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; summ gpurun_out/bench_cfg1.json
ls -la gpurun_out | tail -20
