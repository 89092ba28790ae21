#!/bin/bash
# Run on a B200 box from the repo root (under gpurun): the commands behind the profiles/r1c_* artefacts.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'],2), 'GCUPS', round(d['value']), [ (k['name'][10:],round(k['ms_per_launch'],2)) for k in d.get('kernels',[])], 'e2e', d['e2e'].get('ms_per_step'), 'cpu', round(d['cpu_baseline']['value']), 'traffic', d.get('roofline',{}).get('traffic'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; summ gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:apa_phase -c 3 --csv --log-file gpurun_out/dram_bytes.csv \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu3.log 2>&1
for k in build pass trace; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apa_phase_${k} -c 1 -o gpurun_out/${k}_full_r1c -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 8 > gpurun_out/b_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
