#!/bin/bash
# Run on a B200 box from the repo root (under gpurun): the commands behind the profiles/r1c_* artefacts.
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:apa_phase_pass_coop -c 1 -o gpurun_out/coop_full_r1c -f \
    python bench.py --n 300000 --e 0.15 --pairs 200 --steps 1 --warmup 0 --e2e-steps 0 --cpu-sample 4 > gpurun_out/b_ncu_coop.log 2>&1
ls -la gpurun_out/coop_full_r1c.ncu-rep; tail -2 gpurun_out/b_ncu_coop.log | cut -c1-300
