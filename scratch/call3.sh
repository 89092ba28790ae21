#!/bin/bash
# register-budget experiment per phase + setup/tail timing
mkdir -p gpurun_out
run() { env "$@" APA_DEBUG_TIMING=1 python bench.py --steps 3 --warmup 2 --e2e-steps 0 --cpu-sample 8 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', round(d['ms_per_step'],2), [round(k['ms_per_launch'],2) for k in d['kernels']])"; grep batch_run gpurun_out/err.txt | tail -1; }
run APA_X=0
run APA_BUILD_REGS=56 APA_PASS_REGS=56 APA_TRACE_REGS=56
run APA_BUILD_REGS=64 APA_PASS_REGS=64 APA_TRACE_REGS=64
