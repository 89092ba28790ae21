#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "pair_stats or general" 2>&1 | tail -30 | tee gpurun_out/pytest_general.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not general and not pair_stats" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
