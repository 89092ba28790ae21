#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "general" 2>&1 | tail -40 | tee gpurun_out/pytest_general.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not general" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
