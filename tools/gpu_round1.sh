#!/bin/bash
# first GPU session: correctness first, then quick perf numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py --n 100 > gpurun_out/probe100.log 2>&1
tail -30 gpurun_out/probe100.log
