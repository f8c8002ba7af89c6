#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python tools/perf_probe.py --n 150 --bps 4,8,16 > gpurun_out/probe150.log 2>&1
tail -30 gpurun_out/probe150.log
timeout 600 python tools/perf_probe.py --tet 40 --variants 2 > gpurun_out/probe_tet40.log 2>&1
tail -12 gpurun_out/probe_tet40.log
