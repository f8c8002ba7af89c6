#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/asm_probe.py --n 100 2>&1 | tee gpurun_out/asm_probe.log
timeout 600 python tools/asm_probe.py --tet 40 --probes 0,1 2>&1 | tee -a gpurun_out/asm_probe.log
