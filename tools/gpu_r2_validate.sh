#!/bin/bash
# First GPU call of the next round (1 GPU, ~6 min): validates everything written after GPU access closed in round 1.
#   1. the validated suite (must stay green)        2. the opt-in nonlinear GPU tests (first hardware run)
#   3. assembly probe after the IC-condensation rewrite   4. the bench line   5. ncu of the assembly kernels
#   (+ tools/r2_new_paths_probe.py: device mat_con and BILU(0), the two paths written last)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
FISTR_B200_RUN_NL_GPU=1 timeout 900 python -m pytest tests/test_gpu_z_unvalidated.py -m gpu -q -rA 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_nl.log
timeout 600 python tools/asm_probe.py --n 100 2>&1 | tee gpurun_out/asm_probe.log
timeout 600 python tools/asm_probe.py --tet 40 --probes 0,1 2>&1 | tee -a gpurun_out/asm_probe.log
timeout 600 python tools/post_probe.py --n 100 --tet 40 2>&1 | tee gpurun_out/post_probe.log
timeout 600 python tools/perf_probe.py --tet 40 --variants 2 --sigma 0,256,1024,4096 --solve 0 2>&1 | tee gpurun_out/probe_tet40_sigma.log
timeout 600 python tools/r2_new_paths_probe.py --n 100 --solve-n 60 2>&1 | tee gpurun_out/new_paths_probe.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
for c in 0 2 3; do timeout 900 python tools/run_configs.py --config $c 2>&1 | tail -1 | tee gpurun_out/config_$c.json; done
timeout 900 python tools/run_configs.py --config 3 --sigma 1024 2>&1 | tail -1 | tee gpurun_out/config_3_sigma1024.json
FISTR_B200_RUN_NL_GPU=1 timeout 600 python tools/run_configs.py --config 4 2>&1 | tail -1 | tee gpurun_out/config_4.json
CMD="python tools/asm_probe.py --n 100 --probes 0"
timeout 300 $CMD > gpurun_out/plain_asm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stiff_kernel|gather_sum" -c 4 -o gpurun_out/prof_asm_r2 $CMD > gpurun_out/ncu_asm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_asm.log
