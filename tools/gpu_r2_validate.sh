#!/bin/bash
# SECOND GPU call (1 GPU), after tools/gpu_r2_first.sh was green: first hardware run of everything that never ran on a GPU
# (nonlinear path, linear strain / stress recovery, SELL-C-sigma, BILU(0), device hecmw_mat_con, SCALING, the two new IC
# condensation kernels), then their timings and the remaining BASELINE configs.
#   0. sanitizer gate on the opt-in test file (tools/gpu_gate.sh unvalidated) and on the IC kernel probe at a small size:
#      NOTHING below runs uninstrumented unless both are clean (the gate fails closed)
#   1. the opt-in tests uninstrumented      2. assembly probe (three IC condensation kernels side by side)
#   3. update / nodal-stress timings, SELL-C-sigma timings, BILU / device mat_con probe
#   4. parity on the benched mesh (tools/parity_big.py, H150)      5. configs 0 / 2 / 3 / 4      6. ncu of the assembly kernels
mkdir -p gpurun_out
GATE_TIMEOUT=1800 bash tools/gpu_gate.sh unvalidated || { echo "opt-in tests not clean under compute-sanitizer: stopping"; exit 9; }
timeout 900 compute-sanitizer --error-exitcode 9 --print-limit 5 --report-api-errors no python tools/asm_probe.py --n 12 --probes 0 > gpurun_out/gate_asm_probe.log 2>&1
rc=$?; tail -4 gpurun_out/gate_asm_probe.log
if [ $rc -ne 0 ] || ! grep -q "ERROR SUMMARY: 0 errors" gpurun_out/gate_asm_probe.log; then echo "IC kernel probe not clean under compute-sanitizer (rc=$rc): stopping"; exit 9; fi
export FISTR_B200_RUN_UNVALIDATED=1
timeout 900 python -m pytest tests/test_gpu_z_unvalidated.py -m gpu -q -rA 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_unvalidated.log
timeout 600 python tools/asm_probe.py --n 100 2>&1 | tee gpurun_out/asm_probe.log
timeout 600 python tools/asm_probe.py --tet 40 --probes 0,1 2>&1 | tee -a gpurun_out/asm_probe.log
timeout 600 python tools/post_probe.py --n 100 --tet 40 2>&1 | tee gpurun_out/post_probe.log
timeout 600 python tools/perf_probe.py --tet 40 --variants 2 --sigma 0,256,1024,4096 --solve 0 2>&1 | tee gpurun_out/probe_tet40_sigma.log
timeout 600 python tools/r2_new_paths_probe.py --n 100 --solve-n 60 2>&1 | tee gpurun_out/new_paths_probe.log
timeout 1500 python tools/parity_big.py --n 150 --spread-file profiles/r2_oracle_spread_H150.json > gpurun_out/parity_H150.json 2> gpurun_out/parity_H150.log; echo "parity rc=$?"; cat gpurun_out/parity_H150.log
for c in 0 2 3; do timeout 900 python tools/run_configs.py --config $c 2>&1 | tail -1 | tee gpurun_out/config_$c.json; done
timeout 900 python tools/run_configs.py --config 3 --sigma -1 2>&1 | tail -1 | tee gpurun_out/config_3_sigma_auto.json
timeout 900 python tools/run_configs.py --config 0 --ssor-fused 1 2>&1 | tail -1 | tee gpurun_out/config_0_ssor_fused.json
FISTR_B200_RUN_NL_GPU=1 timeout 600 python tools/run_configs.py --config 4 2>&1 | tail -1 | tee gpurun_out/config_4.json
CMD="python tools/asm_probe.py --n 100 --probes 0"
timeout 300 $CMD > gpurun_out/plain_asm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stiff_kernel|gather_sum" -c 4 -o gpurun_out/prof_asm_r2 $CMD > gpurun_out/ncu_asm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_asm.log
