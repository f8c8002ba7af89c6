#!/bin/bash
# Safety gate before any kernel that has never run on hardware is launched uninstrumented: the small GPU tests that
# exercise it run under compute-sanitizer (memcheck squashes a bad access instead of faulting the GPU).  Only when
# the gate is clean does the validation script proper run.
# Gate 1: assembly + Dirichlet elimination of the validated suite (the element kernels were refactored after their last
# hardware run).  Gate 2: tests/test_gpu_z_unvalidated.py = everything that has never run on a GPU (nonlinear path, linear
# strain / stress recovery, SELL-C-sigma, BILU(0), device hecmw_mat_con, the IC kernel selection and its two new kernels).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
SAN="compute-sanitizer --error-exitcode 9 --print-limit 5"
timeout 1200 $SAN python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assembly or bc_elimination" > gpurun_out/gate_asm.log 2>&1
rc1=$?; tail -5 gpurun_out/gate_asm.log; echo "gate asm rc=$rc1"
FISTR_B200_RUN_NL_GPU=1 timeout 1500 $SAN python -m pytest tests/test_gpu_z_unvalidated.py -m gpu -q -rA -k "not I342 and not cylinder and not B3 and not C3 and not D3 and not E3 and not G3" > gpurun_out/gate_nl.log 2>&1
rc2=$?; tail -30 gpurun_out/gate_nl.log; echo "gate nl rc=$rc2"
if grep -q "Invalid\|Misaligned\|========= Error" gpurun_out/gate_asm.log gpurun_out/gate_nl.log || [ $rc1 -eq 9 ] || [ $rc2 -eq 9 ]; then
  echo "GATE: memory errors found, stopping"; grep -m 20 "=========" gpurun_out/gate_asm.log gpurun_out/gate_nl.log; exit 9
fi
echo "GATE clean"
bash tools/gpu_r2_validate.sh
