#!/bin/bash
# multi-GPU pass: parity of the partitioned solve, then the bench at this N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py --hex 24 2>&1 | grep -E "PASS|FAIL|Error|error" | tee gpurun_out/dist_check_n$N.log
timeout 600 $TR tools/dist_check.py --tet 6 2>&1 | grep -E "PASS|FAIL|Error|error" | tee -a gpurun_out/dist_check_n$N.log
timeout 600 $TR tools/dist_check.py --hex 16 --precond 1 2>&1 | grep -E "PASS|FAIL|Error|error" | tee -a gpurun_out/dist_check_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_n$N.json; grep -v "^W1018\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -20
