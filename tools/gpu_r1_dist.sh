#!/bin/bash
# multi-GPU pass: parity of the partitioned solve (peer-memory and NCCL transports), then the bench at this N
N=${1:-2}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
: > gpurun_out/dist_check_n$N.log
for args in "--hex 24" "--tet 6" "--hex 24 --comm nccl" "--hex 16 --precond 1"; do
  timeout 300 $TR tests/dist_check.py $args 2>&1 | grep -E "PASS|FAIL|Error|error|rror:" | head -5 | tee -a gpurun_out/dist_check_n$N.log
done
for comm in p2p "$@"; do
  timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --comm $comm > gpurun_out/bench_n${N}_$comm.json 2> gpurun_out/bench_n${N}_$comm.err; echo "bench $comm rc=$?"
  cat gpurun_out/bench_n${N}_$comm.json; grep -v "^W1018\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_$comm.err | tail -20
done
