#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_check.py --hex 24 2>&1 | grep -E "PASS|FAIL|rror" | head -5 | tee gpurun_out/dist_check_n8_p2p.log
timeout 300 $TR tests/dist_check.py --tet 6 2>&1 | grep -E "PASS|FAIL|rror" | head -5 | tee -a gpurun_out/dist_check_n8_p2p.log
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --comm p2p > gpurun_out/bench_n8_p2p.json 2> gpurun_out/bench_n8_p2p.err; echo "bench p2p rc=$?"
cat gpurun_out/bench_n8_p2p.json | cut -c1-3000
timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 3 --comm p2p --hex 200 --no-e2e > gpurun_out/bench_n8_p2p_h200.json 2> gpurun_out/bench_n8_p2p_h200.err; echo "bench h200 rc=$?"
cat gpurun_out/bench_n8_p2p_h200.json | cut -c1-3000
grep -v "^W1018\|^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_n8_p2p*.err | tail -8
