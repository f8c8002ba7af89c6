#!/bin/bash
# quick multi-GPU pass: one peer-memory parity check + bench for each transport given
N=${1:-2}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_check.py --hex 24 2>&1 | grep -E "PASS|FAIL|rror" | head -5 | tee gpurun_out/dist_check_n$N.log
for comm in "$@"; do
  timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --comm $comm --no-e2e > gpurun_out/bench_n${N}_$comm.json 2> gpurun_out/bench_n${N}_$comm.err; echo "bench $comm rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_$comm.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step","iteration_profile_us","clocks","gpu_launches")}, d["config"]["comm"], d["roofline"]["achieved"])
PY
  grep -v "^W1018\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_$comm.err | tail -5
done
