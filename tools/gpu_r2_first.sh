#!/bin/bash
# Round 2, first GPU call (ONE GPU).  Order: sanitizer gate on the default suite -> the default suite uninstrumented ->
# smoke -> bench line (H150) -> launch list under ncu.  The opt-in (never-on-hardware) file is gated separately afterwards,
# and runs uninstrumented only if its gate is clean.  Nothing after a failed gate of the default suite runs at all.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
bash tools/gpu_gate.sh validated || { echo "default suite not clean under compute-sanitizer: stopping"; exit 9; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
GATE_TIMEOUT=1500 bash tools/gpu_gate.sh unvalidated -k "not I342 and not cylinder"
echo "unvalidated gate rc=$?"
