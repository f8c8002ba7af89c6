#!/bin/bash
# Safety gate: GPU tests under compute-sanitizer (memcheck squashes a bad access instead of faulting the GPU) on ONE GPU.
#   tools/gpu_gate.sh validated     -> tests/test_gpu_parity.py (everything the driver's `pytest -m gpu` runs by default)
#   tools/gpu_gate.sh unvalidated   -> tests/test_gpu_z_unvalidated.py (opt-in file: kernels that never ran on hardware)
#   tools/gpu_gate.sh <name> <pytest args...>   -> any other selection, logged as gpurun_out/gate_<name>.log
# FAILS CLOSED: exit 0 only if pytest ended with rc 0 or 1 (ran to completion; 1 = assertion failures, reported but not a
# memory fault) AND the log carries "ERROR SUMMARY: 0 errors" for every sanitized process AND no error record.  A timeout
# (124/137), a crash, a collection error or a missing summary line all mean "not clean" (exit 9).
set -u
mkdir -p gpurun_out
what=${1:-validated}; shift || true
case "$what" in
  validated)   sel=(tests/test_gpu_parity.py -m gpu -q -k "not large_properties") ;;
  unvalidated) export FISTR_B200_RUN_UNVALIDATED=1; sel=(tests/test_gpu_z_unvalidated.py -m gpu -q -rA "$@") ;;
  *)           sel=("$@") ;;
esac
log=gpurun_out/gate_$what.log
limit=${GATE_TIMEOUT:-1500}
timeout "$limit" compute-sanitizer --error-exitcode 9 --print-limit 5 --report-api-errors no --target-processes all \
    python -m pytest "${sel[@]}" > "$log" 2>&1
rc=$?
tail -n 12 "$log"
nsum=$(grep -c "ERROR SUMMARY:" "$log")
nzero=$(grep -c "ERROR SUMMARY: 0 errors" "$log")
if [ "$rc" -ne 0 ] && [ "$rc" -ne 1 ]; then echo "GATE $what: NOT CLEAN (rc=$rc)"; exit 9; fi
if [ "$nsum" -lt 1 ] || [ "$nsum" -ne "$nzero" ]; then echo "GATE $what: NOT CLEAN (summaries $nzero/$nsum clean)"; grep -m 20 "=========" "$log"; exit 9; fi
if grep -q "========= Invalid\|========= Misaligned\|========= Error\|========= Program hit" "$log"; then
  echo "GATE $what: NOT CLEAN (error records)"; grep -m 20 "=========" "$log"; exit 9; fi
echo "GATE $what: clean (pytest rc=$rc)"
exit 0
