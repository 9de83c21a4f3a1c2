#!/usr/bin/env bash
# Round-2 GPU run 4: tc2 v2 (lean loaders, byte-perm u8 conversion, conflict-free MN u8 converter,
# relaxed waits, dense-only act' fusion, skinny dW) -- correctness A/B, pipeline trace, full tests
# (incl. baseline-config parity), full bench.
set -u
O=gpurun_out/r2_run4
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run tc2_check 240 python profiles/tc2_check.py
run tc2_trace 240 python profiles/tc2_trace.py
run layer_probe 300 python profiles/layer_probe.py
run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider
run bench 600 python bench.py
run bench_pdl 300 env B200RL_PDL=1 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-200
tail -3 "$O/tc2_check.err"
cat "$O/tc2_trace.out" | cut -c1-900
tail -3 "$O/tc2_trace.err"
cat "$O/layer_probe.out" | cut -c1-110
tail -12 "$O/pytest_gpu.out"
tail -1 "$O/bench.out" | cut -c1-6000
tail -3 "$O/bench.err"
tail -1 "$O/bench_pdl.out" | cut -c1-300
cat gpurun_out/parity_measured.json
