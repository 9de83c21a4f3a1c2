#!/usr/bin/env bash
# Round-2 GPU run 21: Reverb-model table server on the GPU (tests/test_reverb_gpu.py), full GPU
# suite, short bench (watchdog restructuring of bench.py must not change the line).
set -u
O=gpurun_out/r2_run21
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run reverb_gpu 200 python -m pytest tests/test_reverb_gpu.py -m gpu -q -p no:cacheprovider
run pytest_gpu 400 python -m pytest tests -m gpu -q -p no:cacheprovider
run bench 240 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
tail -25 "$O/reverb_gpu.out"
tail -4 "$O/pytest_gpu.out"
tail -1 "$O/bench.out" | cut -c1-700
tail -3 "$O/bench.err"
