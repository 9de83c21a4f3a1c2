#!/usr/bin/env bash
# Round-2 GPU run 6: consistent build of tc2 v3 (TMA tensor tiles, 8 epilogue warps), episodic
# buffer kernel, bucketed all-reduce hook, strided TD-loss inputs, fused e2e graphs.
set -u
O=gpurun_out/r2_run6
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run pytest_gpu_pdl 1500 env B200RL_PDL=1 python -m pytest tests -m gpu -q -p no:cacheprovider
run bench 900 python bench.py
run bench_pdl 300 env B200RL_PDL=1 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
tail -12 "$O/pytest_gpu.out"
tail -6 "$O/pytest_gpu_pdl.out"
tail -1 "$O/bench.out" | cut -c1-4000
tail -3 "$O/bench.err"
tail -1 "$O/bench_pdl.out" | cut -c1-600
