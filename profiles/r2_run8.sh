#!/usr/bin/env bash
# Round-2 GPU run 8: 8 loader + 8 converter warps, 4 epilogue warps; variant with 8 epilogue warps
# (libb200rl_e8.so, -DB200RL_TC2_EPI_WARPS=8) for comparison.
set -u
O=gpurun_out/r2_run8
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
E8=$PWD/agents_b200/lib/libb200rl_e8.so
run tc2_check 300 python profiles/tc2_check.py
run tc2_check_e8 300 env B200RL_LIB=$E8 python profiles/tc2_check.py
run tc2_trace 240 python profiles/tc2_trace.py
run nn_tests 600 python -m pytest tests/test_nn_gpu.py tests/test_baseline_parity_gpu.py tests/test_dqn_gpu.py -m gpu -q -p no:cacheprovider
run bench 300 python bench.py --no-extra --no-cpu-baseline
run bench_e8 300 env B200RL_LIB=$E8 python bench.py --no-extra --no-cpu-baseline
run bench_pdl 300 env B200RL_PDL=1 python bench.py --no-extra --no-cpu-baseline
run ppo 300 python profiles/configs.py ppo
cat "$O/summary.txt"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-220
tail -3 "$O/tc2_check.err"
echo ---- e8
grep -h '"speedup"' "$O/tc2_check_e8.out" | cut -c1-220
cat "$O/tc2_trace.out" | cut -c1-620
tail -6 "$O/nn_tests.out"
for b in bench bench_e8 bench_pdl; do tail -1 "$O/$b.out" | cut -c1-330; tail -2 "$O/$b.err"; done
tail -1 "$O/ppo.out" | cut -c1-600
