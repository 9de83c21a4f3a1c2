#!/usr/bin/env bash
# Round-2 GPU run 5: tc2 v3 (precomputed MMA descriptors, 8 epilogue warps, TMA tensor tiles for
# plain 2-D operands), new row_copy_tma (CTA per row, 4 warps, two-half pipelining), fused e2e
# graphs.  Correctness A/B first; on failure the TMA path is switched off for the rest.
set -u
O=gpurun_out/r2_run5
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run tc2_check_notma 200 python profiles/tc2_check.py --flags 4
run tc2_check 200 python profiles/tc2_check.py
CHECK=$?
if [ $CHECK -ne 0 ] || grep -q '"max_rel_diff": [1-9]' "$O/tc2_check.out" || grep -q 'e-0[1-4]' "$O/tc2_check.out"; then
  echo "tc2_check with TMA failed or inaccurate (rc=$CHECK)" >> "$O/summary.txt"
  for op in fc1.fwd fc1.dX conv2.dX ragged; do
    run "tc2_only_$op" 60 python profiles/tc2_check.py --only "$op" --reps 3
  done
fi
run tc2_trace 240 python profiles/tc2_trace.py
run replay_tests 300 python -m pytest tests/test_replay_gpu.py tests/test_golden_fixtures.py tests/test_zz_late_gpu.py -m gpu -q -p no:cacheprovider
run gather 200 python profiles/configs.py gather
run layer_probe 300 python profiles/layer_probe.py
run pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider
run pytest_gpu_pdl 1200 env B200RL_PDL=1 python -m pytest tests -m gpu -q -p no:cacheprovider -x
run bench 600 python bench.py --no-extra
run bench_pdl 300 env B200RL_PDL=1 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-200
tail -3 "$O/tc2_check.err"
grep -h '"speedup"' "$O/tc2_check_notma.out" | cut -c1-160 | head -14
cat "$O/tc2_trace.out" | cut -c1-620
tail -4 "$O/replay_tests.out"
cat "$O/gather.out" | cut -c1-1500
cat "$O/layer_probe.out" | cut -c1-110
tail -8 "$O/pytest_gpu.out"
tail -4 "$O/pytest_gpu_pdl.out"
tail -1 "$O/bench.out" | cut -c1-1800
tail -3 "$O/bench.err"
tail -1 "$O/bench_pdl.out" | cut -c1-400
