#!/usr/bin/env bash
# Round-2 GPU run 22: neighbour pre-sum in the col2im epilogue (B200RL_COL2IM_MERGE=1): parity of
# the conv / DQN / baseline-config tests with it on, per-layer dX timing and the bench, A/B.
set -u
O=gpurun_out/r2_run22
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run tests_merge 300 env B200RL_COL2IM_MERGE=1 python -m pytest tests/test_nn_gpu.py tests/test_dqn_gpu.py tests/test_baseline_parity_gpu.py -m gpu -q -p no:cacheprovider
run check_merge 120 env B200RL_COL2IM_MERGE=1 python profiles/tc2_check.py --only dX
run check_plain 120 env B200RL_COL2IM_MERGE=0 python profiles/tc2_check.py --only dX
run bench_merge 200 env B200RL_COL2IM_MERGE=1 python bench.py --no-extra --no-cpu-baseline
run bench_plain 200 env B200RL_COL2IM_MERGE=0 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
tail -6 "$O/tests_merge.out"
echo merge; grep -h '"speedup"' "$O/check_merge.out" | cut -c1-160
echo plain; grep -h '"speedup"' "$O/check_plain.out" | cut -c1-160
for f in bench_merge bench_plain; do echo "$f: $(tail -1 $O/$f.out | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['final_loss'], 'frac', d['roofline']['frac'])")"; tail -2 $O/$f.err; done
