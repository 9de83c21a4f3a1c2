#!/usr/bin/env bash
# Round-2 GPU run 19: paired forward (online + target network layers in one launch), ring index
# without divisions; full suite + bench A/B.
set -u
O=gpurun_out/r2_run19
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run pair_test 300 python -m pytest tests/test_nn_gpu.py tests/test_dqn_gpu.py tests/test_baseline_parity_gpu.py -m gpu -q -p no:cacheprovider -x
run pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run bench_pair 300 python bench.py --no-extra --no-cpu-baseline
run bench_nopair 300 env B200RL_DQN_PAIR_FWD=0 python bench.py --no-extra --no-cpu-baseline
run tc2_check 200 python profiles/tc2_check.py --only conv
cat "$O/summary.txt"
tail -6 "$O/pair_test.out"
tail -4 "$O/pytest_gpu.out"
for f in bench_pair bench_nopair; do echo "$f: $(tail -1 $O/$f.out | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'loss', d['final_loss'])")"; tail -2 $O/$f.err; done
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-120
