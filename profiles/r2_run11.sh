#!/usr/bin/env bash
# Round-2 GPU run 11: mbarrier.try_wait (hardware-suspended waits) against test_wait polling (flag 64).
set -u
O=gpurun_out/r2_run11
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run tc2_check 150 python profiles/tc2_check.py
if [ $? -ne 0 ]; then
  echo "try_wait path failed" >> "$O/summary.txt"; tail -5 "$O/tc2_check.out" "$O/tc2_check.err"
  nvidia-smi > "$O/smi.txt" 2>&1
else
  run tc2_check_poll 150 python profiles/tc2_check.py --flags 64
  run nn_tests 600 python -m pytest tests/test_nn_gpu.py tests/test_baseline_parity_gpu.py tests/test_dqn_gpu.py tests/test_ppo_gpu.py tests/test_sac_gpu.py -m gpu -q -p no:cacheprovider
  run bench 300 python bench.py --no-extra --no-cpu-baseline
  run bench_poll 300 env B200RL_TC2_FLAGS=64 python bench.py --no-extra --no-cpu-baseline
  run ppo 300 python profiles/configs.py ppo
fi
cat "$O/summary.txt"
python - <<'PY'
import json
def load(f):
  out = {}
  try:
    for l in open(f):
      d = json.loads(l)
      if 'tc2_us' in d: out[d['layer'] + '.' + d['op']] = (d['tc2_us'], d['max_rel_diff'])
  except Exception as e: print('load', f, e)
  return out
a, b = load('gpurun_out/r2_run11/tc2_check.out'), load('gpurun_out/r2_run11/tc2_check_poll.out')
for k in a: print(k, 'try_wait', a[k][0], 'poll', b.get(k, (None,))[0], 'diff', a[k][1])
PY
tail -4 "$O/nn_tests.out"
for b in bench bench_poll; do tail -1 "$O/$b.out" | cut -c1-330; tail -2 "$O/$b.err"; done
tail -1 "$O/ppo.out" | cut -c1-400
