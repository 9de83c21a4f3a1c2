#!/usr/bin/env bash
# Round-2 GPU run 16: dynamic tile scheduler (global counter + shared-memory ring) against static
# striding (flag 128).
set -u
O=gpurun_out/r2_run16
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
  echo "dynamic scheduler failed" >> "$O/summary.txt"; tail -5 "$O/tc2_check.out" "$O/tc2_check.err"
else
  run tc2_check_static 150 python profiles/tc2_check.py --flags 128
  run pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
  run bench 300 python bench.py --no-extra --no-cpu-baseline
  run bench_static 300 env B200RL_TC2_FLAGS=128 python bench.py --no-extra --no-cpu-baseline
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
a, b = load('gpurun_out/r2_run16/tc2_check.out'), load('gpurun_out/r2_run16/tc2_check_static.out')
for k in a: print(k, 'dynamic', a[k][0], 'static', b.get(k, (None,))[0], 'diff', a[k][1])
PY
tail -4 "$O/pytest_gpu.out"
for b in bench bench_static; do tail -1 "$O/$b.out" | cut -c1-330; tail -2 "$O/$b.err"; done
tail -1 "$O/ppo.out" | cut -c1-300
