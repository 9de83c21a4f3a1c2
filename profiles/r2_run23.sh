#!/usr/bin/env bash
# Round-2 GPU run 23: neighbour pre-sum v2 (own kernel instantiation EPI_COL2IM_MERGE, act' mask
# loads issued before the accumulator reads): per-layer dX timing and bench A/B, then the FULL GPU
# suite with the pre-sum on (the configuration that becomes the default if it wins).
set -u
O=gpurun_out/r2_run23
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run check_merge 120 env B200RL_COL2IM_MERGE=1 python profiles/tc2_check.py --only dX
run check_plain 120 env B200RL_COL2IM_MERGE=0 python profiles/tc2_check.py --only dX
run bench_merge 200 env B200RL_COL2IM_MERGE=1 python bench.py --no-extra --no-cpu-baseline
run bench_plain 200 env B200RL_COL2IM_MERGE=0 python bench.py --no-extra --no-cpu-baseline
run pytest_merge 400 env B200RL_COL2IM_MERGE=1 python -m pytest tests -m gpu -q -p no:cacheprovider
cat "$O/summary.txt"
echo merge; grep -h '"speedup"' "$O/check_merge.out" | grep conv | cut -c1-160
echo plain; grep -h '"speedup"' "$O/check_plain.out" | grep conv | cut -c1-160
for f in bench_merge bench_plain; do echo "$f: $(tail -1 $O/$f.out | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['final_loss'], 'frac', d['roofline']['frac'])")"; tail -2 $O/$f.err; done
tail -5 "$O/pytest_merge.out"
