#!/usr/bin/env bash
# Round-2 GPU run 14: thin_dw at 2 CTAs/SM with ordered in-CTA reduction, tile-based skinny forward,
# L2 prefetch of the act' mask rows in the tc2 epilogue.
set -u
O=gpurun_out/r2_run14
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
run pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run tc2_check 300 python profiles/tc2_check.py --only ppo
run ppo 300 python profiles/configs.py ppo
run sac 300 python profiles/configs.py sac
run bench 300 python bench.py --no-extra --no-cpu-baseline
run ppo_train 900 ncu --metrics $M --clock-control none --profile-from-start off -c 900 --csv \
    --log-file "$O/ppo_train_launches.csv" python profiles/ppo_once.py --epochs 2
cat "$O/summary.txt"
tail -8 "$O/pytest_gpu.out"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-160
for f in ppo sac; do tail -1 "$O/$f.out" | cut -c1-420; done
tail -1 "$O/bench.out" | cut -c1-330
python profiles/launch_summary.py "$O/ppo_train_launches.csv" | grep -v '"dram\|^  }\|^  {' | head -60 | cut -c1-160
