#!/usr/bin/env bash
# Round-2 GPU run 12: ncu launch lists (time + DRAM bytes) of one un-captured DQN step and of one
# 2-epoch PPO train() call.
set -u
O=gpurun_out/r2_run12
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
run dqn_step 600 ncu --metrics $M --clock-control none --profile-from-start off -c 400 --csv \
    --log-file "$O/dqn_step_launches.csv" python bench.py --ncu-step --steps 2 --warmup 3 --no-cpu-baseline --no-extra
run ppo_train 900 ncu --metrics $M --clock-control none --profile-from-start off -c 900 --csv \
    --log-file "$O/ppo_train_launches.csv" python profiles/ppo_once.py --epochs 2
cat "$O/summary.txt"
python profiles/launch_summary.py "$O/dqn_step_launches.csv" | head -120
python profiles/launch_summary.py "$O/ppo_train_launches.csv" | head -150
tail -3 "$O/dqn_step.err" "$O/ppo_train.err"
