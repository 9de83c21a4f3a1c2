#!/usr/bin/env bash
# Round-2 GPU run 15 (2 GPUs): NCCL data-parallel parity (DQN / PPO with normalisers / SAC, parity-mode
# sharded sampler), bench.py at N=2 with the extras (dp_parity, PPO / SAC strong scaling), N=1 beside it.
set -u
O=gpurun_out/r2_run15
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run dist_parity 600 $TR --master-port 29511 tests/dist_parity_main.py
run bench_n2 900 $TR --master-port 29512 bench.py --gpus 2
run bench_n1 300 python bench.py --no-extra --no-cpu-baseline
run bench_n2_nobucket 300 env B200RL_GRAD_BUCKET_BYTES=0 $TR --master-port 29513 bench.py --gpus 2 --no-extra --no-cpu-baseline
cat "$O/summary.txt"
tail -5 "$O/dist_parity.out"; tail -5 "$O/dist_parity.err"
tail -1 "$O/bench_n2.out" | cut -c1-7000
tail -5 "$O/bench_n2.err"
tail -1 "$O/bench_n1.out" | cut -c1-330
tail -1 "$O/bench_n2_nobucket.out" | cut -c1-330
cat gpurun_out/dist_parity.json 2>/dev/null | cut -c1-1500
