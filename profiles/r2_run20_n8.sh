#!/usr/bin/env bash
# Round-2 GPU run 20 (8 GPUs): one full bench line at N = 8 (weak-scaling DQN, dp_parity, PPO / SAC
# sharded 8 ways, gather sweep) -- the command the driver's scaling run uses.
set -u
O=gpurun_out/r2_run20
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
t0=$(date +%s)
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --no-cpu-baseline > "$O/bench_n8.out" 2> "$O/bench_n8.err"
echo "bench_n8 rc=$? secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
cat "$O/summary.txt"
tail -1 "$O/bench_n8.out" | cut -c1-6000
tail -8 "$O/bench_n8.err"
