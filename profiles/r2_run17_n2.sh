#!/usr/bin/env bash
# Round-2 GPU run 17 (2 GPUs): gradient all-reduce overlap with the dynamic tile scheduler against
# static striding, with and without a cap on NCCL's CTAs.
set -u
O=gpurun_out/r2_run17
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
B="bench.py --gpus 2 --no-extra --no-cpu-baseline"
run graph_test 300 python -m pytest tests/test_dqn_gpu.py -m gpu -q -p no:cacheprovider
run n1 300 python bench.py --no-extra --no-cpu-baseline
run n2_dyn 300 $TR --master-port 29521 $B
run n2_static 300 env B200RL_TC2_FLAGS=128 $TR --master-port 29522 $B
run n2_dyn_cta8 300 env NCCL_MAX_CTAS=8 $TR --master-port 29523 $B
run n2_static_cta8 300 env NCCL_MAX_CTAS=8 B200RL_TC2_FLAGS=128 $TR --master-port 29524 $B
run n2_dyn_cta4 300 env NCCL_MAX_CTAS=4 $TR --master-port 29525 $B
run n2_dyn_nobucket 300 env B200RL_GRAD_BUCKET_BYTES=0 $TR --master-port 29526 $B
cat "$O/summary.txt"
tail -2 "$O/graph_test.out"
for f in n1 n2_dyn n2_static n2_dyn_cta8 n2_static_cta8 n2_dyn_cta4 n2_dyn_nobucket; do echo "$f: $(tail -1 $O/$f.out | cut -c1-260)"; done
