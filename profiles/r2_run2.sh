#!/usr/bin/env bash
# Round-2 GPU run 2: first execution of the persistent cp.async GEMM (tc2) -- A/B against the
# first-generation kernel per layer, then the test-suite, the layer table and the bench with it.
set -u
O=gpurun_out/r2_run2
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run tc2_check 240 python profiles/tc2_check.py
CHECK=$?
tail -30 "$O/tc2_check.out"; tail -5 "$O/tc2_check.err"
if [ $CHECK -ne 0 ]; then
  echo "tc2_check failed (rc=$CHECK): remaining steps run per-op to find what works" >> "$O/summary.txt"
  for op in conv1.fwd conv1.dW conv2.fwd conv2.dX conv2.dW fc1.fwd fc1.dX fc1.dW ragged; do
    run "tc2_only_$op" 60 python profiles/tc2_check.py --only "$op" --reps 5
  done
  export B200RL_TC2=0
fi
run tc2_check_rawhi 240 python profiles/tc2_check.py --flags 1
run pytest_gpu 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider
run layer_probe 300 python profiles/layer_probe.py
run bench_tc2 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run bench_tc1 300 env B200RL_TC2=0 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run bench_tc2_pdl 300 env B200RL_PDL=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run ppo_n1 300 python profiles/ppo_bench.py
run sac_n1 300 python profiles/sac_bench.py
cat "$O/summary.txt"
tail -4 "$O/pytest_gpu.out"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-220
grep -h '"speedup"' "$O/tc2_check_rawhi.out" | cut -c1-220 | head -30
for f in bench_tc2 bench_tc1 bench_tc2_pdl; do tail -1 "$O/$f.out" | cut -c1-200; done
cat "$O/ppo_n1.out" "$O/sac_n1.out" | cut -c1-300
