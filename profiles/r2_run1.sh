#!/usr/bin/env bash
# Round-2 GPU run 1: regression of the fused act-mask / bias-grad / PDL build, first per-layer
# GEMM table with phase stamps, ncu captures of the Q-net GEMMs, and the never-run config benches.
# Every step has its own timeout; outputs go to gpurun_out/r2_run1/.
set -u
O=gpurun_out/r2_run1
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  echo "$name rc=$? secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$O/gpu.txt" 2>&1
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run pytest_gpu 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider
run layer_probe 300 python profiles/layer_probe.py --stamps
run bench_pdl0 300 env B200RL_PDL=0 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run bench_pdl1 300 env B200RL_PDL=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run bench_nofuse 300 env B200RL_PDL=0 B200RL_FUSE_ACT_BWD=0 B200RL_FUSE_BIAS_GRAD=0 python bench.py --steps 200 --warmup 5 --no-cpu-baseline
run ncu_net 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:tc_gemm -o "$O/r2_net_v0" python profiles/net_once.py
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv \
    --log-file "$O/launches_step.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph
run ppo_n1 300 python profiles/ppo_bench.py
run sac_n1 300 python profiles/sac_bench.py
run cartpole 300 python profiles/cartpole_bench.py
run gather_sweep 300 python profiles/gather_sweep.py
cat "$O/summary.txt"
tail -3 "$O/pytest_gpu.out"
cat "$O/layer_probe.out" | head -40
tail -1 "$O/bench_pdl0.out" | cut -c1-600
tail -1 "$O/bench_pdl1.out" | cut -c1-600
tail -2 "$O/bench_pdl1.err"
