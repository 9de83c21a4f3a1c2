#!/usr/bin/env bash
# Round-2 GPU run 3: full test-suite on the tc2 build (skinny heads, vectorised RMSProp, raw-hi
# default), the rewritten bench.py end to end, and an ncu capture of the tc2 kernels.
set -u
O=gpurun_out/r2_run3
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
run layer_probe 300 python profiles/layer_probe.py
run bench 600 python bench.py
run bench_noextra_pdl 300 env B200RL_PDL=1 python bench.py --no-extra --no-cpu-baseline
run ncu_net 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:tc2_gemm -o "$O/r2_net_tc2" python profiles/net_once.py
run launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv \
    --log-file "$O/launches_step.csv" python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-graph --no-extra
cat "$O/summary.txt"
tail -15 "$O/pytest_gpu.out"
cat "$O/layer_probe.out" | cut -c1-120
tail -1 "$O/bench.out" | cut -c1-3000
tail -3 "$O/bench.err"
tail -1 "$O/bench_noextra_pdl.out" | cut -c1-300
