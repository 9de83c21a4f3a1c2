#!/usr/bin/env bash
# Round-2 GPU run 10: consolidation of the tc2 v4 build (WIDE B operand, 8 loader + 8 converter
# warps, per-launch 4/8 epilogue warps, PDL on by default): full GPU suite, per-layer A/B, the
# full bench line, ncu launch list with DRAM bytes of one un-captured step, ncu --set full of the
# GEMM kernels of one forward+backward.
set -u
O=gpurun_out/r2_run10
mkdir -p "$O"
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> "$O/summary.txt"
  return $rc
}
run pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
run tc2_check 300 python profiles/tc2_check.py
run bench 900 python bench.py
run bench_pdl0 300 env B200RL_PDL=0 python bench.py --no-extra --no-cpu-baseline
run launches 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 400 -c 200 --csv --log-file "$O/launches_step.csv" \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-graph --no-extra
run ncu_net 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:tc2_gemm -o "$O/r2_net_tc2_v4" python profiles/net_once.py
cat "$O/summary.txt"
tail -8 "$O/pytest_gpu.out"
grep -h '"speedup"' "$O/tc2_check.out" | cut -c1-200
tail -1 "$O/bench.out" | cut -c1-6000
tail -3 "$O/bench.err"
tail -1 "$O/bench_pdl0.out" | cut -c1-330
tail -3 "$O/launches.err" "$O/ncu_net.err"
ls -la "$O"
