#!/usr/bin/env bash
# Round-2 GPU run 18: prefetching input pipeline in bench.py (sample of step i+1 beside train(i))
# with static / dynamic tile scheduling; replay tests for get_next(out=).
set -u
O=gpurun_out/r2_run18
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
run serial 300 python bench.py --no-extra --no-cpu-baseline --no-prefetch
run prefetch_static 300 python bench.py --no-extra --no-cpu-baseline
run prefetch_dynamic 300 env B200RL_TILE_SCHED=1 python bench.py --no-extra --no-cpu-baseline
cat "$O/summary.txt"
tail -4 "$O/pytest_gpu.out"
for f in serial prefetch_static prefetch_dynamic; do echo "$f: $(tail -1 $O/$f.out | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity',{}).get('max_rel_loss_err'), 'loss', d['final_loss'])")"; tail -2 $O/$f.err; done
