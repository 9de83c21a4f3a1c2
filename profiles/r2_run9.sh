#!/usr/bin/env bash
# Round-2 GPU run 9: role ablations of the tc2 kernel (flags 8 = no loads, 16 = no conversion,
# 32 = no MMA issue): which role bounds the K-block period of each layer?
set -u
O=gpurun_out/r2_run9
mkdir -p "$O"
for f in 0 8 16 32 24 40 48 56; do
  timeout 200 python profiles/tc2_check.py --flags $f > "$O/ablate_$f.out" 2> "$O/ablate_$f.err"
  echo "flags=$f rc=$?" >> "$O/summary.txt"
done
cat "$O/summary.txt"
python - <<'PY'
import json, glob, collections
rows = collections.OrderedDict()
for f in (0, 8, 16, 32, 24, 40, 48, 56):
  for l in open(f'gpurun_out/r2_run9/ablate_{f}.out'):
    d = json.loads(l)
    if 'tc2_us' in d:
      rows.setdefault((d['layer'], d['op']), {})[f] = d['tc2_us']
print('layer op | full noload noconv nomma | noload+noconv noload+nomma noconv+nomma none')
for k, v in rows.items():
  print(k[0], k[1], '|', ' '.join(str(v.get(f)) for f in (0, 8, 16, 32)), '|', ' '.join(str(v.get(f)) for f in (24, 40, 48, 56)))
PY
