for c in v3 0 1 3 4 5; do
  if [ $c = v3 ]; then export SATMVS_SWEEP_V3=1; else unset SATMVS_SWEEP_V3; export SATMVS_SWEEP_CFG=$c; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|pack' -s 10 -c 12 --csv --log-file gpurun_out/s2_ncu_$c.csv python tools/tune_sweep.py > /dev/null 2>&1
  echo "== $c"; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/s2_ncu_$c.csv")) if len(r)>5 and r[0].isdigit()]
import collections
d=collections.defaultdict(list)
for r in rows: d[r[4][:60]].append(float(r[-1].replace(',','')))
for k,v in d.items(): print(k, len(v), "median", sorted(v)[len(v)//2], "min", min(v))
PY
done
