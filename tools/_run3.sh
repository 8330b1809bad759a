timeout 300 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_sharded.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -5
ncut() {  # label, args...
  label=$1; shift
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|pack' -s 10 -c 12 --csv --log-file gpurun_out/s3_ncu_$label.csv python tools/tune_sweep.py "$@" > gpurun_out/s3_out_$label.txt 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/s3_ncu_$label.csv")) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4][:70]].append(float(r[-1].replace(',','')))
sha=[l.split('sha')[-1].strip() for l in open("gpurun_out/s3_out_$label.txt") if 'sha' in l]
print("== $label", sha, " | ".join(f"{k.split('(')[0][-40:]}: median {sorted(v)[len(v)//2]/1e3:.1f} us" for k,v in d.items()))
PY
}
export SATMVS_SWEEP_V3=1; ncut v3
unset SATMVS_SWEEP_V3
for c in 0 1 2 3 4 5; do SATMVS_SWEEP_CFG=$c ncut c$c; done
for c in 0 1 2 3; do SATMVS_SWEEP_CFG=$c ncut v5views_c$c 32 192 192 384 5 1; done
SATMVS_SWEEP_V3=1 ncut v5views_v3 32 192 192 384 5 1
ncut st2 16 32 192 384 3 1
SATMVS_SWEEP_V3=1 ncut st2_v3 16 32 192 384 3 1
ncut st3 8 8 384 768 3 1
SATMVS_SWEEP_V3=1 ncut st3_v3 8 8 384 768 3 1
ncut pin 32 64 96 192 3 0 pinhole
SATMVS_SWEEP_V3=1 ncut pin_v3 32 64 96 192 3 0 pinhole
python tools/tune_sweep.py
