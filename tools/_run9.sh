for v in dcB dcC dcD dcE; do
  if [ $v = base ]; then unset SATMVS_B200_LIB; else export SATMVS_B200_LIB=$PWD/satmvs_b200/libsatmvs_b200_$v.so; fi
  python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s9_$v.json 2>>gpurun_out/s9_err.txt
  python bench.py --no-cpu-baseline --steps 20 --workload cfg2_casmvs > gpurun_out/s9_cas_$v.json 2>>gpurun_out/s9_err.txt
done
python - <<PY
import json
for f in ["s9_dcB","s9_dcC","s9_dcD","s9_dcE","s9_cas_dcB","s9_cas_dcC","s9_cas_dcD","s9_cas_dcE"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step", round(d["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}" for k in d["kernels"]))
PY
tail -3 gpurun_out/s9_err.txt
