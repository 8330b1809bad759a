for v in new old new old; do
if [ $v = old ]; then export SATMVS_B200_LIB=$PWD/satmvs_b200/libsatmvs_b200_old.so; else unset SATMVS_B200_LIB; fi
timeout 120 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s22_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/s22_$v.json").read().strip().splitlines()[-1])
print("$v", "ms/step", round(d["ms_per_step"],4), " ".join("%s=%.3f(%d)"%(k["class"],k["ms_per_step"],k["launches_per_step"]) for k in d["kernels"]))
PY
done
