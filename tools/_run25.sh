for v in vec scalar vec scalar; do
if [ $v = scalar ]; then export SATMVS_RED_NO_VEC4=1; else unset SATMVS_RED_NO_VEC4; fi
timeout 120 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/s25_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/s25_$v.json").read().strip().splitlines()[-1])
print("$v", "ms/step", round(d["ms_per_step"],4), " ".join("%s=%.3f(%d)"%(k["class"],k["ms_per_step"],k["launches_per_step"]) for k in d["kernels"]))
PY
done
