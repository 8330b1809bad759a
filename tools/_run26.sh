for v in 0 3 1 2 0; do
export SATMVS_RED_EARLY_WAIT=$v
timeout 120 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/s26_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/s26_$v.json").read().strip().splitlines()[-1])
print("early_wait=$v", "ms/step", round(d["ms_per_step"],4))
PY
done
