timeout 200 python -m pytest tests/test_gpu_red.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -3
timeout 120 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s16.json 2>gpurun_out/s16_err.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s16.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}({k['launches_per_step']:.0f})" for k in d["kernels"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/s16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/s16_launches.csv | grep -E "umma|direct|total"
