timeout 300 python -m pytest tests/test_gpu_red.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -15
timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s11_umma.json 2>gpurun_out/s11_err.txt
SATMVS_NO_UMMA=1 timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s11_noumma.json 2>>gpurun_out/s11_err.txt
python - <<PY
import json
for f in ["s11_umma","s11_noumma"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}({k['launches_per_step']:.0f})" for k in d["kernels"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/s11_err.txt
