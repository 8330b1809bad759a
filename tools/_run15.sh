timeout 200 python -m pytest tests/test_gpu_red.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -4
timeout 120 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s15_fuse.json 2>gpurun_out/s15_err.txt
SATMVS_RED_NO_FUSE=1 timeout 120 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s15_nofuse.json 2>>gpurun_out/s15_err.txt
python - <<'PY'
import json
for f in ["s15_fuse","s15_nofuse"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}({k['launches_per_step']:.0f})" for k in d["kernels"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/s15_err.txt
