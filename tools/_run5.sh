timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s5_bench_pdl.json 2>gpurun_out/s5_err.txt
SATMVS_RED_NO_PDL=1 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s5_bench_nopdl.json 2>>gpurun_out/s5_err.txt
python bench.py --no-cpu-baseline --steps 20 --workload cfg2_build > gpurun_out/s5_bench_build.json 2>>gpurun_out/s5_err.txt
python - <<PY
import json
for f in ["s5_bench_pdl","s5_bench_nopdl","s5_bench_build"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "launches", d["gpu_launches"])
    for k in d["kernels"]: print("   ", k["class"], k["launches_per_step"], round(k["ms_per_step"],4), k.get("roofline",{}).get("frac"))
PY
tail -3 gpurun_out/s5_err.txt
