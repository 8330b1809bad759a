#!/bin/bash
# Round evidence on one B200 (run under gpurun from the repo root): GPU tests, every bench line, the ncu launch list of the
# default bench command and one full ncu capture of the depth-recurrence kernel.  Outputs land in gpurun_out/.
R=${1:-r01}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py > gpurun_out/${R}_bench_stage1.json 2> gpurun_out/${R}_err.txt
for wl in cfg3_cascade cfg2_casmvs cfg2_build cfg5_build; do
  timeout 200 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/${R}_bench_$wl.json 2>> gpurun_out/${R}_err.txt
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_stage1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu.log 2>&1
PYTHONPATH=. timeout 400 ncu --set full --clock-control none --import-source on -k regex:red_cluster_kernel -c 1 \
  -o gpurun_out/${R}_red_cluster python tools/run_red_once.py 2>&1 | tail -2
python - <<PY
import json
for n in ("bench_stage1", "bench_cfg3_cascade", "bench_cfg2_casmvs", "bench_cfg2_build", "bench_cfg5_build", "bench_reference"):
    try:
        j = json.load(open(f"gpurun_out/${R}_{n}.json"))
        print(n, j.get("ms_per_step"), j.get("value"), j.get("e2e", {}).get("value"),
              [(k["class"], round(k["ms_per_step"], 3)) for k in j.get("kernels", [])][:3])
    except Exception as e:
        print(n, "ERR", e)
PY
tail -5 gpurun_out/${R}_err.txt
