#!/bin/bash
# Round evidence on one B200 (run under gpurun from the repo root): smoke, GPU tests, every bench line, the ncu launch list of
# the default bench command, one full ncu capture of the depth-recurrence kernel and the FeatureNet launch list.
# Outputs land in gpurun_out/.
R=${1:-r02}
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/${R}_bench_stage1.json 2> gpurun_out/${R}_err.txt
for wl in cfg3_cascade cfg1_pred cfg2_casmvs cfg2_build cfg5_build cfg2_train cfg2_train_red; do
  timeout 500 python bench.py --workload $wl --no-sharded > gpurun_out/${R}_bench_$wl.json 2>> gpurun_out/${R}_err.txt
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_stage1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sharded > gpurun_out/${R}_ncu.log 2>&1
SATMVS_RED_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:red_tc_kernel -c 1 -f \
  -o gpurun_out/${R}_red_tc python tools/run_red_once.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_featurenet_launches.csv \
  python tools/run_featurenet_once.py > /dev/null 2>&1
python - <<PY
import json
for n in ("bench_stage1", "bench_cfg3_cascade", "bench_cfg1_pred", "bench_cfg2_casmvs", "bench_cfg2_build", "bench_cfg5_build", "bench_cfg2_train",
          "bench_cfg2_train_red", "bench_reference"):
    try:
        j = json.load(open(f"gpurun_out/${R}_{n}.json"))
        print(n, j.get("ms_per_step"), j.get("value"), j.get("e2e", {}).get("value"),
              [(k["class"], round(k["ms_per_step"], 3)) for k in j.get("kernels", [])][:3])
    except Exception as e:
        print(n, "ERR", e)
PY
tail -5 gpurun_out/${R}_err.txt
