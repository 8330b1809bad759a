timeout 300 python -m pytest tests/test_gpu_costreg.py tests/test_gpu_red.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -2
for v in umma noumma; do
if [ $v = noumma ]; then export SATMVS_NO_UMMA=1; fi
timeout 120 python bench.py --no-cpu-baseline --steps 20 --workload cfg2_casmvs > gpurun_out/s19_cas_$v.json 2>>gpurun_out/s19_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 45 --csv --log-file gpurun_out/s19_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg2_casmvs > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/s19_launches_$v.csv | grep -E "conv|total"
done
unset SATMVS_NO_UMMA
timeout 120 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s19_red.json 2>>gpurun_out/s19_err.txt
python - <<'PY'
import json
for f in ["s19_cas_umma","s19_cas_noumma","s19_red"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}({k['launches_per_step']:.0f})" for k in d["kernels"]))
    except Exception as e: print(f, "ERR", e)
PY
