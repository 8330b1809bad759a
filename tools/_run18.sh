timeout 300 python -m pytest tests/test_gpu_costreg.py tests/test_gpu_red.py tests/test_gpu_cascade.py -x -q 2>&1 | tail -4
for v in umma noumma; do
if [ $v = noumma ]; then export SATMVS_NO_UMMA=1; fi
timeout 120 python bench.py --no-cpu-baseline --steps 20 --workload cfg2_casmvs > gpurun_out/s18_cas_$v.json 2>>gpurun_out/s18_err.txt
done
unset SATMVS_NO_UMMA
python - <<'PY'
import json
for f in ["s18_cas_umma","s18_cas_noumma"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), " ".join(f"{k['class']}={k['ms_per_step']:.3f}({k['launches_per_step']:.0f})" for k in d["kernels"]))
    except Exception as e: print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/s18_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload cfg2_casmvs > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/s18_launches.csv | head -24
tail -3 gpurun_out/s18_err.txt
