timeout 300 python -m pytest tests/test_gpu_red.py -x -q 2>&1 | tail -2
for v in 3 2; do
if [ $v = 2 ]; then export SATMVS_UMMA_2CTA=1; fi
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/s13_launches_$v.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/s13_launches_$v.csv | grep -E "umma_conv|total"
done
