timeout 300 python -m pytest tests/test_gpu_red.py -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/s12_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/s12_launches.csv | grep -E "umma|total"
ncu --set full --clock-control none --import-source on -k regex:umma_conv2d -s 9 -c 1 -o gpurun_out/s12_umma python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
