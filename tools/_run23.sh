set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 > gpurun_out/r01_bench_stage1.json 2> gpurun_out/s23_err.txt
python bench.py --steps 30 --warmup 5 --workload cfg2_build --no-cpu-baseline > gpurun_out/r01_bench_build.json 2>> gpurun_out/s23_err.txt
python bench.py --steps 30 --warmup 5 --workload cfg2_casmvs --no-cpu-baseline > gpurun_out/r01_bench_casmvs.json 2>> gpurun_out/s23_err.txt
python bench.py --steps 30 --warmup 5 --workload cfg5_build --no-cpu-baseline > gpurun_out/r01_bench_cfg5.json 2>> gpurun_out/s23_err.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/s23_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r01_launches_stage1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r01_launches_stage1.csv > gpurun_out/r01_launches_stage1.txt
ncu --set full --clock-control none --import-source on -k regex:umma_conv2d -s 6 -c 1 -o gpurun_out/r01_umma_conv3h python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
tail -3 gpurun_out/s23_err.txt
