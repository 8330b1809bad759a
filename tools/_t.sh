timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_red.py -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload cfg2_train_red --no-sharded --steps 5 > gpurun_out/r02_bench_cfg2_train_red.json 2> gpurun_out/err_train.txt; tail -3 gpurun_out/err_train.txt
python -c "
import json; j=json.load(open('gpurun_out/r02_bench_cfg2_train_red.json')); print(j['ms_per_step'], j['value']); print(json.dumps(j.get('parity'))[-600:]); print(json.dumps(j['gpu_eager_baseline'].get('stage'))); print(j['cpu_baseline']['value']); print([(k['class'],k['launches_per_step'],round(k['ms_per_step'],3)) for k in j['kernels']])"
