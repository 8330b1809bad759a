python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s7_scale2.json 2> gpurun_out/s7_err.txt
for g in none nccl fused; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg4_sharded192 --gather $g > gpurun_out/s7_sharded_$g.json 2>> gpurun_out/s7_err.txt
done
python bench.py --steps 10 --warmup 3 --workload cfg4_sharded192 --gather none --no-cpu-baseline > gpurun_out/s7_sharded_1gpu.json 2>> gpurun_out/s7_err.txt
timeout 300 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -2
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s7_*.json")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "ERR", e); continue
    print(f, d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "value %.3e"%d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d["scaling"])
PY
tail -5 gpurun_out/s7_err.txt
