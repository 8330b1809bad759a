set -x
timeout 300 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -5
SATMVS_SWEEP_V3=1 python tools/tune_sweep.py
for c in 0 1 2 3 4 5 6; do SATMVS_SWEEP_CFG=$c timeout 120 python tools/tune_sweep.py; done
python tools/tune_sweep.py 16 32 192 384 3 1
SATMVS_SWEEP_V3=1 python tools/tune_sweep.py 16 32 192 384 3 1
python tools/tune_sweep.py 8 8 384 768 3 1
SATMVS_SWEEP_V3=1 python tools/tune_sweep.py 8 8 384 768 3 1
python tools/tune_sweep.py 32 192 192 384 5 1
SATMVS_SWEEP_V3=1 python tools/tune_sweep.py 32 192 192 384 5 1
python tools/tune_sweep.py 32 64 96 192 3 0 pinhole
SATMVS_SWEEP_V3=1 python tools/tune_sweep.py 32 64 96 192 3 0 pinhole
