cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_ab.sh nopwin 2>&1 | tail -4
