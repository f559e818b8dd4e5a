set -x
cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30
