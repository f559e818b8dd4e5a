cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_linker or larger_than_shared or spans" 2>&1 | tail -6
bash scripts/r02_cli.sh 2>&1 | tail -40
bash scripts/gpu_ab.sh 2>&1 | tail -3
