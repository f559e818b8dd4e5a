cd /root/repo
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/shmem_enabled 2>/dev/null
scripts/micro/alloc_cost 2>&1 | tee gpurun_out/r02_alloc_cost2.txt
timeout 600 python -m pytest tests/test_gpu_stream_cli.py -m gpu -q -k "shorter_than or more_candidates" 2>&1 | tail -4
bash scripts/gpu_ab.sh postalways 2>&1 | tail -4
