cd /root/repo
mkdir -p gpurun_out
scripts/micro/alloc_cost 2>&1 | tee gpurun_out/r02_alloc_cost.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
bash scripts/r02_multi_quick.sh 2 r02q2
