set -x
cd /root/repo
N=${1:-2}
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_device 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/multi_$N.err | tee gpurun_out/multi_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e']['value'], d['kernels_ms'], d['check'])"
tail -3 gpurun_out/multi_$N.err
