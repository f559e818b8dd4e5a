# usage: bash scripts/gpu_quick.sh <tag>  -- GPU parity tests + a short bench (8 waves per step)
set -x
cd /root/repo
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --reads $((75776*8)) --no-cpu-baseline --no-configs > gpurun_out/${TAG}_quick.json 2> gpurun_out/${TAG}_quick.err; echo rc=$?
tail -3 gpurun_out/${TAG}_quick.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_quick.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], {k:v['ms']/v['launches'] for k,v in d['kernels_ms'].items()}, d['check'])
"
