#!/bin/bash
# Host-side check after the worker-pool change: stream / CLI parity tests, the host-only pipeline harness on the box's
# cores, one default bench line.
tag=${1:-r02h}
mkdir -p gpurun_out
nproc
python -m pytest tests -x -q -m gpu -k "stream or cli or dropin or gold or artifact" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
nvcc -x cu -O3 -std=c++17 -Xcompiler -O2 scripts/micro/host_pipeline.cpp -o /tmp/host_pipeline -lpthread 2>/dev/null
mkdir -p /dev/shm/hp
python - <<'P'
import numpy as np, sys
sys.path.insert(0, '.')
import bench
rng = np.random.default_rng(1)
codes = rng.integers(0, 4, (1 << 20, 150), dtype=np.uint8)
bench.write_fastq_fixed("/dev/shm/hp/in.fq", codes, 150, 16)
P
for t in $(nproc) 16 8; do for dev in 8 1; do echo "threads $t devices $dev"; /tmp/host_pipeline /dev/shm/hp/in.fq /dev/shm/hp/out $t $dev | tail -1; done; done 2>&1 | tee gpurun_out/${tag}_harness.log
rm -rf /dev/shm/hp
python bench.py --no-configs > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.loads(open("gpurun_out/${tag}_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.2f M/s e2e %.2f M/s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6))
f = d.get("e2e_files")
print(" files:", f and (f["value"] / 1e6, f.get("cold"), f.get("stage_busy_s")))
P
