# usage: bash scripts/r02_final2.sh <tag>  -- closing 2-GPU run: the two multi-device tests a one-GPU box skips, then the bench line the way the driver launches it
cd /root/repo
TAG=${1:-r02f2}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "two_device or two_devices" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_n2.json").read().strip().splitlines()[-1])
print("N=%d value %.2f M/s e2e %.2f M/s" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6), "frac", round(d["roofline"]["frac"],3), d["clocks"])
f=d.get("e2e_files") or {}
print(" files: %.2f M/s %.2f s" % (f.get("value",0)/1e6, f.get("seconds",0)), f.get("cold"), f.get("stage_busy_s"), f.get("extracted_matches_kernel_run"), f.get("error"))
print(" check:", d.get("multi_device_check"), d.get("e2e_one_context"))
PY
