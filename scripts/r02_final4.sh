# usage: bash scripts/r02_final4.sh <tag>  -- closing 4-GPU run: the bench line the way the driver launches it
cd /root/repo
TAG=${1:-r02f4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 > gpurun_out/${TAG}_n4.json 2> gpurun_out/${TAG}_n4.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_n4.json").read().strip().splitlines()[-1])
print("N=%d value %.2f M/s e2e %.2f M/s" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6), "frac", round(d["roofline"]["frac"],3), d["clocks"])
f=d.get("e2e_files") or {}
print(" files: %.2f M/s %.2f s" % (f.get("value",0)/1e6, f.get("seconds",0)), f.get("cold"), f.get("stage_busy_s"), f.get("extracted_matches_kernel_run"), f.get("error"), f.get("host_threads"))
print(" check:", d.get("multi_device_check"), d.get("e2e_one_context"))
PY
