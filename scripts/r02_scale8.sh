# usage: bash scripts/r02_scale8.sh <tag> [files_reads] ["8 1"]  -- bench at 8 GPUs and at 1 GPU on the same box (strong-scaling file job with stage trace)
cd /root/repo
TAG=${1:-r02s8}; FR=${2:-32000000}; NS=${3:-8 1}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1; df -h /dev/shm | tail -1
for n in $NS; do
	if [ $n -gt 1 ]; then
		TDG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 --files-reads $FR > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
	else
		TDG_TRACE=1 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --files-reads $FR --no-cpu-baseline --no-configs > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
	fi
	echo "rc=$?"; grep -v trace gpurun_out/${TAG}_n$n.err | grep -v convert_chunk | tail -3
	python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_n$n.json"))
print("N=%d value %.2f M/s e2e %.2f M/s" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6), "frac", round(d["roofline"]["frac"],3), d["clocks"])
f=d.get("e2e_files") or {}
print(" files: %.2f M/s %.2f s" % (f.get("value",0)/1e6, f.get("seconds",0)), f.get("stage_busy_s"), f.get("extracted_matches_kernel_run"), f.get("error"), f.get("host_threads"))
print(" check:", d.get("multi_device_check"), d.get("e2e_one_context"))
PY
done
