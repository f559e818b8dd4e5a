# usage: bash scripts/r02_multi.sh <N> <tag> [files_reads]  -- GPU tests (incl. the 2-device ones), then bench at N GPUs and at 1 GPU
cd /root/repo
N=${1:-2}; TAG=${2:-r02m}; FR=${3:-16000000}
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2; df -h /dev/shm | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() {  # run <n>
	local n=$1
	if [ $n -gt 1 ]; then
		timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 --files-reads $FR > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
	else
		TDG_TRACE= timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --files-reads $FR > gpurun_out/${TAG}_n$n.json 2> gpurun_out/${TAG}_n$n.err
	fi
	echo "rc=$?"; tail -4 gpurun_out/${TAG}_n$n.err
	python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_n$n.json"))
print("N=%d value %.2f M/s e2e %.2f M/s" % (d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6), "frac", round(d["roofline"]["frac"],3), d["clocks"])
print(" files:", json.dumps(d.get("e2e_files")))
print(" check:", d.get("multi_device_check"), d.get("cpu_baseline"))
PY
}
run $N
[ $N -gt 1 ] && run 1
