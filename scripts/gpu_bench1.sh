set -x
cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?
tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
nproc; free -g | head -2
