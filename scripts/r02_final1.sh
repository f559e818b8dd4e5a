# usage: bash scripts/r02_final1.sh <tag>  -- the round's closing single-GPU run: the whole GPU test suite, smoke(), the default bench line
cd /root/repo
TAG=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_ref.json
python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.2f M/s e2e %.2f M/s ms/step %.1f frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"]), d["clocks"])
print(" kernels_ms", d.get("kernels_ms"))
f = d.get("e2e_files")
print(" files:", f and (round(f["value"] / 1e6, 2), f.get("cold", {}).get("value"), f.get("stage_busy_s")))
for k, v in (d.get("configs") or {}).items():
    print(" ", k, round(v.get("value", 0) / 1e6, 3), v.get("unit"), v.get("gcups"), v.get("seconds"), v.get("parity_vs_reference_run_pHMM"))
print(" cpu_baseline", d.get("cpu_baseline"))
P
