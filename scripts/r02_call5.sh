cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo rc=$?
tail -12 gpurun_out/r02_bench_a.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_a.json"))
print("value %.2f e2e %.2f files %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_files"]["value"]/1e6), d["roofline"]["frac"], d["clocks"])
print(json.dumps(d.get("configs"), indent=1)[:3000])
PY
