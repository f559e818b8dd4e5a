cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo rc=$?
grep "files\]\|configs\]" gpurun_out/r02_bench_e.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_e.json"))
print("value %.3f e2e %.3f files %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_files"]["value"]/1e6), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["clocks"], {k: round(v["ms"]/v["launches"],3) for k,v in d["kernels_ms"].items()})
PY
bash scripts/gpu_ncu.sh r02c 2>&1 | tail -3
