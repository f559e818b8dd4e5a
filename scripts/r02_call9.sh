cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
TDG_TRACE= python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo rc=$?
grep "files\]" gpurun_out/r02_bench_c.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_c.json"))
f=d["e2e_files"]
print("value %.2f e2e %.2f files warm %.2f cold %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, f["value"]/1e6, f["cold"]["value"]/1e6), f["stage_busy_s"], f["cold"]["stage_busy_s"])
PY
