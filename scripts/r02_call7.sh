cd /root/repo
bash scripts/r02_cli_timeline.sh 2>&1 | tail -40
TDG_TRACE= python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_b.json"))
print("value %.2f e2e %.2f files %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_files"]["value"]/1e6), d["e2e_files"]["stage_busy_s"], d["e2e_files"]["seconds"])
PY
