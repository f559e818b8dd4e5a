# what the driver runs at round end, on a 2-GPU box: reference arm, N=1 and N=2 of the bench with the driver's step counts
cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_drv_ref.json 2> gpurun_out/r02_drv_ref.err; echo "ref rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_drv_n1.json 2> gpurun_out/r02_drv_n1.err; echo "n1 rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_drv_n2.json 2> gpurun_out/r02_drv_n2.err; echo "n2 rc=$? $(( $(date +%s) - T0 )) s"
python - <<PY
import json
for f in ("ref", "n1", "n2"):
    d = json.load(open(f"gpurun_out/r02_drv_{f}.json"))
    print(f, d.get("impl"), "value %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], d.get("clocks"), "launches", d.get("gpu_launches"))
    if "roofline" in d: print("   roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "frac_algorithmic", "traffic")}, "files", (d.get("e2e_files") or {}).get("value"), (d.get("e2e_files") or {}).get("cold", {}).get("value"))
    if "configs" in d: print("   configs", {k: (v.get("value"), v.get("parity_vs_reference_run_pHMM")) for k, v in d["configs"].items() if isinstance(v, dict)})
    if "multi_device_check" in d: print("   ", d["multi_device_check"], d["e2e_one_context"]["value"])
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"])
PY
