set -x
cd /root/repo
cat > /tmp/lab.py <<'PY'
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, torch
from tagdust_b200 import synth
from tagdust_b200.api import MODE_GET_LABEL, Context, compile_architecture
tags = bench.architecture()[1]
adapter = "GGGGGGG"
desc = compile_architecture(["P:" + adapter, "B:" + ",".join(tags), "R:N"], bench.background(), 150.0, 150, five=(7.0, 6.0, 1.5))
n = 148 * 512
codes, lens, _ = synth.make_reads_fast(n, 150, [adapter + t for t in tags], seed=2)
ctx = Context(device_ids=[0]); model = ctx.model(desc, 150); b = ctx.batch(n, 150); b.append(codes, lens); ctx.upload(b)
for _ in range(3):
    ctx.decode_resident(model, b, MODE_GET_LABEL, threshold=1.5, minlen=16, dust=100)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:k_label -s 2 -c 1 -o gpurun_out/label_p7 python /tmp/lab.py > gpurun_out/label_p7.log 2>&1
tail -2 gpurun_out/label_p7.log
