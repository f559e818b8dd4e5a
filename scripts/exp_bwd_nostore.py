import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tagdust_b200 import synth
from tagdust_b200.api import Context, compile_architecture, MODE_GET_LABEL
segs, tags = bench.architecture()
desc = compile_architecture(segs, bench.background(), 150.0, 150)
ctx = Context(device_ids=[0]); model = ctx.model(desc, 150)
n = 75776 * 8
codes, lens, truth = synth.make_reads_fast(n, 150, tags, seed=1)
b = ctx.batch(n, 150); b.append(codes, lens)
for rep in range(2):
    ctx.profile_enable(True)
    ctx.arch_compare([model], b, 1)
    print("no-store backward:", ctx.profile_read(0)["k_backward"])
    ctx.profile_enable(True)
    ctx.run_phmm(model, b, MODE_GET_LABEL, threshold=1.5)
    print("full:", ctx.profile_read(0))
