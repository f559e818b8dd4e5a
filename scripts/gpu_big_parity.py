"""One-off large parity run: GPU vs the oracle (plain-C restatement, bit-identical to the reference) on
N cfg2 reads; counts differing bits in every per-read output.  Usage: python scripts/gpu_big_parity.py [N]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from refharness import Oracle
from tagdust_b200 import synth
from tagdust_b200.api import MODE_GET_LABEL, Context, compile_architecture
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
segs, tags = bench.architecture()
desc = compile_architecture(segs, bench.background(), 150.0, 150)
codes, lens, truth = synth.make_reads_fast(n, 150, tags, error_rate=0.01, random_frac=0.05, seed=4242)
ctx = Context(device_ids=[0]); model = ctx.model(desc, 150); b = ctx.batch(n, 150); b.append(codes, lens)
kw = dict(threshold=1.5, minlen=16, dust=100)
gpu = ctx.run_phmm(model, b, MODE_GET_LABEL, **kw)
t0 = time.time()
ora = Oracle().run(desc, MODE_GET_LABEL, codes, lens, threads=os.cpu_count() or 8, **kw)
print(f"oracle: {n} reads in {time.time() - t0:.1f} s")
bad = {}
for k in ("b_score", "f_score", "r_score", "bar_prob", "mapq"):
    bad[k] = int((gpu[k].view(np.uint32) != ora[k].view(np.uint32)).sum())
for k in ("read_type", "barcode", "fingerprint"):
    bad[k] = int((gpu[k] != ora[k]).sum())
bad["labels"] = int((gpu["labels"][:, :151] != ora["labels"][:, :151]).any(axis=1).sum())
print("reads:", n, "mismatching reads per output:", bad)
