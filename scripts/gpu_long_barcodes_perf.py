"""Throughput with barcodes longer than 8 nt (dynamic-column kernel path) vs 6 and 8 nt (unrolled STDU path)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from tagdust_b200 import synth
from tagdust_b200.api import MODE_GET_LABEL, Context, compile_architecture
import torch
ctx = Context(device_ids=[0])
rng = np.random.default_rng(1)
for bl in (6, 8, 10, 12, 16):
    tags = ["".join(rng.choice(list("ACGT"), size=bl)) for _ in range(48)]
    desc = compile_architecture(["B:" + ",".join(tags), "R:N"], bench.background(), 150.0, 150)
    n = 148 * 512 * 3
    codes, lens, _ = synth.make_reads_fast(n, 150, tags, seed=1)
    model = ctx.model(desc, 150)
    b = ctx.batch(n, 150); b.append(codes, lens); ctx.upload(b)
    kw = dict(threshold=1.5, minlen=16, dust=100)
    ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
    ctx.profile_enable(True)
    ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
    prof = ctx.profile_read(0); ctx.profile_enable(False)
    ms = sum(v["ms"] for v in prof.values())
    cells = 2 * 150 * desc.total_columns
    print(f"barcode length {bl}: C={desc.total_columns}  {n / ms / 1e3:.2f} M reads/s  {n / ms * cells / 1e6:.0f} GCUPS  "
          + "  ".join(f"{k} {v['ms'] / v['launches']:.2f}" for k, v in prof.items()), flush=True)
    b.close(); model.close()

# a long linker in front of the barcodes: its single 24-column HMM runs the column-loop STDU path
tags = bench.architecture()[1]
linker = "ACGTTGCAGTCAGGATCCGATTCA"
desc = compile_architecture(["S:" + linker, "B:" + ",".join(tags), "R:N"], bench.background(), 150.0, 150)
n = 148 * 512 * 3
codes, lens, _ = synth.make_reads_fast(n, 150, [linker + t for t in tags], seed=1)
model = ctx.model(desc, 150)
b = ctx.batch(n, 150); b.append(codes, lens); ctx.upload(b)
kw = dict(threshold=1.5, minlen=16, dust=100)
ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
ctx.profile_enable(True)
ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
prof = ctx.profile_read(0); ctx.profile_enable(False)
ms = sum(v["ms"] for v in prof.values())
cells = 2 * 150 * desc.total_columns
print(f"S:24 + 48 barcodes: C={desc.total_columns}  {n / ms / 1e3:.2f} M reads/s  {n / ms * cells / 1e6:.0f} GCUPS  "
      + "  ".join(f"{k} {v['ms'] / v['launches']:.2f}" for k, v in prof.items()), flush=True)

# partial 5' adapters (generic segments): 7 nt (unrolled, run-time masks) and 18 nt (column loop)
for adapter in ("GGGGGGG", "AGGGAGGACGATGCGGTC"):
    from tagdust_b200.api import compile_architecture as ca
    L = len(adapter)
    desc = ca(["P:" + adapter, "B:" + ",".join(tags), "R:N"], bench.background(), 150.0, 150, five=(float(L), L * 0.85, 1.5))
    codes, lens, _ = synth.make_reads_fast(n, 150, [adapter + t for t in tags], seed=2)
    model = ctx.model(desc, 150)
    b = ctx.batch(n, 150); b.append(codes, lens); ctx.upload(b)
    ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
    ctx.profile_enable(True)
    ctx.decode_resident(model, b, MODE_GET_LABEL, **kw); torch.cuda.synchronize()
    prof = ctx.profile_read(0); ctx.profile_enable(False)
    ms = sum(v["ms"] for v in prof.values())
    cells = 2 * 150 * desc.total_columns
    print(f"P:{L} + 48 barcodes: C={desc.total_columns}  {n / ms / 1e3:.2f} M reads/s  {n / ms * cells / 1e6:.0f} GCUPS  "
          + "  ".join(f"{k} {v['ms'] / v['launches']:.2f}" for k, v in prof.items()), flush=True)
    b.close(); model.close()
