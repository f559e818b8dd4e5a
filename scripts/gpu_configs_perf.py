"""Kernel-only throughput of the BASELINE configs 2-5 on one GPU (informational; bench.py reports config 2).
Prints reads/s, GCUPS (2*L*C cells per read; L*C for the backward-only architecture comparison) and per-kernel ms."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from cases import TAGS6_ED3
from tagdust_b200 import synth
from tagdust_b200.api import MODE_GET_LABEL, Context, compile_architecture

BG = bench.background()
LINKER = "ACGTTGCAGTCA"
WAVE = 148 * 512
ctx = Context(device_ids=[0])


def run(name, segs, make, waves=6):
    desc = compile_architecture(segs, BG, 150.0, 150)
    n = WAVE * waves
    codes, lens = make(n)
    model = ctx.model(desc, 150)
    b = ctx.batch(n, 150); b.append(codes, lens); ctx.upload(b)
    kw = dict(threshold=1.5, minlen=16, dust=100)
    for _ in range(2):
        ctx.decode_resident(model, b, MODE_GET_LABEL, **kw)
    import torch
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.decode_resident(model, b, MODE_GET_LABEL, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    prof = ctx.profile_read(0)
    ctx.profile_enable(False)
    cells = 2 * 150 * desc.total_columns
    print(f"{name}: H={desc.total_hmms} C={desc.total_columns}  {n / dt / 1e6:.2f} M reads/s  {n / dt * cells / 1e9:.0f} GCUPS  "
          + "  ".join(f"{k} {v['ms'] / v['launches']:.2f} ms" for k, v in prof.items()), flush=True)
    b.close(); model.close()


tags48 = TAGS6_ED3[:48]
run("cfg2 B:48 R", ["B:" + ",".join(tags48), "R:N"], lambda n: synth.make_reads_fast(n, 150, tags48, seed=1)[:2])
tags95 = TAGS6_ED3[:95]


def cfg3(n):
    # UMI(8) + linker + barcode + body, vectorised: reuse make_reads_fast with the composite head
    c, l, _ = synth.make_reads_fast(n, 150, [LINKER + t for t in tags95], seed=2)
    rng = np.random.default_rng(3)
    out = np.zeros_like(c)
    out[:, :8] = rng.integers(0, 4, size=(n, 8))
    out[:, 8:150] = c[:, :142]
    return out, l


run("cfg3 F:8 S:12 B:95 R", ["F:NNNNNNNN", "S:" + LINKER, "B:" + ",".join(tags95), "R:N"], cfg3, waves=3)
i7, i5 = TAGS6_ED3[:24], TAGS6_ED3[24:40]


def cfg4(n):
    c, l, _ = synth.make_reads_fast(n, 150, [a + b for a in i7 for b in i5], seed=4)
    return c, l


run("cfg4 B:24 B:16 R", ["B:" + ",".join(i7), "B:" + ",".join(i5), "R:N"], cfg4)

# cfg5: 64 candidate architectures x 100 000 reads, backward only
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_configs import candidate_architectures
archs = candidate_architectures(64)
descs = [compile_architecture(s, BG, 150.0, 150) for s in archs]
n = 100_000
codes, lens, _ = synth.make_reads_fast(n, 150, tags48, seed=5)
models = [ctx.model(d, 150) for d in descs]
b = ctx.batch(n, 150); b.append(codes, lens)
ctx.arch_compare(models[:2], b, 8)
t0 = time.perf_counter()
bs, post = ctx.arch_compare(models, b, 8)
dt = time.perf_counter() - t0
cells = sum(150 * d.total_columns for d in descs) * n
print(f"cfg5 64 architectures x {n} reads: {dt:.2f} s  {64 * n / dt / 1e6:.2f} M (arch,read)/s  {cells / dt / 1e9:.0f} GCUPS  best={int(np.argmax(post))}", flush=True)
