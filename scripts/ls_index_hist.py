#!/usr/bin/env python
"""Where do the logsum table look-ups of the decode kernels land?  

Runs the CPU port (oracle/oracle_hmm.c built with -DORC_LS_TRACE into /tmp) on 32-read groups of the cfg2 workload,
records the table index of every logsum call in call order, lines the 32 traces up as the 32 lanes of a warp (equal read
length => identical call sequence) and reports
  * the index histogram of the calls the kernels execute (calls whose operand is -inf on all 32 lanes are the log(0)
    transitions the kernels skip statically),
  * the shared-memory wavefronts a warp-wide 4-byte gather from the 16 000-entry table needs (32 banks, same address =
    broadcast): the quantity ncu reports as bank conflicts.
Host only; writes profiles/r02_ls_index_hist.json."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from tagdust_b200 import synth
from tagdust_b200._capi import ModelDescC
from tagdust_b200.api import compile_architecture

so = "/tmp/liboracle_trace.so"
subprocess.run(["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", "-ffp-contract=off", "-DORC_LS_TRACE", "-o", so,
                os.path.join(ROOT, "oracle", "oracle_hmm.c"), "-lm", "-lpthread"], check=True)
L = C.CDLL(so)
segs, tags = bench.architecture()
desc = compile_architecture(segs, bench.background(), 150.0, 150)
groups = int(sys.argv[1]) if len(sys.argv) > 1 else 4
codes, lens, _ = synth.make_reads_fast(32 * groups, 150, tags, error_rate=0.01, random_frac=0.05, seed=11)
rng = np.random.default_rng(1)
perm = rng.permutation(len(lens))          # model reads and contaminants mixed, as in a real file
codes, lens = codes[perm], lens[perm]
cap = 2_000_000


class Out(C.Structure):
    _fields_ = [("f", C.c_float * 5), ("i", C.c_int32 * 3)]


L.orc_decode_read.argtypes = [C.POINTER(ModelDescC), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
trace_ptr = C.POINTER(C.c_ushort).in_dll(L, "orc_ls_trace")
hist = np.zeros(16001, np.int64)
wave_hist = np.zeros(33, np.int64)
n_calls = n_exec = 0
inf_lanes = far_lanes = live_lanes = 0
for g in range(groups):
    traces = []
    for r in range(32 * g, 32 * g + 32):
        buf = np.zeros(cap, np.uint16)
        C.c_void_p.in_dll(L, "orc_ls_trace").value = buf.ctypes.data
        C.c_long.in_dll(L, "orc_ls_trace_cap").value = cap
        C.c_long.in_dll(L, "orc_ls_trace_n").value = 0
        out = Out(); lab = np.zeros(160, np.uint8)
        L.orc_decode_read(C.byref(desc.c), codes[r].ctypes.data, int(lens[r]), 1, C.byref(out), lab.ctypes.data)
        n = C.c_long.in_dll(L, "orc_ls_trace_n").value
        traces.append(buf[:n].copy())
    n = min(len(t) for t in traces)
    assert all(len(t) == n for t in traces), "reads of equal length must make the same calls"
    T = np.stack(traces)                               # [32 lanes][calls]
    dead = (T == 0xFFFF).all(axis=0)                   # statically dead terms: never executed on the GPU
    E = T[:, ~dead]
    n_calls += n; n_exec += E.shape[1]
    inf_lanes += int((E == 0xFFFF).sum()); far_lanes += int((E == 0xFFFE).sum()); live_lanes += int((E < 0xFFFE).sum())
    idx = np.where(E >= 0xFFFE, 15999, E).astype(np.int64)   # the kernel clamps these to an entry that is 0
    hist += np.bincount(idx.ravel(), minlength=16001)[:16001]
    # wavefronts: per call, per bank, the number of distinct addresses; the gather needs the maximum over the banks
    order = np.argsort(idx, axis=0, kind="stable")
    srt = np.take_along_axis(idx, order, axis=0)
    first = np.ones_like(srt, bool); first[1:] = srt[1:] != srt[:-1]
    bank = srt % 32
    w = np.zeros(E.shape[1], np.int64)
    for b in range(32):
        w = np.maximum(w, ((bank == b) & first).sum(axis=0))
    wave_hist += np.bincount(w, minlength=33)[:33]
edges = [0, 100, 500, 1000, 2000, 4000, 8000, 12000, 15700, 15999, 16000]
buckets = {f"[{a},{b})": int(hist[a:b].sum()) for a, b in zip(edges, edges[1:])}
tot = int(hist.sum())
res = {"workload": "cfg2 (bench.py), reads permuted; %d warps of 32 reads" % groups,
       "logsum_calls_per_read_reference": n_calls // groups, "executed_per_read_after_static_dead_term_elimination": n_exec // groups,
       "lane_share": {"operand_minus_inf": inf_lanes / tot, "difference_ge_15.7": far_lanes / tot, "table_used": live_lanes / tot},
       "index_histogram_share": {k: v / tot for k, v in buckets.items()},
       "wavefronts_per_warp_gather": {"mean": float((wave_hist * np.arange(33)).sum() / wave_hist.sum()),
                                      "share": {str(k): float(v / wave_hist.sum()) for k, v in enumerate(wave_hist) if v}}}
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_ls_index_hist.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
