#!/usr/bin/env python
"""bench.py -- throughput of the TagDust2 per-read HMM decode path on B200.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
    python bench.py --impl reference ...                   (the reference's CPU run_pHMM)

Workload (BASELINE.json configs[1], "cfg2"): synthetic single-end 150 nt reads, architecture
`-1 B:<48 six-nt barcodes of EDITTAG_6nt_ed_3> -2 R:N`, 1 % substitution errors in the
barcode, 5 % uniform-random contaminant reads; MODE_GET_LABEL (backward + forward/posterior
+ label DP + Q + extraction + dust).  A step = one pass of the hot path over one batch.

Prints ONE JSON line on rank 0 (contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tagdust_b200 import synth  # noqa: E402

READ_LEN = 150
N_BARCODES = 48
TAGS = os.path.join(ROOT, "tests", "golden", "edittag_6nt_ed3.txt")
THRESHOLD = 1.5          # a fixed -Q so that every step does the same work (no calibration phase)
OPS_PER_COLPOS_BWD = 8 * 9 + 18   # 8 logsum (9 FP32-pipe ops each) + 18 adds per column-position
OPS_PER_COLPOS_FWD = 10 * 9 + 17  # forward/posterior: 10 logsum + 17 adds


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def architecture():
    tags = synth.load_tags(TAGS, N_BARCODES)
    return ["B:" + ",".join(tags), "R:N"], tags


def background():
    """ssi->background for uniform A/C/G/T with the +1 pseudocounts of io.c:79-81 on ~1M reads."""
    counts = np.array([37.5e6 + 1, 37.5e6 + 1, 37.5e6 + 1, 37.5e6 + 1, 1.0])
    s = counts.sum()
    return np.array([float(np.float32(np.log(float(np.float32(c / s))))) for c in counts])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (one wave of 75 776 reads) of the dominant
    kernel, from the committed `ncu --set full` capture (profiles/ncu_traffic.json names the .ncu-rep)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None, {}
    with open(p) as fh:
        d = json.load(fh)
    k = d.get("kernels", {}).get(kernel)
    return (k["dram_bytes_read"] + k["dram_bytes_write"] if k else None), d.get("source"), (k or {})


def fastq_block(codes, read_len):
    """Fixed-width FASTQ records (name, 150 nt, '+', qualities) as one uint8 matrix, built with numpy (no Python loop)."""
    n = codes.shape[0]
    name_w = 11
    rec = np.empty((n, 1 + name_w + 1 + read_len + 3 + read_len + 1), np.uint8)
    rec[:, 0] = ord("@")
    idx = np.arange(n)
    rec[:, 1] = ord("r")
    for k in range(name_w - 1):
        rec[:, 1 + name_w - 1 - k] = ord("0") + (idx // 10 ** k) % 10
    o = 1 + name_w
    rec[:, o] = ord("\n"); o += 1
    rec[:, o:o + read_len] = np.frombuffer(b"ACGTN", np.uint8)[codes[:, :read_len]]; o += read_len
    rec[:, o:o + 3] = np.frombuffer(b"\n+\n", np.uint8); o += 3
    rec[:, o:o + read_len] = ord("I"); o += read_len
    rec[:, o] = ord("\n")
    return rec


def write_fastq_fixed(path, codes, read_len, repeats=1):
    rec = fastq_block(codes, read_len)
    with open(path, "wb") as fh:
        for _ in range(repeats):
            rec.tofile(fh)


def files_e2e(ctx, model, tags, codes, n_reads, threads, n_dev, expect_extracted_per_block=None):
    """FASTQ file -> tdg_demux_run (reader, pack, GPU, extraction, demultiplexed FASTQ files), the streaming layer of
    SURVEY 8(f) rank 1, on ONE context over `n_dev` devices: a fixed job, so the N = 1, 2, 4, 8 values form a
    strong-scaling curve.  The file is a block of synthetic cfg2 reads written `repeats` times into /dev/shm."""
    import shutil
    import tempfile
    from tagdust_b200 import stream
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="tdg_bench_", dir=shm)
    try:
        block = min(codes.shape[0], n_reads)
        rec_bytes = 1 + 11 + 1 + READ_LEN + 3 + READ_LEN + 1
        repeats = max(1, -(-n_reads // block))
        free = shutil.disk_usage(tmp).free
        while repeats > 1 and 2.2 * repeats * block * rec_bytes > 0.8 * free:   # input + output must fit
            repeats -= 1
        total = block * repeats
        fq = os.path.join(tmp, "in.fq")
        t0 = time.perf_counter()
        write_fastq_fixed(fq, codes[:block], READ_LEN, repeats)
        log(f"[files] wrote {total} reads ({os.path.getsize(fq) / 1e9:.2f} GB) in {time.perf_counter() - t0:.1f}s")
        runs = []
        for tag in ("cold", "warm"):   # cold: first job of the context (pins its staging batches); warm: the context's pool is filled
            for f in os.listdir(tmp):
                if f.startswith("out"):
                    os.remove(os.path.join(tmp, f))
            t0 = time.perf_counter()
            st = stream.demux_run(ctx, [dict(path=fq, model=model, num_read_segments=1, threshold=THRESHOLD, max_seq_len=READ_LEN)],
                                  os.path.join(tmp, "out"), barcode_input=0, barcode_names=list(tags), minlen=16, dust=100, threads=threads)
            dt = time.perf_counter() - t0
            runs.append((tag, dt, st))
            log(f"[files] {tag}: {total / dt / 1e6:.2f} M reads/s ({dt:.2f} s)")
        out_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp) if f.startswith("out"))
        tag, dt, st = runs[1]
        out = {"value": total / dt, "unit": "reads/s", "reads": total, "seconds": dt, "host_threads": threads, "n_devices": n_dev,
               "scaling": "strong (fixed job, one tdg_context over n_devices)",
               "run": "second job on the same context (staging batches come from the context's pool); the first, cold job is in `cold`",
               "cold": {"value": total / runs[0][1], "seconds": runs[0][1],
                        "stage_busy_s": {k: runs[0][2][k] for k in ("seconds_split", "seconds_parse", "seconds_gpu_wait", "seconds_write")}},
               "input_bytes": os.path.getsize(fq), "output_bytes": out_bytes,
               "stage_busy_s": {k: st[k] for k in ("seconds_split", "seconds_parse", "seconds_gpu_wait", "seconds_write")},
               "extracted": st["num_EXTRACT_SUCCESS"], "total_read": st["total_read"],
               "what": "tdg_demux_run: FASTQ file in /dev/shm -> 49 demultiplexed FASTQ files, byte-identical format to print_all; "
                       "the GPU returns R-run spans, label rows stay on the device"}
        if expect_extracted_per_block is not None:
            out["extracted_matches_kernel_run"] = bool(st["num_EXTRACT_SUCCESS"] == expect_extracted_per_block * repeats
                                                       and st["total_read"] == total)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def other_configs(ctx, threads, quick=False):
    """BASELINE.json configs 3, 4, 5 and one paired-end file run on this GPU (the headline line stays cfg2): kernel-only
    reads/s and GCUPS with CUDA events, the live-op roofline fraction of the dominant kernel, and a parity count of a read
    sample against the UNMODIFIED reference's run_pHMM (oracle/_ref, the checker -- nothing timed here runs on it)."""
    import torch
    from tagdust_b200 import stream
    from tagdust_b200.api import MODE_GET_LABEL, compile_architecture, live_ops
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness
    R = refharness.RefHarness() if refharness.have_ref() else None
    tags_all = synth.load_tags(TAGS)
    pk, _ = peaks()
    fp32_peak = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6
    WAVE = 148 * 512
    kw = dict(threshold=THRESHOLD, minlen=16, dust=100)
    out = {}

    def label_config(name, what, segs, codes, lens, waves):
        n = WAVE * waves
        desc = compile_architecture(segs, background(), float(READ_LEN), READ_LEN)
        model = ctx.model(desc, READ_LEN)
        b = ctx.batch(n, READ_LEN)
        b.append(codes[:n], lens[:n])
        ctx.upload(b)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            ctx.decode_resident(model, b, MODE_GET_LABEL, stream=st, **kw)
        torch.cuda.synchronize()
        ctx.profile_enable(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ev0.record()
        for _ in range(reps):
            ctx.decode_resident(model, b, MODE_GET_LABEL, stream=st, **kw)
        ev1.record()
        torch.cuda.synchronize()
        dt = ev0.elapsed_time(ev1) / 1000.0 / reps
        prof = ctx.profile_read(0)
        ctx.profile_enable(False)
        res = ctx.download(b)
        lo = live_ops(desc)
        live = {"k_backward": 9 * lo["ls_bwd"] + lo["add_bwd"], "k_forward": 9 * lo["ls_fwd"] + lo["add_fwd"], "k_label": 0.0}
        dom = max(prof, key=lambda k: prof[k]["ms"])
        dom_s = prof[dom]["ms"] / 1000.0 / reps
        cells = 2 * READ_LEN * desc.total_columns
        rec = {"what": what, "hmms": desc.total_hmms, "columns": desc.total_columns, "reads": n, "value": n / dt, "unit": "reads/s",
               "gcups": n / dt * cells / 1e9, "cells_per_read": cells,
               "kernels_ms_per_wave": {k: v["ms"] / v["launches"] for k, v in prof.items()},
               "roofline": {"kernel": dom, "frac": n * READ_LEN * live[dom] / dom_s / fp32_peak,
                            "live_logsums_per_hmm_position": (lo["ls_fwd"] if dom == "k_forward" else lo["ls_bwd"]) / desc.total_hmms},
               "read_type_counts": np.bincount(res["read_type"], minlength=7).tolist()}
        if R is not None:
            k = 600 if quick else 1500
            idx = np.linspace(0, n - 1, k).astype(np.int64)
            p = R.param_new(segs, threshold=THRESHOLD, minlen=16, dust=100, threads=threads)
            mb = R.model_new(p, background=background(), average_length=float(READ_LEN), max_seq_len=READ_LEN)
            want = R.run_phmm(mb, p, 1, codes[idx], lens[idx])
            R.model_free(mb); R.param_free(p)
            bad = res["mapq"][idx].view(np.uint32) != want["mapq"].view(np.uint32)
            for key in ("read_type", "barcode", "fingerprint"):
                bad |= res[key][idx] != want[key]
            bad |= (res["labels"][idx][:, :READ_LEN + 1] != want["labels"][:, :READ_LEN + 1]).any(axis=1)
            rec["parity_vs_reference_run_pHMM"] = {"reads": int(k), "mismatches": int(bad.sum()),
                                                   "compared": "mapq bits, read_type, barcode, fingerprint, labels"}
        b.close(); model.close()
        out[name] = rec
        log(f"[configs] {name}: {rec['value'] / 1e6:.2f} M reads/s, {rec['gcups']:.0f} GCUPS, parity {rec.get('parity_vs_reference_run_pHMM')}")
        return desc

    waves = 2 if quick else 4
    segs3, tags3, codes3, lens3, _ = synth.cfg3_workload(tags_all, WAVE * waves)
    label_config("cfg3", "read 1 of the paired-end UMI run: -1 F:NNNNNNNN -2 S:<12 nt> -3 B:<95 x 6 nt> -4 R:N, 150 nt", segs3, codes3, lens3, waves)
    segs4, _, codes4, lens4, _ = synth.cfg4_workload(tags_all, WAVE * waves)
    label_config("cfg4", "dual index on read 1: -1 B:<24 I7> -2 B:<16 I5> -3 R:N (384 combinations), 150 nt", segs4, codes4, lens4, waves)

    # cfg5: library-prep auto-detection, 64 candidate architectures x 100 000 reads, backward only (MODE_ARCH_COMP)
    archs = synth.candidate_architectures(tags_all, 16 if quick else 64)
    descs = [compile_architecture(a, background(), float(READ_LEN), READ_LEN) for a in archs]
    n5 = 20_000 if quick else 100_000
    _, tags2 = architecture()
    codes5, lens5, _ = synth.make_reads_fast(n5, READ_LEN, tags2, error_rate=0.01, random_frac=0.05, seed=5)
    models = [ctx.model(d, READ_LEN) for d in descs]
    b = ctx.batch(n5, READ_LEN)
    b.append(codes5, lens5)
    ctx.arch_compare(models[:2], b, threads)
    t0 = time.perf_counter()
    _, post = ctx.arch_compare(models, b, threads)
    dt = time.perf_counter() - t0
    cells5 = float(sum(READ_LEN * d.total_columns for d in descs)) * n5
    rec = {"what": f"{len(archs)} candidate architectures x {n5} reads, backward() of every read under every architecture, "
                   f"posteriors summed in {threads} reference thread slices (host arrays in, posteriors out)",
           "seconds": dt, "value": len(archs) * n5 / dt, "unit": "(architecture, read) pairs/s", "gcups": cells5 / dt / 1e9,
           "best_architecture": int(np.argmax(post)), "true_architecture": 5}
    b.close()
    if R is not None:
        k = 300 if quick else 1000
        idx = np.linspace(0, n5 - 1, k).astype(np.int64)
        bs = ctx.batch(k, READ_LEN)
        bs.append(codes5[idx], lens5[idx])
        _, post_s = ctx.arch_compare(models, bs, threads)
        bs.close()
        p = R.param_new(archs[0], threads=threads)
        mbs = []
        for a in archs:
            pa = R.param_new(a, threads=threads)
            mbs.append((R.model_new(pa, background=background(), average_length=float(READ_LEN), max_seq_len=READ_LEN), pa))
        want = R.run_arch_comp([m for m, _ in mbs], p, codes5[idx], lens5[idx])
        for m, pa in mbs:
            R.model_free(m); R.param_free(pa)
        R.param_free(p)
        rec["parity_vs_reference_run_pHMM"] = {"reads": int(k), "architectures": len(archs),
                                               "posterior_bit_mismatches": int((post_s.view(np.uint32) != want.view(np.uint32)).sum()),
                                               "same_best": bool(int(np.argmax(post_s)) == int(np.argmax(want)))}
    for m in models:
        m.close()
    out["cfg5"] = rec
    log(f"[configs] cfg5: {rec['seconds']:.2f} s, {rec['gcups']:.0f} GCUPS, {rec.get('parity_vs_reference_run_pHMM')}")

    # paired-end through the streaming demultiplexer: read 1 = cfg3 architecture, read 2 = R:N (run_rna_dust path, host)
    import shutil
    import tempfile
    npair = (300_000 if quick else 2_000_000)
    tmp = tempfile.mkdtemp(prefix="tdg_bench_pe_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        blk = min(npair, codes3.shape[0])
        reps = max(1, npair // blk)
        write_fastq_fixed(os.path.join(tmp, "r1.fq"), codes3[:blk], READ_LEN, reps)
        rng = np.random.default_rng(9)
        r2 = rng.integers(0, 4, size=(blk, READ_LEN + 1), dtype=np.uint8)
        write_fastq_fixed(os.path.join(tmp, "r2.fq"), r2, READ_LEN, reps)
        desc3 = compile_architecture(segs3, background(), float(READ_LEN), READ_LEN)
        model3 = ctx.model(desc3, READ_LEN)
        t0 = time.perf_counter()
        st = stream.demux_run(ctx, [dict(path=os.path.join(tmp, "r1.fq"), model=model3, num_read_segments=1, threshold=THRESHOLD, max_seq_len=READ_LEN),
                                    dict(path=os.path.join(tmp, "r2.fq"), model=None, num_read_segments=1, threshold=0.0, max_seq_len=READ_LEN)],
                              os.path.join(tmp, "out"), barcode_input=0, barcode_names=list(tags3), minlen=16, dust=100, threads=threads)
        dt = time.perf_counter() - t0
        model3.close()
        out["paired_end_files"] = {"what": "2 x 150 nt FASTQ files in /dev/shm -> 2 x 96 demultiplexed files through tdg_demux_run; read 1 carries "
                                           "UMI + linker + 95 barcodes (cfg3), read 2 is R:N",
                                   "pairs": int(blk * reps), "seconds": dt, "value": blk * reps / dt, "unit": "read pairs/s",
                                   "extracted": st["num_EXTRACT_SUCCESS"], "total_read": st["total_read"],
                                   "stage_busy_s": {k: st[k] for k in ("seconds_split", "seconds_parse", "seconds_gpu_wait", "seconds_write")}}
        log(f"[configs] paired-end files: {blk * reps / dt / 1e6:.2f} M pairs/s")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def one_context_e2e(ctxN, modelN, codes, lens, kw, steps, n_dev):
    """The C-ABI call with host arrays (pack -> H2D -> kernels -> D2H, double buffered) on ONE context that shards every
    batch over all devices (tdg_plan_shards), driven by one host thread: the library's own multi-device path."""
    from tagdust_b200.api import MODE_GET_LABEL
    reps = max(1, min(n_dev, 4))            # a larger batch per step so that every device gets several waves
    big_codes = np.concatenate([codes] * reps) if reps > 1 else codes
    big_lens = np.concatenate([lens] * reps) if reps > 1 else lens
    n = big_codes.shape[0]
    os.environ["TDG_PACK_THREADS"] = str(max(4, host_threads() // 2))
    bs = [ctxN.batch(n, READ_LEN) for _ in range(2)]

    def run(k):
        pending = None
        for s in range(k):
            b = bs[s % 2]
            b.clear()
            b.append(big_codes, big_lens)
            ctxN.submit(modelN, b, MODE_GET_LABEL, **kw)
            if pending is not None:
                ctxN.wait(pending, copy=False)
            pending = b
        ctxN.wait(pending, copy=False)

    run(2)
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    for b in bs:
        b.close()
    os.environ.pop("TDG_PACK_THREADS", None)
    return {"value": n * steps / dt, "unit": "reads/s", "n_devices": n_dev, "reads_per_step": int(n), "steps": steps,
            "pack_threads": max(4, host_threads() // 2),
            "what": "tdg_batch_append_codes + tdg_submit + tdg_wait on one tdg_context over all devices, one host thread, host arrays"}


def multi_device_check(ctx1, model1, ctxN, modelN, codes, lens, kw, n):
    """The same reads through a one-device context and through one context sharded over all devices
    (tdg_plan_shards: contiguous, tile-aligned shards; output order = input order): every output must have the same bits."""
    from tagdust_b200.api import MODE_GET_LABEL
    n = min(n, codes.shape[0])
    out = []
    for ctx, model in ((ctx1, model1), (ctxN, modelN)):
        b = ctx.batch(n, READ_LEN)
        b.append(codes[:n], lens[:n])
        out.append(ctx.run_phmm(model, b, MODE_GET_LABEL, want_spans=True, **kw))
        b.close()
    a, c = out
    bad = np.zeros(n, bool)
    for k in ("b_score", "f_score", "r_score", "bar_prob", "mapq"):
        bad |= a[k].view(np.uint32) != c[k].view(np.uint32)
    for k in ("read_type", "barcode", "fingerprint", "extracted"):
        bad |= a[k] != c[k]
    bad |= (a["labels"] != c["labels"]).any(axis=1)
    bad |= (a["spans"] != c["spans"]).reshape(n, -1).any(axis=1)
    return {"reads": int(n), "devices": ctxN.device_count, "mismatches": int(bad.sum()),
            "what": "one tdg_context over all devices vs a one-device context, all outputs compared bit for bit"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference's run_pHMM on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(n_reads, threads, seed, steps=1, warmup=0):
    """Times run_pHMM(MODE_GET_LABEL) of oracle/_ref/libtagdust_ref.so (reference compiled from
    its own sources) on `n_reads` reads of the bench workload with `threads` pthreads.
    Falls back to the plain-C port (oracle/liboracle.so) when _ref is absent."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness
    segs, tags = architecture()
    codes, lens, _ = synth.make_reads_fast(n_reads, READ_LEN, tags, error_rate=0.01, random_frac=0.05, seed=seed)
    times = []
    if refharness.have_ref():
        kind = "reference"
        R = refharness.RefHarness()
        p = R.param_new(segs, threshold=THRESHOLD, minlen=16, dust=100, threads=threads)
        mb = R.model_new(p, background=background(), average_length=float(READ_LEN), max_seq_len=READ_LEN)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            R.run_phmm(mb, p, 1, codes, lens)
            times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        refharness.build_oracle()
        from tagdust_b200.api import compile_architecture
        desc = compile_architecture(segs, background(), float(READ_LEN), READ_LEN)
        O = refharness.Oracle()
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            O.run(desc, 1, codes, lens, threshold=THRESHOLD, minlen=16, dust=100, threads=threads)
            times.append(time.perf_counter() - t0)
    times = times[warmup:]
    return kind, n_reads, times


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n = max(threads * 250, 2000)
    kind, n, times = cpu_reference_rate(n, threads, seed=1234, steps=args.steps, warmup=args.warmup)
    total = sum(times)
    value = n * len(times) / total
    segs, _ = architecture()
    C = 48 * 6 + 6 + 1
    line = {
        "impl": "reference", "metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 150 nt single-end, -1 B:<48 x 6 nt EDITTAG_6nt_ed_3> -2 R:N, 1% error, 5% random",
                   "mode": "MODE_GET_LABEL", "threshold": THRESHOLD},
        "gcups": value * 2 * READ_LEN * C / 1e9,
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": kind,
                         "sample": f"{n} reads per step, run_pHMM(MODE_GET_LABEL) with {threads} pthreads"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    from tagdust_b200 import dist_util
    from tagdust_b200.api import MODE_GET_LABEL, Context, compile_architecture, live_ops

    rank, world, local = dist_util.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist_util.init("nccl")

    segs, tags = architecture()
    desc = compile_architecture(segs, background(), float(READ_LEN), READ_LEN)
    ctx = Context(device_ids=[local])
    model = ctx.model(desc, READ_LEN)
    n_reads = args.reads
    t0 = time.perf_counter()
    codes, lens, truth = synth.make_reads_fast(n_reads, READ_LEN, tags, error_rate=0.01, random_frac=0.05, seed=100 + rank)
    log(f"[rank {rank}] generated {n_reads} reads in {time.perf_counter() - t0:.1f}s")
    batches = [ctx.batch(n_reads, READ_LEN) for _ in range(2)]
    kw = dict(threshold=THRESHOLD, minlen=16, dust=100)

    barrier = dist_util.barrier

    # ---- (1) kernel-only: inputs resident in HBM, CUDA events on the launching stream
    batches[0].append(codes, lens)
    ctx.upload(batches[0])
    stream = torch.cuda.current_stream().cuda_stream
    launches_per_step = 0
    for _ in range(args.warmup):
        launches_per_step = ctx.decode_resident(model, batches[0], MODE_GET_LABEL, stream=stream, **kw)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ctx.decode_resident(model, batches[0], MODE_GET_LABEL, stream=stream, **kw)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    prof = ctx.profile_read(0)
    ctx.profile_enable(False)
    clocks = sampler.finish()
    res = ctx.download(batches[0])
    assigned_ok = int(((res["barcode"] & 0xFFFF) == truth)[truth >= 0].sum())
    n_model = int((truth >= 0).sum())

    # ---- (2) end to end through the C ABI with HOST buffers: pack -> H2D -> kernels -> D2H,
    #          double buffered (pack of step k+1 overlaps the GPU work of step k)
    for b in batches:
        b.clear()
    h2d = n_reads // 32 * 32 * ((READ_LEN + 8) // 8) * 4 + n_reads * 4
    d2h = n_reads * (5 * 4 + 3 * 4 + 1 + ((READ_LEN + 8) // 8 * 8))

    def e2e_pass(steps):
        pending = None
        for s in range(steps):
            b = batches[s % 2]
            b.clear()
            b.append(codes, lens)                      # host pack into pinned 4-bit tiles
            ctx.submit(model, b, MODE_GET_LABEL, **kw)  # async H2D + kernels + D2H
            if pending is not None:
                ctx.wait(pending, copy=False)
            pending = b
        ctx.wait(pending, copy=False)

    e2e_pass(min(args.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_pass(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks (max time), rank 0 prints
    ms_total, e2e_ms = dist_util.max_over_ranks([ms_total, e2e_s * 1000.0])
    tallies = dist_util.merge_tallies(np.bincount(res["read_type"], minlength=7))   # merged on the host side
    assigned_ok, n_model = [int(x) for x in dist_util.merge_tallies([assigned_ok, n_model])]
    total_reads = n_reads * world
    value = total_reads * args.steps / (ms_total / 1000.0)
    e2e_value = total_reads * args.steps / (e2e_ms / 1000.0)
    C = desc.total_columns
    cells_per_read = 2 * READ_LEN * C

    pk, pk_src = peaks()
    fp32_peak = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12   # T lane-ops/s
    dom = max(prof, key=lambda k: prof[k]["ms"])
    ops_colpos = {"k_backward": OPS_PER_COLPOS_BWD, "k_forward": OPS_PER_COLPOS_FWD, "k_label": 0}
    # live work: the logsums / adds the kernels execute (log(0) terms are skipped exactly), from the compiled model;
    # a logsum is costed at SURVEY 8d's 9 FP32-pipe ops
    lo = live_ops(desc)
    live_pos = {"k_backward": 9 * lo["ls_bwd"] + lo["add_bwd"], "k_forward": 9 * lo["ls_fwd"] + lo["add_fwd"], "k_label": 0.0}
    dom_s = prof[dom]["ms"] / 1000.0
    achieved_alg = n_reads * args.steps * READ_LEN * C * ops_colpos[dom] / dom_s / 1e12 if dom_s > 0 else 0.0
    achieved = n_reads * args.steps * READ_LEN * live_pos[dom] / dom_s / 1e12 if dom_s > 0 else 0.0
    # algorithmic HBM bytes of the backward->forward hand-off: Mb,Ib of every (column, position), written once, read once
    hbm_bytes = n_reads * args.steps * READ_LEN * C * 8 * 2
    line = {
        "metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 150 nt single-end, -1 B:<48 x 6 nt EDITTAG_6nt_ed_3> -2 R:N, 1% error, 5% random",
                   "mode": "MODE_GET_LABEL", "threshold": THRESHOLD, "reads_per_step_per_gpu": n_reads,
                   "cells_per_read": cells_per_read,
                   "l2": "per-step scratch working set (~30 GB) and packed input both exceed the 126 MB L2"},
        "gcups": value * cells_per_read / 1e9,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "gcups": e2e_value * cells_per_read / 1e9,
                "what": "tdg_batch_append_codes (host pack) + tdg_submit + tdg_wait, double buffered, host arrays in and out"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp32", "kernel": dom, "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": None,
                     "achieved_algorithmic": achieved_alg, "frac_algorithmic": achieved_alg / fp32_peak,
                     "live_ops_per_hmm_position": live_pos[dom] / desc.total_hmms,
                     "live_logsums_per_hmm_position": (lo["ls_fwd"] if dom == "k_forward" else lo["ls_bwd"]) / desc.total_hmms,
                     "ops_per_column_position_algorithmic": ops_colpos[dom],
                     "peak_source": f"148 SM x 128 FP32 lanes x sm_max_mhz, {pk_src}",
                     "note": "frac = LIVE FP32-pipe ops (logsums and adds the kernel executes; a logsum = 9 ops as in SURVEY 8d) "
                             "/ duration / lane peak; frac_algorithmic counts SURVEY 8d's 8+10 logsums per (column, position) "
                             "including the log(0) terms that are never evaluated, so it can exceed 1; the posterior window of k_forward "
                             "skips a data-dependent share of the live posterior logsums at run time (they exp() to exactly 0), "
                             "which this count keeps; T lane-ops/s in the TFLOP/s unit"},
        "roofline_whole_path": {"achieved": value / world * READ_LEN * (live_pos["k_backward"] + live_pos["k_forward"]) / 1e12,
                                "achieved_algorithmic": value / world * READ_LEN * C * (OPS_PER_COLPOS_BWD + OPS_PER_COLPOS_FWD) / 1e12,
                                "peak": fp32_peak, "unit": "TFLOP/s"},
        "hbm": {"achieved_gbs": hbm_bytes / ((prof["k_backward"]["ms"] + prof["k_forward"]["ms"]) / 1000.0) / 1e9,
                "peak_gbs": pk.get("hbm_gbs"), "what": "Mb/Ib scratch write+read over k_backward+k_forward time"},
        "kernels_ms": prof,
        "check": {"reads_with_true_barcode_assigned": assigned_ok, "model_reads": n_model,
                  "read_type_counts": tallies.tolist()},
    }
    line["roofline_whole_path"]["frac"] = line["roofline_whole_path"]["achieved"] / fp32_peak
    line["roofline_whole_path"]["frac_algorithmic"] = line["roofline_whole_path"]["achieved_algorithmic"] / fp32_peak
    if not os.environ.get("TDG_LIB"):   # experiment builds (scripts/build_variant.sh) may skip work on purpose
        assert line["roofline"]["frac"] <= 1.0 and line["roofline_whole_path"]["frac"] <= 1.0, "live-op roofline fraction above 1"
    traffic, traffic_src, ncu_k = ncu_traffic(dom)
    line["roofline"]["traffic"] = traffic
    line["roofline"]["traffic_source"] = traffic_src
    for key in ("issue_active_pct", "lsu_wavefronts_pct", "pipe_fma_pct", "pipe_alu_pct", "pipe_lsu_pct",
                "shared_bank_conflict_share"):   # ncu: what the kernel is actually bound by per SM
        if key in ncu_k:
            line["roofline"][key] = ncu_k[key]
    # the same kernel against the HBM roofline: algorithmic bytes = Mb/Ib of every stored (column, position), moved once
    # (246 of the 295 columns are stored), per launch of one wave, over the event-timed duration
    alg_bytes = 148 * 512 * READ_LEN * (C - 49) * 8
    full_wave_s = dom_s / (args.steps * n_reads / (148 * 512)) if n_reads else 0.0   # duration per full wave of 75 776 reads
    hbm_achieved = alg_bytes / full_wave_s / 1e9 if full_wave_s > 0 else 0.0
    line["roofline_hbm"] = {"bound": "hbm", "kernel": dom, "achieved": hbm_achieved, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                            "frac": hbm_achieved / pk.get("hbm_gbs") if pk.get("hbm_gbs") else None, "traffic": traffic,
                            "algorithmic_bytes_per_launch": alg_bytes, "peak_source": pk_src,
                            "note": "k_backward writes and k_forward reads this stream once; ncu traffic adds the silent-state arrays"}
    # ---- strong scaling, file to files: rank 0 alone drives ONE context over all N devices (the other ranks have
    #      released their device memory and wait at the barrier below)
    for b in batches:
        b.close()
    batches = []
    if rank != 0:
        model.close(); ctx.close()
    dist_util.cpu_barrier()   # host-side from here on: an NCCL barrier would park a spinning kernel on the other ranks' GPUs
    if rank == 0 and not args.no_files:
        try:
            ctxN = ctx if world == 1 else Context(n_devices=world)
            modelN = model if world == 1 else ctxN.model(desc, READ_LEN)
            if world > 1:
                line["multi_device_check"] = multi_device_check(ctx, model, ctxN, modelN, codes, lens, kw, 2 * 148 * 512 * world + 1000)
                line["e2e_one_context"] = one_context_e2e(ctxN, modelN, codes, lens, kw, args.steps, world)
            extracted_block = int((res["read_type"][:min(n_reads, args.files_reads)] == 0).sum())
            line["e2e_files"] = files_e2e(ctxN, modelN, tags, codes, args.files_reads, host_threads(), world, extracted_block)
            if world > 1:
                modelN.close(); ctxN.close()
        except Exception as exc:  # the streaming layer is an extra line, never a reason to lose the bench
            line["e2e_files"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            line["configs"] = other_configs(ctx, host_threads(), quick=args.quick_configs)
        except Exception as exc:
            import traceback
            line["configs"] = {"error": repr(exc), "trace": traceback.format_exc()[-600:]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_cpu = max(threads * 2500, 20000)   # ~10 s of host work
        kind, n_cpu, times = cpu_reference_rate(n_cpu, threads, seed=1234)
        cpu_value = n_cpu / times[0]
        line["cpu_baseline"] = {"value": cpu_value, "unit": "reads/s", "cores": threads, "kind": kind,
                                "sample": f"{n_cpu} reads of the same workload, run_pHMM(MODE_GET_LABEL), {threads} pthreads, {times[0]:.1f} s"}
    if rank == 0:
        emit(line)
        model.close()
        ctx.close()
    dist_util.cpu_barrier()
    dist_util.finalize()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints to fd 1
    (e.g. NCCL's version banner) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=32 * 148 * 512, help="reads per step per GPU (default 32 waves)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3/4/5 + paired-end block")
    ap.add_argument("--quick-configs", action="store_true", help="smaller read sets for the configs block")
    ap.add_argument("--no-files", action="store_true", help="skip the FASTQ-file -> demultiplexed-files measurement")
    ap.add_argument("--files-reads", type=int, default=32_000_000, help="reads of the file-to-files job (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
