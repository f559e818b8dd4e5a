"""Shared architecture / read-set definitions for the parity tests (TEST INFRASTRUCTURE)."""
import os

import numpy as np

from tagdust_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS6_ED3 = synth.load_tags(os.path.join(GOLDEN, "edittag_6nt_ed3.txt"))
TAGS6_ED4 = synth.load_tags(os.path.join(GOLDEN, "edittag_6nt_ed4.txt"))

# name -> dict(segments, gen kwargs, read_len, five/three partial stats)
CASES = {
    # dev/bar_read_test.sh case 1 shape: 4 barcodes + 20 nt read
    "b4_r": dict(segments=["B:" + ",".join(TAGS6_ED4[:4]), "R:N"], barcodes=TAGS6_ED4[:4], read_len=26,
                 gen=dict(error_rate=0.02, random_frac=0.1)),
    # BASELINE cfg2 shape (shortened reads for CPU tests are set by the caller)
    "b48_r": dict(segments=["B:" + ",".join(TAGS6_ED3[:48]), "R:N"], barcodes=TAGS6_ED3[:48], read_len=150,
                  gen=dict(error_rate=0.01, random_frac=0.05)),
    # cfg3 shape: UMI + linker + barcodes + read  (16 barcodes keeps the CPU oracle quick)
    "f_s_b_r": dict(segments=["F:NNNNNNNN", "S:ACGTTGCAGTCA", "B:" + ",".join(TAGS6_ED3[:16]), "R:N"],
                    barcodes=TAGS6_ED3[:16], read_len=100,
                    gen=dict(error_rate=0.01, random_frac=0.05, umi_len=8, linker5=""), linker_after_umi="ACGTTGCAGTCA"),
    # cfg4 shape: two barcode segments
    "b_b_r": dict(segments=["B:" + ",".join(TAGS6_ED3[:6]), "B:" + ",".join(TAGS6_ED3[6:10]), "R:N"],
                  barcodes=TAGS6_ED3[:6], second=TAGS6_ED3[6:10], read_len=80, gen=dict(error_rate=0.01, random_frac=0.05)),
    # optional G-addition + short barcode + linker (the manual's example architecture)
    "o_b_s_r": dict(segments=["O:N", "B:ACGT,TTGA,GGCA,CATG", "S:GGG", "R:N"], barcodes=["ACGT", "TTGA", "GGCA", "CATG"],
                    read_len=60, gen=dict(error_rate=0.02, random_frac=0.1), linker_after_bc="GGG"),
    # partial 5'/3' adapters around barcode + read (bar_read_test.sh case 2 shape)
    "p_b_r_p": dict(segments=["P:GGGGGGG", "B:" + ",".join(TAGS6_ED4[:4]), "R:N", "P:TTTTTTT"], barcodes=TAGS6_ED4[:4],
                    read_len=40, gen=dict(error_rate=0.02, random_frac=0.1, linker5="GGGGGGG", linker3="TTTTTTT"),
                    five=(7.0, 6.2, 1.1), three=(7.0, 5.9, 1.3)),
    # G segment and 2-column / 1-column HMMs
    "g_b2_r": dict(segments=["G:G", "B:AC,GT,TG", "S:T", "R:N"], barcodes=["AC", "GT", "TG"], read_len=40,
                   gen=dict(error_rate=0.02, random_frac=0.1), linker_after_bc="T"),
    # long linker (dynamic column path, > 8 columns)
    "s20_b_r": dict(segments=["S:ACGTACGGTTCAGCATGCAA", "B:" + ",".join(TAGS6_ED3[:8]), "R:N"], barcodes=TAGS6_ED3[:8],
                    read_len=70, gen=dict(error_rate=0.01, random_frac=0.05, linker5="ACGTACGGTTCAGCATGCAA")),
    # 12-nt UMI + 12-nt barcodes (unrolled standard-pattern kernels for 9..16 columns)
    "f12_b12_r": dict(segments=["F:NNNNNNNNNNNN", "B:ACGTTGCAACGT,TTGACCATGCAA,GGCATTACCGTA,CATGCATGGTAC,TACGGATCTAGC", "R:N"],
                      barcodes=["ACGTTGCAACGT", "TTGACCATGCAA", "GGCATTACCGTA", "CATGCATGGTAC", "TACGGATCTAGC"], read_len=70,
                      gen=dict(error_rate=0.02, random_frac=0.1, umi_len=12)),
    # 5-, 7- and 8-nt barcode sets: the grouped backward loop (4 HMMs per position loop up to 6 columns, 2 beyond) with
    # leftover HMMs that do not fill a group
    "b5x7_r": dict(segments=["B:ACGTA,TTGAC,GGCAT,CATGC,TACGG,AGTCT,CCATA", "R:N"],
                   barcodes=["ACGTA", "TTGAC", "GGCAT", "CATGC", "TACGG", "AGTCT", "CCATA"], read_len=50, gen=dict(error_rate=0.02, random_frac=0.1)),
    "b7x6_r": dict(segments=["B:ACGTACG,TTGACCA,GGCATTA,CATGCAT,TACGGAT,AGTCTGA", "R:N"],
                   barcodes=["ACGTACG", "TTGACCA", "GGCATTA", "CATGCAT", "TACGGAT", "AGTCTGA"], read_len=50, gen=dict(error_rate=0.02, random_frac=0.1)),
    "f4_b8x9_r": dict(segments=["F:NNNN", "B:ACGTACGT,TTGACCAT,GGCATTAC,CATGCATG,TACGGATC,AGTCTGAA,CCATAGGC,GTTCAAGT,TGCAGTCA", "R:N"],
                      barcodes=["ACGTACGT", "TTGACCAT", "GGCATTAC", "CATGCATG", "TACGGATC", "AGTCTGAA", "CCATAGGC", "GTTCAAGT", "TGCAGTCA"],
                      read_len=60, gen=dict(error_rate=0.02, random_frac=0.1, umi_len=4)),
    # long partial adapters on both ends (column-loop kernel path, generic segments)
    "p18_b_r_p14": dict(segments=["P:AGGGAGGACGATGCGGTC", "B:" + ",".join(TAGS6_ED4[:4]), "R:N", "P:GATCGGAAGAGCAC"], barcodes=TAGS6_ED4[:4],
                        read_len=80, gen=dict(error_rate=0.02, random_frac=0.1, linker5="AGGGAGGACGATGCGGTC", linker3="GATCGGAAGAGCAC"),
                        five=(18.0, 15.2, 2.1), three=(14.0, 11.9, 1.7)),
}


def make_case_reads(name, n, seed=11, read_len=None, len_jitter=3, n_frac=0.01):
    c = CASES[name]
    L = read_len or c["read_len"]
    gen = dict(c["gen"])
    kw = dict(seed=seed, len_jitter=len_jitter, n_frac=n_frac)
    kw.update(gen)
    bcs = list(c["barcodes"])
    if "linker_after_umi" in c:  # UMI + linker + barcode: fold the linker into every barcode string
        bcs = [c["linker_after_umi"] + b for b in bcs]
    if "linker_after_bc" in c:
        bcs = [b + c["linker_after_bc"] for b in bcs]
    second = c.get("second")
    codes, lens, truth = synth.make_reads(n, L, bcs, second_barcodes=second, **kw)
    return codes, lens, truth


def build_ref_model(R, name, threshold=0.0, minlen=16, dust=100, avg_len=None, max_len=None, threads=1, background=None):
    from refharness import background_logp
    c = CASES[name]
    L = c["read_len"]
    p = R.param_new(c["segments"], threshold=threshold, minlen=minlen, dust=dust, threads=threads)
    bg = background_logp((2501.0, 2480.0, 2510.0, 2492.0, 21.0)) if background is None else background
    mb = R.model_new(p, background=bg, average_length=float(avg_len or L), max_seq_len=int(max_len or (L + 8)),
                     five=c.get("five", (0.0, 0.0, 0.0)), three=c.get("three", (0.0, -1.0, -1.0)))
    return p, mb, R.flatten(mb, p)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a
