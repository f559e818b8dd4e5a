"""CPU tests of the streaming layer (include/tagdust_b200_stream.h): the FASTQ/FASTA reader against
the reference's own io_handler + read_fasta_fastq (through oracle/_ref), the %0.2f formatter
against printf, and the host-only path of the demultiplexer (architectures that are a single R
segment never reach the GPU: run_rna_dust, barcode_hmm.c:312-318) against the reference CLI."""
import ctypes as C
import filecmp
import glob
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from refharness import REF_SO, have_ref
from tagdust_b200 import _capi
from tagdust_b200.api import TagdustError
from tagdust_b200.stream import FastqReader, demux_run, format_rq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def test_stream_header_symbols_exported():
    lib = _capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "tagdust_b200_stream.h")).read()
    declared = set(re.findall(r"\b(tdg_[a-z0-9_]+)\s*\(", hdr))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tagdust_b200_stream.h but not exported"
    assert declared == set(_capi.STREAM_PROTOTYPES), declared ^ set(_capi.STREAM_PROTOTYPES)


def test_format_rq_matches_printf():
    rng = np.random.default_rng(3)
    vals = list(rng.uniform(0, 40, 20000).astype(np.float32)) + list(rng.uniform(-2, 2, 2000).astype(np.float32))
    # exact ties of the third decimal, signed zeros, large and tiny values, the reference's constants
    vals += [np.float32(x) for x in (0.125, 0.375, 0.625, 0.875, 2.5, 1.005, 39.995, 40.0, 0.0, -0.0, -1.0, -0.001, 0.005,
                                     0.015, 0.025, 1e-9, 123456.789, 9999999.0, 3.0386538505554199, 1.7987838983535767)]
    for k in range(0, 4000):
        vals.append(np.float32(k / 8.0 + 0.125))
    for v in vals:
        assert format_rq(float(v)) == "%0.2f" % float(v), float(v)


def write_fastq(path, recs, crlf=False, trailing_newline=True):
    eol = "\r\n" if crlf else "\n"
    txt = eol.join(f"@{n}{eol}{s}{eol}+{eol}{q}" for n, s, q in recs)
    if trailing_newline:
        txt += eol
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt", newline="") as fh:
        fh.write(txt)


def random_records(rng, n, lo=1, hi=200, exotic=False):
    alpha = "ACGTN" + ("acgtnRYK.U" if exotic else "")
    recs = []
    for k in range(n):
        L = int(rng.integers(lo, hi + 1))
        seq = "".join(rng.choice(list(alpha), size=L))
        qual = "".join(chr(int(c)) for c in rng.integers(33, 75, size=L))
        if exotic and k % 7 == 0:
            qual = "@" + qual[1:]          # quality lines may start with '@' or '+'
        if exotic and k % 11 == 0:
            qual = "+" + qual[1:]
        name = f"read{k}:{rng.integers(1, 9)}:{rng.integers(1000, 9999)} 1:N:0:{k % 5}" if k % 3 else f"r{k};x=1\tafter_tab"
        recs.append((name, seq, qual))
    return recs


class RefReader:
    def __init__(self):
        self.L = C.CDLL(REF_SO)
        self.L.refh_read_file_chunk.restype = C.c_int

    def chunk(self, path, num_query, index, stride=400):
        lens = np.zeros(num_query, np.int32)
        codes = np.zeros((num_query, stride), np.uint8)
        quals = np.zeros((num_query, stride), np.uint8)
        hq = C.c_int(0)
        names = C.create_string_buffer(num_query * 200 + 16)
        n = self.L.refh_read_file_chunk(str(path).encode(), num_query, index, stride, lens.ctypes.data_as(C.c_void_p),
                                        codes.ctypes.data_as(C.c_void_p), quals.ctypes.data_as(C.c_void_p), C.byref(hq),
                                        names, C.c_size_t(len(names)))
        assert n >= 0
        nm = names.value.split(b"\n")[:n]
        return n, lens[:n], codes[:n], quals[:n], bool(hq.value), nm


def assert_chunk_equal(mine, ref):
    n, lens, codes, quals, hq, names = ref
    assert mine["n"] == n
    assert np.array_equal(mine["len"], lens)
    assert mine["names"] == names
    for r in range(n):
        o, L = int(mine["off"][r]), int(lens[r])
        assert np.array_equal(mine["codes"][o:o + L + 1], codes[r, :L + 1]), r
        if hq:
            assert np.array_equal(mine["qual"][o:o + L + 1], quals[r, :L + 1]), r
        else:
            assert mine["qual"] is None


@pytest.mark.parametrize("variant", ["plain", "crlf", "gz", "no_trailing_newline", "exotic"])
def test_reader_matches_reference_reader(tmp_path, variant):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    recs = random_records(rng, 2500, exotic=(variant == "exotic"))
    path = tmp_path / ("x.fq.gz" if variant == "gz" else "x.fq")
    write_fastq(path, recs, crlf=(variant == "crlf"), trailing_newline=(variant != "no_trailing_newline"))
    R = RefReader()
    rd = FastqReader(path)
    # chunk size that does not divide the file, several chunks, then end of input
    for k in range(4):
        mine = rd.next(1000, threads=3)
        ref = R.chunk(path, 1000, k)
        if ref[0] == 0:
            assert mine is None
            break
        assert_chunk_equal(mine, ref)
    rd.close()


def read_all_chunks(path, chunk, threads):
    rd = FastqReader(path)
    out = []
    while True:
        try:
            c = rd.next(chunk, threads=threads)
        except TagdustError as e:
            out.append(("error", str(e)))
            break
        if c is None:
            break
        # only len + 1 bytes of a read's slot are defined (the slot is sized for the raw line, e.g. with its '\r')
        rows = [(int(o), int(L)) for o, L in zip(c["off"], c["len"])]
        out.append((c["n"], c["len"].tobytes(), c["off"].tobytes(), b"".join(c["codes"][o:o + L + 1].tobytes() for o, L in rows),
                    None if c["qual"] is None else b"".join(c["qual"][o:o + L + 1].tobytes() for o, L in rows), tuple(c["names"])))
    rd.close()
    return out


@pytest.mark.parametrize("variant", ["plain", "crlf", "exotic", "no_trailing_newline", "long_lines", "fasta", "double_plus",
                                     "truncated", "junk_between"])
def test_parallel_line_pass_equals_serial(tmp_path, variant):
    """Files of a few MB go through the multi-threaded line pass (stretches cut at line starts, three start-state
    hypotheses per stretch, stitched in order); it must hand out exactly the chunks of the sequential pass, which the
    tests above pin to the reference's read_fasta_fastq -- including malformed input, where it has to step aside."""
    rng = np.random.default_rng(11)
    n = 14000
    path = tmp_path / ("x.fa" if variant == "fasta" else "x.fq")
    if variant == "fasta":
        with open(path, "w") as fh:
            for k in range(n * 2):
                fh.write(f">seq{k} d\n{'ACGTNNACGT' * (1 + k % 19)}\n")
                if k % 10 == 0:
                    fh.write("\n")
                if k % 25 == 0:
                    fh.write("GGGGGGGG\n>not a header: the flag is clear, so it IS one\nAC\n")
    else:
        recs = random_records(rng, n, exotic=(variant in ("exotic", "junk_between")))
        if variant == "long_lines":
            for k in range(0, n, 900):
                L = 10050 + k
                recs[k] = (recs[k][0], "ACGT" * (L // 4), "I" * (L // 4 * 4))
        write_fastq(path, recs, crlf=(variant == "crlf"), trailing_newline=(variant != "no_trailing_newline"))
        if variant in ("double_plus", "truncated", "junk_between"):
            txt = open(path).read().split("\n")
            if variant == "double_plus":      # an entry with two quality blocks, far into the file
                i = 4 * 9001
                txt[i + 3:i + 3] = [txt[i + 3], "+"]
            elif variant == "truncated":      # the last entry has no quality line
                txt = txt[:-2]
            else:                             # stray lines between entries, some looking like headers
                for i in range(4 * 13000, 4 * 100, -4 * 777):
                    txt[i:i] = ["junk line", "+", "@IIIIIII"]
            open(path, "w").write("\n".join(txt))
    assert os.path.getsize(path) > 2 << 20
    for chunk in (1000, 5003, 100000):
        serial = read_all_chunks(path, chunk, 1)
        for threads in (2, 5):
            par = read_all_chunks(path, chunk, threads)
            assert len(par) == len(serial)
            for a, b in zip(par, serial):
                assert a == b
    if variant in ("plain", "exotic", "fasta") and have_ref():
        R = RefReader()
        rd = FastqReader(path)
        for k in range(3):
            mine = rd.next(6000, threads=4)
            ref = R.chunk(path, 6000, k)
            assert_chunk_equal(mine, ref)
        rd.close()


def test_reader_fasta_and_blank_lines(tmp_path):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    path = tmp_path / "x.fa"
    with open(path, "w") as fh:
        for k in range(300):
            fh.write(f">seq{k} d\nACGTNNACGT{'ACGT' * (k % 9)}\n")
            if k % 10 == 0:
                fh.write("\n")                      # blank line
            if k % 25 == 0:
                fh.write("GGGGGGGG\n")              # second sequence line: ignored by the reference's reader
    R = RefReader()
    rd = FastqReader(path)
    assert_chunk_equal(rd.next(1000, threads=2), R.chunk(path, 1000, 0))
    assert rd.next(1000) is None
    rd.close()


def test_reader_errors(tmp_path):
    with pytest.raises(TagdustError):
        FastqReader(tmp_path / "missing.fq")
    p = tmp_path / "bad.fq"
    p.write_text("@r1\nACGT\n+\nIII\n")           # io.c:1770 "Length of sequence and base qualities differ"
    rd = FastqReader(p)
    with pytest.raises(TagdustError) as e:
        rd.next(10)
    assert "differ" in str(e.value)
    rd.close()
    with pytest.raises(TagdustError):
        FastqReader(tmp_path / "x.bam") if (tmp_path / "x.bam").write_text("x") else None


def test_empty_file(tmp_path):
    p = tmp_path / "empty.fq"
    p.write_text("")
    rd = FastqReader(p)
    assert rd.next(10) is None
    rd.close()


def run_ref_cli(tmp, args, prefix, tag="cpu"):
    d = os.path.join(tmp, tag)
    os.makedirs(d, exist_ok=True)
    r = subprocess.run(f"{REFDIR}/tagdust_rtest -seed 42 {args} -o {d}/{prefix}", cwd=tmp, shell=True, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return d


def test_demux_read_only_architecture_matches_reference_cli(tmp_path):
    """-1 R:N on two paired files (no HMM, no GPU): dust filter, cross-file merge of read_type,
    READ1/READ2 + un files byte-identical to the reference CLI."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    tmp = str(tmp_path)
    rng = np.random.default_rng(9)
    recs1 = random_records(rng, 3000, lo=30, hi=80)
    recs2 = []
    for k, (n, s, q) in enumerate(recs1):
        L = int(rng.integers(30, 81))
        s2 = "".join(rng.choice(list("ACGT"), size=L)) if k % 17 else "A" * L      # low complexity -> read_type 6
        recs2.append((n, s2, "".join(chr(int(c)) for c in rng.integers(33, 75, size=L))))
    write_fastq(os.path.join(tmp, "r1.fq"), recs1)
    write_fastq(os.path.join(tmp, "r2.fq"), recs2)
    cpu = run_ref_cli(tmp, "-1 R:N r1.fq r2.fq", "out")
    mine = os.path.join(tmp, "mine"); os.makedirs(mine)
    st = demux_run(None, [dict(path=os.path.join(tmp, "r1.fq"), model=None, num_read_segments=1),
                          dict(path=os.path.join(tmp, "r2.fq"), model=None, num_read_segments=1)],
                   os.path.join(mine, "out"), dust=100, threads=4, chunk_reads=700)
    a = sorted(glob.glob(os.path.join(cpu, "out*.fq"))); b = sorted(glob.glob(os.path.join(mine, "out*.fq")))
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and len(a) == 4
    for x, y in zip(a, b):
        assert filecmp.cmp(x, y, shallow=False), os.path.basename(x)
    assert st["total_read"] == 3000
    log = open(os.path.join(cpu, "out_logfile.txt")).read()
    assert f"{st['num_EXTRACT_SUCCESS']}\tsuccessfully extracted" in log
    assert f"{st['num_EXTRACT_FAIL_LOW_COMPLEXITY']}\tlow complexity" in log


@pytest.mark.parametrize("alpha", ["ACGTNacgtnRYKUu*-"])
def test_demux_odd_characters_match_reference_cli(tmp_path, alpha):
    """-1 R:N on one file whose reads hold lower-case bases, IUPAC letters, U and other characters: the writer's
    sixteen-at-a-time base copy against the reference CLI's print_all, byte for byte (no GPU involved).  ('.' is left out:
    the reference binary crashes on reads that contain it, with or without -Q.)"""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    tmp = str(tmp_path)
    rng = np.random.default_rng(21)
    recs = []
    for k in range(2500):
        L = int(rng.integers(20, 140))
        pool = list(alpha) if k % 3 == 0 else list("ACGT")
        seq = "".join(rng.choice(pool, size=L))
        recs.append((f"r{k} odd", seq, "".join(chr(int(c)) for c in rng.integers(33, 75, size=L))))
    write_fastq(os.path.join(tmp, "r1.fq"), recs)
    cpu = run_ref_cli(tmp, "-1 R:N r1.fq", "out")
    mine = os.path.join(tmp, "mine"); os.makedirs(mine)
    st = demux_run(None, [dict(path=os.path.join(tmp, "r1.fq"), model=None, num_read_segments=1)],
                   os.path.join(mine, "out"), dust=100, threads=3, chunk_reads=600)
    a = sorted(glob.glob(os.path.join(cpu, "out*.fq"))); b = sorted(glob.glob(os.path.join(mine, "out*.fq")))
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and len(a) >= 2
    for x, y in zip(a, b):
        assert filecmp.cmp(x, y, shallow=False), os.path.basename(x)
    assert st["total_read"] == 2500


def test_demux_unequal_files_error(tmp_path):
    rng = np.random.default_rng(1)
    recs = random_records(rng, 50, lo=30, hi=40)
    write_fastq(tmp_path / "a.fq", recs)
    write_fastq(tmp_path / "b.fq", recs[:40])
    with pytest.raises(TagdustError) as e:
        demux_run(None, [dict(path=tmp_path / "a.fq", model=None), dict(path=tmp_path / "b.fq", model=None)], tmp_path / "o", threads=2)
    assert "differ in number of entries" in str(e.value)


def derived_stats(st, five_len, three_len):
    """io.c:216-270 on the raw sums of tdg_sequence_stats (the same arithmetic integration/stats_fast.c does)."""
    import math
    out = {}
    bg = [1.0 + st.base_count[k] for k in range(5)]
    s = sum(bg)
    out["background"] = [float(np.float32(math.log(float(np.float32(b / s))))) for b in bg]
    def part(s0, s1, s2, L):
        if not L:
            return -1.0, -1.0
        if s0 <= 1:
            return float(L), 1.0
        sd = math.sqrt((s0 * s2 - s1 ** 2.0) / (s0 * (s0 - 1.0)))
        return s1 / s0, (sd if sd else 10000.0)
    out["five"] = part(st.five_s0, st.five_s1, st.five_s2, five_len)
    out["three"] = part(st.three_s0, st.three_s1, st.three_s2, three_len)
    out["average_length"] = float(int(math.floor(st.sum_len / st.total_read + 0.5)))
    out["max_seq_len"] = st.max_seq_len
    return out


@pytest.mark.parametrize("num_query", [1000, 1000001])
def test_sequence_stats_match_reference(tmp_path, ref, num_query):
    """tdg_sequence_stats (+ the derivation of io.c:216-270) vs the reference's get_sequence_stats, with 5' and 3'
    partial segments whose matched length varies from read to read."""
    from tagdust_b200.synth import encode
    rng = np.random.default_rng(12)
    five, three = "GGGTACGTAGG", "TTTCAGGCATT"
    recs = []
    for k in range(4200):
        a = int(rng.integers(0, len(five) + 1)); b = int(rng.integers(0, len(three) + 1))
        body = "".join(rng.choice(list("ACGTN"), size=int(rng.integers(20, 60)), p=[0.3, 0.2, 0.2, 0.29, 0.01]))
        seq = five[a:] + body + three[:len(three) - b]
        recs.append((f"r{k}", seq, "I" * len(seq)))
    path = tmp_path / "p.fq"
    write_fastq(path, recs)
    p = ref.param_new(["P:" + five, "B:ACGT,TTGA", "R:N", "P:" + three])
    out = np.zeros(13, np.float64)
    ref.L.refh_sequence_stats.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    assert ref.L.refh_sequence_stats(p, str(path).encode(), num_query, out.ctypes.data_as(C.c_void_p)) == 0
    lib = _capi.load_library()
    st = _capi.SeqStatsC()
    f5, t3 = encode(five), encode(three)
    rc = lib.tdg_sequence_stats(str(path).encode(), 0, num_query, f5.ctypes.data_as(_capi.c_uint8_p), len(five),
                                t3.ctypes.data_as(_capi.c_uint8_p), len(three), 3, C.byref(st))
    assert rc == 0 and st.total_read == 4200
    d = derived_stats(st, len(five), len(three))
    assert d["background"] == list(out[:5])
    assert d["five"] == (out[7], out[8]) and d["three"] == (out[9], out[10])
    assert d["average_length"] == out[11] and d["max_seq_len"] == int(out[12])
    ref.param_free(p)


@pytest.mark.parametrize("gz", [False, True])
def test_demux_tiny_chunks_and_pipes(tmp_path, gz):
    """Chunk boundaries in the middle of blocks, pipe input (zcat) with carried-over tails, three slots recycled
    many times: the output is still the reference CLI's, byte for byte."""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    tmp = str(tmp_path)
    rng = np.random.default_rng(17)
    recs = random_records(rng, 4321, lo=17, hi=140)
    suffix = ".fq.gz" if gz else ".fq"
    write_fastq(os.path.join(tmp, "in" + suffix), recs)
    cpu = run_ref_cli(tmp, f"-Q 3 -1 R:N in{suffix}", "out")
    mine = os.path.join(tmp, "mine"); os.makedirs(mine)
    st = demux_run(None, [dict(path=os.path.join(tmp, "in" + suffix), model=None, num_read_segments=1)],
                   os.path.join(mine, "out"), dust=100, threads=3, chunk_reads=97)
    assert st["total_read"] == 4321
    for name in ("out.fq", "out_un.fq"):
        assert filecmp.cmp(os.path.join(cpu, name), os.path.join(mine, name), shallow=False), name


def test_dropin_binary_host_only_run_with_interposed_reader(tmp_path):
    """The drop-in binary (reference CLI + integration/*.c) on an architecture that never reaches the HMM (-1 R:N on two
    paired files, fixed -Q): no GPU is needed, so this runs here.  It exercises the interposed io_handler /
    read_fasta_fastq (integration/reader_fast.c: the read-name order check of the controller reads its 1000 entries
    through tdg_fastq_next), the fast get_sequence_stats and the host-only path of tdg_demux_run, byte for byte against
    the reference CLI."""
    gpu_bin = os.path.join(ROOT, "integration", "_build", "tagdust_gpu_rtest")
    if not have_ref() or not os.path.exists(gpu_bin):
        pytest.skip("oracle/_ref or integration/_build not built")
    rng = np.random.default_rng(21)
    recs1 = random_records(rng, 3300, lo=30, hi=90)
    recs2 = [(n.replace(" 1:N", " 2:N"), s, q) for n, s, q in random_records(rng, 3300, lo=30, hi=90)]
    recs2 = [(a[0], b[1], b[2]) for a, b in zip(recs1, recs2)]            # same names (paired files)
    write_fastq(tmp_path / "r1.fq", recs1)
    write_fastq(tmp_path / "r2.fq", recs2)
    outs = {}
    for tag, binary in (("cpu", f"{REFDIR}/tagdust_rtest"), ("gpu", gpu_bin)):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run(f"{binary} -seed 42 -t 3 -Q 10 -1 R:N r1.fq r2.fq -o {d}/out", cwd=tmp_path, shell=True, capture_output=True, text=True)
        assert r.returncode == 0, f"{tag}: {r.stderr}"
        outs[tag] = sorted(glob.glob(str(d / "out*.fq")))
    assert [os.path.basename(x) for x in outs["cpu"]] == [os.path.basename(x) for x in outs["gpu"]] and len(outs["cpu"]) == 4
    for a, b in zip(outs["cpu"], outs["gpu"]):
        assert filecmp.cmp(a, b, shallow=False), os.path.basename(a)
    # files in different order are refused by both (compare_read_names on the interposed reader's entries)
    write_fastq(tmp_path / "r3.fq", [(f"other{k}", s, q) for k, (n, s, q) in enumerate(recs2)])
    codes = []
    for binary in (f"{REFDIR}/tagdust_rtest", gpu_bin):
        r = subprocess.run(f"{binary} -seed 42 -t 3 -Q 10 -1 R:N r1.fq r3.fq -o {tmp_path}/bad", cwd=tmp_path, shell=True, capture_output=True, text=True)
        codes.append(r.returncode != 0)
    assert codes[0] == codes[1]
