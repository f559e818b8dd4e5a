"""Command-line scenarios through the drop-in binary with the streaming controller
(integration/controller_gpu.c + tdg_demux_run) against the CPU reference CLI: every output file
must be byte-identical and the run summaries equal.  Covers the writer's branches: several R
segments per input (READ1/READ2 from one file), UMI headers (`FP:` as number and as sequence),
FASTA input ('.' qualities), gz input, -start/-end windows, too-short / low-complexity reads,
barcodes on the first of two paired files, an index-only file (no R segment)."""
import filecmp
import glob
import gzip
import os
import subprocess

import numpy as np
import pytest

from cases import TAGS6_ED4
from test_gold_dropin import GPU_BIN, REF, need_bins, summary_lines

pytestmark = pytest.mark.gpu


def make_fastq(path, n, layout, seed, fasta=False, read_len=(40, 60), name_fmt="M1:7:FC:1:{t}:{x}:{y} 1:N:0:1", short=True):
    """layout: list of ('B', [tags]) / ('F', k) / ('S', seq) / ('R', None) parts, in read order."""
    rng = np.random.default_rng(seed)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt") as fh:
        for r in range(n):
            parts = []
            rnd = rng.random() < 0.08
            for kind, arg in layout:
                if kind == "B":
                    s = arg[int(rng.integers(0, len(arg)))]
                    if rng.random() < 0.05:
                        k = int(rng.integers(0, len(s))); s = s[:k] + "ACGT"[int(rng.integers(0, 4))] + s[k + 1:]
                    parts.append(s)
                elif kind == "F":
                    parts.append("".join(rng.choice(list("ACGT"), size=arg)))
                elif kind == "S":
                    parts.append(arg)
                else:
                    L = int(rng.integers(read_len[0], read_len[1] + 1))
                    if r % 23 == 0:
                        parts.append("A" * L)                       # low complexity
                    elif short and r % 29 == 0:
                        parts.append("".join(rng.choice(list("ACGT"), size=9)))   # shorter than minlen
                    else:
                        parts.append("".join(rng.choice(list("ACGTN"), size=L, p=[.248, .248, .248, .248, .008])))
            seq = "".join(parts)
            if rnd:
                seq = "".join(rng.choice(list("ACGT"), size=len(seq)))
            name = name_fmt.format(t=1100 + r % 9, x=1000 + r, y=2000 + 3 * r)
            if fasta:
                fh.write(f">{name}\n{seq}\n")
            else:
                q = "".join(chr(int(c)) for c in rng.integers(35, 74, size=len(seq)))
                fh.write(f"@{name}\n{seq}\n+\n{q}\n")


def run_pair(tmp, args, prefix="out"):
    outs = {}
    for tag, binary in (("cpu", os.path.join(REF, "tagdust_rtest")), ("gpu", "TDG_CHUNK_READS=900 " + GPU_BIN)):
        d = os.path.join(tmp, tag)
        os.makedirs(d, exist_ok=True)
        r = subprocess.run(f"{binary} -seed 42 -t 4 {args} -o {d}/{prefix}", cwd=tmp, shell=True, capture_output=True, text=True)
        assert r.returncode == 0, f"{tag}: {r.stdout}\n{r.stderr}"
        outs[tag] = d
    a = sorted(glob.glob(os.path.join(outs["cpu"], prefix + "*.fq")))
    b = sorted(glob.glob(os.path.join(outs["gpu"], prefix + "*.fq")))
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and a
    for x, y in zip(a, b):
        assert filecmp.cmp(x, y, shallow=False), f"{os.path.basename(x)} differs"
    assert summary_lines(f"{outs['cpu']}/{prefix}_logfile.txt") == summary_lines(f"{outs['gpu']}/{prefix}_logfile.txt")
    return a


TAGS = TAGS6_ED4[:5]
BARC = "B:" + ",".join(TAGS)


def test_two_read_segments_from_one_file(tmp_path):
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq"), 2500, [("R", None), ("S", "GGTCTCGG"), ("B", TAGS), ("R", None)], seed=1, read_len=(25, 35))
    files = run_pair(tmp, f"-1 R:N -2 S:GGTCTCGG -3 {BARC} -4 R:N in.fq")
    assert any(f.endswith("_READ2.fq") for f in files)


@pytest.mark.parametrize("show", ["", "-show_finger_seq"])
def test_umi_headers(tmp_path, show):
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq"), 2500, [("F", 6), ("B", TAGS), ("R", None)], seed=2)
    run_pair(tmp, f"{show} -1 F:NNNNNN -2 {BARC} -3 R:N in.fq")


def test_gz_input(tmp_path):
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq.gz"), 2000, [("B", TAGS), ("R", None)], seed=4)
    run_pair(tmp, f"-1 {BARC} -2 R:N in.fq.gz", prefix="gz")


def test_fasta_input_runs(tmp_path):
    """The CPU reference segfaults on FASTA input with an HMM architecture (make_extracted_read writes
    ri->qual, which read_fasta_fastq leaves NULL, barcode_hmm.c:3336); the streaming path keeps
    print_all's '.' qualities (io.c:940-944).  No byte comparison is possible: check the shape."""
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fa"), 2000, [("B", TAGS), ("R", None)], seed=3, fasta=True)
    r = subprocess.run(f"{GPU_BIN} -seed 42 -t 4 -1 {BARC} -2 R:N in.fa -o {tmp}/fa", cwd=tmp, shell=True, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    recs = 0
    for f in glob.glob(os.path.join(tmp, "fa*.fq")):
        lines = open(f).read().splitlines()
        assert len(lines) % 4 == 0
        for k in range(0, len(lines), 4):
            assert lines[k].startswith("@") and lines[k + 2] == "+" and set(lines[k + 3]) <= {"."} and len(lines[k + 1]) == len(lines[k + 3])
        recs += len(lines) // 4
    assert recs == 2000


def test_window_minlen_dust(tmp_path):
    tmp = str(tmp_path)
    # every read reaches the -end position: the reference reads past shorter reads (it crashes on them)
    make_fastq(os.path.join(tmp, "w.fq"), 2500, [("S", "TT"), ("B", TAGS), ("R", None)], seed=5, read_len=(45, 45), short=False)
    # -Q: threshold calibration under -start/-end makes the reference read past the end of its simulated
    # reads (do_probability_estimation :2196-2199 uses seq + matchstart for matchend - matchstart bases
    # whatever ri->len is); the GPU path refuses such reads instead, so the window is tested with a fixed threshold
    run_pair(tmp, f"-Q 3 -start 3 -end 40 -1 {BARC} -2 R:N w.fq", prefix="win")
    make_fastq(os.path.join(tmp, "in.fq"), 2500, [("S", "TT"), ("B", TAGS), ("R", None)], seed=5, read_len=(30, 55))
    run_pair(tmp, f"-minlen 30 -dust 20 -1 S:TT -2 {BARC} -3 R:N in.fq", prefix="flt")


def test_paired_barcode_in_first_file(tmp_path):
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "r1.fq"), 2200, [("B", TAGS), ("R", None)], seed=6)
    make_fastq(os.path.join(tmp, "r2.fq"), 2200, [("R", None)], seed=7, name_fmt="M1:7:FC:1:{t}:{x}:{y} 2:N:0:1")
    files = run_pair(tmp, f"-1 {BARC} -2 R:N r1.fq r2.fq")
    assert sum(f.endswith("_READ2.fq") for f in files) == len(TAGS) + 1


def test_index_only_file_via_arch_file(tmp_path):
    """casava layout: read 1, 6-nt index read (architecture without an R segment), read 2."""
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "r1.fq"), 2000, [("R", None)], seed=8)
    make_fastq(os.path.join(tmp, "i1.fq"), 2000, [("B", TAGS)], seed=9, name_fmt="M1:7:FC:1:{t}:{x}:{y} 2:N:0:1")
    make_fastq(os.path.join(tmp, "r2.fq"), 2000, [("R", None)], seed=10, name_fmt="M1:7:FC:1:{t}:{x}:{y} 3:N:0:1")
    with open(os.path.join(tmp, "arch.txt"), "w") as fh:
        fh.write("tagdust -1 R:N\n")
        fh.write(f"tagdust -1 {BARC}\n")
    run_pair(tmp, "-arch arch.txt r1.fq i1.fq r2.fq")


def test_architecture_selection_from_arch_file(tmp_path):
    """-arch with several candidate lines for one input: test_architectures.c scores every candidate with
    backward() over the first chunk (MODE_ARCH_COMP, per-thread float sums) and continues with the best one;
    the log carries the selected architecture and its confidence, the output files its barcodes."""
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq"), 2400, [("F", 4), ("B", TAGS), ("R", None)], seed=11)
    with open(os.path.join(tmp, "arch.txt"), "w") as fh:
        fh.write(f"tagdust -1 {BARC} -2 R:N\n")
        fh.write(f"tagdust -1 F:NNNN -2 {BARC} -3 R:N\n")
        fh.write(f"tagdust -1 F:NNNNNNNN -2 {BARC} -3 R:N\n")
        fh.write(f"tagdust -1 O:N -2 B:{','.join(TAGS[:3])} -3 S:GG -4 R:N\n")
        fh.write("tagdust -1 R:N\n")
    files = run_pair(tmp, "-arch arch.txt in.fq")
    assert len(files) == len(TAGS) + 1
    log = open(os.path.join(tmp, "gpu", "out_logfile.txt")).read()
    assert "F:NNNN" in log and "Confidence" in log


def test_arch_file_with_more_candidates_than_the_model_cache(tmp_path):
    """30 candidate architectures (the shim's content-keyed model cache holds 24): every model of ab->archs[] must stay
    alive until tdg_arch_compare returns (the cache used to evict -- and destroy -- models still in use).
    test_architectures allows up to 99 candidates (MAX_NUM_ARCH 100)."""
    from cases import TAGS6_ED3
    from tagdust_b200 import synth
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq"), 2000, [("B", TAGS), ("R", None)], seed=12)
    archs = synth.candidate_architectures(TAGS6_ED3, 30)
    archs[7] = [BARC, "R:N"]                      # the true one
    with open(os.path.join(tmp, "arch.txt"), "w") as fh:
        for a in archs:
            fh.write("tagdust " + " ".join(f"-{k + 1} {seg}" for k, seg in enumerate(a)) + "\n")
    files = run_pair(tmp, "-arch arch.txt in.fq")
    assert len(files) == len(TAGS) + 1
    log = open(os.path.join(tmp, "gpu", "out_logfile.txt")).read()
    assert ",".join(TAGS) in log and "Confidence" in log


def test_reads_shorter_than_the_architecture(tmp_path):
    """Reads of 2-5 nt under a 6-nt barcode + read architecture: the profile HMMs can
    delete columns, so the scores stay finite and the reference classifies them (too short / mismatch); the drop-in must
    write the same files.  (1-nt reads make the CPU reference itself segfault; reads for which b_score is -inf make it
    index its logsum table with (int)NaN, DESIGN.md section 2: both are outside the contract.)"""
    tmp = str(tmp_path)
    rng = np.random.default_rng(13)
    with open(os.path.join(tmp, "in.fq"), "w") as fh:
        for r in range(1500):
            if r % 4 == 0:
                seq = "".join(rng.choice(list("ACGT"), size=int(rng.integers(2, 6))))
            else:
                seq = TAGS[r % len(TAGS)] + "".join(rng.choice(list("ACGT"), size=int(rng.integers(20, 50))))
            fh.write(f"@s{r}\n{seq}\n+\n{'H' * len(seq)}\n")
    run_pair(tmp, f"-1 {BARC} -2 R:N in.fq", prefix="short")   # (with F + S + B + R the CPU reference segfaults on these reads)


def write_reference_fasta(path, fq_files, rng, n_from_reads=30):
    """Contaminant sequences: windows of some reads (either strand, some with an edit) plus random ones."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    seqs = []
    for fq in fq_files:
        lines = open(fq).read().splitlines()
        reads = lines[1::4]
        for k in range(n_from_reads):
            s = reads[int(rng.integers(0, len(reads)))][:90]
            if k % 2:
                s = "".join(comp[c] for c in reversed(s))
            if k % 3 == 0:
                s = s[:20] + "ACGT"[int(rng.integers(0, 4))] + s[21:]
            pad = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 25))))
            seqs.append(pad + s + pad[::-1])
    for _ in range(6):
        seqs.append("".join(rng.choice(list("ACGT"), size=int(rng.integers(30, 400)))))
    with open(path, "w") as fh:
        for k, s in enumerate(seqs):
            fh.write(f">contaminant_{k} some description\n")
            for i in range(0, len(s), 60):
                fh.write(s[i:i + 60] + "\n")


@pytest.mark.parametrize("threads", [1, 4, 3])
def test_ref_artifact_filter_cli(tmp_path, threads):
    """-ref: reads within -fe edits of a known contaminant (either strand) go to the `un` file and are counted per
    contaminant in the log.  The read comes first in the architecture and is longer than 63 nt (with 5' segments the
    rewritten read starts with spacers, which the reference's filter can never match).  The second file is a plain
    R:N read (run_rna_dust path).  The reference matches per thread slice (groups of four / the rest), so -t matters."""
    tmp = str(tmp_path)
    rng = np.random.default_rng(41)
    make_fastq(os.path.join(tmp, "r1.fq"), 2300, [("R", None), ("B", TAGS)], seed=31, read_len=(70, 90), short=False)
    make_fastq(os.path.join(tmp, "r2.fq"), 2300, [("R", None)], seed=32, read_len=(50, 80), short=False, name_fmt="M1:7:FC:1:{t}:{x}:{y} 2:N:0:1")
    write_reference_fasta(os.path.join(tmp, "contaminants.fa"), [os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")], rng)
    outs = {}
    for tag, binary in (("cpu", os.path.join(REF, "tagdust_rtest")), ("gpu", GPU_BIN)):
        d = os.path.join(tmp, tag)
        os.makedirs(d, exist_ok=True)
        r = subprocess.run(f"{binary} -seed 42 -t {threads} -ref contaminants.fa -fe 2 -1 R:N -2 {BARC} r1.fq r2.fq -o {d}/out", cwd=tmp, shell=True,
                           capture_output=True, text=True)
        assert r.returncode == 0, f"{tag}: {r.stdout}\n{r.stderr}"
        outs[tag] = d
    a = sorted(glob.glob(os.path.join(outs["cpu"], "out*.fq")))
    b = sorted(glob.glob(os.path.join(outs["gpu"], "out*.fq")))
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and a
    for x, y in zip(a, b):
        assert filecmp.cmp(x, y, shallow=False), f"{os.path.basename(x)} differs"
    cpu_log, gpu_log = summary_lines(f"{outs['cpu']}/out_logfile.txt"), summary_lines(f"{outs['gpu']}/out_logfile.txt")
    assert cpu_log == gpu_log
    assert any("contaminant_" in line for line in gpu_log), "no artifact was counted: the test set is too easy"


@pytest.mark.parametrize("ref", [False, True])
def test_chunk_loop_route_matches_reference(tmp_path, ref):
    """The route SAM/BAM input takes (integration/controller_gpu.c chunk_loop: the reference's reader and print_all around
    the GPU run_pHMM, no per-read model rebuild), forced on FASTQ files with TDG_CONTROLLER=chunks because samtools is not
    in this image.  All reads have the same length -- the case in which the reference's own loop rebuilds the model for
    every read -- and the -t 1000-read chunks of the rtest build make several chunks."""
    import time
    tmp = str(tmp_path)
    rng = np.random.default_rng(77)
    make_fastq(os.path.join(tmp, "r1.fq"), 3300, [("R", None), ("B", TAGS)], seed=51, read_len=(72, 72), short=False)
    make_fastq(os.path.join(tmp, "r2.fq"), 3300, [("R", None)], seed=52, read_len=(64, 64), short=False, name_fmt="M1:7:FC:1:{t}:{x}:{y} 2:N:0:1")
    extra = ""
    if ref:
        write_reference_fasta(os.path.join(tmp, "contaminants.fa"), [os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")], rng)
        extra = "-ref contaminants.fa -fe 2 "
    outs, secs = {}, {}
    for tag, binary in (("cpu", os.path.join(REF, "tagdust_rtest")), ("gpu", "TDG_CONTROLLER=chunks TDG_VERBOSE=1 " + GPU_BIN)):
        d = os.path.join(tmp, tag)
        os.makedirs(d, exist_ok=True)
        t0 = time.time()
        r = subprocess.run(f"{binary} -seed 42 -t 4 {extra}-1 R:N -2 {BARC} r1.fq r2.fq -o {d}/out", cwd=tmp, shell=True, capture_output=True, text=True)
        secs[tag] = time.time() - t0
        assert r.returncode == 0, f"{tag}: {r.stdout}\n{r.stderr}"
        if tag == "gpu":
            assert "chunk loop starts" in r.stderr, "the run did not take the chunk-loop route"
        outs[tag] = d
    a = sorted(glob.glob(os.path.join(outs["cpu"], "out*.fq")))
    b = sorted(glob.glob(os.path.join(outs["gpu"], "out*.fq")))
    assert [os.path.basename(x) for x in a] == [os.path.basename(x) for x in b] and a
    for x, y in zip(a, b):
        assert filecmp.cmp(x, y, shallow=False), f"{os.path.basename(x)} differs"
    assert summary_lines(f"{outs['cpu']}/out_logfile.txt") == summary_lines(f"{outs['gpu']}/out_logfile.txt")
    print(f"chunk-loop route: reference {secs['cpu']:.1f} s, drop-in {secs['gpu']:.1f} s")


def test_streaming_on_two_devices_matches_one(tmp_path):
    """tdg_demux_run over a 2-GPU context (every chunk sharded contiguously over the devices) writes the
    same bytes as over one GPU."""
    import filecmp as fc
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cases import TAGS6_ED3
    from refharness import background_logp
    from tagdust_b200 import stream, synth
    from tagdust_b200.api import Context, compile_architecture
    import bench
    tags = TAGS6_ED3[:24]
    desc = compile_architecture(["B:" + ",".join(tags), "R:N"], background_logp((2.5e6, 2.5e6, 2.5e6, 2.5e6, 1.0)), 150.0, 150)
    n = 400_000
    codes, lens, _ = synth.make_reads_fast(n, 150, tags, seed=3)
    fq = str(tmp_path / "in.fq")
    bench.write_fastq_fixed(fq, codes, 150)
    outs = []
    for nd in (1, 2):
        ctx = Context(nd)
        model = ctx.model(desc, 150)
        d = tmp_path / f"d{nd}"
        d.mkdir()
        st = stream.demux_run(ctx, [dict(path=fq, model=model, num_read_segments=1, threshold=1.5, max_seq_len=150)], str(d / "out"),
                              barcode_input=0, barcode_names=list(tags), threads=8, chunk_reads=70_000)
        assert st["total_read"] == n
        outs.append(sorted(glob.glob(str(d / "out*.fq"))))
        model.close(); ctx.close()
    assert len(outs[0]) == 25 and [os.path.basename(x) for x in outs[0]] == [os.path.basename(x) for x in outs[1]]
    for a, b in zip(*outs):
        assert fc.cmp(a, b, shallow=False), os.path.basename(a)


def test_existing_output_files_are_refused_like_the_reference(tmp_path):
    """Second run into the same prefix: the reference refuses (check_for_existing_demultiplexed_files_multiple,
    barcode_hmm.c:150-158); so does the streaming controller, and the first run's files stay untouched."""
    tmp = str(tmp_path)
    make_fastq(os.path.join(tmp, "in.fq"), 1500, [("B", TAGS), ("R", None)], seed=21)
    files = run_pair(tmp, f"-1 {BARC} -2 R:N in.fq")
    before = {f: open(f, "rb").read() for f in files}
    msgs = {}
    for tag, binary in (("cpu", os.path.join(REF, "tagdust_rtest")), ("gpu", GPU_BIN)):
        r = subprocess.run(f"{binary} -seed 42 -t 4 -1 {BARC} -2 R:N in.fq -o {tmp}/{tag}/out", cwd=tmp, shell=True, capture_output=True, text=True)
        msgs[tag] = (r.returncode, "already exists" in (r.stdout + r.stderr))
    assert msgs["cpu"] == msgs["gpu"] and msgs["gpu"][1]
    gpu_files = sorted(glob.glob(os.path.join(tmp, "gpu", "out*.fq")))
    cpu_files = sorted(glob.glob(os.path.join(tmp, "cpu", "out*.fq")))
    for a, b in zip(cpu_files, gpu_files):
        assert open(a, "rb").read() == open(b, "rb").read() == before[a]


def test_reads_longer_than_announced_grow_model_and_batches(tmp_path, gpu_ctx):
    """A file whose later reads are longer than what the set-up announced (get_sequence_stats only looks at the
    first million reads): the pipeline re-sizes the staging batch of that slot and the model's scratch layout
    (the reference re-allocates its model_bag, barcode_hmm.c:293-309) and writes the same bytes as a run that
    knew the lengths from the start."""
    import filecmp as fc
    from refharness import background_logp
    from tagdust_b200 import stream
    from tagdust_b200.api import compile_architecture
    tags = TAGS
    desc = compile_architecture([BARC, "R:N"], background_logp((2.5e6, 2.5e6, 2.5e6, 2.5e6, 1.0)), 60.0, 150)
    rng = np.random.default_rng(5)
    fq = str(tmp_path / "grow.fq")
    with open(fq, "w") as fh:
        for r in range(6000):
            L = 44 if r < 2500 else int(rng.integers(60, 131))
            seq = tags[r % len(tags)] + "".join(rng.choice(list("ACGT"), size=L))
            fh.write(f"@g{r}\n{seq}\n+\n{'F' * len(seq)}\n")
    outs = []
    for announced, model_len in ((50, 60), (136, 146)):
        model = gpu_ctx.model(desc, model_len)
        d = tmp_path / f"run{announced}"
        d.mkdir()
        st = stream.demux_run(gpu_ctx, [dict(path=fq, model=model, num_read_segments=1, threshold=1.0, max_seq_len=announced,
                                             expected_len=announced)], str(d / "out"), barcode_input=0, barcode_names=list(tags),
                              threads=4, chunk_reads=800)
        assert st["total_read"] == 6000
        outs.append(sorted(glob.glob(str(d / "out*.fq"))))
        model.close()
    assert len(outs[0]) == len(tags) + 1
    for a, b in zip(*outs):
        assert fc.cmp(a, b, shallow=False), os.path.basename(a)
