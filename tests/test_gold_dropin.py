"""The reference's own gold tests (dev/bar_read_test.sh, dev/casava_test.sh) through the drop-in
binary: the UNMODIFIED reference CLI with run_pHMM() interposed by integration/run_phmm_gpu.c and
hmm_controller_multiple() by integration/controller_gpu.c (both controllers are exercised).
Needs a GPU and the prebuilt integration/_build + oracle/_ref (they travel with the snapshot)."""
import filecmp
import glob
import gzip
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GPU_BIN = os.path.join(ROOT, "integration", "_build", "tagdust_gpu_rtest")
GOLD = os.path.join(ROOT, "tests", "golden")
TAGS = os.path.join(GOLD, "edittag_6nt_ed4.txt")


def sh(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, shell=True, capture_output=True, text=True)
    assert r.returncode == 0, f"{cmd}\n{r.stdout}\n{r.stderr}"
    return r


def need_bins():
    for p in (GPU_BIN, os.path.join(REF, "simreads_rtest"), os.path.join(REF, "evalres_rtest"), os.path.join(REF, "tagdust_rtest")):
        if not os.path.exists(p):
            pytest.skip(f"{p} not built")


SIM = ("{ref}/simreads_rtest {tags} -seed 42 -sim_barnum {bn} {link} -sim_readlen 20 -sim_readlen_mod 0 -sim_numseq 10000 "
       "-sim_endloss 0 -sim_random_frac {rf} -o {out} -sim_error_rate 0.02")


def sorted_lines(path):
    with open(path) as fh:
        return sorted(fh.readlines())


#: "stream": hmm_controller_multiple replaced by integration/controller_gpu.c (streaming ingest + demux writer);
#: "reference": the reference's own controller loop with only run_pHMM interposed (integration/run_phmm_gpu.c)
CONTROLLERS = ["stream", "reference"]


def run_both(tmp, args, prefix, controller="stream"):
    """Run the CPU reference and the GPU drop-in with identical arguments; return their output dirs."""
    outs = {}
    env = "TDG_REFERENCE_CONTROLLER=1 " if controller == "reference" else "TDG_CHUNK_READS=700 "
    for tag, binary in (("cpu", os.path.join(REF, "tagdust_rtest")), ("gpu", env + GPU_BIN)):
        d = os.path.join(tmp, tag)
        os.makedirs(d, exist_ok=True)
        sh(f"{binary} -seed 42 {args} -o {d}/{prefix}", tmp)
        outs[tag] = d
    return outs


def summary_lines(path):
    """The run summary of a log file without timestamps, the command line and the per-read
    'Long sequence found' lines (documented difference of the streaming controller)."""
    out = []
    with open(path) as fh:
        for line in fh:
            t = line.split("\t", 1)[1] if line.startswith("[") and "\t" in line else line
            if t.startswith("cmd:") or t.startswith("Long sequence found"):
                continue
            out.append(t)
    return out


def assert_same_outputs(outs, prefix):
    cpu = sorted(glob.glob(os.path.join(outs["cpu"], prefix + "*.fq")))
    gpu = sorted(glob.glob(os.path.join(outs["gpu"], prefix + "*.fq")))
    assert [os.path.basename(x) for x in cpu] == [os.path.basename(x) for x in gpu] and cpu
    for a, b in zip(cpu, gpu):
        assert filecmp.cmp(a, b, shallow=False), f"{os.path.basename(a)} differs between CPU reference and GPU drop-in"


@pytest.mark.parametrize("controller", CONTROLLERS)
@pytest.mark.parametrize("case", ["barread1", "barread2"])
def test_bar_read_single_end(tmp_path, case, controller):
    need_bins()
    tmp = str(tmp_path)
    link = "" if case == "barread1" else "-sim_5seq GGGGGGG -sim_3seq TTTTTTT"
    sh(SIM.format(ref=REF, tags=TAGS, bn=4, link=link, rf=0.1, out=f"{case}.fq"), tmp)
    outs = run_both(tmp, f"{case}.fq -arch {case}.fq_tagdust_arch.txt", f"{case}_tagdust", controller)
    assert_same_outputs(outs, f"{case}_tagdust")
    assert summary_lines(f"{outs['cpu']}/{case}_tagdust_logfile.txt") == summary_lines(f"{outs['gpu']}/{case}_tagdust_logfile.txt")
    sh(f"{REF}/evalres_rtest -name tagdust {outs['gpu']}/{case}_tagdust*.fq -o {outs['gpu']}/{case}_tagdust", tmp)
    assert sorted_lines(f"{outs['gpu']}/{case}_tagdust_results.txt") == sorted_lines(os.path.join(GOLD, f"{case}_tagdust_results_gold.txt"))


@pytest.mark.parametrize("controller", CONTROLLERS)
@pytest.mark.parametrize("case", ["read_paired", "barread_paired"])
def test_bar_read_paired(tmp_path, case, controller):
    need_bins()
    tmp = str(tmp_path)
    bn = 0 if case == "read_paired" else 4
    sh(SIM.format(ref=REF, tags=TAGS, bn=bn, link="-sim_5seq GGGGGGG -sim_3seq TTTTTTT", rf=0.1, out="r1.fq"), tmp)
    sh(SIM.format(ref=REF, tags=TAGS, bn=0, link="", rf="0.00", out="r2.fq"), tmp)
    sh("cat r1.fq_tagdust_arch.txt r2.fq_tagdust_arch.txt > combo_arch.txt", tmp)
    outs = run_both(tmp, "-sim_numseq 1 r1.fq r2.fq -arch combo_arch.txt", f"{case}_tagdust", controller)
    assert_same_outputs(outs, f"{case}_tagdust")
    assert summary_lines(f"{outs['cpu']}/{case}_tagdust_logfile.txt") == summary_lines(f"{outs['gpu']}/{case}_tagdust_logfile.txt")
    sh(f"{REF}/evalres_rtest -name tagdust {outs['gpu']}/{case}_tagdust_*READ1.fq -o {outs['gpu']}/{case}_tagdust", tmp)
    assert sorted_lines(f"{outs['gpu']}/{case}_tagdust_results.txt") == sorted_lines(os.path.join(GOLD, f"{case}_tagdust_results_gold.txt"))


@pytest.mark.parametrize("controller", CONTROLLERS)
def test_casava_derived(tmp_path, controller):
    """dev/casava_test.sh: READ1/READ3 blobs are missing upstream, so constant 76-nt
    stand-ins carrying read2's names are used; the 1582 gold TTAGGC names must all be assigned to
    TTAGGC, and CPU reference and GPU drop-in must write identical files."""
    need_bins()
    tmp = str(tmp_path)
    shutil.copy(os.path.join(GOLD, "casava_read2.fastq.gz"), tmp)
    shutil.copy(os.path.join(GOLD, "casava_arch.txt"), tmp)
    body = "ACGTTGCAAGTCCGATAGCTTAGGCATCGATCGGATCCTAGCTAGGATCGATTAGCGCTAGGCTAACGTAGCTAGCA"[:76]
    with gzip.open(os.path.join(tmp, "casava_read2.fastq.gz"), "rt") as fh, \
            open(os.path.join(tmp, "r1.fq"), "w") as o1, open(os.path.join(tmp, "r3.fq"), "w") as o3:
        for k, line in enumerate(fh):
            if k % 4 == 0:
                name = line.rstrip("\n")
                base, rest = name.split(" ", 1)
                o1.write(f"{base} 1{rest[1:]}\n{body}\n+\n{'I' * 76}\n")
                o3.write(f"{base} 3{rest[1:]}\n{body}\n+\n{'I' * 76}\n")
    outs = run_both(tmp, "-arch casava_arch.txt r1.fq casava_read2.fastq.gz r3.fq", "casava_out", controller)
    assert_same_outputs(outs, "casava_out")
    assert summary_lines(f"{outs['cpu']}/casava_out_logfile.txt") == summary_lines(f"{outs['gpu']}/casava_out_logfile.txt")
    got = set()
    with open(os.path.join(outs["gpu"], "casava_out_BC_TTAGGC_READ2.fq")) as fh:
        for k, line in enumerate(fh):
            if k % 4 == 0:
                got.add(line.split(" ")[0].split(";")[0])
    with open(os.path.join(GOLD, "casava_gold_TTAGGC_names.txt")) as fh:
        gold = {x.strip() for x in fh if x.strip()}
    assert len(gold) == 1582 and gold <= got
    assert len(got - gold) <= 2   # the two NNNNNC index reads the real R1/R3 would have dust-filtered
