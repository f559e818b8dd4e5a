# usage: bash scripts/r02_cli.sh   -- CLI-level timings of round 2 (drop-in binary = reference CLI + integration/*.c):
#   (1) whole tool on 8 M cfg2 reads, fixed threshold and with calibration
#   (2) -ref artifact filter on 2 M reads (R-first architecture so that the filter can fire), byte comparison with the
#       CPU reference on a prefix
#   (3) cfg5 from the CLI: -arch file with 64 candidate architectures on a 1 M-read cfg2 file
# Output: gpurun_out/r02_cli.txt
cd /root/repo
W=/dev/shm/r02_cli; rm -rf $W; mkdir -p $W gpurun_out
REF=oracle/_ref
OUT=gpurun_out/r02_cli.txt; : > $OUT
wall() { python3 -c "import time,sys; print(round(time.time()-float(sys.argv[1]),2))" $1; }
$REF/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq 8000000 -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq > /dev/null 2>&1
ARCH=$W/syn48.fq_tagdust_arch.txt
echo "host cores $(nproc); input $(ls -la $W/syn48.fq | awk '{print $5}') bytes" | tee -a $OUT
for Q in "-Q 1.5" ""; do
	T0=$(date +%s.%N)
	timeout 600 env TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) $Q -arch $ARCH $W/syn48.fq -o $W/full > $W/full.log 2>&1
	echo "(1) rc=$? 8M cfg2 reads, Q='$Q': wall $(wall $T0) s" | tee -a $OUT; grep "reads in" $W/full.log | tee -a $OUT
	rm -f $W/full*
done
# (2) -ref
python - <<PY
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from tagdust_b200 import synth
segs, tags = bench.architecture()
rng = np.random.default_rng(3)
n = 2_000_000
codes = np.zeros((n, 160), np.uint8)
codes[:, :150] = rng.integers(0, 4, size=(n, 150), dtype=np.uint8)
bc = np.stack([synth.encode(t) for t in tags])
codes[:, 144:150] = bc[rng.integers(0, len(tags), size=n)]
bench.write_fastq_fixed("$W/rfirst.fq", codes, 150)
alpha = np.frombuffer(b"ACGT", np.uint8)
with open("$W/contaminants.fa", "w") as fh:
    for k in range(40):   # 40 sequences, ~12 kb in total: some are windows of reads, some random
        if k < 20:
            r = int(rng.integers(0, n)); s = alpha[codes[r, :110]].tobytes().decode()
        else:
            s = alpha[rng.integers(0, 4, size=int(rng.integers(200, 800)))].tobytes().decode()
        fh.write(f">contaminant_{k}\n{s}\n")
open("$W/arch_rfirst.txt", "w").write("tagdust -1 R:N -2 B:" + ",".join(tags) + "\n")
PY
T0=$(date +%s.%N)
timeout 600 env TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -ref $W/contaminants.fa -arch $W/arch_rfirst.txt $W/rfirst.fq -o $W/refrun > $W/ref.log 2>&1
echo "(2) rc=$? -ref on 2M reads (40 contaminant sequences): wall $(wall $T0) s" | tee -a $OUT; grep "reads in" $W/ref.log | tee -a $OUT; grep -E "match artifacts|contaminant_" $W/refrun_logfile.txt | head -5 | tee -a $OUT
T0=$(date +%s.%N)
timeout 600 env TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $W/arch_rfirst.txt $W/rfirst.fq -o $W/norefrun > $W/noref.log 2>&1
echo "(2) rc=$? same run without -ref: wall $(wall $T0) s" | tee -a $OUT; grep "reads in" $W/noref.log | tee -a $OUT
head -n 80000 $W/rfirst.fq > $W/small.fq
mkdir -p $W/cpu $W/gpu
T0=$(date +%s.%N); timeout 900 $REF/tagdust -t $(nproc) -Q 1.5 -ref $W/contaminants.fa -arch $W/arch_rfirst.txt $W/small.fq -o $W/cpu/out > /dev/null 2>&1; echo "(2) rc=$? cpu reference -ref on 20000 reads: wall $(wall $T0) s" | tee -a $OUT
timeout 300 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -ref $W/contaminants.fa -arch $W/arch_rfirst.txt $W/small.fq -o $W/gpu/out > /dev/null 2>&1
nd=0; for f in $W/cpu/*.fq; do cmp -s $f $W/gpu/$(basename $f) || { echo DIFF $(basename $f); nd=$((nd+1)); }; done
echo "(2) files compared: $(ls $W/cpu/*.fq | wc -l), differing: $nd" | tee -a $OUT
grep -E "match artifacts" $W/cpu/out_logfile.txt $W/gpu/out_logfile.txt | tee -a $OUT
# (3) cfg5 from the CLI
python - <<PY
import sys
sys.path.insert(0, "/root/repo")
import bench
from tagdust_b200 import synth
tags = synth.load_tags(bench.TAGS)
with open("$W/arch64.txt", "w") as fh:
    for a in synth.candidate_architectures(tags, 64):
        fh.write("tagdust " + " ".join(f"-{k + 1} {s}" for k, s in enumerate(a)) + "\n")
PY
head -n 4000000 $W/syn48.fq > $W/syn1m.fq
T0=$(date +%s.%N)
timeout 900 env TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $W/arch64.txt $W/syn1m.fq -o $W/a64 > $W/a64.log 2>&1
echo "(3) rc=$? cfg5 from the CLI: 64-line -arch file, 1M-read cfg2 file, whole tool: wall $(wall $T0) s" | tee -a $OUT
grep -E "run_pHMM mode 5|reads in" $W/a64.log | tee -a $OUT; grep -E "Confidence|Using" -A1 $W/a64_logfile.txt | head -6 | tee -a $OUT
rm -rf $W
