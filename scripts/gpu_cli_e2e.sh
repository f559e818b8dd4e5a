# usage: bash scripts/gpu_cli_e2e.sh <nreads> [ncmp]
# CLI-level end to end on a simreads cfg2 file: FASTQ file -> drop-in binary (reference CLI + interposed
# run_pHMM and hmm_controller_multiple) -> demultiplexed files; plus a byte comparison with the CPU reference
# on a prefix of the file.
set -x
cd /root/repo
N=${1:-4000000}
NCMP=${2:-60000}
W=/tmp/cli_e2e; rm -rf $W; mkdir -p $W gpurun_out
REF=oracle/_ref
$REF/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq $N -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq > /dev/null 2>&1
ls -la $W/syn48.fq; nproc
ARCH=$W/syn48.fq_tagdust_arch.txt
# (1) throughput of the whole tool, fixed threshold (no calibration phase) and with calibration
for Q in "-Q 1.5" ""; do
  T0=$(date +%s.%N)
  timeout 600 env TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) $Q -arch $ARCH $W/syn48.fq -o $W/gpu_full > $W/full.log 2>&1
  echo "rc=$? Q='$Q' wall $(python3 -c "import time,sys; print(round(time.time()-float(sys.argv[1]),2))" $T0) s for $N reads"; tail -4 $W/full.log; grep -E "total input|extracted|Threshold|threshold" $W/gpu_full_logfile.txt
  rm -f $W/gpu_full*
done
# (2) byte comparison with the CPU reference on a prefix
head -n $((NCMP*4)) $W/syn48.fq > $W/small.fq
mkdir -p $W/cpu $W/gpu
T0=$(date +%s.%N); timeout 900 $REF/tagdust -t $(nproc) -Q 1.5 -arch $ARCH $W/small.fq -o $W/cpu/out > /dev/null 2>&1; echo "rc=$? cpu reference wall $(python3 -c "import time,sys; print(round(time.time()-float(sys.argv[1]),2))" $T0) s for $NCMP reads"
T0=$(date +%s.%N); timeout 300 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $ARCH $W/small.fq -o $W/gpu/out > /dev/null 2>&1; echo "rc=$? gpu drop-in wall $(python3 -c "import time,sys; print(round(time.time()-float(sys.argv[1]),2))" $T0) s for $NCMP reads"
nd=0; for f in $W/cpu/*.fq; do cmp -s $f $W/gpu/$(basename $f) || { echo DIFF $(basename $f); nd=$((nd+1)); }; done
echo "files compared: $(ls $W/cpu/*.fq | wc -l), differing: $nd"
grep -E "total input|successfully" $W/cpu/out_logfile.txt $W/gpu/out_logfile.txt
