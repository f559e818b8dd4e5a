# usage: bash scripts/r02_chunkroute.sh  -- the SAM/BAM route of the drop-in (controller_gpu.c chunk_loop), forced on a FASTQ
# file with TDG_CONTROLLER=chunks: parity tests, then 2 M fixed-length cfg2 reads through (a) the streaming controller,
# (b) the chunk loop (reference reader + print_all around the GPU run_pHMM), (c) the reference's own loop with the GPU
# run_pHMM (TDG_REFERENCE_CONTROLLER=1; one model rebuild per read), bounded to 120 s.   Output: gpurun_out/r02_chunkroute.txt
cd /root/repo
W=/dev/shm/r02_cr; rm -rf $W; mkdir -p $W gpurun_out
OUT=gpurun_out/r02_chunkroute.txt; : > $OUT
python -m pytest tests/test_gpu_stream_cli.py tests/test_gold_dropin.py -x -q -m gpu > gpurun_out/r02_chunkroute_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT; tail -3 gpurun_out/r02_chunkroute_tests.log | tee -a $OUT
wall() { python3 -c "import time,sys; print(round(time.time()-float(sys.argv[1]),2))" $1; }
oracle/_ref/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq 2000000 -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq > /dev/null 2>&1
ARCH=$W/syn48.fq_tagdust_arch.txt
echo "host cores $(nproc); input $(ls -la $W/syn48.fq | awk '{print $5}') bytes" | tee -a $OUT
for mode in stream chunks reference; do
	case $mode in stream) E="TDG_VERBOSE=1";; chunks) E="TDG_VERBOSE=1 TDG_CONTROLLER=chunks";; reference) E="TDG_VERBOSE=1 TDG_REFERENCE_CONTROLLER=1";; esac
	mkdir -p $W/$mode
	T0=$(date +%s.%N)
	timeout 120 env $E integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $ARCH $W/syn48.fq -o $W/$mode/out > $W/$mode.log 2>&1
	echo "$mode: rc=$? (124 = stopped after 120 s) wall $(wall $T0) s; $(cat $W/$mode/out*.fq 2>/dev/null | wc -l | awk '{print $1/4}') reads written" | tee -a $OUT
	grep -E "chunk loop|streaming job|reads in" $W/$mode.log | tee -a $OUT
done
nd=0; for f in $W/stream/*.fq; do cmp -s $f $W/chunks/$(basename $f) || { echo DIFF $(basename $f); nd=$((nd+1)); }; done
echo "stream vs chunks: files $(ls $W/stream/*.fq | wc -l), differing $nd" | tee -a $OUT
rm -rf $W
