set -x
cd /root/repo
N=${1:-4000000}
W=/tmp/cli_e2e; rm -rf $W; mkdir -p $W gpurun_out
oracle/_ref/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq $N -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq > /dev/null 2>&1
for T in 16 8; do
TDG_TRACE=1 TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $T -Q 1.5 -arch $W/syn48.fq_tagdust_arch.txt $W/syn48.fq -o $W/o$T 2>&1 | grep -E "trace|tagdust_b200:" > gpurun_out/stream_trace_$T.txt
tail -1 gpurun_out/stream_trace_$T.txt
rm -f $W/o$T*
done
dd if=$W/syn48.fq of=/dev/null bs=16M 2>&1 | tail -1
dd if=$W/syn48.fq of=$W/copy bs=16M 2>&1 | tail -1
