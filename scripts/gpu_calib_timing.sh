set -x
cd /root/repo
W=/tmp/calib; rm -rf $W; mkdir -p $W
oracle/_ref/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq 200000 -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq > /dev/null 2>&1
date +%s.%N
TDG_TRACE=1 TDG_VERBOSE=1 integration/_build/tagdust_gpu -t 16 -arch $W/syn48.fq_tagdust_arch.txt $W/syn48.fq -o $W/o 2>&1 | grep -E "tagdust_b200|Threshold|threshold|trace" | head -40
date +%s.%N
