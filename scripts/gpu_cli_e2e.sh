# usage: bash scripts/gpu_cli_e2e.sh <nreads>  -- CLI-level end to end: reference CLI + interposed run_pHMM on a simreads cfg2 file
set -x
cd /root/repo
N=${1:-2000000}
W=/tmp/cli_e2e; mkdir -p $W gpurun_out
REF=oracle/_ref
( time $REF/simreads tests/golden/edittag_6nt_ed3.txt -seed 7 -sim_barnum 48 -sim_readlen 144 -sim_readlen_mod 0 -sim_numseq $N -sim_endloss 0 -sim_random_frac 0.05 -sim_error_rate 0.01 -o $W/syn48.fq ) 2>&1 | tail -4
ls -la $W | head; nproc
head -c 400 $W/syn48.fq_tagdust_arch.txt; echo
( time integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $W/syn48.fq_tagdust_arch.txt $W/syn48.fq -o $W/gpu_out ) > gpurun_out/cli_gpu.log 2>&1
tail -5 gpurun_out/cli_gpu.log
cat $W/gpu_out_logfile.txt | tail -30
head -n 200000 $W/syn48.fq > $W/small.fq
( time $REF/tagdust -t $(nproc) -Q 1.5 -arch $W/syn48.fq_tagdust_arch.txt $W/small.fq -o $W/cpu_out ) > gpurun_out/cli_cpu.log 2>&1
tail -5 gpurun_out/cli_cpu.log
cat $W/cpu_out_logfile.txt | tail -12
