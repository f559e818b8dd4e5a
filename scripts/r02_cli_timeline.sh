# usage: bash scripts/r02_cli_timeline.sh [nreads]  -- where the wall time of one CLI run goes (TDG_VERBOSE phase marks)
cd /root/repo
N=${1:-8000000}
W=/dev/shm/r02_tl; rm -rf $W; mkdir -p $W gpurun_out
python - <<PY
import sys
sys.path.insert(0, "/root/repo")
import bench
from tagdust_b200 import synth
segs, tags = bench.architecture()
codes, lens, _ = synth.make_reads_fast($N // 4, 150, tags, seed=3)
bench.write_fastq_fixed("$W/in.fq", codes, 150, 4)
open("$W/arch.txt", "w").write("tagdust -1 " + segs[0] + " -2 R:N\n")
PY
for rep in 1 2; do
T0=$(date +%s.%N)
TDG_VERBOSE=1 integration/_build/tagdust_gpu -t $(nproc) -Q 1.5 -arch $W/arch.txt $W/in.fq -o $W/out$rep > $W/log$rep.txt 2>&1
T1=$(date +%s.%N)
echo "run $rep: shell start epoch $T0, end $T1, wall $(python3 -c "print(round($T1-$T0,3))") s for $N reads"
grep "tagdust_b200" $W/log$rep.txt | grep -v "calibration chunk"
done 2>&1 | tee gpurun_out/r02_cli_timeline.txt
rm -rf $W
