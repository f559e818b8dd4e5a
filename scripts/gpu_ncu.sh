# usage: bash scripts/gpu_ncu.sh <tag>   -- launch list + full capture of the decode kernels (1 wave)
set -x
cd /root/repo
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --reads 75776 --no-cpu-baseline --no-configs --no-files"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_forward|k_backward|k_label' -s 9 -c 3 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/
