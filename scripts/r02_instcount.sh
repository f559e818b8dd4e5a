cd /root/repo
CMD="python bench.py --steps 1 --warmup 3 --reads 75776 --no-cpu-baseline --no-configs --no-files"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 9 -c 3 --csv --log-file gpurun_out/r02_instcount.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_instcount.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:28], r[-3], r[-2], r[-1])
PY
