cd /root/repo
for k in 1 2 3; do
for L in tagdust_b200/libtagdust_b200.so tagdust_b200/libtagdust_b200_pf.so; do
TDG_LIB=$PWD/$L python bench.py --steps 3 --warmup 3 --reads $((75776*8)) --no-cpu-baseline --no-files 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$L', round(d[\"value\"]/1e6,3), {k:round(v[\"ms\"]/v[\"launches\"],3) for k,v in d[\"kernels_ms\"].items()})"
done; done
