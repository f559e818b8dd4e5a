# same-box A/B of several builds of the library: bash scripts/gpu_ab.sh <suffix> [<suffix> ...]  (libtagdust_b200_<suffix>.so)
cd /root/repo
for k in 1 2; do
for S in "" "$@"; do
L=tagdust_b200/libtagdust_b200${S:+_$S}.so
TDG_LIB=$PWD/$L python bench.py --steps 3 --warmup 3 --reads $((75776*8)) --no-cpu-baseline --no-files --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${S:-base}', round(d[\"value\"]/1e6,3), {k:round(v[\"ms\"]/v[\"launches\"],3) for k,v in d[\"kernels_ms\"].items()})"
done; done
