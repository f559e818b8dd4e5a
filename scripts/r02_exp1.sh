# round 2, GPU call 1: parity after the ADVICE fixes, then the compute-only bounds of both decode kernels and the
# co-resident two-lane experiment.  Output: gpurun_out/r02_exp1_*.  usage: bash scripts/r02_exp1.sh
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
one() {  # one <tag> <lib suffix or ""> [env...]
	local tag=$1 suf=$2; shift 2
	env "$@" TDG_LIB=$PWD/tagdust_b200/libtagdust_b200${suf:+_$suf}.so timeout 600 python bench.py --steps 3 --warmup 3 --reads $((75776*8)) --no-cpu-baseline --no-files --no-configs 2>gpurun_out/r02_exp1_$tag.err | tee gpurun_out/r02_exp1_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$tag', 'value %.3f M/s' % (d['value']/1e6), 'e2e %.3f' % (d['e2e']['value']/1e6), {k:round(v['ms']/v['launches'],3) for k,v in d['kernels_ms'].items()}, 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['check']['read_type_counts'][:2])"
}
for rep in 1 2; do
one base$rep ""
one noload$rep noload
one b256x1_$rep b256
one b256x2_$rep b256 TDG_LANES=2
done
timeout 300 python scripts/exp_bwd_nostore.py 2>&1 | tail -4 | tee gpurun_out/r02_exp1_bwd_nostore.txt
