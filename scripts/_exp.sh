cd $GRAFT_REPO_ROOT
for p in 1 0; do KVHBM_BENCH_PRIORITY=$p python bench.py --steps 192 --warmup 10 --no-cpu --no-check 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('prio $p', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['stages'].items()})
"; done
