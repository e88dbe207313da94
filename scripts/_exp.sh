cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for p in 1 0; do KVHBM_PDL=$p python bench.py --steps 192 --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pdl $p', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['stages'].items()}, d['parity_check']['ok'])
"; done
