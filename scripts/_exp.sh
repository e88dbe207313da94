cd $GRAFT_REPO_ROOT
python bench.py --config dcn --steps 200 --warmup 10 > gpurun_out/r02_bench_dcn.json 2> gpurun_out/cfg_dcn.err || tail -5 gpurun_out/cfg_dcn.err
python bench.py --config streaming --steps 200 --warmup 5 > gpurun_out/r02_bench_streaming.json 2> gpurun_out/cfg_streaming.err || tail -5 gpurun_out/cfg_streaming.err
for c in dcn streaming; do python -c "
import json,sys
d=json.load(open('gpurun_out/r02_bench_$c.json'))
print('$c', round(d['value']/1e6,2),'M keys/s', round(d['ms_per_step']*1e3,1),'us frac', round(d['roofline']['frac'],4), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('checkpoint_round_trips'), d.get('keys_evicted'), d.get('table_size_after'), d.get('wall_ms_per_step'))
"; done
