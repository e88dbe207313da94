cd $GRAFT_REPO_ROOT
for c in ncf dcn streaming sharded200m; do
  f=gpurun_out/r02_bench_$c.json; [ $c = sharded200m ] && f=gpurun_out/r02_bench_sharded200m_n1.json
  timeout 600 python bench.py --config $c > $f 2> gpurun_out/cfg_$c.err || tail -3 gpurun_out/cfg_$c.err
  python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['value']/1e6,2), 'M keys/s', round(d['ms_per_step']*1e3,1), 'us', d.get('cpu_baseline',{}).get('value'))
PY
done
