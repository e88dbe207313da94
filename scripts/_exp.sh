cd $GRAFT_REPO_ROOT
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err || tail -5 gpurun_out/n1.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_k20.json 2> gpurun_out/n1b.err || tail -5 gpurun_out/n1b.err
python - <<'PY'
import json
for f in ('r02_bench_n1','r02_bench_n1_k20'):
  d=json.load(open('gpurun_out/%s.json'%f))
  print(f, round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), round(d['e2e']['value']/1e6,1), d['parity_check']['ok'], d['roofline']['frac'], d['cpu_baseline']['value'], d['schedule'][-80:])
PY
python -c "import __graft_entry__ as g; g.smoke()"
