cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err || tail -5 gpurun_out/n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print('n1', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), round(d['e2e']['value']/1e6,1), d['parity_check']['ok'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
