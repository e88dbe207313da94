cd $GRAFT_REPO_ROOT
K='regex:apply_plan_kernel|stage_heavy_kernel|gather_kernel|expand_plan_kernel|unique_insert_kernel|unique_rank_kernel|unique_index_kernel|plan_sort_kernel'
ncu --set full --clock-control none --import-source on -k "$K" -s 44 -c 8 -f -o gpurun_out/r02_ncu_step python scripts/profile_step.py --steps 5 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log | cut -c1-200
python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err || tail -5 gpurun_out/n1.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_k20.json 2> gpurun_out/n1b.err || tail -5 gpurun_out/n1b.err
python - <<'PY'
import json
for f in ('r02_bench_n1','r02_bench_n1_k20'):
  d=json.load(open('gpurun_out/%s.json'%f))
  print(f, round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), round(d['e2e']['value']/1e6,1), d['parity_check']['ok'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
