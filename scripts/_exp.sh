cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/b.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r)>10]
h=rows[0]; iN=h.index('Kernel Name'); iV=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]:
  try: d[r[iN].split('(')[0][-60:]].append(float(r[iV].replace(',','')))
  except: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])):
  print('%-62s n=%4d  median %8.1f us  total %9.1f' % (k, len(v), sorted(v)[len(v)//2]/1e3 if max(v)>1e3 else sorted(v)[len(v)//2], sum(v)))
PY
