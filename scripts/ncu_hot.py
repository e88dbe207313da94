#!/usr/bin/env python
"""Top stall sites (SASS) of a kernel in an ncu report: python scripts/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"]).decode()
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "# Samples" in r)
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows if len(r) > si and r[si].isdigit()]
seen, uniq = set(), []
for r in data:
  k = (r[0], r[src])
  if k in seen: continue
  seen.add(k); uniq.append(r)
tot = sum(int(r[si]) for r in uniq)
print("total samples", tot, "instructions", len(uniq))
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ci = [hdr.index(c) for c in cols]
agg = {c: sum(int(r[i] or 0) for r in uniq) for c, i in zip(cols, ci)}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(uniq, key=lambda r: -int(r[si]))[:topn]:
  print(r[si].rjust(6), r[ie].rjust(8), r[src][:86].ljust(86),
        {c[6:]: r[i] for c, i in zip(cols, ci) if r[i] not in ("0", "")})
