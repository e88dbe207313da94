#!/usr/bin/env bash
# What the round's evidence was produced with, on a GPU box (from the repo root):
#   bash scripts/run_on_box.sh [N]      N = 1 (default), 2, 4 or 8 GPUs
# GPU tests, smoke(), the bench line for N GPUs (-> gpurun_out/r02_bench_nN.json) and, at N = 1,
# the ncu launch list and one `ncu --set full` capture of every kernel of a step.
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  python -m pytest tests -x -q -m gpu 2>&1 | tail -2
  python -c "import __graft_entry__ as g; g.smoke()"
  python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err || tail -5 gpurun_out/n1.err
  K='regex:apply_plan_kernel|stage_heavy_kernel|gather_kernel|expand_plan_kernel|unique_insert_kernel|unique_rank_kernel|unique_index_kernel|plan_sort_kernel'
  ncu --set full --clock-control none --import-source on -k "$K" -s 44 -c 8 -f -o gpurun_out/r02_ncu_step \
      python scripts/profile_step.py --steps 5 > gpurun_out/ncu_step.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/launches.log 2>&1
else
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
  $T bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/n$N.err || tail -5 gpurun_out/n$N.err
  python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -2
fi
python - "$N" <<'PY'
import json, sys
d = json.load(open("gpurun_out/r02_bench_n%s.json" % sys.argv[1]))
print("n" + sys.argv[1], "%.3f G keys/s" % (d["value"] / 1e9), "%.1f us/step" % (d["ms_per_step"] * 1e3),
      "strict %.1f" % (d["strict_per_step"]["ms_per_step"] * 1e3), "e2e %.1f M keys/s" % (d["e2e"]["value"] / 1e6),
      "parity", d["parity_check"]["ok"], "roofline frac %.3f" % d["roofline"]["frac"])
PY
