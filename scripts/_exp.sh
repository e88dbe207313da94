cd $GRAFT_REPO_ROOT
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 208 --warmup 16 --no-cpu --no-check > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err || tail -3 gpurun_out/exp_$name.err
  python - $name <<'PY'
import json,sys
d=json.load(open('gpurun_out/exp_%s.json'%sys.argv[1]))
st=d['roofline']['stages']
print(sys.argv[1], 'step', round(d['ms_per_step']*1e3,1), 'strict', round(d['strict_per_step']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in st.items()})
PY
}
run ring9
run ring9_k8 KVHBM_APPLYP_KPW=8
python -m pytest tests/test_gpu_plan.py -x -q -m gpu 2>&1 | tail -1
cp tfplus_b200/build/libkvhbm_trace.so tfplus_b200/libkvhbm.so
python scripts/trace_apply_plan.py > gpurun_out/trace_ring9.log 2>&1
tail -12 gpurun_out/trace_ring9.log
