cd $GRAFT_REPO_ROOT
python scripts/plan_stage.py --stages plan,gather,apply,chain --tag auto 2>&1 | tail -1
for k in 10 12 16; do KVHBM_APPLYP_KPW=$k python scripts/plan_stage.py --stages apply --tag kpw$k 2>&1 | tail -1; done
python bench.py --steps 64 --warmup 5 > gpurun_out/bench2.log 2>&1; tail -c 3000 gpurun_out/bench2.log
