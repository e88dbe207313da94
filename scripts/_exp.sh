cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_plan.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python scripts/plan_stage.py --stages apply,chain --tag guided 2>&1 | tail -1
KVHBM_APPLYP_KPW=8 python scripts/plan_stage.py --stages apply --tag guided_kpw8 2>&1 | tail -1
KVHBM_APPLYP_KPW=6 python scripts/plan_stage.py --stages apply --tag guided_kpw6 2>&1 | tail -1
