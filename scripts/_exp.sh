cd $GRAFT_REPO_ROOT
python scripts/plan_stage.py --stages plan,gather,apply --tag h32 2>&1 | tail -1
KVHBM_PLAN_HEAVY=16 python scripts/plan_stage.py --stages plan,apply --tag h16 2>&1 | tail -1
KVHBM_PLAN_HEAVY=8 python scripts/plan_stage.py --stages plan,apply --tag h8 2>&1 | tail -1
KVHBM_PLAN_HEAVY=64 python scripts/plan_stage.py --stages plan,apply --tag h64 2>&1 | tail -1
