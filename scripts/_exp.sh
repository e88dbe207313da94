cd $GRAFT_REPO_ROOT
for k in 4 8 16; do KVHBM_APPLYP_KPW=$k python scripts/plan_stage.py --stages apply --tag kpw$k 2>&1 | tail -1; done
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
