cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_plan.py -x -q -m gpu 2>&1 | tail -3
python scripts/plan_stage.py --stages segsum,apply --tag tight 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:apply_plan_kernel -s 20 -c 1 -o gpurun_out/r02_apply_plan_v1 python scripts/plan_stage.py --stages apply --eager --steps 8 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
