cd $GRAFT_REPO_ROOT
mkdir -p /tmp/ncu
for k in apply_plan_kernel stage_heavy_kernel expand_plan_kernel unique_insert_kernel plan_sort_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -o /tmp/ncu/r02_ncu_$k python scripts/profile_step.py --steps 5 > /tmp/ncu/ncu_$k.log 2>&1
done
ncu --set full --clock-control none -k regex:gather_kernel -s 30 -c 1 -o /tmp/ncu/r02_ncu_gather_kernel python scripts/profile_step.py --steps 5 > /tmp/ncu/ncu_gather_kernel.log 2>&1
python scripts/ncu_summary.py --json /tmp/ncu/*.ncu-rep > gpurun_out/r02_ncu_kernels.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
cp /tmp/ncu/r02_ncu_apply_plan_kernel.ncu-rep gpurun_out/
tail -3 gpurun_out/r02_ncu_kernels.txt; du -sh gpurun_out
