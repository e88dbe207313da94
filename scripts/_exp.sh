cd $GRAFT_REPO_ROOT
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$T bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/n$N.err || tail -5 gpurun_out/n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.load(open('gpurun_out/r02_bench_n%s.json'%N))
print('n'+N, round(d['value']/1e9,3), round(d['ms_per_step']*1e3,1), round(d['strict_per_step']['ms_per_step']*1e3,1), round(d['e2e']['value']/1e6,1), d['parity_check']['ok'], d['nvlink']['bus_gbs_per_gpu'])
PY
python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -2
KVHBM_SHARDED_GRAPH=1 PIPE=1 KEYS=10000000 timeout 300 $T scripts/profile_sharded.py > gpurun_out/r02_sharded_timeline.txt 2>&1
