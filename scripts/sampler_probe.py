"""How much does NVML polling disturb the sharded step?  (torchrun, N >= 2)

Replays the captured sharded step while rank 0 polls NVML in a thread, one query kind at a
time, and prints the step time next to the host duration of the queries.
"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from tfplus_b200 import ops, sharded


def main():
  rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
  lr = int(os.environ.get("LOCAL_RANK", rank))
  torch.cuda.set_device(lr)
  dev = torch.device("cuda", lr)
  dist.init_process_group("nccl", device_id=dev)
  ops.set_today(bench.TODAY)
  keys, B, D = 2_000_000, 65536, 64
  st = sharded.ShardedStepper(keys, D, B, bench.HP, dev, rank, world)
  st.populate()
  ids_np, grads_np = bench.make_batches(bench.N_BATCHES, keys * world, B, D, seed_ids=2024 + rank,
                                        seed_grad=7 + rank)
  st.prepare([torch.from_numpy(x).to(dev) for x in ids_np],
             [torch.from_numpy(x).to(dev) for x in grads_np])
  import pynvml as nv
  nv.nvmlInit()
  h = nv.nvmlDeviceGetHandleByIndex(lr)
  kinds = {
      "none": [],
      "clock_info": [lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)],
      "clock_current": [lambda: nv.nvmlDeviceGetClock(h, nv.NVML_CLOCK_SM, nv.NVML_CLOCK_ID_CURRENT)],
      "reasons": [lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)],
      "power": [lambda: nv.nvmlDeviceGetPowerUsage(h)],
      "both": [lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
               lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)],
      "sleep_only": [lambda: None],
  }
  for name, fns in kinds.items():
    for period in (0.002, 0.02):
      if name == "none" and period != 0.002:
        continue
      stop = threading.Event()
      durs = []
      def run():
        while not stop.is_set():
          for f in fns:
            t0 = time.perf_counter(); f(); durs.append(time.perf_counter() - t0)
          time.sleep(period)
      th = None
      dist.barrier(); torch.cuda.synchronize()
      if rank == 0 and fns:
        th = threading.Thread(target=run, daemon=True); th.start()
      t = st.stage_times(400)["sharded step"]
      stop.set()
      if th: th.join()
      if rank == 0:
        print("%-14s period %4.0f ms: %7.1f us/step  samples %3d  mean call %6.1f us  max %7.1f us"
              % (name, period * 1e3, t * 1e3, len(durs), (np.mean(durs) if durs else 0) * 1e6,
                 (np.max(durs) if durs else 0) * 1e6), flush=True)
  st.release()
  torch.cuda.synchronize()
  dist.barrier()
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
