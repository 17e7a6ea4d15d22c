"""Per-rank pinned host-to-device bandwidth with N ranks copying at once (why does bench.py's e2e stop scaling past 2 GPUs?).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_bw.py [--bind]
--bind: before allocating, pin the process to the CPU cores local to its GPU (sysfs local_cpulist) so that the pinned buffer is
first-touched on the GPU's NUMA node."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from lv_slam_b200 import dist as D

rank, local, world = D.env_rank()
torch.cuda.set_device(local)
bind = "--bind" in sys.argv
info = ""
if bind:
    info = D.bind_to_gpu_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)                                  # touch every page on this rank's node
d = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
for _ in range(3):
    with torch.cuda.stream(st): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
with torch.cuda.stream(st):
    ev0.record(st)
    for _ in range(reps): d.copy_(h, non_blocking=True)
    ev1.record(st)
torch.cuda.synchronize()
gbs = reps * n / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
allv = [None] * world
if world > 1: dist.all_gather_object(allv, (rank, gbs, info))
else: allv = [(rank, gbs, info)]
if rank == 0:
    print("ranks %d  bind %s  per-rank H2D GB/s: %s  | sum %.1f  min %.1f" % (world, bind, " ".join("%.1f" % v[1] for v in sorted(allv)), sum(v[1] for v in allv), min(v[1] for v in allv)))
    if bind:
        for v in sorted(allv): print("   rank %d: %s" % (v[0], v[2]))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
