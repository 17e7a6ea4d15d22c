"""Where one bench step spends its time: phases of the batched NDT step timed separately (host clock around synchronised phases)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import lv_slam_b200 as L
import bench as Bn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scans, poses, plan, keys = Bn.make_workload(B, 0)
key_slot = {k: i for i, k in enumerate(keys)}
stream = torch.cuda.Stream()
nb = L.NdtBatch(len(keys), B, device=0, stream=stream.cuda_stream, transformation_epsilon=0.01, max_iterations=64, variant=0, search_method=2)
src_slots = list(range(B)); tgt_slots = [key_slot[k] for _, k, _ in plan]; guesses = [g for _, _, g in plan]
dev_src = [torch.from_numpy(scans[f]).cuda() for f, _, _ in plan]; dev_tgt = [torch.from_numpy(scans[k]).cuda() for k in keys]
pin_src = [torch.from_numpy(scans[f]).pin_memory() for f, _, _ in plan]; pin_tgt = [torch.from_numpy(scans[k]).pin_memory() for k in keys]
sync = torch.cuda.synchronize

def phases(src, tgt, reps=10):
    acc = np.zeros(4)
    for r in range(reps + 2):
        sync(); t0 = time.perf_counter()
        nb.set_targets(list(range(len(keys))), tgt); t1h = time.perf_counter(); sync(); t1 = time.perf_counter()
        nb.set_sources(src_slots, src); t2h = time.perf_counter(); sync(); t2 = time.perf_counter()
        nb.align(src_slots, tgt_slots, guesses); t3 = time.perf_counter()
        if r >= 2: acc += [t1 - t0, t2 - t1, t3 - t2, (t1h - t0) + (t2h - t1)]
    return acc / reps * 1e3

for name, s, t in (("resident", dev_src, dev_tgt), ("pinned host", pin_src, pin_tgt)):
    p = phases(s, t)
    print("%-12s set_targets %.3f ms | set_sources %.3f ms | align %.3f ms | (host time inside the two setters %.3f ms)" % (name, *p))
st = nb.last_stats()
print("last align: device %.3f ms, %d launches, eval kernels %.3f ms over %d" % (st["device_ms"], st["launches"], st["deriv_kernel_ms"], st["deriv_launches"]))
for grp in (64, 32, 16, 8):
    sync(); t0 = time.perf_counter()
    for r in range(10):
        nb.set_targets(list(range(len(keys))), pin_tgt); nb.set_sources(src_slots, pin_src)
        for a in range(0, B, grp): nb.align(src_slots[a:a + grp], tgt_slots[a:a + grp], guesses[a:a + grp])
    sync(); print("pinned, aligns in groups of %2d: %.3f ms per step" % (grp, (time.perf_counter() - t0) / 10 * 1e3))
