"""Point-sharded NDT across GPUs (BASELINE configs[2]: 128-beam scan, 0.5 m voxels; also the 64-beam pair).
   torchrun --nproc-per-node N tools/shard_check.py      (one process per GPU, NCCL only for the handle exchange / barriers)
Checks: every rank returns the same bits; the sharded result agrees with the same align on one GPU; prints timings."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import lv_slam_b200 as L
from lv_slam_b200 import dist as D, synth

rank, local, world = D.env_rank()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, (nb_, naz, res, variant, search) in {"64-beam 1.0 m pclomp/DIRECT7": (64, 2000, 1.0, 0, 2), "128-beam 0.5 m pclomp/DIRECT7": (128, 1875, 0.5, 0, 2),
                                               "128-beam 0.5 m pclpca/DIRECT1": (128, 1875, 0.5, 1, 3)}.items():
    tgt, src, guess, truth = synth.config1_pair(n_beams=nb_, n_az=naz, seed=42)
    kw = dict(resolution=res, transformation_epsilon=0.01, max_iterations=64, variant=variant, search_method=search)
    single = L.NdtBatch(1, 1, device=local, **kw)
    shard = L.NdtBatch(1, 1, device=local, **kw)
    shard.enable_point_sharding(rank, world, 1, lambda blob: D.all_gather_bytes(blob, world))
    dt, ds = torch.from_numpy(tgt).cuda(), torch.from_numpy(src).cuda()
    out = {}
    for label, nb in (("1 GPU", single), ("%d GPUs point-sharded" % world, shard)):
        nb.set_target(0, dt); nb.set_source(0, ds)
        for rep in range(3):
            r = nb.align([0], [0], [guess])[0]
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        reps = 10
        for rep in range(reps):
            r = nb.align([0], [0], [guess])[0]
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        ms = D.max_over_ranks(ms, world, "cuda")
        out[label] = (r, ms, nb.last_stats()["device_ms"])
    (r1, ms1, dev1), (rs, mss, devs) = out["1 GPU"], out["%d GPUs point-sharded" % world]
    fin = torch.from_numpy(rs["final"].copy()).cuda()
    same = True
    if world > 1:
        allf = [torch.empty_like(fin) for _ in range(world)]
        dist.all_gather(allf, fin)
        same = all(bool(torch.equal(allf[0], f)) for f in allf)
    dt_ = float(np.abs(rs["final"][:3, 3] - r1["final"][:3, 3]).max()); dr_ = float(np.abs(rs["final"][:3, :3] - r1["final"][:3, :3]).max())
    good = same and rs["iterations"] == r1["iterations"] and dt_ <= 1e-4 and dr_ <= 1e-5
    ok = ok and good
    if rank == 0:
        print("%-32s %6d pts | 1 GPU: %.3f ms/align (device %.3f), %d iters | %d GPUs sharded: %.3f ms/align (device %.3f), %d iters | ranks bit-identical %s, |dt| %.1e m |dR| %.1e vs 1 GPU -> %s"
              % (name, src.shape[0], ms1, dev1, r1["iterations"], world, mss, devs, rs["iterations"], same, dt_, dr_, "OK" if good else "MISMATCH"), flush=True)
    single.close(); shard.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
