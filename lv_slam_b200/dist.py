"""Multi-GPU plumbing of the benchmark / replay driver: one process per GPU, work split by independent scan pairs or pose
graphs (the stream shards by frame, no collective on the data path), device timings combined as the max over ranks."""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def frame_range(rank, pairs_per_rank):
    """Frames of the synthetic drive owned by `rank`: two lead-in frames (keyframe 0 and the first constant-velocity
    predecessor) followed by `pairs_per_rank` matched frames.  Ranges of different ranks are disjoint."""
    span = pairs_per_rank + 2
    return rank * span, span


def max_over_ranks(value, world, device=None):
    """MAX all-reduce of a scalar (timings are reported as the slowest rank's)."""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, world, device=None):
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_range(n, rank, world):
    """Contiguous chunk [lo, hi) of an n-point source evaluated by `rank` when one align is point-sharded (the library computes the
    same split: rank*n // world)."""
    return rank * n // world, (rank + 1) * n // world


def all_gather_bytes(blob, world):
    """-> [blob of rank 0, ..., blob of rank world-1] over the default torch.distributed group (gloo or nccl)."""
    if world <= 1:
        return [bytes(blob)]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, bytes(blob))
    return out


def bind_to_gpu_numa(local_rank):
    """Pins this process to the CPU cores that are local to GPU `local_rank` (sysfs local_cpulist of its PCI device), so that pinned
    host buffers allocated afterwards are first-touched on that NUMA node and the copy engine does not cross the socket interconnect.
    Returns a one-line description; does nothing (and says so) where the topology cannot be read."""
    try:
        import subprocess
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip()
        dev = bus.lower()
        if dev.startswith("00000000:"):
            dev = "0000:" + dev[9:]
        path = "/sys/bus/pci/devices/%s/" % dev
        cpus = open(path + "local_cpulist").read().strip()
        node = open(path + "numa_node").read().strip()
        ids = []
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids += list(range(int(a), int(b) + 1))
            elif part:
                ids.append(int(part))
        allowed = sorted(set(ids) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "gpu %s numa %s cpus %s: none of them allowed for this process, not bound" % (dev, node, cpus)
        os.sched_setaffinity(0, allowed)
        return "gpu %s numa %s bound to %d cpus (%s)" % (dev, node, len(allowed), cpus)
    except Exception as e:                       # containers often hide the topology
        return "not bound (%s)" % e
