"""Builds liblvslam_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

The shared library is git-ignored but travels to the GPU box with the repository snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only container (``__graft_entry__.build()``).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblvslam_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: the NDT kernels reproduce the reference's float32 operation order bit for bit (the reference is built
# with SSE4.2 and no FMA, /root/reference/CMakeLists.txt:6); contraction would change voxel indices on cell faces.
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
NO_FMA = ["-fmad=false"]
# (source, extra flags): the pose-graph code is fp64 throughout and compared to tolerance, so it keeps FMA contraction
SOURCES = [("ndt_voxel.cu", NO_FMA), ("ndt_eval.cu", NO_FMA), ("ndt_eval_fast.cu", NO_FMA), ("ndt_eval_cold.cu", NO_FMA), ("ndt_api.cu", NO_FMA), ("ndt_fitness.cu", NO_FMA), ("prefilter.cu", NO_FMA), ("pgo.cu", []), ("pgo_chol.cu", [])]


def _deps():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "lvslam_b200.h"))
    return hs


def _stamp(src, extra=()):
    h = hashlib.sha1()
    for p in [src] + _deps():
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(COMMON + list(extra)).encode())
    return h.hexdigest()


def _compile(item, verbose):
    name, extra = item
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src, extra)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False
    cmd = [NVCC] + COMMON + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed on " + name)
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s[0]))]
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
