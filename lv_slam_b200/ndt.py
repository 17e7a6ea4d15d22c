"""Host mirror of the reference registration classes over the C-ABI.

``NormalDistributionsTransform`` carries the method names of ``pclomp::NormalDistributionsTransform``
(/root/reference/include/ndt_omp/ndt_omp.h:69-551) and ``pclpca::NormalDistributionsTransform``
(include/ndt_pca/ndt_pca.h); clouds are numpy float32 [n, >=3] arrays (row stride = point stride) or CUDA torch
tensors.  All arithmetic happens in liblvslam_b200.so on the GPU.
"""
import ctypes

import numpy as np

from . import _capi as C


def _colmajor16(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T.reshape(16))


def _from_colmajor16(v):
    return np.asarray(v, dtype=np.float32).reshape(4, 4).T.copy()


def _cloud_args(xyz):
    """-> (pointer, n, stride_bytes, on_device, keepalive)"""
    if hasattr(xyz, "is_cuda"):
        import torch
        t = xyz
        if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] < 3 or (t.shape[0] > 0 and t.stride(1) != 1):
            raise ValueError("cloud tensor must be float32 [n, >=3] with unit inner stride")
        stride = t.stride(0) * 4 if t.shape[0] > 1 else t.shape[1] * 4
        return t.data_ptr(), t.shape[0], stride, 1 if t.is_cuda else 0, t
    a = np.asarray(xyz)
    if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3:
        a = np.ascontiguousarray(a, dtype=np.float32)
    if a.shape[0] > 0 and a.strides[1] != 4:
        a = np.ascontiguousarray(a)
    stride = a.strides[0] if a.shape[0] > 1 else a.shape[1] * 4
    return a.ctypes.data, a.shape[0], stride, 0, a


def _params(**kw):
    p = C.NdtParams()
    C.lib().lvs_ndt_default_params(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class CloudBatch:
    """Clouds marshalled once for the plural setters (pointers, counts, stride, residency): a caller that hands the same buffers to
    set_sources / set_targets step after step keeps the per-call host work to the C call itself."""

    def __init__(self, clouds):
        args = [_cloud_args(c) for c in clouds]
        self.n = len(args)
        self.stride, self.on_device = (args[0][2], args[0][3]) if args else (12, 0)
        if any(a[2] != self.stride or a[3] != self.on_device for a in args):
            raise ValueError("set_sources / set_targets need clouds of one stride and one residency")
        self.ptrs = (ctypes.c_void_p * max(self.n, 1))(*[a[0] for a in args])
        self.counts = (ctypes.c_size_t * max(self.n, 1))(*[a[1] for a in args])
        self._keep = [a[4] for a in args]


_RESULT_DTYPE = np.dtype({"names": ["final_transformation", "converged", "iterations", "trans_probability", "n_eval", "n_hess", "score"],
                          "formats": [(np.float32, 16), np.int32, np.int32, np.float64, np.int32, np.int32, np.float64],
                          "offsets": [C.NdtResult.final_transformation.offset, C.NdtResult.converged.offset, C.NdtResult.iterations.offset,
                                      C.NdtResult.trans_probability.offset, C.NdtResult.n_eval.offset, C.NdtResult.n_hess.offset, C.NdtResult.score.offset],
                          "itemsize": ctypes.sizeof(C.NdtResult)})


class AlignResults:
    """The lvs_ndt_result records of one batched align as numpy views (no per-pair Python objects on the hot path); indexing or
    iterating yields the per-pair dicts the single-object API returns."""

    def __init__(self, raw, n):
        self._raw = raw
        self._a = np.frombuffer(raw, dtype=_RESULT_DTYPE, count=n) if n else np.zeros(0, _RESULT_DTYPE)

    def __len__(self):
        return self._a.shape[0]

    @property
    def finals(self):
        """[n, 4, 4] float32 final transformations (row-major views of the column-major records)"""
        return self._a["final_transformation"].reshape(-1, 4, 4).transpose(0, 2, 1)

    @property
    def n_eval(self):
        return self._a["n_eval"]

    @property
    def iterations(self):
        return self._a["iterations"]

    @property
    def converged(self):
        return self._a["converged"] != 0

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        r = self._a[i]
        return dict(final=r["final_transformation"].reshape(4, 4).T.copy(), iterations=int(r["iterations"]), converged=bool(r["converged"]),
                    trans_probability=float(r["trans_probability"]), n_eval=int(r["n_eval"]), n_hess=int(r["n_hess"]), score=float(r["score"]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __add__(self, other):                      # results of consecutive align groups concatenate
        out = AlignResults.__new__(AlignResults)
        out._raw = (self._raw, getattr(other, "_raw", None))
        out._a = np.concatenate([self._a, other._a]) if len(other) else self._a
        return out


def pack_guesses(guesses):
    """[n, 16] float32 column-major guesses, the layout lvs_ndt_batch_align takes"""
    return np.ascontiguousarray(np.stack([_colmajor16(G) for G in guesses])) if len(guesses) else np.zeros((0, 16), np.float32)


class NormalDistributionsTransform:
    """One registration object.  ``variant`` selects pclomp (LVS_NDT_OMP), pclpca (LVS_NDT_PCA) or
    pclomp_ground::NormalDistributionsTransformGround (LVS_NDT_GROUND, include/ndt_omp/ndt_ground.h)."""

    def __init__(self, variant=C.LVS_NDT_OMP, device=0, stream=None):
        self._L = C.lib()
        self._p = _params(variant=variant)
        h = ctypes.c_void_p()
        C.check(self._L.lvs_ndt_create(ctypes.byref(self._p), device, stream, ctypes.byref(h)))
        self._h = h
        self._res = None
        self._n_src = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_ndt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setters of the reference class
    def _push(self):
        C.check(self._L.lvs_ndt_set_params(self._h, ctypes.byref(self._p)))

    def setResolution(self, r):
        self._p.resolution = float(r); self._push()

    def setStepSize(self, s):
        self._p.step_size = float(s); self._push()

    def setOulierRatio(self, o):  # [sic] the reference's spelling (ndt_omp.h:177)
        self._p.outlier_ratio = float(o); self._push()

    def setTransformationEpsilon(self, e):
        self._p.transformation_epsilon = float(e); self._push()

    def setMaximumIterations(self, n):
        self._p.max_iterations = int(n); self._push()

    def setNeighborhoodSearchMethod(self, m):
        self._p.search_method = int(m); self._push()

    def setNumThreads(self, n):
        pass  # OpenMP team size of the reference; meaningless on the device

    def setAccumulation(self, mode):
        """Not a reference setter: LVS_ACC_EXACT (default, the reference's arithmetic) or LVS_ACC_FAST (tolerance mode)."""
        self._p.accumulation = int(mode); self._push()

    def setLeanFinalEvaluation(self, on):
        """Not a reference setter: skip the Hessian of the derivative pass that ends an align (nothing reads it); results unchanged."""
        self._p.lean_final_evaluation = 1 if on else 0; self._push()

    def getResolution(self):
        return self._p.resolution

    def getStepSize(self):
        return self._p.step_size

    def getOulierRatio(self):
        return self._p.outlier_ratio

    # ---- clouds
    def setInputTarget(self, xyz):
        ptr, n, stride, dev, keep = _cloud_args(xyz)
        C.check(self._L.lvs_ndt_set_target(self._h, ptr, n, stride, dev))

    def setInputSource(self, xyz):
        ptr, n, stride, dev, keep = _cloud_args(xyz)
        self._n_src = n
        C.check(self._L.lvs_ndt_set_source(self._h, ptr, n, stride, dev))

    # ---- registration
    def align(self, guess=None, want_cloud=False):
        g = _colmajor16(np.eye(4) if guess is None else guess)
        r = C.NdtResult()
        C.check(self._L.lvs_ndt_align(self._h, g.ctypes.data, ctypes.byref(r)))
        self._res = r
        if want_cloud:
            out = np.empty((self._n_src, 3), np.float32)
            C.check(self._L.lvs_ndt_get_aligned_cloud(self._h, out.ctypes.data, 0))
            return out
        return None

    def getFinalTransformation(self):
        return _from_colmajor16(np.frombuffer(self._res.final_transformation, dtype=np.float32))

    def hasConverged(self):
        return bool(self._res.converged)

    def getFinalNumIteration(self):
        return int(self._res.iterations)

    def getTransformationProbability(self):
        return float(self._res.trans_probability)

    def result(self):
        r = self._res
        return dict(final=self.getFinalTransformation(), iterations=int(r.iterations), converged=bool(r.converged),
                    trans_probability=float(r.trans_probability), n_eval=int(r.n_eval), n_hess=int(r.n_hess), score=float(r.score),
                    trace=self.trace())

    def trace(self):
        recs = (C.NdtTraceRec * 80)()
        n = ctypes.c_int(0)
        C.check(self._L.lvs_ndt_get_trace(self._h, recs, 80, ctypes.byref(n)))
        out = np.zeros((min(n.value, 80), 22))
        for k in range(out.shape[0]):
            t = recs[k]
            out[k, 0:6] = t.p_before; out[k, 6:12] = t.dir; out[k, 12] = t.step; out[k, 13] = t.score
            out[k, 14:20] = t.p_after; out[k, 20] = t.trials; out[k, 21] = t.hessian_recomputed
        return out

    def getFitnessScore(self, max_range=np.finfo(np.float64).max, T=None, with_count=False):
        """pcl::Registration::getFitnessScore(max_range): mean squared nearest-neighbour distance of the aligned source (T = None:
        the final transformation of the last align) to the target points, over squared distances <= max_range."""
        s, n = ctypes.c_double(0), ctypes.c_int(0)
        Tm = _colmajor16(T) if T is not None else None
        C.check(self._L.lvs_ndt_fitness_score(self._h, Tm.ctypes.data if Tm is not None else None, float(max_range), ctypes.byref(s), ctypes.byref(n)))
        return (s.value, n.value) if with_count else s.value

    def calculateScore(self, T):
        g = _colmajor16(T)
        s = ctypes.c_double(0)
        C.check(self._L.lvs_ndt_calculate_score(self._h, g.ctypes.data, ctypes.byref(s)))
        return s.value

    # ---- parity taps
    def eval_derivatives(self, p6, T=None, compute_hessian=True):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        Tm = _colmajor16(T) if T is not None else None
        s = ctypes.c_double(0)
        g, H = np.zeros(6), np.zeros((6, 6))
        C.check(self._L.lvs_ndt_eval_derivatives(self._h, p.ctypes.data, Tm.ctypes.data if Tm is not None else None,
                                                 int(compute_hessian), ctypes.byref(s), g.ctypes.data, H.ctypes.data))
        return s.value, g, H

    def eval_hessian(self, p6, T=None):
        p = np.ascontiguousarray(p6, dtype=np.float64)
        Tm = _colmajor16(T) if T is not None else None
        H = np.zeros((6, 6))
        C.check(self._L.lvs_ndt_eval_hessian(self._h, p.ctypes.data, Tm.ctypes.data if Tm is not None else None, H.ctypes.data))
        return H

    def grid(self):
        mn, mx, dv = (np.zeros(3, np.int32) for _ in range(3))
        C.check(self._L.lvs_ndt_get_grid(self._h, mn.ctypes.data, mx.ctypes.data, dv.ctypes.data))
        return mn, mx, dv

    def cells(self):
        n = ctypes.c_int(0)
        C.check(self._L.lvs_ndt_num_cells(self._h, ctypes.byref(n)))
        n = n.value
        out = dict(keys=np.zeros(n, np.int32), nr_points=np.zeros(n, np.int32), mean=np.zeros((n, 3)), icov=np.zeros((n, 3, 3)),
                   evals=np.zeros((n, 3)), centroid=np.zeros((n, 3), np.float32), weight=np.zeros(n, np.int32))
        C.check(self._L.lvs_ndt_get_cells(self._h, *[out[k].ctypes.data for k in ("keys", "nr_points", "mean", "icov", "evals", "centroid", "weight")]))
        return out

    def cell_horizontal(self):
        """pclomp_ground: 1 per occupied cell (the order of cells()) whose normal is within 10 degrees of the z axis
        (ndt_ground_impl.hpp:507-511,533); all 0 for the other variants."""
        n = ctypes.c_int(0)
        C.check(self._L.lvs_ndt_num_cells(self._h, ctypes.byref(n)))
        out = np.zeros(n.value, np.int32)
        C.check(self._L.lvs_ndt_get_cell_horizontal(self._h, out.ctypes.data))
        return out

    def lookup_keys(self, T):
        g = _colmajor16(T)
        keys = np.zeros(self._n_src, np.int32)
        C.check(self._L.lvs_ndt_lookup_keys(self._h, g.ctypes.data, keys.ctypes.data))
        return keys

    def enable_point_sharding(self, rank, world, exchange):
        """Point-sharded align() of this registration object across `world` GPUs (see NdtBatch.enable_point_sharding)."""
        _shard_setup(self._L, self.batch_handle(), rank, world, 1, exchange)

    def batch_handle(self):
        b = ctypes.c_void_p()
        C.check(self._L.lvs_ndt_handle_batch(self._h, ctypes.byref(b)))
        return b


def _shard_setup(L, batch_handle, rank, world, max_pairs, exchange):
    mine = (ctypes.c_ubyte * 64)()
    C.check(L.lvs_ndt_batch_shard_init(batch_handle, rank, world, max_pairs, mine))
    blobs = exchange(bytes(mine))
    if len(blobs) != world or any(len(x) != 64 for x in blobs):
        raise ValueError("exchange() must return one 64-byte handle per rank")
    allh = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(blobs))
    C.check(L.lvs_ndt_batch_shard_connect(batch_handle, allh))


class NormalDistributionsTransformGround(NormalDistributionsTransform):
    """pclomp_ground::NormalDistributionsTransformGround (include/ndt_omp/ndt_ground.h:69, ndt_ground_impl.hpp): the registration
    object scan_matching_odom_nodelet.cpp:121-126 configures as ``ground_s2k`` (resolution 10, DIRECT1, epsilon 0.01, 64 iterations)."""

    def __init__(self, device=0, stream=None):
        super().__init__(variant=C.LVS_NDT_GROUND, device=device, stream=stream)


class NdtBatch:
    """Many (source, target, guess) pairs advanced together on one device (loop-closure candidates, stream replay)."""

    def __init__(self, n_target_slots, n_source_slots, device=0, stream=None, **params):
        self._L = C.lib()
        self._p = _params(**params)
        h = ctypes.c_void_p()
        C.check(self._L.lvs_ndt_batch_create(ctypes.byref(self._p), device, stream, n_target_slots, n_source_slots, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_ndt_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_target(self, slot, xyz):
        ptr, n, stride, dev, keep = _cloud_args(xyz)
        C.check(self._L.lvs_ndt_batch_set_target(self._h, slot, ptr, n, stride, dev))

    def set_source(self, slot, xyz):
        ptr, n, stride, dev, keep = _cloud_args(xyz)
        C.check(self._L.lvs_ndt_batch_set_source(self._h, slot, ptr, n, stride, dev))

    def _set_many(self, fn, slots, clouds):
        cb = clouds if isinstance(clouds, CloudBatch) else CloudBatch(clouds)
        if cb.n == 0:
            return
        sl = slots if isinstance(slots, ctypes.Array) else (ctypes.c_int32 * cb.n)(*slots)
        C.check(fn(self._h, cb.n, sl, cb.ptrs, cb.counts, cb.stride, cb.on_device))

    def set_targets(self, slots, clouds):
        self._set_many(self._L.lvs_ndt_batch_set_targets, slots, clouds)

    def set_sources(self, slots, clouds):
        self._set_many(self._L.lvs_ndt_batch_set_sources, slots, clouds)

    def enable_point_sharding(self, rank, world, max_pairs, exchange):
        """Splits every source cloud over `world` ranks (one process per GPU); the 43 sums of each evaluation are exchanged through
        NVLink peer memory inside the evaluation kernel.  `exchange(blob) -> [blob of rank 0, ..., blob of rank world-1]` moves the
        64-byte IPC handles between the processes (lv_slam_b200.dist.all_gather_bytes under torch.distributed).  Call before
        set_source; afterwards every rank must issue the same calls with the same (whole) clouds, pairs and guesses."""
        _shard_setup(self._L, self._h, rank, world, max_pairs, exchange)

    def wait_uploads(self):
        """Blocks until every host cloud handed to set_target / set_source has reached the device (pinned buffers are free again)."""
        C.check(self._L.lvs_ndt_batch_wait_uploads(self._h))

    def align(self, source_slots, target_slots, guesses):
        """guesses: a list of 4x4 matrices, or an [n, 16] float32 array from pack_guesses.  Returns AlignResults."""
        s = np.ascontiguousarray(source_slots, dtype=np.int32)
        t = np.ascontiguousarray(target_slots, dtype=np.int32)
        n = s.shape[0]
        if isinstance(guesses, np.ndarray) and guesses.dtype == np.float32 and guesses.ndim == 2 and guesses.shape[1] == 16:
            g = np.ascontiguousarray(guesses)
        else:
            g = pack_guesses(guesses)
        if g.shape[0] != n or t.shape[0] != n:
            raise ValueError("source slots, target slots and guesses must have one length")
        res = (C.NdtResult * max(n, 1))()
        C.check(self._L.lvs_ndt_batch_align(self._h, n, s.ctypes.data, t.ctypes.data, g.ctypes.data, res))
        return AlignResults(res, n)

    def align_begin(self, source_slots, target_slots, guesses):
        """First half of align: queues the work and returns; set_targets / set_sources of OTHER slots may follow before align_end."""
        s = np.ascontiguousarray(source_slots, dtype=np.int32)
        t = np.ascontiguousarray(target_slots, dtype=np.int32)
        n = s.shape[0]
        g = np.ascontiguousarray(guesses) if (isinstance(guesses, np.ndarray) and guesses.dtype == np.float32 and guesses.ndim == 2 and guesses.shape[1] == 16) \
            else pack_guesses(guesses)
        if g.shape[0] != n or t.shape[0] != n:
            raise ValueError("source slots, target slots and guesses must have one length")
        C.check(self._L.lvs_ndt_batch_align_begin(self._h, n, s.ctypes.data, t.ctypes.data, g.ctypes.data))
        self._pending = (n, s, t, g)

    def align_end(self):
        n = self._pending[0] if getattr(self, "_pending", None) else 0
        self._pending = None
        res = (C.NdtResult * max(n, 1))()
        C.check(self._L.lvs_ndt_batch_align_end(self._h, res))
        return AlignResults(res, n)

    def fitness_score(self, source_slot, target_slot, T, max_range=np.finfo(np.float64).max):
        """getFitnessScore of one (source, target) pair under T -> (score, correspondences)"""
        s, n = ctypes.c_double(0), ctypes.c_int(0)
        Tm = _colmajor16(T)
        C.check(self._L.lvs_ndt_batch_fitness_score(self._h, int(source_slot), int(target_slot), Tm.ctypes.data, float(max_range), ctypes.byref(s), ctypes.byref(n)))
        return s.value, n.value

    def set_profiling(self, on):
        C.check(self._L.lvs_ndt_batch_set_profiling(self._h, int(on)))

    def set_tuning(self, blocks_per_pair=0, chunk_first=0, chunk_next=0):
        C.check(self._L.lvs_ndt_batch_set_tuning(self._h, blocks_per_pair, chunk_first, chunk_next))

    def last_stats(self):
        ms, dms = ctypes.c_double(0), ctypes.c_double(0)
        nl, dl = ctypes.c_int(0), ctypes.c_int(0)
        C.check(self._L.lvs_ndt_batch_last_stats(self._h, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(dms), ctypes.byref(dl)))
        return dict(device_ms=ms.value, launches=nl.value, deriv_kernel_ms=dms.value, deriv_launches=dl.value)

    def total_launches(self):
        n = ctypes.c_longlong(0)
        C.check(self._L.lvs_ndt_batch_total_launches(self._h, ctypes.byref(n)))
        return n.value

    def transfer_bytes(self):
        a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
        C.check(self._L.lvs_ndt_batch_transfer_bytes(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def num_cells(self, slot):
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        C.check(self._L.lvs_ndt_batch_num_cells(self._h, slot, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value
