"""Host mirror of lidar_odometry/PrefilteringNodelet's filter chain (/root/reference/src/lidar_odometry/prefiltering_nodelet.cpp:
30-98 parameters, 117-128 chain): distance filter then VoxelGrid downsampling, on the GPU through lvs_prefilter_run."""
import ctypes

import numpy as np

from . import _capi as C
from .ndt import _cloud_args


class Prefilter:
    """Parameter names and defaults of the nodelet (prefiltering_nodelet.cpp:40-41, 85-87).  ``downsample_method`` "VOXELGRID" or
    "NONE" (the approximate grid and the outlier filters are not on the benchmark path: the launch file's RADIUS filter is never
    installed by the reference, prefiltering_nodelet.cpp:76-83)."""

    def __init__(self, downsample_method="VOXELGRID", downsample_resolution=0.1, use_distance_filter=True, distance_near_thresh=1.0,
                 distance_far_thresh=100.0, device=0, stream=None):
        if downsample_method not in ("VOXELGRID", "NONE"):
            raise ValueError("downsample_method must be VOXELGRID or NONE")
        self.downsample_method, self.downsample_resolution = downsample_method, float(downsample_resolution)
        self.use_distance_filter = bool(use_distance_filter)
        self.distance_near_thresh, self.distance_far_thresh = float(distance_near_thresh), float(distance_far_thresh)
        self._L = C.lib()
        self._h = ctypes.c_void_p()
        C.check(self._L.lvs_prefilter_create(device, stream, ctypes.byref(self._h)))
        self.last_flags = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_prefilter_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def filter(self, cloud):
        """cloud: float32 [n, 3 or more]; columns 0-2 are x y z, column 3 (if present) the intensity.  Returns float32 [m, 3 or 4]."""
        ptr, n, stride, on_device, keep = _cloud_args(cloud)
        nf = 4 if (keep.shape[1] >= 4) else 3
        leaf = self.downsample_resolution if self.downsample_method == "VOXELGRID" else 0.0
        m, fl = ctypes.c_size_t(0), ctypes.c_int(0)
        if on_device:
            import torch
            out = torch.empty((max(n, 1), nf), dtype=torch.float32, device=keep.device)
            optr = out.data_ptr()
        else:
            out = np.empty((max(n, 1), nf), np.float32)
            optr = out.ctypes.data
        C.check(self._L.lvs_prefilter_run(self._h, ptr, n, stride, nf, on_device, self.distance_near_thresh, self.distance_far_thresh,
                                          int(self.use_distance_filter), leaf, optr, n, on_device, ctypes.byref(m), ctypes.byref(fl)))
        self.last_flags = fl.value
        return out[:m.value]


class WindowMap:
    """The window cloud of GlobalGraphNodelet::cloud_callback (src/global_graph/global_graph_nodelet.cpp:199-243): scans between two
    keyframes are transformed into the window's frame (double matrix, pcl::transformPointCloud) and appended; flush() applies the
    0.1 m VoxelGrid that turns the window into the keyframe's cloud."""

    def __init__(self, leaf=0.1, device=0, stream=None):
        self.leaf = float(leaf)
        self._L = C.lib()
        self._h = ctypes.c_void_p()
        C.check(self._L.lvs_prefilter_create(device, stream, ctypes.byref(self._h)))
        self._nf, self._n = None, 0
        self.last_flags = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.lvs_prefilter_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def start(self, cloud):
        """w_cloud = *cloud (the window's first scan, its own frame)"""
        C.check(self._L.lvs_prefilter_accumulate_begin(self._h))
        self._nf, self._n = None, 0
        self.add(cloud, None)

    def add(self, cloud, T):
        """w_cloud += transformPointCloud(cloud, T), T = w_odom^-1 * odom as a 4x4 double matrix (None: no transform)"""
        ptr, n, stride, on_device, keep = _cloud_args(cloud)
        nf = 4 if keep.shape[1] >= 4 else 3
        if self._nf is None:
            self._nf = nf
        elif nf != self._nf:
            raise ValueError("all clouds of a window need the same fields")
        Tm = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T.reshape(16)) if T is not None else None
        C.check(self._L.lvs_prefilter_accumulate_add(self._h, ptr, n, stride, nf, on_device, Tm.ctypes.data if Tm is not None else None))
        self._n += n

    def flush(self):
        """VoxelGrid(leaf) of the window -> float32 [m, 3 or 4]"""
        nf = self._nf or 3
        out = np.empty((max(self._n, 1), nf), np.float32)
        m, fl = ctypes.c_size_t(0), ctypes.c_int(0)
        C.check(self._L.lvs_prefilter_accumulate_flush(self._h, nf, self.leaf, out.ctypes.data, self._n, 0, ctypes.byref(m), ctypes.byref(fl)))
        self.last_flags = fl.value
        return out[:m.value].copy()
