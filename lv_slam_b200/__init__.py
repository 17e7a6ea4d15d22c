"""lv_slam_b200 — B200-native (sm_100a) NDT scan matching and pose-graph optimisation behind the C-ABI of
``include/lvslam_b200.h``.  This package is the thin Python host mirror used by the tests and the bench; the product
is the shared library ``liblvslam_b200.so`` built from ``csrc/``."""
from ._capi import (LVS_ACC_EXACT, LVS_ACC_FAST, LVS_DIRECT1, LVS_DIRECT7, LVS_DIRECT26, LVS_KDTREE, LVS_NDT_GROUND, LVS_NDT_OMP, LVS_NDT_PCA, LvsError, lib)  # noqa: F401
from .ndt import NdtBatch, NormalDistributionsTransform, NormalDistributionsTransformGround  # noqa: F401
from .graph_slam import GraphSLAM, PoseGraph  # noqa: F401,E402
from .information_matrix import InformationMatrixCalculator  # noqa: F401,E402
from .prefilter import Prefilter, WindowMap  # noqa: F401,E402
from . import keyframe_io  # noqa: F401,E402
