#!/bin/sh
# Builds the reference-derived checker objects into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
# Only runs where the reference tree is mounted.  The one piece of the pose-graph path the reference vendors in compilable
# form is CSparse (the sparse Cholesky behind g2o's `lm_var` / `gn_var` solvers), inside 3rdtools/g2o-a48ff8c.zip under
# g2o/EXTERNAL/csparse/.  The sources are unpacked to a temporary directory and compiled from there with gcc directly
# (no CMake); nothing from the reference is copied into the repository.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ZIP=/root/reference/3rdtools/g2o-a48ff8c.zip
OUT="$HERE/_ref"
[ -f "$ZIP" ] || { echo "build_ref.sh: $ZIP not found, skipping"; exit 0; }
SZIP=/root/reference/3rdtools/Sophus-a621ff2-ubuntu18.04.zip
if [ -f "$OUT/libcsparse_ref.so" ] && [ "$OUT/libcsparse_ref.so" -nt "$ZIP" ] && [ -f "$OUT/libsophus_ref.so" ] && [ "$OUT/libsophus_ref.so" -nt "$SZIP" ] &&
   [ "$OUT/libsophus_ref.so" -nt "$HERE/sophus_ref_api.cpp" ] && [ "$OUT/libsophus_ref.so" -nt "$HERE/ref_stubs/eigen_min.h" ] &&
   [ -f "$OUT/libndt_ref.so" ] && [ "$OUT/libndt_ref.so" -nt "$HERE/ndt_ref_harness.cpp" ] && [ "$OUT/libndt_ref.so" -nt "$HERE/ref_stubs/eigen_min.h" ] &&
   [ "$OUT/libndt_ref.so" -nt "$HERE/extract_ref_functions.py" ] && [ "$OUT/libndt_ref.so" -nt "$HERE/olin.h" ] &&
   [ -f "$OUT/libndt_pca_ref.so" ] && [ "$OUT/libndt_pca_ref.so" -nt "$OUT/libndt_ref.so" ] &&
   [ -f "$OUT/libndt_ground_ref.so" ] && [ "$OUT/libndt_ground_ref.so" -nt "$OUT/libndt_ref.so" ] &&
   [ -f "$OUT/libvoxel_ref.so" ] && [ "$OUT/libvoxel_ref.so" -nt "$HERE/voxel_ref_harness.cpp" ] && [ "$OUT/libvoxel_ref.so" -nt "$OUT/libndt_ref.so" ] &&
   [ -f "$OUT/libvoxel_pca_ref.so" ] && [ "$OUT/libvoxel_pca_ref.so" -nt "$OUT/libvoxel_ref.so" ] &&
   [ -f "$OUT/libinfo_ref.so" ] && [ "$OUT/libinfo_ref.so" -nt "$HERE/info_ref_api.cpp" ] && [ "$OUT/libinfo_ref.so" -nt "$OUT/libvoxel_ref.so" ] &&
   [ -f "$OUT/libprior_ref.so" ] && [ "$OUT/libprior_ref.so" -nt "$HERE/prior_ref_api.cpp" ] && [ "$OUT/libprior_ref.so" -nt "$OUT/libvoxel_ref.so" ] &&
   [ -f "$OUT/libdquat_ref.so" ] && [ "$OUT/libdquat_ref.so" -nt "$HERE/dquat_ref_api.cpp" ] && [ "$OUT/libdquat_ref.so" -nt "$OUT/libinfo_ref.so" ] &&
   [ -f "$OUT/libg2o_ref.so" ] && [ "$OUT/libg2o_ref.so" -nt "$HERE/g2o_ref_harness.cpp" ] && [ "$OUT/libg2o_ref.so" -nt "$OUT/libdquat_ref.so" ] &&
   [ -f "$OUT/liblm_ref.so" ] && [ "$OUT/liblm_ref.so" -nt "$HERE/lm_ref_harness.cpp" ] && [ "$OUT/liblm_ref.so" -nt "$HERE/pgo_oracle.cpp" ] &&
   [ "$OUT/liblm_ref.so" -nt "$OUT/libg2o_ref.so" ]; then exit 0; fi
# the Eigen stand-in on its own (no reference source involved): tests hold its semantics against numpy
mkdir -p "$OUT"
/usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -I"$HERE/ref_stubs" -o "$OUT/libeigen_min_selftest.so" "$HERE/ref_stubs/selftest_api.cpp"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in z.namelist():
    if n.startswith("g2o/EXTERNAL/csparse/") and (n.endswith(".c") or n.endswith(".h")):
        z.extract(n, sys.argv[2])
PY
mkdir -p "$TMP/stub/g2o" "$OUT"
: > "$TMP/stub/g2o/config.h"      # cs_api.h includes g2o/config.h, which CMake would generate; nothing in it is needed
/usr/bin/gcc -O2 -fPIC -shared -I"$TMP/stub" -I"$TMP/g2o/EXTERNAL/csparse" -o "$OUT/libcsparse_ref.so" "$TMP"/g2o/EXTERNAL/csparse/*.c -lm
echo "built $OUT/libcsparse_ref.so"

# The NDT path's Lie-group code: Sophus a621ff2 (so3.cpp, se3.cpp), vendored by the reference as a zip.  It needs Eigen, which this image
# does not have: the two files are compiled as they are against oracle/ref_stubs/eigen_min.h, a stand-in for the handful of Eigen
# operations they use (see its header for what that does and does not pin), together with the C entry points of sophus_ref_api.cpp.
if [ -f "$SZIP" ]; then
  python3 - "$SZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in ("Sophus/sophus/so3.h", "Sophus/sophus/so3.cpp", "Sophus/sophus/se3.h", "Sophus/sophus/se3.cpp"):
    z.extract(n, sys.argv[2])
PY
  /usr/bin/g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -I"$HERE/ref_stubs" -I"$TMP/Sophus/sophus" -o "$OUT/libsophus_ref.so" \
      "$TMP/Sophus/sophus/so3.cpp" "$TMP/Sophus/sophus/se3.cpp" "$HERE/sophus_ref_api.cpp"
  echo "built $OUT/libsophus_ref.so"
  # The NDT path itself: the member functions of pclomp::NormalDistributionsTransform, taken verbatim from the reference's
  # include/ndt_omp/ndt_omp_impl2.hpp at build time (into the temporary directory) and compiled inside oracle/ndt_ref_harness.cpp, which
  # supplies the class declaration, the voxel-grid adapter and pcl::transformPointCloud (see its header for what this pins).
  FUNCS="computeTransformation computePointDerivatives_AngleAxisd updateDerivatives computeHessian updateHessian updateIntervalMT trialValueSelectionMT computeStepLengthMT calculateScore"
  build_ndt() {      # <impl file> <qualified class> <derivative pass> <extra define> <output>
    [ -f "$1" ] || return 0
    python3 "$HERE/extract_ref_functions.py" "$1" "$TMP/bodies_$5.inc" "$2" $FUNCS "$3"
    /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -Wno-unknown-pragmas $4 -DREF_NDT_BODIES="\"$TMP/bodies_$5.inc\"" -I"$HERE/ref_stubs" \
        -I"$TMP/Sophus/sophus" -o "$OUT/$5" "$TMP/Sophus/sophus/so3.cpp" "$TMP/Sophus/sophus/se3.cpp" "$HERE/ndt_ref_harness.cpp"
    echo "built $OUT/$5"
  }
  INC=/root/reference/include
  build_ndt "$INC/ndt_omp/ndt_omp_impl2.hpp" pclomp::NormalDistributionsTransform computeDerivatives "" libndt_ref.so
  build_ndt "$INC/ndt_pca/ndt_pca_impl2.hpp" pclpca::NormalDistributionsTransform computeDerivatives -DREF_PCA libndt_pca_ref.so
  build_ndt "$INC/ndt_omp/ndt_ground_impl.hpp" pclomp_ground::NormalDistributionsTransformGround computeDerivatives_seg -DREF_GROUND libndt_ground_ref.so
  # The voxel build and the direct searches: VoxelGridCovariance<PointT>::applyFilter / getNeighborhoodAtPoint{,7,1}, taken the same way from
  # voxel_grid_covariance_omp_impl.hpp and compiled in oracle/voxel_ref_harness.cpp (class declaration incl. the Leaf constructor's values).
  VIMPL="$INC/ndt_omp/voxel_grid_covariance_omp_impl.hpp"
  if [ -f "$VIMPL" ]; then
    python3 "$HERE/extract_ref_functions.py" "$VIMPL" "$TMP/voxel_bodies.inc" "pclomp::VoxelGridCovariance<PointT>" applyFilter getNeighborhoodAtPoint \
        getNeighborhoodAtPoint7 getNeighborhoodAtPoint1
    /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -DREF_VOXEL_BODIES="\"$TMP/voxel_bodies.inc\"" -I"$HERE/ref_stubs" -o "$OUT/libvoxel_ref.so" \
        "$HERE/voxel_ref_harness.cpp"
    echo "built $OUT/libvoxel_ref.so"
  fi
  PIMPL="$INC/ndt_pca/voxel_grid_covariance_pca_impl.hpp"
  if [ -f "$PIMPL" ]; then
    python3 "$HERE/extract_ref_functions.py" "$PIMPL" "$TMP/voxel_pca_bodies.inc" "pclpca::VoxelGridCovariance<PointT>" applyFilter getNeighborhoodAtPoint \
        getNeighborhoodAtPoint7 getNeighborhoodAtPoint1
    /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -DREF_PCA -DREF_VOXEL_BODIES="\"$TMP/voxel_pca_bodies.inc\"" -I"$HERE/ref_stubs" \
        -o "$OUT/libvoxel_pca_ref.so" "$HERE/voxel_ref_harness.cpp"
    echo "built $OUT/libvoxel_pca_ref.so"
  fi
  # The edge information matrix of the pose-graph nodelet: the reference's own src/global_graph/information_matrix_calculator.cpp, whole and as
  # it lies, against stand-ins for the <ros/ros.h>, <pcl/...> and Eigen headers it includes (oracle/ref_stubs/).
  ICPP=/root/reference/src/global_graph/information_matrix_calculator.cpp
  if [ -f "$ICPP" ]; then
    /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -I"$HERE/ref_stubs" -I/root/reference/include -o "$OUT/libinfo_ref.so" "$ICPP" "$HERE/info_ref_api.cpp"
    echo "built $OUT/libinfo_ref.so"
  fi
  # The reference's own unary edges of the global graph (GPS / IMU priors, floor plane): include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp and
  # edge_se3_plane.hpp as they are (the latter with g2o's own plane3d.h from the zip), against
  # stand-ins for the g2o / Eigen headers they include
  if [ -f /root/reference/include/g2o/edge_se3_priorvec.hpp ]; then
    python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in ("g2o/g2o/types/slam3d_addons/plane3d.h",          # g2o's own Plane3D, for edge_se3_plane.hpp
          "g2o/g2o/core/base_unary_edge.hpp", "g2o/g2o/core/base_binary_edge.hpp",      # the numeric linearizeOplus of both edge bases
          "g2o/g2o/core/robust_kernel_impl.cpp",             # the Huber kernel constructQuadraticForm weighs with
          "g2o/g2o/types/slam3d/isometry3d_mappings.cpp"):   # fromVectorMQT behind VertexSE3::oplus
    z.extract(n, sys.argv[2])
PY
    python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/base_unary_edge.hpp" "$TMP/g2o_unary.inc" "BaseUnaryEdge<D, E, VertexXiType>" "linearizeOplus()" \
        constructQuadraticForm
    python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/base_binary_edge.hpp" "$TMP/g2o_binary.inc" "BaseBinaryEdge<D, E, VertexXiType, VertexXjType>" \
        "linearizeOplus()" constructQuadraticForm
    python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/robust_kernel_impl.cpp" "$TMP/g2o_huber_p.inc" "=RobustKernelHuber" robustify
    python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/types/slam3d/isometry3d_mappings.cpp" "$TMP/g2o_map_p.inc" - normalize toCompactQuaternion \
        fromCompactQuaternion toVectorMQT fromVectorMQT
    /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -DG2O_MAP_BODIES="\"$TMP/g2o_map_p.inc\"" -DG2O_UNARY_BODIES="\"$TMP/g2o_unary.inc\"" \
        -DG2O_BINARY_BODIES="\"$TMP/g2o_binary.inc\"" -DG2O_HUBER_BODIES="\"$TMP/g2o_huber_p.inc\"" -I"$HERE/ref_stubs" -I"$HERE/ref_stubs/g2o_api" \
        -I"$TMP/g2o/g2o/types/slam3d_addons" \
        -I/root/reference/include -o "$OUT/libprior_ref.so" "$HERE/prior_ref_api.cpp"
    echo "built $OUT/libprior_ref.so"
  fi
fi
# g2o's own compute_dq_dR (dquat2mat.cpp + its Maxima-generated cases), as they are in the zip: the table behind EdgeSE3::linearizeOplus
if [ -f "$ZIP" ]; then
  python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in ("g2o/g2o/types/slam3d/dquat2mat.cpp", "g2o/g2o/types/slam3d/dquat2mat.h", "g2o/g2o/types/slam3d/dquat2mat_maxima_generated.cpp"):
    z.extract(n, sys.argv[2])
PY
  /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -I"$HERE/ref_stubs" -I"$HERE/ref_stubs/g2o_api" -I"$TMP/g2o/g2o/types/slam3d" \
      -o "$OUT/libdquat_ref.so" "$TMP/g2o/g2o/types/slam3d/dquat2mat.cpp" "$HERE/dquat_ref_api.cpp"
  echo "built $OUT/libdquat_ref.so"
  # g2o's own slam3d edge math (error vector, analytic Jacobians, oplus): functions of isometry3d_gradients.h / isometry3d_mappings.cpp taken at
  # build time and compiled in oracle/g2o_ref_harness.cpp with dquat2mat.cpp
  python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
for n in ("g2o/g2o/types/slam3d/isometry3d_gradients.h", "g2o/g2o/types/slam3d/isometry3d_mappings.cpp", "g2o/g2o/core/robust_kernel_impl.cpp"):
    z.extract(n, sys.argv[2])
PY
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/types/slam3d/isometry3d_gradients.h" "$TMP/g2o_grad.inc" - skew skewT computeEdgeSE3Gradient
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/types/slam3d/isometry3d_mappings.cpp" "$TMP/g2o_map.inc" - normalize toCompactQuaternion \
      fromCompactQuaternion toVectorMQT fromVectorMQT
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/robust_kernel_impl.cpp" "$TMP/g2o_huber.inc" "=RobustKernelHuber" robustify
  /usr/bin/g++ -O2 -std=gnu++17 -ffp-contract=off -fPIC -shared -DG2O_GRAD_BODIES="\"$TMP/g2o_grad.inc\"" -DG2O_MAP_BODIES="\"$TMP/g2o_map.inc\"" \
      -DG2O_HUBER_BODIES="\"$TMP/g2o_huber.inc\"" \
      -I"$HERE/ref_stubs" -I"$HERE/ref_stubs/g2o_api" -I"$TMP/g2o/g2o/types/slam3d" -o "$OUT/libg2o_ref.so" "$TMP/g2o/g2o/types/slam3d/dquat2mat.cpp" \
      "$HERE/g2o_ref_harness.cpp"
  echo "built $OUT/libg2o_ref.so"
  # g2o's own Levenberg-Marquardt control flow (solve, computeLambdaInit, computeScale) over the restatement's building blocks:
  # oracle/lm_ref_harness.cpp includes oracle/pgo_oracle.cpp, so the library is a second copy of the oracle whose optimiser is g2o's code
  python3 - "$ZIP" "$TMP" <<'PY'
import sys, zipfile
z = zipfile.ZipFile(sys.argv[1])
z.extract("g2o/g2o/core/optimization_algorithm_levenberg.cpp", sys.argv[2])
z.extract("g2o/g2o/core/optimization_algorithm_gauss_newton.cpp", sys.argv[2])
z.extract("g2o/g2o/solvers/pcg/linear_solver_pcg.hpp", sys.argv[2])
PY
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/solvers/pcg/linear_solver_pcg.hpp" "$TMP/g2o_pcg.inc" "LinearSolverPCG<MatrixType>" solve multDiag mult
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/optimization_algorithm_gauss_newton.cpp" "$TMP/g2o_gn.inc" "=OptimizationAlgorithmGaussNewton" solve
  python3 "$HERE/extract_ref_functions.py" "$TMP/g2o/g2o/core/optimization_algorithm_levenberg.cpp" "$TMP/g2o_lm.inc" "=OptimizationAlgorithmLevenberg" \
      solve computeLambdaInit computeScale
  /usr/bin/g++ -O3 -fopenmp -msse4.2 -ffp-contract=off -fPIC -std=gnu++17 -shared -DG2O_LM_BODIES="\"$TMP/g2o_lm.inc\"" -DG2O_GN_BODIES="\"$TMP/g2o_gn.inc\"" \
      -DG2O_PCG_BODIES="\"$TMP/g2o_pcg.inc\"" -I"$HERE" -I"$HERE/ref_stubs" -o "$OUT/liblm_ref.so" \
      "$HERE/lm_ref_harness.cpp" -ldl
  echo "built $OUT/liblm_ref.so"
fi
