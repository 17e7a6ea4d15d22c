// Scan prefilter on the device — the stage right before the NDT path: PrefilteringNodelet's
// distance_filter (src/lidar_odometry/prefiltering_nodelet.cpp:164-181) followed by downsample() = pcl::VoxelGrid with
// downsample_resolution 0.1 m (:41-47, 138-148; launch/dlo_lfa_ggo_kitti.launch:30-36).  The outlier filter of that launch file
// ("RADIUS") is constructed but never installed by the reference (:76-83), so the chain ends here.
//
//   distance filter  d = float norm of (x, y, z) (Eigen: ((x*x + y*y) + z*z), sqrtf), kept when d > near && d < far as doubles;
//                    order-preserving compaction (flag, scan, scatter)
//   VoxelGrid        PCL 1.8 voxel_grid.hpp applyFilter: bounding box of the finite points, min_b/max_b = floor(p * inv_leaf),
//                    leaf index per point with the same float arithmetic as the NDT target grid, points grouped by index, one
//                    output point per occupied leaf in ascending index order = float sums of x, y, z (and intensity) divided by
//                    float(count).  PCL orders the points of a leaf with an unstable std::sort, so its float sums are
//                    reproducible only up to summation order; here the order is the input order (stable radix sort).
//                    When dx*dy*dz overflows int32 PCL warns and returns the input unchanged: so does this (status flag 1).
// Compiled with -fmad=false.
#include <climits>
#include <new>
#include "ndt_internal.cuh"

namespace lvs {

constexpr int kPfThreads = 256;

__global__ void pf_flag_kernel(const float* __restrict__ in, size_t stride_floats, int n, int n_fields, double near_t, double far_t, int use_filter,
                               float4* __restrict__ pts, int* __restrict__ flags) {
  const int i = blockIdx.x * kPfThreads + threadIdx.x;
  if (i >= n) return;
  const float* p = in + (size_t)i * stride_floats;
  const float4 v = make_float4(p[0], p[1], p[2], n_fields > 3 ? p[3] : 0.0f);
  pts[i] = v;
  int keep = 1;
  if (use_filter) {
    const double d = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
    keep = (d > near_t && d < far_t) ? 1 : 0;
  }
  flags[i] = keep;
}

__global__ void pf_compact_kernel(const float4* __restrict__ pts, const int* __restrict__ flags, const int* __restrict__ pos, int n,
                                  float4* __restrict__ out) {
  const int i = blockIdx.x * kPfThreads + threadIdx.x;
  if (i < n && flags[i]) out[pos[i]] = pts[i];
}

// one thread per occupied leaf: float sums in input order, divided by float(count)
__global__ void pf_centroid_kernel(const float4* __restrict__ pts, const int* __restrict__ sorted, const int* __restrict__ cell_start,
                                   const int* __restrict__ n_seg_p, int n_fields, float* __restrict__ out) {
  const int seg = blockIdx.x * kPfThreads + threadIdx.x;
  if (seg >= *n_seg_p) return;
  const int s0 = cell_start[seg], s1 = cell_start[seg + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (int k = s0; k < s1; k++) {
    const float4 p = __ldg(pts + __ldg(sorted + k));
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
  }
  const float fn = (float)(s1 - s0);
  float* o = out + (size_t)seg * n_fields;
  o[0] = __fdiv_rn(sx, fn); o[1] = __fdiv_rn(sy, fn); o[2] = __fdiv_rn(sz, fn);
  if (n_fields > 3) o[3] = __fdiv_rn(si, fn);
}

__global__ void pf_unpack_kernel(const float4* __restrict__ pts, int n, int n_fields, float* __restrict__ out) {
  const int i = blockIdx.x * kPfThreads + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  float* o = out + (size_t)i * n_fields;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
  if (n_fields > 3) o[3] = p.w;
}

struct PfMat { double m[16]; };       // column-major 4x4

// pcl::transformPointCloud with a double matrix (PCL 1.8 common/impl/transforms.hpp): each output coordinate is
// float(m_r0*x + m_r1*y + m_r2*z + m_r3) evaluated left to right in double; the intensity rides along.
__global__ void pf_transform_append_kernel(const float* __restrict__ in, size_t stride_floats, int n, int n_fields, PfMat M, int use_T,
                                           float4* __restrict__ out) {
  const int i = blockIdx.x * kPfThreads + threadIdx.x;
  if (i >= n) return;
  const float* p = in + (size_t)i * stride_floats;
  float4 v = make_float4(p[0], p[1], p[2], n_fields > 3 ? p[3] : 0.0f);
  if (use_T) {
    const double x = (double)v.x, y = (double)v.y, z = (double)v.z;
    v.x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[0], x), __dmul_rn(M.m[4], y)), __dmul_rn(M.m[8], z)), M.m[12]);
    v.y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[1], x), __dmul_rn(M.m[5], y)), __dmul_rn(M.m[9], z)), M.m[13]);
    v.z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M.m[2], x), __dmul_rn(M.m[6], y)), __dmul_rn(M.m[10], z)), M.m[14]);
  }
  out[i] = v;
}

}  // namespace lvs

using namespace lvs;

struct lvs_prefilter {
  int device = 0;
  cudaStream_t st = nullptr;
  bool own_stream = false;
  BuildScratch ws;
  size_t cap = 0;
  float* d_in = nullptr;           // staged host input
  size_t in_cap = 0;
  float4 *d_pts = nullptr, *d_kept = nullptr;
  int *d_sorted = nullptr, *d_cell_start = nullptr;
  float* d_out = nullptr;          // packed result
  GridParams* d_gp = nullptr;
  GridParams* h_gp = nullptr;      // pinned
  int* h_counts = nullptr;         // pinned: [0] kept points, [1] leaves
  long long launches = 0;
  float4* d_acc = nullptr;         // the window map under accumulation (w_cloud)
  size_t acc_n = 0, acc_cap = 0;
};

static int pf_reserve(lvs_prefilter* p, size_t n) {
  CUDA_TRY(p->ws.reserve((int)n));
  if (n <= p->cap) return LVS_OK;
  for (void* q : {(void*)p->d_pts, (void*)p->d_kept, (void*)p->d_sorted, (void*)p->d_cell_start, (void*)p->d_out}) if (q) cudaFree(q);
  p->d_pts = p->d_kept = nullptr; p->d_sorted = p->d_cell_start = nullptr; p->d_out = nullptr; p->cap = 0;
  const size_t cap = n + n / 8 + 1024;
  CUDA_TRY(cudaMalloc(&p->d_pts, cap * sizeof(float4)));
  CUDA_TRY(cudaMalloc(&p->d_kept, cap * sizeof(float4)));
  CUDA_TRY(cudaMalloc(&p->d_sorted, cap * sizeof(int)));
  CUDA_TRY(cudaMalloc(&p->d_cell_start, (cap + 2) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&p->d_out, cap * 4 * sizeof(float)));
  p->cap = cap;
  return LVS_OK;
}

// The tail both entry points share: pcl::VoxelGrid of `kept` resident points (leaf_size <= 0: pass-through), result copied out.
static int pf_emit(lvs_prefilter* p, const float4* cloud, int kept, int n_fields, float leaf_size, float* out, size_t capacity, int out_on_device,
                   size_t* n_out, int* flags_out) {
  cudaStream_t st = p->st;
  int rc;
  const bool downsample = leaf_size > 0.0f;
  if (downsample && kept > 0) {
    vox_bbox(st, cloud, kept, p->ws, p->d_gp, leaf_size, (long long)INT_MAX);
    CUDA_TRY(cudaMemcpyAsync(p->h_gp, p->d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->launches += 1;
  }
  size_t n_result = (size_t)kept;
  const float* d_result = nullptr;
  bool passthrough = !downsample || kept == 0;
  if (downsample && kept > 0) {
    const GridParams& g = *p->h_gp;
    if (g.status == LVS_ERR_GRID_OVERFLOW) {           // "Leaf size is too small for the input dataset": output = input
      passthrough = true;
      if (flags_out) *flags_out |= 1;
    } else if (g.status == kStatusEmpty) {              // no finite point survives: VoxelGrid emits an empty cloud
      n_result = 0;
    } else {
      const int passes = vox_radix_passes(g.total_cells);
      if ((rc = vox_sort_segments(st, cloud, kept, p->d_gp, passes, p->ws, p->d_sorted, p->d_cell_start))) return rc;
      pf_centroid_kernel<<<(kept + kPfThreads - 1) / kPfThreads, kPfThreads, 0, st>>>(cloud, p->d_sorted, p->d_cell_start, p->ws.d_nseg, n_fields, p->d_out);
      CUDA_TRY(cudaMemcpyAsync(&p->h_counts[1], p->ws.d_nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      p->launches += 2 + passes * 4 + 1 + 2 + 1 + 1;
      n_result = (size_t)p->h_counts[1];
      d_result = p->d_out;
    }
  }
  if (passthrough && n_result > 0) {
    pf_unpack_kernel<<<((int)n_result + kPfThreads - 1) / kPfThreads, kPfThreads, 0, st>>>(cloud, (int)n_result, n_fields, p->d_out);
    p->launches += 1;
    d_result = p->d_out;
  }
  CUDA_TRY(cudaGetLastError());
  *n_out = n_result;
  if (n_result > capacity) return fail(LVS_ERR_INVALID_ARG, "output capacity %zu too small for %zu points", capacity, n_result);
  if (n_result > 0)
    CUDA_TRY(cudaMemcpyAsync(out, d_result, n_result * n_fields * sizeof(float), out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return LVS_OK;
}

extern "C" {

int lvs_prefilter_create(int device, void* stream, lvs_prefilter_t** out) {
  if (!out) return fail(LVS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { (void)cudaGetLastError(); return fail(LVS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback"); }
  if (device < 0 || device >= count) return fail(LVS_ERR_INVALID_ARG, "device %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  lvs_prefilter* p = new (std::nothrow) lvs_prefilter();
  if (!p) return fail(LVS_ERR_OOM, "out of host memory");
  p->device = device;
  if (stream) p->st = (cudaStream_t)stream;
  else { if (cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking) != cudaSuccess) { delete p; return fail(LVS_ERR_CUDA, "cudaStreamCreate failed"); } p->own_stream = true; }
  if (cudaMalloc(&p->d_gp, sizeof(GridParams)) != cudaSuccess || cudaMallocHost(&p->h_gp, sizeof(GridParams)) != cudaSuccess ||
      cudaMallocHost(&p->h_counts, 4 * sizeof(int)) != cudaSuccess) { delete p; return fail(LVS_ERR_OOM, "prefilter allocation failed"); }
  *out = p;
  return LVS_OK;
}

int lvs_prefilter_destroy(lvs_prefilter_t* p) {
  if (!p) return LVS_OK;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->st);
  p->ws.release();
  for (void* q : {(void*)p->d_in, (void*)p->d_pts, (void*)p->d_kept, (void*)p->d_sorted, (void*)p->d_cell_start, (void*)p->d_out, (void*)p->d_gp, (void*)p->d_acc}) if (q) cudaFree(q);
  if (p->h_gp) cudaFreeHost(p->h_gp);
  if (p->h_counts) cudaFreeHost(p->h_counts);
  if (p->own_stream) cudaStreamDestroy(p->st);
  delete p;
  return LVS_OK;
}

int lvs_prefilter_run(lvs_prefilter_t* p, const float* xyz, size_t n, size_t stride_bytes, int n_fields, int on_device, double distance_near,
                      double distance_far, int use_distance_filter, float leaf_size, float* out, size_t capacity, int out_on_device, size_t* n_out,
                      int* flags_out) {
  if (!p || !n_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *n_out = 0;
  if (flags_out) *flags_out = 0;
  if (n_fields != 3 && n_fields != 4) return fail(LVS_ERR_INVALID_ARG, "n_fields must be 3 (xyz) or 4 (xyz + intensity)");
  if (stride_bytes < (size_t)n_fields * 4 || (stride_bytes % 4) != 0) return fail(LVS_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and cover n_fields floats");
  if (n > (size_t)0x7fffff00) return fail(LVS_ERR_INVALID_ARG, "too many points");
  if (n > 0 && (!xyz || !out)) return fail(LVS_ERR_INVALID_ARG, "NULL cloud");
  CUDA_TRY(cudaSetDevice(p->device));
  if (n == 0) return LVS_OK;
  int rc = pf_reserve(p, n);
  if (rc) return rc;
  const size_t stride_floats = stride_bytes / 4;
  const float* d_in = xyz;
  if (!on_device) {
    const size_t bytes = (n - 1) * stride_bytes + (size_t)n_fields * 4;
    if (bytes > p->in_cap) {
      if (p->d_in) cudaFree(p->d_in);
      p->d_in = nullptr; p->in_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_in, bytes + bytes / 8 + 4096));
      p->in_cap = bytes + bytes / 8 + 4096;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_in, xyz, bytes, cudaMemcpyHostToDevice, p->st));
    d_in = p->d_in;
  }
  cudaStream_t st = p->st;
  const int ni = (int)n, gb = (ni + kPfThreads - 1) / kPfThreads;
  // ---- distance filter: flag, scan, scatter (order preserving)
  pf_flag_kernel<<<gb, kPfThreads, 0, st>>>(d_in, stride_floats, ni, n_fields, distance_near, distance_far, use_distance_filter ? 1 : 0, p->d_pts, p->ws.d_flags);
  CUDA_TRY(cudaMemsetAsync(p->ws.d_nseg, 0, 2 * sizeof(int), st));
  exclusive_scan(st, p->ws.d_flags, p->ws.d_pos, ni, p->ws.d_nseg, p->ws.d_tile_tot);
  pf_compact_kernel<<<gb, kPfThreads, 0, st>>>(p->d_pts, p->ws.d_flags, p->ws.d_pos, ni, p->d_kept);
  CUDA_TRY(cudaMemcpyAsync(&p->h_counts[0], p->ws.d_nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
  p->launches += 4;
  CUDA_TRY(cudaStreamSynchronize(st));        // the kept count shapes the launches that follow: one small read-back
  return pf_emit(p, p->d_kept, p->h_counts[0], n_fields, leaf_size, out, capacity, out_on_device, n_out, flags_out);
}

// ---- window map of the global-graph nodelet (src/global_graph/global_graph_nodelet.cpp:199-243): between two keyframes every scan
// is moved into the frame of the window's first scan (pcl::transformPointCloud with the DOUBLE matrix w_odom^-1 * odom) and
// appended to w_cloud; when the next keyframe is declared the window is downsampled with a 0.1 m VoxelGrid and becomes the
// keyframe's cloud.  accumulate_begin = w_cloud.clear(), accumulate_add = w_cloud += transformed, accumulate_flush = the VoxelGrid.
int lvs_prefilter_accumulate_begin(lvs_prefilter_t* p) {
  if (!p) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  p->acc_n = 0;
  return LVS_OK;
}

int lvs_prefilter_accumulate_add(lvs_prefilter_t* p, const float* xyz, size_t n, size_t stride_bytes, int n_fields, int on_device, const double* T16) {
  if (!p) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n_fields != 3 && n_fields != 4) return fail(LVS_ERR_INVALID_ARG, "n_fields must be 3 (xyz) or 4 (xyz + intensity)");
  if (stride_bytes < (size_t)n_fields * 4 || (stride_bytes % 4) != 0) return fail(LVS_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and cover n_fields floats");
  if (p->acc_n + n > (size_t)0x7fffff00) return fail(LVS_ERR_INVALID_ARG, "too many points");
  if (n == 0) return LVS_OK;
  if (!xyz) return fail(LVS_ERR_INVALID_ARG, "NULL cloud");
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t st = p->st;
  if (p->acc_n + n > p->acc_cap) {
    const size_t cap = (p->acc_n + n) * 2 + 4096;
    float4* bigger = nullptr;
    CUDA_TRY(cudaMalloc(&bigger, cap * sizeof(float4)));
    if (p->acc_n) CUDA_TRY(cudaMemcpyAsync(bigger, p->d_acc, p->acc_n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (p->d_acc) cudaFree(p->d_acc);
    p->d_acc = bigger; p->acc_cap = cap;
  }
  const float* d_in = xyz;
  if (!on_device) {
    const size_t bytes = (n - 1) * stride_bytes + (size_t)n_fields * 4;
    if (bytes > p->in_cap) {
      CUDA_TRY(cudaStreamSynchronize(st));
      if (p->d_in) cudaFree(p->d_in);
      p->d_in = nullptr; p->in_cap = 0;
      CUDA_TRY(cudaMalloc(&p->d_in, bytes + bytes / 8 + 4096));
      p->in_cap = bytes + bytes / 8 + 4096;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_in, xyz, bytes, cudaMemcpyHostToDevice, st));
    d_in = p->d_in;
  }
  PfMat M;
  for (int k = 0; k < 16; k++) M.m[k] = T16 ? T16[k] : ((k % 5) == 0 ? 1.0 : 0.0);
  pf_transform_append_kernel<<<((int)n + kPfThreads - 1) / kPfThreads, kPfThreads, 0, st>>>(d_in, stride_bytes / 4, (int)n, n_fields, M, T16 ? 1 : 0, p->d_acc + p->acc_n);
  CUDA_TRY(cudaGetLastError());
  if (!on_device) CUDA_TRY(cudaStreamSynchronize(st));     // the staging buffer is reused by the next call
  p->acc_n += n;
  p->launches += 1;
  return LVS_OK;
}

int lvs_prefilter_accumulate_flush(lvs_prefilter_t* p, int n_fields, float leaf_size, float* out, size_t capacity, int out_on_device, size_t* n_out,
                                   int* flags_out) {
  if (!p || !n_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *n_out = 0;
  if (flags_out) *flags_out = 0;
  if (n_fields != 3 && n_fields != 4) return fail(LVS_ERR_INVALID_ARG, "n_fields must be 3 (xyz) or 4 (xyz + intensity)");
  if (p->acc_n == 0) return LVS_OK;
  if (!out) return fail(LVS_ERR_INVALID_ARG, "NULL output");
  CUDA_TRY(cudaSetDevice(p->device));
  int rc = pf_reserve(p, p->acc_n);
  if (rc) return rc;
  return pf_emit(p, p->d_acc, (int)p->acc_n, n_fields, leaf_size, out, capacity, out_on_device, n_out, flags_out);
}

}  // extern "C"
