// C-ABI of the NDT path (include/lvslam_b200.h).  Host-side plumbing only: every number the reference's
// computeTransformation produces is computed by the kernels in ndt_voxel.cu / ndt_eval.cu and the device-resident
// state machine in ndt_state.cuh.  No CPU fallback: without a usable sm_100 device every entry point fails.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <new>
#include "ndt_internal.cuh"
#include "ndt_state.cuh"

namespace lvs {

static thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return status;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  (void)cudaGetLastError();
  int st = (e == cudaErrorMemoryAllocation) ? LVS_ERR_OOM
           : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) ? LVS_ERR_NO_DEVICE
                                                                                                          : LVS_ERR_CUDA;
  return fail(st, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
}

static AlignConsts make_consts(const lvs_ndt_params& p) {
  AlignConsts c;
  // gauss constants, eq. 6.8 [Magnusson 2009] as in ndt_omp_impl2.hpp:93-100
  double c1 = 10 * (1 - p.outlier_ratio);
  double c2 = p.outlier_ratio / std::pow((double)p.resolution, 3);
  c.gauss_d3 = -std::log(c2);
  c.gauss_d1 = -std::log(c1 + c2) - c.gauss_d3;
  c.gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - c.gauss_d3) / c.gauss_d1);
  c.step_size = p.step_size;
  c.trans_eps = p.transformation_epsilon;
  c.max_iter = p.max_iterations;
  c.search = p.search_method;
  c.variant = p.variant;
  c.resolution = p.resolution;
  c.fast = p.accumulation == LVS_ACC_FAST;
  c.lean_final = p.lean_final_evaluation != 0;
  return c;
}

static int check_params(const lvs_ndt_params* p) {
  if (!p) return fail(LVS_ERR_INVALID_ARG, "params is NULL");
  if (!(p->resolution > 0.0f)) return fail(LVS_ERR_INVALID_ARG, "resolution must be > 0");
  if (p->search_method < LVS_KDTREE || p->search_method > LVS_DIRECT1) return fail(LVS_ERR_INVALID_ARG, "unknown search_method %d", p->search_method);
  if (p->variant != LVS_NDT_OMP && p->variant != LVS_NDT_PCA && p->variant != LVS_NDT_GROUND) return fail(LVS_ERR_INVALID_ARG, "unknown variant %d", p->variant);
  if (p->max_iterations < 0 || p->max_iterations > kMaxTrace - 4) return fail(LVS_ERR_INVALID_ARG, "max_iterations must be in [0, %d]", kMaxTrace - 4);
  if (p->min_points_per_voxel < 1) return fail(LVS_ERR_INVALID_ARG, "min_points_per_voxel must be >= 1");
  if (p->accumulation != LVS_ACC_EXACT && p->accumulation != LVS_ACC_FAST) return fail(LVS_ERR_INVALID_ARG, "unknown accumulation %d", p->accumulation);
  if (p->variant == LVS_NDT_GROUND && p->accumulation != LVS_ACC_EXACT) return fail(LVS_ERR_INVALID_ARG, "LVS_NDT_GROUND runs with LVS_ACC_EXACT only");
  return LVS_OK;
}

struct CloudSlot {
  float4* d_pts = nullptr;
  size_t cap = 0;
  int n = 0;
  bool set = false;
  // Host clouds are copied and repacked on the batch's upload stream; `ready` marks the end of that work and is waited for by
  // the compute stream the first time the slot is consumed.  `used` marks the last compute-stream reader that was queued
  // without a host synchronisation (a voxelisation), so that a re-upload cannot overwrite points that are still being read.
  cudaEvent_t ready = nullptr, used = nullptr;
  bool ready_pending = false, used_pending = false;
  int n_total = 0;     // points of the whole cloud (> n when the batch keeps only its shard of every source)
  int up_lane = 0;     // upload stream of this slot (fixed, so that re-records of `ready` stay ordered behind earlier uploads)
};

constexpr int kBuildLanes = 4;
struct BuildLane {
  cudaStream_t st = nullptr;
  BuildScratch ws;
};
struct TargetBuildState {
  int lane = 0;
  cudaEvent_t built = nullptr;       // end of the queued voxelisation on its lane
  bool built_pending = false;
};

// Upload staging buffers.  The pool grows while the budget allows instead of making the host wait for a buffer, so a whole
// batch of clouds can be queued at once and the host is free to issue the aligns that consume the first ones.
constexpr size_t kStageBudgetBytes = (size_t)1 << 30;
constexpr int kStageGrow = 16;
constexpr int kUploadLanes = 2;
constexpr int kDeferGroup = 16;     // host clouds per repack launch of the plural setters
struct StageBuf {
  float* d = nullptr;
  size_t cap = 0;
  cudaEvent_t copied_ev = nullptr;   // end of the host-to-device copy into this buffer (copy stream)
  cudaEvent_t free_ev = nullptr;     // the buffer's own event, recorded after the repack kernel that read it (a slot's `ready` event
                                     // would not do: it is re-recorded by the slot's NEXT upload, which may already be queued when
                                     // batches are pipelined, and the buffer would look busy until that one completes)
  bool in_flight = false;
  bool pack_pending = false;         // copied, its repack waits for the group launch (plural setters)
};

constexpr int kMaxEvents = 2048;
constexpr int kWinRing = 64;          // per-launch events kept for the window bookkeeping (>= launches ever in flight)
constexpr int kWindow = 4;            // evaluation launches align_end keeps in flight while it polls the completion flag
constexpr int kMaxFirst = 32;         // launches align_begin may queue up front

}  // namespace lvs

using namespace lvs;

struct PendingAlign {          // the align between align_begin and align_end
  bool active = false;
  std::vector<int> src_slots, tgt_slots;     // slots the pairs in flight read: the setters refuse them until align_end
  int n_pairs = 0, launches = 0, max_launches = 0;
  int completed = 0;             // launches known to have finished (window bookkeeping of align_end)
  bool need_cold = false, prof = false;
  EvalLaunch L;
};

struct lvs_ndt_batch {
  int device = 0;
  PendingAlign pend;
  cudaStream_t st = nullptr;
  bool own_stream = false;
  lvs_ndt_params prm{};
  // voxelisations run on a few build lanes (stream + scratch each) so that the ~16 small kernels of one keyframe overlap
  // those of the next and the repacking of the scans; the compute stream waits for a grid the first time it is consumed
  BuildLane lanes[kBuildLanes];
  int lane_next = 0;
  BuildScratch batch_ws[kVoxBatch];   // scratch of the batched voxelisation (set_targets): one per cloud of a launch group, used on lanes[0].st
  cudaEvent_t ev_mark = nullptr;     // position of the compute stream, for resident inputs produced on it
  // plural setters with host clouds: the repacks of a group of clouds go into ONE launch behind the group's last copy (72 repack launches of
  // ~3 us each per bench step were the whole gap between the host-buffer and the resident step time)
  bool defer_packs = false;
  PackMany defer_pm;
  int defer_max_n = 0;
  CloudSlot* defer_slot[kPackMany];
  int defer_stage[kPackMany];
  cudaEvent_t ev_group = nullptr;
  std::vector<TargetGrid> targets;
  std::vector<TargetBuildState> tstate;
  std::vector<CloudSlot> target_pts, sources;
  // upload path: its own stream, a ring of staging buffers, and one more staging buffer for results on the compute stream
  cudaStream_t up[kUploadLanes] = {};   // [0] host-to-device copies only, [1] the repack kernels (each waits for its own copy)
  std::vector<StageBuf> ring;
  size_t ring_bytes = 0;
  int ring_next = 0;
  cudaEvent_t ev_up_all[kUploadLanes] = {};
  bool uploads_in_flight = false;
  float* d_stage = nullptr;
  size_t stage_cap = 0;
  // pair state
  int pair_cap = 0, bpp_cap = 0;
  PairDesc *d_pairs = nullptr, *h_pairs = nullptr;
  AlignState *d_states = nullptr, *h_states = nullptr;
  TraceRec *d_trace = nullptr;
  double* d_partials = nullptr;
  unsigned int* d_tickets = nullptr;
  int *d_done = nullptr, *h_done = nullptr;
  // completion without a stream synchronisation: the device stores the align's serial into this host-mapped word when the last pair
  // finishes (eval_finish); align_end polls it and keeps only a small window of launches in flight meanwhile
  volatile int* h_flag = nullptr;
  int* d_flag_alias = nullptr;
  int align_serial = 0;
  long long* d_dbg = nullptr;        // LVS_DEBUG_TIMING=1: managed buffer of tail clock stamps (diagnostics, tools/tail_timing.py)
  cudaStream_t rb = nullptr;         // read-back stream: results are copied out as soon as the flag is seen, ahead of idle launches still queued on st
  std::vector<cudaEvent_t> ev_win;   // ring of per-launch events (window bookkeeping)
  GridParams* h_gp_all = nullptr;   // pinned, one per target slot: batched geometry read-back
  float* d_T16 = nullptr;        // scratch 16 floats
  double *d_scalar = nullptr, *h_scalar = nullptr;
  // fitness score scratch (ndt_fitness.cu): best squared distance per source point, undecided list, CTA partials
  float* d_fit_best = nullptr; int* d_fit_list = nullptr; double* d_fit_partials = nullptr; unsigned int* d_fit_ticket = nullptr;
  size_t fit_cap = 0;
  int trace_on = 0;
  int last_n_pairs = 0;
  float last_final_T[16] = {};     // pair 0 of the last align (the parity taps reuse h_states[0], so the getters read these copies)
  int last_n_trace = 0;
  // stats
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  bool ev_end_pending = false;
  std::vector<cudaEvent_t> ev_pool;
  int profiling = 0;
  double last_device_ms = 0, last_deriv_ms = 0;
  int last_launches = 0, last_deriv_launches = 0;
  long total_launches = 0;
  long long h2d_bytes = 0, d2h_bytes = 0;   // bytes this object copied across PCIe since creation
  // point sharding (lvs_ndt_batch_shard_*)
  bool shard_on = false;
  int shard_rank = 0, shard_world = 1, shard_cap = 0;
  double* d_mail = nullptr;            // this rank's mailbox
  double** d_peers = nullptr;          // device array of every rank's mailbox as mapped here
  std::vector<double*> peer_ptrs;      // the same on the host; entries != d_mail were opened with cudaIpcOpenMemHandle
  long long shard_serial = 0;
  long long shard_timeout_cycles = 60000000000LL;   // ~30 s at 1.9 GHz (LVS_SHARD_TIMEOUT_S overrides): ranks are separate processes, a first-call
                                                     // cudaMalloc, a sanitizer or a loaded host may delay a peer's launch by seconds
  int *d_shard_error = nullptr, *h_shard_error = nullptr;
  int blocks_per_pair_override = 0;
  int chunk_first = 6, chunk_next = 4;
  int learned_first = 0;      // evaluations the previous align needed + 1: how many launches the next align queues up front (unless set_tuning fixed it)
  bool chunk_fixed = false;
};

namespace lvs {

static int set_device(lvs_ndt_batch* b) {
  CUDA_TRY(cudaSetDevice(b->device));
  return LVS_OK;
}

// Makes the compute stream wait for the slot's upload (once; later compute work is ordered behind that wait).
static int wait_ready(lvs_ndt_batch* b, CloudSlot& slot) {
  if (slot.ready_pending) {
    CUDA_TRY(cudaStreamWaitEvent(b->st, slot.ready, 0));
    slot.ready_pending = false;
  }
  return LVS_OK;
}

// Makes the compute stream wait for every upload queued so far (single-handle taps and getters).
static int wait_all_uploads(lvs_ndt_batch* b) {
  if (!b->uploads_in_flight) return LVS_OK;
  for (int l = 0; l < kUploadLanes; l++) {
    CUDA_TRY(cudaEventRecord(b->ev_up_all[l], b->up[l]));
    CUDA_TRY(cudaStreamWaitEvent(b->st, b->ev_up_all[l], 0));
  }
  b->uploads_in_flight = false;
  for (auto& c : b->target_pts) c.ready_pending = false;
  for (auto& c : b->sources) c.ready_pending = false;
  return LVS_OK;
}

// Launches the deferred repacks of the plural setters' host clouds: one kernel behind the last copy of the group, then every slot's
// `ready` and every staging buffer's `free` event.
static int flush_packs(lvs_ndt_batch* b) {
  PackMany& pm = b->defer_pm;
  if (pm.count == 0) return LVS_OK;
  cudaStream_t cp = b->up[0], pk = b->up[1];
  if (!b->ev_group) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_group, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(b->ev_group, cp));
  CUDA_TRY(cudaStreamWaitEvent(pk, b->ev_group, 0));
  int rc = pack_many(pk, pm, b->defer_max_n);
  if (rc) return rc;
  b->total_launches++;
  for (int k = 0; k < pm.count; k++) {
    StageBuf& sb = b->ring[b->defer_stage[k]];
    CUDA_TRY(cudaEventRecord(b->defer_slot[k]->ready, pk));
    CUDA_TRY(cudaEventRecord(sb.free_ev, pk));
    sb.in_flight = true; sb.pack_pending = false;
    b->defer_slot[k]->ready_pending = true;
  }
  pm.count = 0; b->defer_max_n = 0;
  b->uploads_in_flight = true;
  return LVS_OK;
}

static int upload_cloud(lvs_ndt_batch* b, CloudSlot& slot, const float* xyz, size_t n, size_t stride_bytes, int on_device,
                        cudaStream_t resident_stream) {
  if (n > 0 && !xyz) return fail(LVS_ERR_INVALID_ARG, "xyz is NULL");
  if (stride_bytes < 12 || (stride_bytes % 4) != 0) return fail(LVS_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and >= 12");
  if (n > (size_t)0x7fffff00) return fail(LVS_ERR_INVALID_ARG, "too many points");
  if (n > slot.cap) {
    if (slot.d_pts) cudaFree(slot.d_pts);      // cudaFree synchronises the device: no reader or writer of the old buffer is left
    slot.d_pts = nullptr; slot.cap = 0;
    size_t cap = n + n / 8 + 256;
    CUDA_TRY(cudaMalloc(&slot.d_pts, cap * sizeof(float4)));
    slot.cap = cap;
  }
  if (!slot.ready) {
    CUDA_TRY(cudaEventCreateWithFlags(&slot.ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&slot.used, cudaEventDisableTiming));
  }
  slot.n = (int)n;
  slot.n_total = (int)n;
  slot.set = true;
  if (n == 0) return LVS_OK;
  if (on_device) {
    // resident input: repack on the consumer's stream, behind any upload of this slot that is still in flight and behind
    // whatever the caller has queued on the handle's stream to produce the cloud
    if (slot.ready_pending) { CUDA_TRY(cudaStreamWaitEvent(resident_stream, slot.ready, 0)); slot.ready_pending = false; }
    if (resident_stream != b->st) {
      // while an align is in flight the handle's stream is full of its evaluation launches: order the repack behind the point
      // where align_begin found the stream (whatever the caller had queued by then), not behind the aligns themselves
      if (!b->pend.active) CUDA_TRY(cudaEventRecord(b->ev_mark, b->st));
      CUDA_TRY(cudaStreamWaitEvent(resident_stream, b->ev_mark, 0));
      if (slot.used_pending) { CUDA_TRY(cudaStreamWaitEvent(resident_stream, slot.used, 0)); slot.used_pending = false; }
    }
    int rc = pack_points(resident_stream, xyz, stride_bytes / 4, (int)n, slot.d_pts);
    b->total_launches++;
    return rc;
  }
  const size_t bytes = (n - 1) * stride_bytes + 12;
  // staging buffer: the oldest one if its repack has finished, else a new one while the budget allows, else wait for the oldest
  int pick = -1;
  if (!b->ring.empty() && b->ring[b->ring_next].pack_pending) { int rcf = flush_packs(b); if (rcf) return rcf; }   // the ring came round inside a group
  if (!b->ring.empty()) {
    StageBuf& old = b->ring[b->ring_next];
    if (!old.in_flight || cudaEventQuery(old.free_ev) == cudaSuccess) { old.in_flight = false; pick = b->ring_next; }
    else (void)cudaGetLastError();
  }
  if (pick < 0 && b->ring.size() < 4096 && b->defer_pm.count > 0) { int rcf = flush_packs(b); if (rcf) return rcf; }   // growing the ring renumbers its buffers
  if (pick < 0 && b->ring.size() < 4096) {
    // grow by several buffers at once (allocated here, not lazily): cudaMalloc stalls the pipeline, so the pool should reach
    // its steady size within the first batch instead of one buffer per batch
    const size_t cap = bytes + bytes / 8 + 4096;
    int added = 0;
    for (int k = 0; k < kStageGrow && b->ring_bytes + cap <= kStageBudgetBytes; k++) {
      StageBuf nb;
      if (cudaMalloc(&nb.d, cap) != cudaSuccess) { (void)cudaGetLastError(); break; }
      nb.cap = cap;
      b->ring_bytes += cap;
      b->ring.insert(b->ring.begin() + b->ring_next, nb);          // in front of the oldest: keeps the ring in age order
      added++;
    }
    if (added) pick = b->ring_next;
  }
  if (pick < 0) pick = b->ring_next;
  StageBuf& sb = b->ring[pick];
  b->ring_next = (pick + 1) % (int)b->ring.size();
  if (sb.in_flight) { CUDA_TRY(cudaEventSynchronize(sb.free_ev)); sb.in_flight = false; }
  if (bytes > sb.cap) {
    if (sb.d) cudaFree(sb.d);
    b->ring_bytes -= sb.cap;
    sb.d = nullptr; sb.cap = 0;
    size_t cap = bytes + bytes / 8 + 4096;
    CUDA_TRY(cudaMalloc(&sb.d, cap));
    sb.cap = cap;
    b->ring_bytes += cap;
  }
  // Copies and repacks live on different streams: the copy stream carries nothing but DMA transfers, so it never waits for an SM
  // (while a batch is being aligned the evaluation kernels hold every SM for the length of a launch, and a repack kernel queued
  // between two copies would stall all the copies behind it); the repack of a cloud waits for its own copy only.
  cudaStream_t cp = b->up[0], pk = b->up[1];
  if (slot.used_pending) { CUDA_TRY(cudaStreamWaitEvent(pk, slot.used, 0)); slot.used_pending = false; }
  CUDA_TRY(cudaMemcpyAsync(sb.d, xyz, bytes, cudaMemcpyHostToDevice, cp));
  b->h2d_bytes += (long long)bytes;
  if (!sb.free_ev) {
    CUDA_TRY(cudaEventCreateWithFlags(&sb.free_ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&sb.copied_ev, cudaEventDisableTiming));
  }
  if (b->defer_packs) {
    PackMany& pm = b->defer_pm;
    PackOne& c = pm.c[pm.count];
    c.in = sb.d; c.out = slot.d_pts; c.n = (int)n; c.stride_floats = (int)(stride_bytes / 4);
    b->defer_slot[pm.count] = &slot; b->defer_stage[pm.count] = pick;
    pm.count++;
    b->defer_max_n = std::max(b->defer_max_n, (int)n);
    sb.pack_pending = true;
    if (pm.count == kDeferGroup) return flush_packs(b);
    return LVS_OK;
  }
  CUDA_TRY(cudaEventRecord(sb.copied_ev, cp));
  CUDA_TRY(cudaStreamWaitEvent(pk, sb.copied_ev, 0));
  int rc = pack_points(pk, sb.d, stride_bytes / 4, (int)n, slot.d_pts);
  b->total_launches++;
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(slot.ready, pk));
  CUDA_TRY(cudaEventRecord(sb.free_ev, pk));
  sb.in_flight = true;
  slot.ready_pending = true;
  b->uploads_in_flight = true;
  return LVS_OK;
}

// Completes every queued voxelisation among `slots` with ONE synchronisation (plus a rebuild for the rare slot whose
// bounding box outgrew its index grid).
static int finish_targets(lvs_ndt_batch* b, const int32_t* slots, int n) {
  bool any = false;
  for (int i = 0; i < n; i++) {
    TargetGrid& t = b->targets[slots[i]];
    TargetBuildState& ts = b->tstate[slots[i]];
    if (ts.built_pending) { CUDA_TRY(cudaStreamWaitEvent(b->st, ts.built, 0)); ts.built_pending = false; }
    if (t.pending && !t.fetch_queued) {
      CUDA_TRY(cudaMemcpyAsync(&b->h_gp_all[slots[i]], t.d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, b->st));
      b->d2h_bytes += sizeof(GridParams);
      t.fetch_queued = true;
      any = true;
    }
  }
  if (!any) return LVS_OK;
  CUDA_TRY(cudaStreamSynchronize(b->st));
  CUDA_TRY(cudaGetLastError());
  for (int i = 0; i < n; i++) {
    TargetGrid& t = b->targets[slots[i]];
    if (!t.fetch_queued) continue;
    t.fetch_queued = false;
    // the builds this loop waits for have completed, so the compute stream and the lane's scratch are free for a re-build
    BuildLane& ln = b->lanes[b->tstate[slots[i]].lane];
    if (b->h_gp_all[slots[i]].status == kStatusNeedsGrow) CUDA_TRY(cudaStreamSynchronize(ln.st));   // a later build may be using the scratch
    int rc = t.accept(b->h_gp_all[slots[i]], b->st, ln.ws);
    if (rc) return rc;
    if (t.pending) { b->total_launches += t.launches_last_build; if ((rc = t.finish(b->st, ln.ws))) return rc; }
  }
  return LVS_OK;
}

static int finish_target(lvs_ndt_batch* b, int slot) { const int32_t s = slot; return finish_targets(b, &s, 1); }

static int reserve_pairs(lvs_ndt_batch* b, int n_pairs, int bpp) {
  if (n_pairs > b->pair_cap) {
    if (b->d_pairs) cudaFree(b->d_pairs);
    if (b->d_states) cudaFree(b->d_states);
    if (b->d_trace) cudaFree(b->d_trace);
    if (b->d_tickets) cudaFree(b->d_tickets);
    if (b->h_pairs) cudaFreeHost(b->h_pairs);
    if (b->h_states) cudaFreeHost(b->h_states);
    b->d_pairs = nullptr; b->d_states = nullptr; b->d_trace = nullptr; b->d_tickets = nullptr; b->h_pairs = nullptr; b->h_states = nullptr;
    b->pair_cap = 0;
    int cap = n_pairs + n_pairs / 4 + 4;
    CUDA_TRY(cudaMalloc(&b->d_pairs, cap * sizeof(PairDesc)));
    CUDA_TRY(cudaMalloc(&b->d_states, cap * sizeof(AlignState)));
    if (b->trace_on) CUDA_TRY(cudaMalloc(&b->d_trace, (size_t)cap * kMaxTrace * sizeof(TraceRec)));
    CUDA_TRY(cudaMalloc(&b->d_tickets, cap * sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(b->d_tickets, 0, cap * sizeof(unsigned int), b->st));
    CUDA_TRY(cudaMallocHost(&b->h_pairs, cap * sizeof(PairDesc)));
    CUDA_TRY(cudaMallocHost(&b->h_states, cap * sizeof(AlignState)));
    b->pair_cap = cap;
    b->bpp_cap = 0;
  }
  if ((size_t)b->pair_cap * bpp > (size_t)b->bpp_cap) {
    if (b->d_partials) cudaFree(b->d_partials);
    b->d_partials = nullptr;
    CUDA_TRY(cudaMalloc(&b->d_partials, (size_t)b->pair_cap * bpp * kPartialStride * sizeof(double)));
    b->bpp_cap = b->pair_cap * bpp;
  }
  return LVS_OK;
}

static PairDesc make_pair(lvs_ndt_batch* b, int src_slot, int tgt_slot) {
  PairDesc P;
  const TargetGrid& tg = b->targets[tgt_slot];
  P.src = b->sources[src_slot].d_pts;
  P.n_src = b->sources[src_slot].n;
  P.n_total = b->sources[src_slot].n_total;
  P.grid = tg.d_grid;
  P.recs = tg.d_recs;
  P.frecs = tg.d_frecs;
  P.centroids = tg.d_centroids;
  P.icov64 = tg.d_icov64;
  P.gp = tg.d_gp;
  return P;
}

// CTAs per pair.  All CTAs of a launch are identical in cost to first order, so what matters is wave quantisation: the launch
// runs ceil(total / resident) waves and the last one should be full.  A CTA should also iterate a few times (its prologue and
// the 43-double reduction are fixed costs) unless that would leave SMs idle.
static int choose_bpp(lvs_ndt_batch* b, int n_pairs, int max_src) {
  if (b->blocks_per_pair_override > 0) return b->blocks_per_pair_override;
  const int ppi = eval_points_per_cta_iteration();
  const int by_points = std::max(1, (max_src + ppi - 1) / ppi);          // one full iteration per CTA
  const int resident = 148 * eval_max_resident_ctas_per_sm();
  if ((long long)n_pairs * by_points <= resident) {
    // cannot fill one wave with full iterations: spread the points over every resident CTA instead (the warps of a pair take equal
    // contiguous ranges), down to 64 points per CTA - what a single-pair align waits for is the slowest warp
    const int finest = std::max(1, (max_src + 63) / 64);
    return std::max(1, std::min(resident / n_pairs, finest));
  }
  const int coarse = std::max(1, by_points / 4);                          // >= 4 iterations per CTA
  int best = 1;
  double best_eff = -1.0;
  for (int bpp = 1; bpp <= std::max(coarse, (resident + n_pairs - 1) / n_pairs); bpp++) {
    if (bpp > by_points) break;
    const long long total = (long long)n_pairs * bpp;
    const long long waves = (total + resident - 1) / resident;
    const double eff = (double)total / (double)(waves * resident);
    if (eff >= best_eff - 1e-9) { best_eff = std::max(eff, best_eff); best = bpp; }
  }
  return best;
}

static void shard_view(lvs_ndt_batch* b, EvalLaunch& L) {
  if (!b->shard_on) return;
  L.shard.peers = b->d_peers; L.shard.mine = b->d_mail;
  L.shard.rank = b->shard_rank; L.shard.world = b->shard_world; L.shard.cap = b->shard_cap;
  L.shard.timeout_cycles = b->shard_timeout_cycles;
  L.shard.d_error = b->d_shard_error;
}

static int shard_check(lvs_ndt_batch* b) {
  if (!b->shard_on) return LVS_OK;
  if (*b->h_shard_error) return fail(LVS_ERR_PEER, "point-sharded evaluation: a peer rank did not deliver its sums within the time-out");
  return LVS_OK;
}

// Runs the evaluation launches of one batch until every pair's state machine reports done.
// One evaluation launch (plus the cold kernel when reachable) of the align in flight, bracketed by profiling events on request.
static int align_launch_one(lvs_ndt_batch* b) {
  PendingAlign& P = b->pend;
  int rc;
  const bool ev = P.prof && P.launches < kMaxEvents;
  if (ev) CUDA_TRY(cudaEventRecord(b->ev_pool[2 * P.launches], b->st));
  if ((rc = launch_eval(b->st, P.L))) return rc;
  if (P.need_cold) { if ((rc = launch_eval_cold(b->st, P.L))) return rc; }
  if (ev) CUDA_TRY(cudaEventRecord(b->ev_pool[2 * P.launches + 1], b->st));
  CUDA_TRY(cudaEventRecord(b->ev_win[P.launches % kWinRing], b->st));
  P.launches++;
  return LVS_OK;
}

// First half of an align: uploads the pair states and queues the first chunk of evaluation launches (enough for a typical align;
// finished pairs leave their CTAs at once), then returns without waiting.  The Newton / More-Thuente state machine lives on the
// device, so nothing else is needed from the host until align_end collects the results.
static int align_begin(lvs_ndt_batch* b, int n_pairs, const int32_t* src_slot, const int32_t* tgt_slot, const float* guesses16) {
  int rc = set_device(b);
  if (rc) return rc;
  if (b->pend.active) return fail(LVS_ERR_INVALID_ARG, "an align is already in flight: call align_end first");
  if (n_pairs <= 0) { b->pend = PendingAlign(); b->pend.active = true; return LVS_OK; }
  int max_src = 0;
  for (int i = 0; i < n_pairs; i++) {
    int s = src_slot[i], t = tgt_slot[i];
    if (s < 0 || s >= (int)b->sources.size() || t < 0 || t >= (int)b->targets.size()) return fail(LVS_ERR_BAD_SLOT, "pair %d: slot out of range", i);
    if (!b->target_pts[t].set) return fail(LVS_ERR_NO_TARGET, "pair %d: target slot %d has no cloud (setInputTarget not called)", i, t);
    if (!b->sources[s].set) return fail(LVS_ERR_NO_SOURCE, "pair %d: source slot %d has no cloud (setInputSource not called)", i, s);
    max_src = std::max(max_src, b->sources[s].n);
  }
  for (int i = 0; i < n_pairs; i++)
    if ((rc = wait_ready(b, b->sources[src_slot[i]]))) return rc;
  if ((rc = finish_targets(b, tgt_slot, n_pairs))) return rc;
  const int bpp = choose_bpp(b, n_pairs, max_src);
  if ((rc = reserve_pairs(b, n_pairs, bpp))) return rc;
  for (int i = 0; i < n_pairs; i++) {
    b->h_pairs[i] = make_pair(b, src_slot[i], tgt_slot[i]);
    align_state_init(b->h_states[i], guesses16 + 16 * i, b->trace_on);
  }
  CUDA_TRY(cudaEventRecord(b->ev_begin, b->st));
  CUDA_TRY(cudaEventRecord(b->ev_mark, b->st));      // position of the stream before this align (see upload_cloud)
  CUDA_TRY(cudaMemcpyAsync(b->d_pairs, b->h_pairs, n_pairs * sizeof(PairDesc), cudaMemcpyHostToDevice, b->st));
  CUDA_TRY(cudaMemcpyAsync(b->d_states, b->h_states, n_pairs * sizeof(AlignState), cudaMemcpyHostToDevice, b->st));
  b->h2d_bytes += (long long)n_pairs * (sizeof(PairDesc) + sizeof(AlignState));
  CUDA_TRY(cudaMemsetAsync(b->d_done, 0, sizeof(int), b->st));
  PendingAlign& P = b->pend;
  P = PendingAlign();
  P.src_slots.assign(src_slot, src_slot + n_pairs);
  P.tgt_slots.assign(tgt_slot, tgt_slot + n_pairs);
  EvalLaunch& L = P.L;
  L.d_pairs = b->d_pairs; L.d_states = b->d_states; L.d_trace = b->trace_on ? b->d_trace : nullptr;
  L.d_partials = b->d_partials; L.d_tickets = b->d_tickets; L.d_done_count = b->d_done;
  L.n_pairs = n_pairs; L.blocks_per_pair = bpp; L.advance = 1;
  L.consts = make_consts(b->prm);
  if (b->shard_on) { CUDA_TRY(cudaMemsetAsync(b->d_shard_error, 0, sizeof(int), b->st)); *b->h_shard_error = 0; }
  if (b->shard_on && n_pairs > b->shard_cap) return fail(LVS_ERR_INVALID_ARG, "%d pairs exceed the sharded batch's max_pairs %d", n_pairs, b->shard_cap);
  shard_view(b, L);
  L.h_done_flag = b->d_flag_alias; L.align_serial = ++b->align_serial;
  L.d_dbg = b->d_dbg;
  if (L.consts.fast && getenv("LVS_STAGE_RECORDS")) {      // experiment switch: voxel records of the pair staged in shared memory by a bulk copy
    int max_cells = 0;
    for (int i = 0; i < n_pairs; i++) max_cells = std::max(max_cells, b->targets[tgt_slot[i]].n_cells);
    const long long bytes = (long long)max_cells * (long long)sizeof(FastRec);
    L.stage_bytes = (bytes > 0 && bytes <= 160 * 1024) ? (int)((bytes + 15) & ~15LL) : 0;
  }
  // worst case: initial pass + (max_iter + 2) outer iterations of (first + 10 trials + Hessian pass)
  P.max_launches = 1 + (b->prm.max_iterations + 2) * 12 + 8;
  L.shard.serial = b->shard_serial;            // base of this align; every pair adds its own evaluation count (eval_finish)
  b->shard_serial += 2 * P.max_launches + 2;
  P.n_pairs = n_pairs;
  *b->h_done = 0;
  P.prof = b->profiling != 0;
  if (P.prof && (int)b->ev_pool.size() < 2 * kMaxEvents) {
    size_t old = b->ev_pool.size();
    b->ev_pool.resize(2 * kMaxEvents);
    for (size_t k = old; k < b->ev_pool.size(); k++) CUDA_TRY(cudaEventCreate(&b->ev_pool[k]));
  }
  // The radius-search passes are only reachable in KDTREE mode or when the reference's More-Thuente loop can run, i.e.
  // when `interval_converged = (step_max - step_min) > 0` (ndt_omp_impl2.hpp:888) is false.
  P.need_cold = b->prm.search_method == LVS_KDTREE || !((b->prm.step_size - b->prm.transformation_epsilon / 2) > 0);
  // launches queued before the first completion check: every one beyond what the pairs need is a grid of CTAs that find nothing
  // to do, so the count follows the previous align (streams are self-similar) unless the caller fixed it
  const int first = std::min(kMaxFirst, (!b->chunk_fixed && b->learned_first > 0) ? b->learned_first : b->chunk_first);
  for (int k = 0; k < first && P.launches < P.max_launches; k++)
    if ((rc = align_launch_one(b))) return rc;
  P.active = true;
  return LVS_OK;
}

// Second half: waits for the queued launches, keeps launching in chunks while pairs are still iterating, reads the results back.
static int align_end(lvs_ndt_batch* b, lvs_ndt_result* results) {
  int rc = set_device(b);
  if (rc) return rc;
  PendingAlign& P = b->pend;
  if (!P.active) return fail(LVS_ERR_INVALID_ARG, "no align in flight: call align_begin first");
  P.active = false;
  const int n_pairs = P.n_pairs;
  if (n_pairs <= 0) return LVS_OK;
  const int max_launches = P.max_launches;
  const bool need_cold = P.need_cold, prof = P.prof;
  // Poll the host-mapped completion word instead of synchronising the stream; meanwhile keep a small window of evaluation launches
  // in flight (a launch that finds every pair finished returns at once, so the few that outlive the align cost microseconds).
  const int serial = P.L.align_serial;
  bool done = false;
  for (long spins = 0;; spins++) {
    if (*b->h_flag == serial) { done = true; break; }
    while (P.completed < P.launches) {
      cudaError_t q = cudaEventQuery(b->ev_win[P.completed % kWinRing]);
      if (q == cudaSuccess) P.completed++;
      else if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }
      else return cuda_fail(q, "cudaEventQuery", __FILE__, __LINE__);
    }
    if (P.launches - P.completed < std::max(kWindow, b->chunk_next) && P.launches < max_launches) {
      if ((rc = align_launch_one(b))) return rc;
      continue;
    }
    if (P.completed == P.launches && P.launches >= max_launches) { done = (*b->h_flag == serial); break; }
    if (b->shard_on && (spins & 0xfff) == 0xfff) {      // a lost peer never completes: look at the error word from time to time
      CUDA_TRY(cudaMemcpyAsync(b->h_shard_error, b->d_shard_error, sizeof(int), cudaMemcpyDeviceToHost, b->rb));
      CUDA_TRY(cudaStreamSynchronize(b->rb));
      if ((rc = shard_check(b))) return rc;
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  const int launches = P.launches;
  // every state is complete and fenced once the flag is visible: copy the results on the read-back stream, ahead of whatever idle
  // launches are still queued on the compute stream
  CUDA_TRY(cudaMemcpyAsync(b->h_states, b->d_states, n_pairs * sizeof(AlignState), cudaMemcpyDeviceToHost, b->rb));
  b->d2h_bytes += (long long)n_pairs * sizeof(AlignState);
  if (b->shard_on) CUDA_TRY(cudaMemcpyAsync(b->h_shard_error, b->d_shard_error, sizeof(int), cudaMemcpyDeviceToHost, b->rb));
  CUDA_TRY(cudaEventRecord(b->ev_end, b->st));
  b->ev_end_pending = true;
  CUDA_TRY(cudaStreamSynchronize(b->rb));
  if ((rc = shard_check(b))) return rc;
  if (!done) return fail(LVS_ERR_CUDA, "align state machine did not finish within %d evaluation launches", max_launches);
  b->last_device_ms = -1.0;                      // computed on demand (lvs_ndt_batch_last_stats): needs ev_end, i.e. the idle launches too
  b->last_launches = launches;
  b->total_launches += launches * (need_cold ? 2 : 1);
  b->last_n_pairs = n_pairs;
  memcpy(b->last_final_T, b->h_states[0].final_T, sizeof b->last_final_T);
  b->last_n_trace = b->h_states[0].n_trace;
  int active = 0;
  for (int i = 0; i < n_pairs; i++) {
    const AlignState& s = b->h_states[i];
    active = std::max(active, s.n_eval + s.n_hess);
    if (results) {
      lvs_ndt_result& r = results[i];
      memcpy(r.final_transformation, s.final_T, sizeof r.final_transformation);
      r.converged = s.converged;
      r.iterations = s.nr_iterations;
      r.trans_probability = s.trans_probability;
      r.n_eval = s.n_eval;
      r.n_hess = s.n_hess;
      r.score = s.score;
    }
  }
  b->last_deriv_launches = active;
  b->learned_first = std::max(2, std::min(active + 1, kMaxFirst));
  b->last_deriv_ms = 0;
  if (prof) {
    CUDA_TRY(cudaEventSynchronize(b->ev_end));
    double sum = 0;
    for (int k = 0; k < std::min(active, kMaxEvents); k++) {
      float t = 0;
      CUDA_TRY(cudaEventElapsedTime(&t, b->ev_pool[2 * k], b->ev_pool[2 * k + 1]));
      sum += t;
    }
    b->last_deriv_ms = sum;
  }
  return LVS_OK;
}

static int run_align(lvs_ndt_batch* b, int n_pairs, const int32_t* src_slot, const int32_t* tgt_slot, const float* guesses16, lvs_ndt_result* results) {
  int rc = align_begin(b, n_pairs, src_slot, tgt_slot, guesses16);
  if (rc) return rc;          // begin raises `active` only when everything is queued
  return align_end(b, results);
}

// Tap: one evaluation of pair 0 = (source 0, target 0) with an explicit state.
static int run_tap(lvs_ndt_batch* b, int kind, const double p[6], const float* T16, double* score, double* g, double* H36) {
  int rc = set_device(b);
  if (rc) return rc;
  if (b->pend.active) return fail(LVS_ERR_INVALID_ARG, "an align is in flight: call align_end first");
  if (!b->target_pts[0].set) return fail(LVS_ERR_NO_TARGET, "setInputTarget not called");
  if (!b->sources[0].set) return fail(LVS_ERR_NO_SOURCE, "setInputSource not called");
  if ((rc = wait_all_uploads(b))) return rc;
  if ((rc = finish_target(b, 0))) return rc;
  const int bpp = choose_bpp(b, 1, b->sources[0].n);
  if ((rc = reserve_pairs(b, 1, bpp))) return rc;
  b->h_pairs[0] = make_pair(b, 0, 0);
  AlignState& s = b->h_states[0];
  const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  align_state_init(s, I4, 0);
  state_set_eval_point(s, p);            // T = float(SE3::exp(p)), R from p
  if (T16) memcpy(s.T, T16, sizeof s.T);  // explicit transformed cloud
  s.eval_kind = kind;
  CUDA_TRY(cudaMemcpyAsync(b->d_pairs, b->h_pairs, sizeof(PairDesc), cudaMemcpyHostToDevice, b->st));
  CUDA_TRY(cudaMemcpyAsync(b->d_states, b->h_states, sizeof(AlignState), cudaMemcpyHostToDevice, b->st));
  EvalLaunch L;
  L.d_pairs = b->d_pairs; L.d_states = b->d_states; L.d_trace = nullptr;
  L.d_partials = b->d_partials; L.d_tickets = b->d_tickets; L.d_done_count = b->d_done;
  L.n_pairs = 1; L.blocks_per_pair = bpp; L.advance = 0;
  L.consts = make_consts(b->prm);
  shard_view(b, L);
  L.shard.serial = b->shard_serial; b->shard_serial += 4;
  if (kind == EVAL_HESS27 || b->prm.search_method == LVS_KDTREE) rc = launch_eval_cold(b->st, L);
  else rc = launch_eval(b->st, L);
  if (rc) return rc;
  b->total_launches++;
  CUDA_TRY(cudaMemcpyAsync(b->h_states, b->d_states, sizeof(AlignState), cudaMemcpyDeviceToHost, b->st));
  if (b->shard_on) CUDA_TRY(cudaMemcpyAsync(b->h_shard_error, b->d_shard_error, sizeof(int), cudaMemcpyDeviceToHost, b->st));
  CUDA_TRY(cudaStreamSynchronize(b->st));
  if ((rc = shard_check(b))) return rc;
  if (score) *score = s.score;
  if (g) memcpy(g, s.g, sizeof s.g);
  if (H36) memcpy(H36, s.H, sizeof s.H);
  return LVS_OK;
}

static int batch_create(const lvs_ndt_params* params, int device, void* stream, int n_t, int n_s, int trace_on, lvs_ndt_batch** out) {
  if (!out) return fail(LVS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  lvs_ndt_params p;
  if (params) p = *params; else lvs_ndt_default_params(&p);
  int rc = check_params(&p);
  if (rc) return rc;
  if (n_t < 1 || n_s < 1) return fail(LVS_ERR_INVALID_ARG, "slot counts must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { (void)cudaGetLastError(); return fail(LVS_ERR_NO_DEVICE, "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)); }
  if (device < 0 || device >= ndev) return fail(LVS_ERR_NO_DEVICE, "device %d out of range (have %d)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(LVS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  lvs_ndt_batch* b = new (std::nothrow) lvs_ndt_batch();
  if (!b) return fail(LVS_ERR_OOM, "host allocation failed");
  b->device = device;
  b->prm = p;
  b->trace_on = trace_on;
  b->targets.resize(n_t); b->target_pts.resize(n_t); b->sources.resize(n_s); b->tstate.resize(n_t);
  auto bail = [&](int st) { lvs_ndt_batch_destroy(b); return st; };
  if (stream) b->st = (cudaStream_t)stream;
  else {
    e = cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
    b->own_stream = true;
  }
  for (int l = 0; l < kUploadLanes; l++) {
    if ((e = cudaStreamCreateWithFlags(&b->up[l], cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
    if ((e = cudaEventCreateWithFlags(&b->ev_up_all[l], cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__));
  }
  for (int i = 0; i < n_t; i++) b->target_pts[i].up_lane = i % kUploadLanes;
  for (int i = 0; i < n_s; i++) b->sources[i].up_lane = (i + n_t) % kUploadLanes;
  if ((e = cudaEventCreateWithFlags(&b->ev_mark, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__));
  for (auto& ln : b->lanes)
    if ((e = cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
  if ((e = cudaMalloc(&b->d_done, 4 * sizeof(int))) != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc", __FILE__, __LINE__));
  if ((e = cudaMallocHost(&b->h_done, 4 * sizeof(int))) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__));
  {
    int* hf = nullptr;
    if ((e = cudaHostAlloc(&hf, 64, cudaHostAllocMapped)) != cudaSuccess) return bail(cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__));
    hf[0] = 0;
    b->h_flag = hf;
    if ((e = cudaHostGetDevicePointer((void**)&b->d_flag_alias, hf, 0)) != cudaSuccess) return bail(cuda_fail(e, "cudaHostGetDevicePointer", __FILE__, __LINE__));
  }
  if ((e = cudaStreamCreateWithFlags(&b->rb, cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
  if (getenv("LVS_DEBUG_TIMING")) {
    if ((e = cudaMallocManaged(&b->d_dbg, 8 * 64 * sizeof(long long))) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocManaged", __FILE__, __LINE__));
    memset(b->d_dbg, 0, 8 * 64 * sizeof(long long));
  }
  b->ev_win.assign(kWinRing, nullptr);
  for (auto& ev : b->ev_win)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__));
  if ((e = cudaMallocHost(&b->h_gp_all, n_t * sizeof(GridParams))) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__));
  if ((e = cudaMalloc(&b->d_T16, 16 * sizeof(float))) != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc", __FILE__, __LINE__));
  if ((e = cudaMalloc(&b->d_scalar, 8 * sizeof(double))) != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc", __FILE__, __LINE__));
  if ((e = cudaMallocHost(&b->h_scalar, 8 * sizeof(double))) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__));
  if ((e = cudaEventCreate(&b->ev_begin)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__));
  if ((e = cudaEventCreate(&b->ev_end)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__));
  *out = b;
  return LVS_OK;
}

// A cloud that the align in flight reads must not be replaced before align_end.
static int slot_in_flight(const lvs_ndt_batch* b, bool target, int slot) {
  if (!b->pend.active) return LVS_OK;
  const std::vector<int>& v = target ? b->pend.tgt_slots : b->pend.src_slots;
  if (std::find(v.begin(), v.end(), slot) != v.end())
    return fail(LVS_ERR_INVALID_ARG, "%s slot %d is read by the align in flight: call align_end first", target ? "target" : "source", slot);
  return LVS_OK;
}

static int set_target(lvs_ndt_batch* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  int rc = set_device(b);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)b->targets.size()) return fail(LVS_ERR_BAD_SLOT, "target slot %d out of range", slot);
  if ((rc = slot_in_flight(b, true, slot))) return rc;
  CloudSlot& cs = b->target_pts[slot];
  TargetBuildState& ts = b->tstate[slot];
  // a grid that was re-queued on the compute stream (grown index grid) or consumed there is complete by now: every consumer
  // synchronises the host.  A build still in flight on another lane has to finish before this one touches the same arrays.
  const int lane = b->lane_next;
  b->lane_next = (b->lane_next + 1) % kBuildLanes;
  BuildLane& ln = b->lanes[lane];
  if (ts.built_pending && ts.lane != lane) CUDA_TRY(cudaStreamWaitEvent(ln.st, ts.built, 0));
  if ((rc = upload_cloud(b, cs, xyz, n, stride_bytes, on_device, ln.st))) return rc;
  if (cs.ready_pending) { CUDA_TRY(cudaStreamWaitEvent(ln.st, cs.ready, 0)); cs.ready_pending = false; }
  rc = b->targets[slot].build(ln.st, cs.d_pts, (int)n, b->prm, ln.ws);
  b->total_launches += b->targets[slot].launches_last_build;
  if (rc) return rc;
  if (!ts.built) CUDA_TRY(cudaEventCreateWithFlags(&ts.built, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(ts.built, ln.st));
  ts.built_pending = true;
  ts.lane = lane;
  if (cs.used) {
    CUDA_TRY(cudaEventRecord(cs.used, ln.st));
    cs.used_pending = true;
  }
  return LVS_OK;
}

// Plural setInputTarget: the clouds are uploaded / repacked one by one, then voxelised TOGETHER by one set of batched launches
// (TargetGrid::enqueue_many) on the first build lane's stream - about 20 launches per group of kVoxBatch clouds instead of 20 per
// cloud.  A slot that is built for the first time (no index grid yet) sizes its grid inside prepare(), exactly like set_target.
static int set_targets_many(lvs_ndt_batch* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts, size_t stride_bytes,
                            int on_device) {
  int rc = set_device(b);
  if (rc) return rc;
  for (int i = 0; i < n; i++) {
    if (slots[i] < 0 || slots[i] >= (int)b->targets.size()) return fail(LVS_ERR_BAD_SLOT, "target slot %d out of range", slots[i]);
    if ((rc = slot_in_flight(b, true, slots[i]))) return rc;
  }
  BuildLane& ln = b->lanes[0];
  for (int base = 0; base < n; base += kVoxBatch) {
    const int cnt = std::min(kVoxBatch, n - base);
    TargetGrid* grids[kVoxBatch];
    BuildScratch* wss[kVoxBatch];
    int slot_of[kVoxBatch];
    int m = 0;
    // host clouds: all copies of the group first, their repacks in one launch (flush_packs), then the voxelisation behind them
    for (int k = 0; k < cnt; k++)
      for (int q = 0; q < k; q++)
        if (slots[base + q] == slots[base + k]) return fail(LVS_ERR_INVALID_ARG, "target slot %d appears twice in one set_targets call", slots[base + k]);
    b->defer_packs = !on_device;
    for (int k = 0; k < cnt && !rc; k++) {
      const int slot = slots[base + k];
      TargetBuildState& ts = b->tstate[slot];
      if (ts.built_pending && ts.lane != 0) { cudaError_t e = cudaStreamWaitEvent(ln.st, ts.built, 0); if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamWaitEvent", __FILE__, __LINE__); }
      if (!rc) rc = upload_cloud(b, b->target_pts[slot], xyz[base + k], counts[base + k], stride_bytes, on_device, ln.st);
    }
    b->defer_packs = false;
    { const int rcf = flush_packs(b); if (!rc) rc = rcf; }
    if (rc) return rc;
    for (int k = 0; k < cnt; k++) {
      const int slot = slots[base + k];
      CloudSlot& cs = b->target_pts[slot];
      TargetBuildState& ts = b->tstate[slot];
      if (cs.ready_pending) { CUDA_TRY(cudaStreamWaitEvent(ln.st, cs.ready, 0)); cs.ready_pending = false; }
      rc = b->targets[slot].prepare(ln.st, cs.d_pts, (int)counts[base + k], b->prm, b->batch_ws[m]);
      if (rc < 0) return rc;
      ts.lane = 0;
      if (rc == 0) { ts.built_pending = false; continue; }      // empty cloud: nothing queued
      grids[m] = &b->targets[slot]; wss[m] = &b->batch_ws[m]; slot_of[m] = slot;
      m++;
    }
    if (m == 0) continue;
    if ((rc = TargetGrid::enqueue_many(ln.st, m, grids, wss, b->prm))) return rc;
    for (int q = 0; q < m; q++) {
      const int slot = slot_of[q];
      TargetBuildState& ts = b->tstate[slot];
      CloudSlot& cs = b->target_pts[slot];
      b->total_launches += b->targets[slot].launches_last_build;
      if (!ts.built) CUDA_TRY(cudaEventCreateWithFlags(&ts.built, cudaEventDisableTiming));
      CUDA_TRY(cudaEventRecord(ts.built, ln.st));
      ts.built_pending = true;
      if (cs.used) { CUDA_TRY(cudaEventRecord(cs.used, ln.st)); cs.used_pending = true; }
    }
  }
  return LVS_OK;
}

// The chunk of an n-point source this rank evaluates: [rank*n/world, (rank+1)*n/world).
static void shard_range(const lvs_ndt_batch* b, size_t n, size_t* lo, size_t* cnt) {
  if (!b->shard_on) { *lo = 0; *cnt = n; return; }
  const size_t a = (size_t)b->shard_rank * n / (size_t)b->shard_world, e = (size_t)(b->shard_rank + 1) * n / (size_t)b->shard_world;
  *lo = a; *cnt = e - a;
}

static int set_source(lvs_ndt_batch* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  int rc = set_device(b);
  if (rc) return rc;
  if (slot < 0 || slot >= (int)b->sources.size()) return fail(LVS_ERR_BAD_SLOT, "source slot %d out of range", slot);
  if ((rc = slot_in_flight(b, false, slot))) return rc;
  size_t lo = 0, cnt = n;
  shard_range(b, n, &lo, &cnt);
  rc = upload_cloud(b, b->sources[slot], xyz ? (const float*)((const char*)xyz + lo * stride_bytes) : xyz, cnt, stride_bytes, on_device, b->st);
  b->sources[slot].n_total = (int)n;
  return rc;
}

// Batched setInputSource: host clouds go through the upload stream one by one (each keeps its own ready event); resident
// clouds are repacked by one launch per kPackMany clouds instead of one launch each.
static int set_sources(lvs_ndt_batch* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts, size_t stride_bytes,
                       int on_device) {
  int rc = set_device(b);
  if (rc) return rc;
  for (int i = 0; i < n; i++) {
    if (slots[i] < 0 || slots[i] >= (int)b->sources.size()) return fail(LVS_ERR_BAD_SLOT, "source slot %d out of range", slots[i]);
    if ((rc = slot_in_flight(b, false, slots[i]))) return rc;
  }
  if (!on_device) {
    b->defer_packs = true;
    for (int i = 0; i < n && !rc; i++) rc = set_source(b, slots[i], xyz[i], counts[i], stride_bytes, 0);
    b->defer_packs = false;
    const int rcf = flush_packs(b);
    return rc ? rc : rcf;
  }
  if (stride_bytes < 12 || (stride_bytes % 4) != 0) return fail(LVS_ERR_INVALID_ARG, "stride_bytes must be a multiple of 4 and >= 12");
  PackMany pm;
  pm.count = 0;
  int max_n = 0;
  for (int i = 0; i < n; i++) {
    CloudSlot& slot = b->sources[slots[i]];
    // allocation and bookkeeping exactly as upload_cloud, the repack itself is deferred to the fused launch
    if ((rc = upload_cloud(b, slot, xyz[i], 0, stride_bytes, 1, b->st))) return rc;
    size_t lo = 0, cnt = counts[i];
    shard_range(b, counts[i], &lo, &cnt);
    slot.n_total = (int)counts[i];
    if (cnt > 0 && !xyz[i]) return fail(LVS_ERR_INVALID_ARG, "xyz is NULL");
    if (cnt > (size_t)0x7fffff00) return fail(LVS_ERR_INVALID_ARG, "too many points");
    if (cnt > slot.cap) {
      if (slot.d_pts) cudaFree(slot.d_pts);
      slot.d_pts = nullptr; slot.cap = 0;
      size_t cap = cnt + cnt / 8 + 256;
      CUDA_TRY(cudaMalloc(&slot.d_pts, cap * sizeof(float4)));
      slot.cap = cap;
    }
    slot.n = (int)cnt;
    if (cnt == 0) continue;
    if ((rc = wait_ready(b, slot))) return rc;
    PackOne& c = pm.c[pm.count++];
    c.in = (const float*)((const char*)xyz[i] + lo * stride_bytes); c.out = slot.d_pts; c.n = (int)cnt; c.stride_floats = (int)(stride_bytes / 4);
    max_n = std::max(max_n, (int)cnt);
    if (pm.count == kPackMany || i == n - 1) {
      if ((rc = pack_many(b->st, pm, max_n))) return rc;
      b->total_launches++;
      pm.count = 0; max_n = 0;
    }
  }
  if (pm.count > 0) {
    if ((rc = pack_many(b->st, pm, max_n))) return rc;
    b->total_launches++;
  }
  return LVS_OK;
}

}  // namespace lvs

struct lvs_ndt {
  lvs_ndt_batch* b = nullptr;
};

extern "C" {

void lvs_ndt_default_params(lvs_ndt_params* p) {
  if (!p) return;
  p->resolution = 1.0f;            // ndt_omp_impl2.hpp:54-83
  p->step_size = 0.1;
  p->outlier_ratio = 0.55;
  p->transformation_epsilon = 0.1;
  p->max_iterations = 35;
  p->search_method = LVS_DIRECT7;
  p->variant = LVS_NDT_OMP;
  p->min_points_per_voxel = 6;
  p->min_covar_eigvalue_mult = 0.01;
  p->accumulation = LVS_ACC_EXACT;
  p->lean_final_evaluation = 0;
}

const char* lvs_status_string(int status) {
  switch (status) {
    case LVS_OK: return "ok";
    case LVS_ERR_INVALID_ARG: return "invalid argument";
    case LVS_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case LVS_ERR_CUDA: return "CUDA error";
    case LVS_ERR_OOM: return "out of memory";
    case LVS_ERR_NO_TARGET: return "no target cloud";
    case LVS_ERR_NO_SOURCE: return "no source cloud";
    case LVS_ERR_GRID_OVERFLOW: return "voxel grid index overflow";
    case LVS_ERR_BAD_SLOT: return "bad slot";
    case LVS_ERR_NOT_SPD: return "matrix not positive definite";
    case LVS_ERR_EMPTY_GRAPH: return "empty graph";
    default: return "unknown status";
  }
}

const char* lvs_last_error(void) { return g_err; }

int lvs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;
}

// ---- batch object
int lvs_ndt_batch_create(const lvs_ndt_params* params, int device, void* stream, int n_target_slots, int n_source_slots, lvs_ndt_batch_t** out) {
  return batch_create(params, device, stream, n_target_slots, n_source_slots, 0, out);
}

int lvs_ndt_batch_destroy(lvs_ndt_batch_t* b) {
  if (!b) return LVS_OK;
  cudaSetDevice(b->device);
  for (auto u : b->up) if (u) cudaStreamSynchronize(u);
  for (auto& ln : b->lanes) if (ln.st) cudaStreamSynchronize(ln.st);
  if (b->st) cudaStreamSynchronize(b->st);
  for (auto& w : b->batch_ws) { w.release(); if (w.h_gp) cudaFreeHost(w.h_gp); }
  for (auto& t : b->targets) t.release();
  for (auto* v : {&b->target_pts, &b->sources})
    for (auto& c : *v) {
      if (c.d_pts) cudaFree(c.d_pts);
      if (c.ready) cudaEventDestroy(c.ready);
      if (c.used) cudaEventDestroy(c.used);
    }
  for (auto& sb : b->ring) {
    if (sb.d) cudaFree(sb.d);
    if (sb.free_ev) cudaEventDestroy(sb.free_ev);
    if (sb.copied_ev) cudaEventDestroy(sb.copied_ev);
  }
  for (int r = 0; r < (int)b->peer_ptrs.size(); r++)
    if (b->peer_ptrs[r] && b->peer_ptrs[r] != b->d_mail) cudaIpcCloseMemHandle(b->peer_ptrs[r]);
  if (b->d_mail) cudaFree(b->d_mail);
  if (b->d_peers) cudaFree(b->d_peers);
  if (b->d_shard_error) cudaFree(b->d_shard_error);
  if (b->h_shard_error) cudaFreeHost(b->h_shard_error);
  for (auto ev : b->ev_up_all) if (ev) cudaEventDestroy(ev);
  for (auto u : b->up) if (u) cudaStreamDestroy(u);
  for (auto& ln : b->lanes) {
    ln.ws.release();
    if (ln.ws.h_gp) cudaFreeHost(ln.ws.h_gp);
    if (ln.st) cudaStreamDestroy(ln.st);
  }
  for (auto& ts : b->tstate) if (ts.built) cudaEventDestroy(ts.built);
  if (b->ev_mark) cudaEventDestroy(b->ev_mark);
  if (b->ev_group) cudaEventDestroy(b->ev_group);
  if (b->d_stage) cudaFree(b->d_stage);
  if (b->d_pairs) cudaFree(b->d_pairs);
  if (b->d_states) cudaFree(b->d_states);
  if (b->d_trace) cudaFree(b->d_trace);
  if (b->d_partials) cudaFree(b->d_partials);
  if (b->d_tickets) cudaFree(b->d_tickets);
  if (b->d_done) cudaFree(b->d_done);
  if (b->d_T16) cudaFree(b->d_T16);
  if (b->d_scalar) cudaFree(b->d_scalar);
  if (b->d_fit_best) cudaFree(b->d_fit_best);
  if (b->d_fit_list) cudaFree(b->d_fit_list);
  if (b->d_fit_partials) cudaFree(b->d_fit_partials);
  if (b->d_fit_ticket) cudaFree(b->d_fit_ticket);
  if (b->h_pairs) cudaFreeHost(b->h_pairs);
  if (b->h_states) cudaFreeHost(b->h_states);
  if (b->h_done) cudaFreeHost(b->h_done);
  if (b->h_flag) cudaFreeHost((void*)b->h_flag);
  if (b->d_dbg) cudaFree(b->d_dbg);
  if (b->rb) { cudaStreamSynchronize(b->rb); cudaStreamDestroy(b->rb); }
  for (auto ev : b->ev_win) if (ev) cudaEventDestroy(ev);
  if (b->h_gp_all) cudaFreeHost(b->h_gp_all);
  if (b->h_scalar) cudaFreeHost(b->h_scalar);
  if (b->ev_begin) cudaEventDestroy(b->ev_begin);
  if (b->ev_end) cudaEventDestroy(b->ev_end);
  for (auto e : b->ev_pool) cudaEventDestroy(e);
  if (b->own_stream && b->st) cudaStreamDestroy(b->st);
  (void)cudaGetLastError();
  delete b;
  return LVS_OK;
}

int lvs_ndt_batch_set_target(lvs_ndt_batch_t* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  return set_target(b, slot, xyz, n, stride_bytes, on_device);
}

int lvs_ndt_batch_set_source(lvs_ndt_batch_t* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  return set_source(b, slot, xyz, n, stride_bytes, on_device);
}

int lvs_ndt_batch_set_sources(lvs_ndt_batch_t* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts, size_t stride_bytes,
                              int on_device) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n < 0 || (n > 0 && (!slots || !xyz || !counts))) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return set_sources(b, n, slots, xyz, counts, stride_bytes, on_device);
}

int lvs_ndt_batch_set_targets(lvs_ndt_batch_t* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts, size_t stride_bytes,
                              int on_device) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n < 0 || (n > 0 && (!slots || !xyz || !counts))) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (n == 1) return set_target(b, slots[0], xyz[0], counts[0], stride_bytes, on_device);
  return set_targets_many(b, n, slots, xyz, counts, stride_bytes, on_device);
}

int lvs_ndt_batch_align(lvs_ndt_batch_t* b, int n_pairs, const int32_t* source_slot, const int32_t* target_slot, const float* guesses16,
                        lvs_ndt_result* results) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n_pairs < 0 || (n_pairs > 0 && (!source_slot || !target_slot || !guesses16 || !results))) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return run_align(b, n_pairs, source_slot, target_slot, guesses16, results);
}

int lvs_ndt_batch_align_begin(lvs_ndt_batch_t* b, int n_pairs, const int32_t* source_slot, const int32_t* target_slot, const float* guesses16) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n_pairs < 0 || (n_pairs > 0 && (!source_slot || !target_slot || !guesses16))) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return align_begin(b, n_pairs, source_slot, target_slot, guesses16);
}

int lvs_ndt_batch_align_end(lvs_ndt_batch_t* b, lvs_ndt_result* results) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  return align_end(b, results);
}

int lvs_ndt_batch_last_stats(lvs_ndt_batch_t* b, double* device_ms, int* launches, double* deriv_kernel_ms, int* deriv_launches) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (b->ev_end_pending && b->last_device_ms < 0) {
    float ms = 0;
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaEventSynchronize(b->ev_end));
    CUDA_TRY(cudaEventElapsedTime(&ms, b->ev_begin, b->ev_end));
    b->last_device_ms = ms;
  }
  if (device_ms) *device_ms = b->last_device_ms;
  if (launches) *launches = b->last_launches;
  if (deriv_kernel_ms) *deriv_kernel_ms = b->last_deriv_ms;
  if (deriv_launches) *deriv_launches = b->last_deriv_launches;
  return LVS_OK;
}

int lvs_ndt_batch_set_profiling(lvs_ndt_batch_t* b, int on) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  b->profiling = on;
  return LVS_OK;
}

int lvs_ndt_batch_set_tuning(lvs_ndt_batch_t* b, int blocks_per_pair, int chunk_first, int chunk_next) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  b->blocks_per_pair_override = blocks_per_pair > 0 ? blocks_per_pair : 0;
  if (chunk_first > 0) { b->chunk_first = chunk_first; b->chunk_fixed = true; }
  if (chunk_next > 0) b->chunk_next = chunk_next;
  return LVS_OK;
}

int lvs_ndt_batch_total_launches(lvs_ndt_batch_t* b, long long* launches) {
  if (!b || !launches) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *launches = b->total_launches;
  return LVS_OK;
}

int lvs_ndt_batch_transfer_bytes(lvs_ndt_batch_t* b, long long* h2d, long long* d2h) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (h2d) *h2d = b->h2d_bytes;
  if (d2h) *d2h = b->d2h_bytes;
  return LVS_OK;
}

int lvs_ndt_batch_shard_init(lvs_ndt_batch_t* b, int rank, int world, int max_pairs, unsigned char handle_out[LVS_IPC_HANDLE_BYTES]) {
  if (!b || !handle_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (world < 1 || world > 64 || rank < 0 || rank >= world || max_pairs < 1) return fail(LVS_ERR_INVALID_ARG, "bad rank / world / max_pairs");
  if (b->d_mail) return fail(LVS_ERR_INVALID_ARG, "shard_init was already called on this object");
  static_assert(sizeof(cudaIpcMemHandle_t) == LVS_IPC_HANDLE_BYTES, "IPC handle size");
  int rc = set_device(b);
  if (rc) return rc;
  const size_t bytes = (size_t)2 * world * max_pairs * kMailStride * sizeof(double);
  CUDA_TRY(cudaMalloc(&b->d_mail, bytes));
  CUDA_TRY(cudaMemset(b->d_mail, 0, bytes));
  CUDA_TRY(cudaMalloc(&b->d_peers, world * sizeof(double*)));
  CUDA_TRY(cudaMalloc(&b->d_shard_error, sizeof(int)));
  CUDA_TRY(cudaMemset(b->d_shard_error, 0, sizeof(int)));
  CUDA_TRY(cudaMallocHost(&b->h_shard_error, sizeof(int)));
  *b->h_shard_error = 0;
  b->shard_rank = rank; b->shard_world = world; b->shard_cap = max_pairs;
  if (const char* ts = getenv("LVS_SHARD_TIMEOUT_S")) { const double sec = atof(ts); if (sec > 0) b->shard_timeout_cycles = (long long)(sec * 1.9e9); }
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, b->d_mail));
  memcpy(handle_out, &h, sizeof h);
  return LVS_OK;
}

int lvs_ndt_batch_shard_connect(lvs_ndt_batch_t* b, const unsigned char* handles) {
  if (!b || !handles) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (!b->d_mail) return fail(LVS_ERR_INVALID_ARG, "shard_init has not been called");
  if (b->shard_on) return fail(LVS_ERR_INVALID_ARG, "already connected");
  int rc = set_device(b);
  if (rc) return rc;
  b->peer_ptrs.assign(b->shard_world, nullptr);
  for (int r = 0; r < b->shard_world; r++) {
    if (r == b->shard_rank) { b->peer_ptrs[r] = b->d_mail; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * LVS_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    b->peer_ptrs[r] = (double*)p;
  }
  CUDA_TRY(cudaMemcpy(b->d_peers, b->peer_ptrs.data(), b->shard_world * sizeof(double*), cudaMemcpyHostToDevice));
  b->shard_on = true;
  // sources set before this point hold whole clouds: they have to be set again
  for (auto& c : b->sources) c.set = false;
  return LVS_OK;
}

int lvs_ndt_batch_wait_uploads(lvs_ndt_batch_t* b) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  int rc = set_device(b);
  if (rc) return rc;
  for (auto u : b->up) CUDA_TRY(cudaStreamSynchronize(u));
  return LVS_OK;
}

int lvs_ndt_batch_num_cells(lvs_ndt_batch_t* b, int slot, int* n_cells, int* n_valid) {
  if (!b) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (slot < 0 || slot >= (int)b->targets.size()) return fail(LVS_ERR_BAD_SLOT, "target slot %d out of range", slot);
  int rc = set_device(b);
  if (!rc) rc = finish_target(b, slot);
  if (rc) return rc;
  if (n_cells) *n_cells = b->targets[slot].n_cells;
  if (n_valid) *n_valid = b->targets[slot].gp.n_valid;
  return LVS_OK;
}

// ---- single registration object
int lvs_ndt_create(const lvs_ndt_params* params, int device, void* stream, lvs_ndt_t** out) {
  if (!out) return fail(LVS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  lvs_ndt_batch* b = nullptr;
  int rc = batch_create(params, device, stream, 1, 1, 1, &b);
  if (rc) return rc;
  lvs_ndt* h = new (std::nothrow) lvs_ndt();
  if (!h) { lvs_ndt_batch_destroy(b); return fail(LVS_ERR_OOM, "host allocation failed"); }
  h->b = b;
  *out = h;
  return LVS_OK;
}

int lvs_ndt_destroy(lvs_ndt_t* h) {
  if (!h) return LVS_OK;
  lvs_ndt_batch_destroy(h->b);
  delete h;
  return LVS_OK;
}

int lvs_ndt_set_params(lvs_ndt_t* h, const lvs_ndt_params* params) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  int rc = check_params(params);
  if (rc) return rc;
  lvs_ndt_batch* b = h->b;
  const bool revox = b->target_pts[0].set && (params->resolution != b->prm.resolution || params->variant != b->prm.variant ||
                                              params->min_points_per_voxel != b->prm.min_points_per_voxel ||
                                              params->min_covar_eigvalue_mult != b->prm.min_covar_eigvalue_mult);
  b->prm = *params;
  if (revox) {   // setResolution re-runs init() when the value changed and a target is set (ndt_omp.h:126-136)
    if ((rc = set_device(b))) return rc;
    if ((rc = wait_all_uploads(b))) return rc;
    if ((rc = finish_target(b, 0))) return rc;     // the queued build (if any) is complete and its lane idle after this
    BuildLane& ln = b->lanes[b->tstate[0].lane];
    CUDA_TRY(cudaStreamSynchronize(ln.st));
    rc = b->targets[0].build(b->st, b->target_pts[0].d_pts, b->target_pts[0].n, b->prm, ln.ws);
    b->total_launches += b->targets[0].launches_last_build;
    if (rc) return rc;
    // this rebuild runs on the compute stream with the lane's scratch and is tracked by no event: finish it here, so that a
    // following setInputTarget (another lane, same grid arrays and point buffer) cannot run against it
    CUDA_TRY(cudaStreamSynchronize(b->st));
    return LVS_OK;
  }
  return LVS_OK;
}

int lvs_ndt_get_params(const lvs_ndt_t* h, lvs_ndt_params* out) {
  if (!h || !out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *out = h->b->prm;
  return LVS_OK;
}

int lvs_ndt_set_target(lvs_ndt_t* h, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  return set_target(h->b, 0, xyz, n, stride_bytes, on_device);
}

int lvs_ndt_set_source(lvs_ndt_t* h, const float* xyz, size_t n, size_t stride_bytes, int on_device) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  return set_source(h->b, 0, xyz, n, stride_bytes, on_device);
}

int lvs_ndt_align(lvs_ndt_t* h, const float guess[16], lvs_ndt_result* out) {
  if (!h || !guess || !out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  const int32_t zero = 0;
  return run_align(h->b, 1, &zero, &zero, guess, out);
}

int lvs_ndt_get_aligned_cloud(lvs_ndt_t* h, float* xyz_out, int on_device) {
  if (!h || !xyz_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (rc) return rc;
  if (!b->sources[0].set) return fail(LVS_ERR_NO_SOURCE, "setInputSource not called");
  if (b->last_n_pairs < 1) return fail(LVS_ERR_INVALID_ARG, "align() has not run");
  const int n = b->sources[0].n;
  if (n == 0) return LVS_OK;
  if ((rc = wait_all_uploads(b))) return rc;
  CUDA_TRY(cudaMemcpyAsync(b->d_T16, b->last_final_T, 16 * sizeof(float), cudaMemcpyHostToDevice, b->st));
  float* d_out = xyz_out;
  if (!on_device) {
    size_t bytes = (size_t)n * 12;
    if (bytes > b->stage_cap) {
      if (b->d_stage) cudaFree(b->d_stage);
      b->d_stage = nullptr; b->stage_cap = 0;
      CUDA_TRY(cudaMalloc(&b->d_stage, bytes + 4096));
      b->stage_cap = bytes + 4096;
    }
    d_out = b->d_stage;
  }
  if ((rc = launch_transform(b->st, b->sources[0].d_pts, n, b->d_T16, d_out))) return rc;
  b->total_launches++;
  if (!on_device) CUDA_TRY(cudaMemcpyAsync(xyz_out, d_out, (size_t)n * 12, cudaMemcpyDeviceToHost, b->st));
  CUDA_TRY(cudaStreamSynchronize(b->st));
  return LVS_OK;
}

int lvs_ndt_get_trace(lvs_ndt_t* h, lvs_ndt_trace_rec* recs, int capacity, int* n_out) {
  if (!h || !n_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (rc) return rc;
  if (b->last_n_pairs < 1) { *n_out = 0; return LVS_OK; }
  int n = std::min(b->last_n_trace, (int)kMaxTrace);
  *n_out = n;
  n = std::min(n, capacity);
  if (n > 0 && recs) {
    static_assert(sizeof(lvs_ndt_trace_rec) == sizeof(TraceRec), "trace record layout");
    CUDA_TRY(cudaMemcpyAsync(recs, b->d_trace, (size_t)n * sizeof(TraceRec), cudaMemcpyDeviceToHost, b->st));
    CUDA_TRY(cudaStreamSynchronize(b->st));
  }
  return LVS_OK;
}

int lvs_ndt_eval_derivatives(lvs_ndt_t* h, const double p[6], const float* T16, int compute_hessian, double* score, double g[6], double H36[36]) {
  if (!h || !p) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return run_tap(h->b, compute_hessian ? EVAL_DERIV_H : EVAL_DERIV_NOH, p, T16, score, g, H36);
}

int lvs_ndt_eval_hessian(lvs_ndt_t* h, const double p[6], const float* T16, double H36[36]) {
  if (!h || !p) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return run_tap(h->b, EVAL_HESS27, p, T16, nullptr, nullptr, H36);
}

int lvs_ndt_calculate_score(lvs_ndt_t* h, const float T16[16], double* score) {
  if (!h || !T16 || !score) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (rc) return rc;
  if (!b->target_pts[0].set) return fail(LVS_ERR_NO_TARGET, "setInputTarget not called");
  if (!b->sources[0].set) return fail(LVS_ERR_NO_SOURCE, "setInputSource not called");
  if (b->shard_on) return fail(LVS_ERR_INVALID_ARG, "calculateScore is not available on a point-sharded object");
  if (b->sources[0].n == 0) { *score = NAN; return LVS_OK; }   // 0/0 in the reference
  if ((rc = wait_all_uploads(b))) return rc;
  if ((rc = finish_target(b, 0))) return rc;
  const int bpp = choose_bpp(b, 1, b->sources[0].n);
  if ((rc = reserve_pairs(b, 1, bpp))) return rc;
  PairDesc P = make_pair(b, 0, 0);
  CUDA_TRY(cudaMemcpyAsync(b->d_T16, T16, 16 * sizeof(float), cudaMemcpyHostToDevice, b->st));
  if ((rc = launch_calc_score(b->st, P, b->d_T16, make_consts(b->prm), b->d_partials, bpp * kAcc, b->d_tickets, b->d_scalar))) return rc;
  b->total_launches++;
  CUDA_TRY(cudaMemcpyAsync(b->h_scalar, b->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, b->st));
  CUDA_TRY(cudaStreamSynchronize(b->st));
  *score = b->h_scalar[0];
  return LVS_OK;
}

// getFitnessScore of the pair (source slot, target slot) of a batch object under T16.
static int fitness_score(lvs_ndt_batch* b, int src_slot, int tgt_slot, const float* T16, double max_range, double* score, int* n_corr) {
  int rc = set_device(b);
  if (rc) return rc;
  if (tgt_slot < 0 || tgt_slot >= (int)b->targets.size() || src_slot < 0 || src_slot >= (int)b->sources.size()) return fail(LVS_ERR_BAD_SLOT, "slot out of range");
  if (!b->target_pts[tgt_slot].set) return fail(LVS_ERR_NO_TARGET, "setInputTarget not called");
  if (!b->sources[src_slot].set) return fail(LVS_ERR_NO_SOURCE, "setInputSource not called");
  if (b->shard_on) return fail(LVS_ERR_INVALID_ARG, "getFitnessScore is not available on a point-sharded object");
  if (b->pend.active) return fail(LVS_ERR_INVALID_ARG, "an align is in flight: call align_end first");
  if ((rc = wait_all_uploads(b))) return rc;
  if ((rc = finish_target(b, tgt_slot))) return rc;
  const CloudSlot& src = b->sources[src_slot];
  TargetGrid& tg = b->targets[tgt_slot];
  if (!tg.sorted_pts_valid && tg.n_cells > 0) {      // first query against this build: put the target points into cell and slab order
    if ((size_t)tg.n_cells > tg.slab_cells) {
      if (tg.d_slab) cudaFree(tg.d_slab);
      tg.d_slab = nullptr; tg.slab_cells = 0;
      const size_t cells = (size_t)tg.n_cells + tg.n_cells / 4 + 64;
      CUDA_TRY(cudaMalloc(&tg.d_slab, cells * kFitSlabsPerCell * sizeof(int)));
      tg.slab_cells = cells;
    }
    int gl = 0;
    if ((rc = launch_fitness_gather(b->st, b->target_pts[tgt_slot].d_pts, b->target_pts[tgt_slot].n, tg.n_cells, tg.d_sorted_idx, tg.d_cell_start, tg.d_gp,
                                    tg.d_slab, tg.d_sorted_pts, &gl))) return rc;
    b->total_launches += gl;
    tg.sorted_pts_valid = true;
  }
  if ((size_t)src.n + 1 > b->fit_cap) {
    if (b->d_fit_best) cudaFree(b->d_fit_best);
    if (b->d_fit_list) cudaFree(b->d_fit_list);
    b->d_fit_best = nullptr; b->d_fit_list = nullptr; b->fit_cap = 0;
    const size_t cap = (size_t)src.n + src.n / 8 + 64;
    CUDA_TRY(cudaMalloc(&b->d_fit_best, cap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b->d_fit_list, 2 * (cap + 1) * sizeof(int)));      // list 1 (undecided after the 27-cell block), list 2 (brute force)
    b->fit_cap = cap;
  }
  if (!b->d_fit_partials) {
    CUDA_TRY(cudaMalloc(&b->d_fit_partials, 2 * 64 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&b->d_fit_ticket, sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(b->d_fit_ticket, 0, sizeof(unsigned int), b->st));
  }
  CUDA_TRY(cudaMemcpyAsync(b->d_T16, T16, 16 * sizeof(float), cudaMemcpyHostToDevice, b->st));
  FitnessArgs a;
  a.src = src.d_pts; a.n_src = src.n;
  a.tgt = b->target_pts[tgt_slot].d_pts; a.n_tgt = b->target_pts[tgt_slot].n;
  a.grid = tg.d_grid; a.gp = tg.d_gp; a.cell_start = tg.d_cell_start; a.sorted_idx = tg.d_sorted_idx; a.tgt_sorted = tg.d_sorted_pts; a.slab_end = tg.d_slab;
  a.T16 = b->d_T16; a.max_range = max_range;
  a.best = b->d_fit_best; a.list = b->d_fit_list; a.list2 = b->d_fit_list + (b->fit_cap + 1); a.partials = b->d_fit_partials; a.ticket = b->d_fit_ticket; a.out = b->d_scalar;
  int launches = 0;
  if ((rc = launch_fitness(b->st, a, &launches))) return rc;
  b->total_launches += launches;
  CUDA_TRY(cudaMemcpyAsync(b->h_scalar, b->d_scalar, 2 * sizeof(double), cudaMemcpyDeviceToHost, b->st));
  b->d2h_bytes += 2 * sizeof(double);
  CUDA_TRY(cudaStreamSynchronize(b->st));
  if (score) *score = b->h_scalar[0];
  if (n_corr) *n_corr = (int)b->h_scalar[1];
  return LVS_OK;
}

int lvs_ndt_fitness_score(lvs_ndt_t* h, const float* T16, double max_range, double* score, int* n_correspondences) {
  if (!h || !score) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  if (!T16) {
    if (b->last_n_pairs < 1) return fail(LVS_ERR_INVALID_ARG, "align() has not run and no transformation was given");
    T16 = b->last_final_T;
  }
  return fitness_score(b, 0, 0, T16, max_range, score, n_correspondences);
}

int lvs_ndt_batch_fitness_score(lvs_ndt_batch_t* b, int source_slot, int target_slot, const float T16[16], double max_range, double* score,
                                int* n_correspondences) {
  if (!b || !T16 || !score) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  return fitness_score(b, source_slot, target_slot, T16, max_range, score, n_correspondences);
}

int lvs_ndt_get_grid(lvs_ndt_t* h, int32_t min_b[3], int32_t max_b[3], int32_t div_b[3]) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  { int rc = set_device(h->b); if (!rc) rc = finish_target(h->b, 0); if (rc) return rc; }
  const GridParams& g = h->b->targets[0].gp;
  for (int a = 0; a < 3; a++) {
    if (min_b) min_b[a] = g.min_b[a];
    if (max_b) max_b[a] = g.max_b[a];
    if (div_b) div_b[a] = g.div_b[a];
  }
  return LVS_OK;
}

int lvs_ndt_num_cells(lvs_ndt_t* h, int* n_cells) {
  if (!h || !n_cells) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  { int rc = set_device(h->b); if (!rc) rc = finish_target(h->b, 0); if (rc) return rc; }
  *n_cells = h->b->targets[0].n_cells;
  return LVS_OK;
}

int lvs_ndt_get_cells(lvs_ndt_t* h, int32_t* keys, int32_t* nr_points, double* mean3, double* icov9, double* evals3, float* centroid3, int32_t* weight) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (!rc) rc = finish_target(b, 0);
  if (rc) return rc;
  const TargetGrid& t = b->targets[0];
  const int n = t.n_cells;
  if (n == 0) return LVS_OK;
  CUDA_TRY(cudaStreamSynchronize(b->st));
  if (keys) CUDA_TRY(cudaMemcpy(keys, t.d_cell_keys, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (nr_points) CUDA_TRY(cudaMemcpy(nr_points, t.d_cell_npts, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (icov9) CUDA_TRY(cudaMemcpy(icov9, t.d_icov64, (size_t)n * 72, cudaMemcpyDeviceToHost));
  if (evals3) CUDA_TRY(cudaMemcpy(evals3, t.d_cell_evals, (size_t)n * 24, cudaMemcpyDeviceToHost));
  if (mean3 || weight) {
    std::vector<VoxelRec> recs(n);
    CUDA_TRY(cudaMemcpy(recs.data(), t.d_recs, (size_t)n * sizeof(VoxelRec), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
      if (mean3) for (int a = 0; a < 3; a++) mean3[i * 3 + a] = recs[i].mean[a];
      if (weight) weight[i] = recs[i].meta & kMetaWeightMask;
    }
  }
  if (centroid3) {
    std::vector<float4> c(n);
    CUDA_TRY(cudaMemcpy(c.data(), t.d_centroids, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) { centroid3[i * 3] = c[i].x; centroid3[i * 3 + 1] = c[i].y; centroid3[i * 3 + 2] = c[i].z; }
  }
  return LVS_OK;
}

int lvs_ndt_get_cell_horizontal(lvs_ndt_t* h, int32_t* horizontal) {
  if (!h || !horizontal) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (!rc) rc = finish_target(b, 0);
  if (rc) return rc;
  const TargetGrid& t = b->targets[0];
  const int n = t.n_cells;
  if (n == 0) return LVS_OK;
  CUDA_TRY(cudaStreamSynchronize(b->st));
  std::vector<VoxelRec> recs(n);
  CUDA_TRY(cudaMemcpy(recs.data(), t.d_recs, (size_t)n * sizeof(VoxelRec), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) horizontal[i] = (recs[i].meta & kMetaHorizBit) ? 1 : 0;
  return LVS_OK;
}

int lvs_ndt_lookup_keys(lvs_ndt_t* h, const float T16[16], int32_t* keys_out) {
  if (!h || !T16 || !keys_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  lvs_ndt_batch* b = h->b;
  int rc = set_device(b);
  if (rc) return rc;
  if (!b->target_pts[0].set) return fail(LVS_ERR_NO_TARGET, "setInputTarget not called");
  if (!b->sources[0].set) return fail(LVS_ERR_NO_SOURCE, "setInputSource not called");
  const int n = b->sources[0].n;
  if (n == 0) return LVS_OK;
  if ((rc = wait_all_uploads(b))) return rc;
  if ((rc = finish_target(b, 0))) return rc;
  PairDesc P = make_pair(b, 0, 0);
  int* d_keys = nullptr;
  CUDA_TRY(cudaMalloc(&d_keys, (size_t)n * 4));
  cudaError_t e = cudaMemcpyAsync(b->d_T16, T16, 16 * sizeof(float), cudaMemcpyHostToDevice, b->st);
  if (e == cudaSuccess) { rc = launch_lookup_keys(b->st, P, b->d_T16, d_keys); b->total_launches++; }
  if (e == cudaSuccess && rc == LVS_OK) e = cudaMemcpyAsync(keys_out, d_keys, (size_t)n * 4, cudaMemcpyDeviceToHost, b->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->st);
  cudaFree(d_keys);
  if (e != cudaSuccess) return cuda_fail(e, "lookup_keys", __FILE__, __LINE__);
  return rc;
}

// Diagnostics, not part of the public header: the 8 tail stamps of pair 0's last evaluation (LVS_DEBUG_TIMING=1), in SM clocks.
__attribute__((visibility("default"))) int lvs_ndt_batch_debug_stamps(lvs_ndt_batch_t* b, long long out8[8]) {
  if (!b || !out8 || !b->d_dbg) return fail(LVS_ERR_INVALID_ARG, "debug timing is off");
  cudaSetDevice(b->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 8; i++) out8[i] = b->d_dbg[i];
  return LVS_OK;
}

int lvs_ndt_handle_batch(lvs_ndt_t* h, lvs_ndt_batch_t** out) {
  if (!h || !out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *out = h->b;
  return LVS_OK;
}

}  // extern "C"
