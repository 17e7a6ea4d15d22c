// Supernodal multifrontal block Cholesky for the pose-graph normal equations (see pgo_chol.cuh for what it replaces).
#include <algorithm>
#include <cstring>
#include <set>
#include "ndt_internal.cuh"
#include "pgo_chol.cuh"

namespace lvs {

// =====================================================================================================================
// Host: symbolic analysis
// =====================================================================================================================

// Minimum-degree ordering on the quotient graph.  A variable keeps its not-yet-absorbed variable neighbours and the elements
// (eliminated pivots) it touches; eliminating p forms the element L_p = reach(p), which is exactly the below-diagonal pattern
// of p's column in L.  Degrees are the approximate external degrees of AMD, recomputed for the members of the new element only
// (exact degrees cost a walk over every element list of every member; with them and one ordered set as the queue the ordering
// took 2.4 s on the 50 000-vertex sphere, now 0.27 s, with identical fill on the sphere graphs).
static void minimum_degree(int n, const std::vector<std::vector<int>>& adj, std::vector<int>& order, std::vector<std::vector<int>>& pattern) {
  std::vector<std::vector<int>> av(adj), ae(n), el(n);
  std::vector<char> gone(n, 0), dead(n, 0);
  std::vector<int> mark(n, -1), deg(n), w(n, 0), wmark(n, -1);
  // One min-heap of variable indices per degree, with lazy deletion: an entry of bucket d is live while the variable is still
  // there and its degree is still d.  The pivot is the smallest index of the lowest non-empty degree.
  std::vector<std::vector<int>> bucket(n + 1);
  int mindeg = n;
  auto push = [&](int d, int i) {
    std::vector<int>& b = bucket[d];
    b.push_back(i);
    std::push_heap(b.begin(), b.end(), std::greater<int>());
    mindeg = std::min(mindeg, d);
  };
  for (int i = 0; i < n; i++) {
    std::sort(av[i].begin(), av[i].end());
    av[i].erase(std::unique(av[i].begin(), av[i].end()), av[i].end());
    deg[i] = (int)av[i].size();
    push(deg[i], i);
  }
  order.clear(); order.reserve(n);
  pattern.assign(n, {});
  int tag = 0, wtag = 0;
  std::vector<int> Lp;
  for (int step = 0; step < n; step++) {
    int p = -1;
    while (p < 0) {
      std::vector<int>& b = bucket[mindeg];
      if (b.empty()) { mindeg++; continue; }
      const int c = b.front();
      std::pop_heap(b.begin(), b.end(), std::greater<int>());
      b.pop_back();
      if (!gone[c] && deg[c] == mindeg) p = c;
    }
    gone[p] = 1;
    order.push_back(p);
    // L_p = (A_p u U_{e in E_p} L_e) \ {p}
    Lp.clear();
    ++tag;
    mark[p] = tag;
    for (int v : av[p]) if (!gone[v] && mark[v] != tag) { mark[v] = tag; Lp.push_back(v); }
    for (int e : ae[p]) {
      if (dead[e]) continue;
      for (int v : el[e]) if (!gone[v] && mark[v] != tag) { mark[v] = tag; Lp.push_back(v); }
      dead[e] = 1;                          // absorbed into the new element
      std::vector<int>().swap(el[e]);
    }
    std::vector<int>().swap(av[p]);
    std::vector<int>().swap(ae[p]);
    el[p] = Lp;
    pattern[p] = Lp;
    for (int i : Lp) {
      // variable neighbours now covered by the new element are dropped, absorbed elements too
      std::vector<int>& a = av[i];
      size_t k = 0;
      for (int v : a) if (!gone[v] && mark[v] != tag) a[k++] = v;
      a.resize(k);
      std::vector<int>& e = ae[i];
      k = 0;
      for (int x : e) if (!dead[x]) e[k++] = x;
      e.resize(k);
      e.push_back(p);
    }
    // approximate external degree (the bound of the AMD algorithm): |A_i| + |L_p minus i| + sum over the other elements e of i of
    // |L_e minus L_p|.  One sweep over the element lists of the members gives every |L_e minus L_p| (w[e] starts at |L_e| and loses one per
    // member of L_p that e contains); overlaps between different elements are counted twice, which is what makes it a bound.
    ++wtag;
    for (int i : Lp)
      for (int x : ae[i]) {
        if (x == p) continue;
        if (wmark[x] != wtag) { wmark[x] = wtag; w[x] = (int)el[x].size(); }
        w[x]--;
      }
    const int remaining = n - step - 1;
    for (int i : Lp) {
      long long d = (long long)av[i].size() + (long long)Lp.size() - 1;
      // an element with nothing outside L_p is covered by the new one: absorbed (dropped from the lists at the next visit)
      for (int x : ae[i]) if (x != p) { if (w[x] == 0 && !dead[x]) { dead[x] = 1; std::vector<int>().swap(el[x]); } d += w[x]; }
      d = std::min<long long>(d, remaining - 1);
      d = std::min<long long>(d, (long long)deg[i] + (long long)Lp.size() - 1);
      const int nd = (int)std::max<long long>(d, 0);
      if (nd != deg[i]) { deg[i] = nd; push(nd, i); }
    }
  }
}

void chol_analyze(int n, int n_off, const int* off_ij, CholSymbolic& S) {
  S = CholSymbolic();
  S.n = n;
  if (n == 0) return;
  std::vector<std::vector<int>> adj(n);
  for (int o = 0; o < n_off; o++) {
    const int r = off_ij[2 * o], c = off_ij[2 * o + 1];
    adj[r].push_back(c); adj[c].push_back(r);
  }
  std::vector<int> order;
  std::vector<std::vector<int>> pat;
  minimum_degree(n, adj, order, pat);
  // elimination tree in elimination positions, then a postorder so that every subtree is contiguous
  std::vector<int> pos(n);
  for (int k = 0; k < n; k++) pos[order[k]] = k;
  std::vector<int> parent(n, -1);
  std::vector<std::vector<int>> kids(n);
  for (int k = 0; k < n; k++) {
    int best = -1;
    for (int v : pat[order[k]]) if (best < 0 || pos[v] < best) best = pos[v];
    parent[k] = best;
    if (best >= 0) kids[best].push_back(k);
  }
  std::vector<int> post(n), label(n);     // post[new] = elimination position
  {
    int cnt = 0;
    std::vector<std::pair<int, size_t>> stack;
    for (int root = 0; root < n; root++) {
      if (parent[root] >= 0) continue;
      stack.push_back({root, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second < kids[top.first].size()) { const int c = kids[top.first][top.second++]; stack.push_back({c, 0}); }
        else { label[top.first] = cnt; post[cnt++] = top.first; stack.pop_back(); }
      }
    }
  }
  S.perm.resize(n); S.iperm.resize(n);
  for (int c = 0; c < n; c++) { S.perm[c] = order[post[c]]; S.iperm[S.perm[c]] = c; }
  // column patterns in the final numbering
  std::vector<std::vector<int>> col(n);
  std::vector<int> par(n, -1);
  for (int c = 0; c < n; c++) {
    const std::vector<int>& p = pat[S.perm[c]];
    col[c].resize(p.size());
    for (size_t k = 0; k < p.size(); k++) col[c][k] = S.iperm[p[k]];
    std::sort(col[c].begin(), col[c].end());
    if (!col[c].empty()) par[c] = col[c][0];
  }
  // fundamental supernodes: column c joins c-1 when pattern(c-1) = {c} u pattern(c)
  S.col_front.assign(n, -1);
  for (int c = 0; c < n; c++) {
    const bool join = c > 0 && par[c - 1] == c && col[c - 1].size() == col[c].size() + 1;
    if (join) { S.fronts.back().w++; }
    else { CholFront f; memset(&f, 0, sizeof f); f.c0 = c; f.w = 1; f.parent = -1; S.fronts.push_back(f); }
    S.col_front[c] = (int)S.fronts.size() - 1;
  }
  const int nf = (int)S.fronts.size();
  std::vector<std::vector<int>> fkids(nf);
  for (int s = 0; s < nf; s++) {
    CholFront& f = S.fronts[s];
    const std::vector<int>& R = col[f.c0 + f.w - 1];
    f.r = (int)R.size();
    f.F = 6 * (f.w + f.r) + 1;
    f.rows_off = (int)S.rows.size();
    S.rows.insert(S.rows.end(), R.begin(), R.end());
    f.parent = f.r ? S.col_front[R[0]] : -1;
    if (f.parent >= 0) fkids[f.parent].push_back(s);
    f.off = S.arena;
    S.arena += (long long)f.F * f.F;
    S.max_front = std::max(S.max_front, f.F);
    S.nnz_l_blocks += (long long)f.w * (f.w + 1) / 2 + (long long)f.w * f.r;
    for (int k = 0; k < f.w; k++) { const double m = 6.0 * (f.w - k - 1 + f.r) + 1; S.flops += 3.0 * m * m; }   // 6 pivot columns x m^2 / 2
  }
  // relative indices: position of every row of R_S inside the parent's index set (its pivots, then its R)
  S.rel.assign(S.rows.size(), -1);
  for (int s = 0; s < nf; s++) {
    const CholFront& f = S.fronts[s];
    if (f.parent < 0) continue;
    const CholFront& P = S.fronts[f.parent];
    int q = 0;
    for (int k = 0; k < f.r; k++) {
      const int row = S.rows[f.rows_off + k];
      if (row < P.c0 + P.w) { S.rel[f.rows_off + k] = row - P.c0; continue; }
      while (q < P.r && S.rows[P.rows_off + q] < row) q++;
      S.rel[f.rows_off + k] = P.w + q;      // containment: pattern(child) \ pivots(parent) is a subset of pattern(parent)
    }
  }
  // children lists, levels
  int max_level = 0;
  for (int s = 0; s < nf; s++) {
    CholFront& f = S.fronts[s];
    f.child_begin = (int)S.child_idx.size();
    S.child_idx.insert(S.child_idx.end(), fkids[s].begin(), fkids[s].end());
    f.child_end = (int)S.child_idx.size();
    int lv = 0;
    for (int c : fkids[s]) lv = std::max(lv, S.fronts[c].level + 1);   // children precede parents (postorder)
    f.level = lv;
    max_level = std::max(max_level, lv);
  }
  S.level_ptr.assign(max_level + 2, 0);
  for (int s = 0; s < nf; s++) S.level_ptr[S.fronts[s].level + 1]++;
  for (int l = 0; l <= max_level; l++) S.level_ptr[l + 1] += S.level_ptr[l];
  S.level_fronts.resize(nf);
  {
    std::vector<int> fill(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int s = 0; s < nf; s++) S.level_fronts[fill[S.fronts[s].level]++] = s;
  }
  // scatter maps
  S.diag_dst.resize(n); S.diag_ld.resize(n); S.rhs_dst.resize(n);
  for (int v = 0; v < n; v++) {
    const int c = S.iperm[v];
    const CholFront& f = S.fronts[S.col_front[c]];
    const int k = c - f.c0;
    S.diag_dst[v] = f.off + (long long)(6 * k) * f.F + 6 * k;
    S.diag_ld[v] = f.F;
    S.rhs_dst[v] = f.off + (long long)(6 * k) * f.F + (f.F - 1);
  }
  S.off_dst.resize(n_off); S.off_ld.resize(n_off); S.off_tr.resize(n_off);
  for (int o = 0; o < n_off; o++) {
    const int pr = S.iperm[off_ij[2 * o]], pc = S.iperm[off_ij[2 * o + 1]];
    const int lo = std::min(pr, pc), hi = std::max(pr, pc);
    const CholFront& f = S.fronts[S.col_front[lo]];
    int rp;
    if (hi < f.c0 + f.w) rp = hi - f.c0;
    else {
      const int* b = S.rows.data() + f.rows_off;
      rp = f.w + (int)(std::lower_bound(b, b + f.r, hi) - b);
    }
    S.off_dst[o] = f.off + (long long)(6 * (lo - f.c0)) * f.F + 6 * rp;
    S.off_ld[o] = f.F;
    S.off_tr[o] = pr < pc;       // the stored block is H[row][col]; the lower triangle wants H[hi][lo]
  }
}

// =====================================================================================================================
// Device
// =====================================================================================================================
constexpr int kCholThreads = 256;
constexpr int kCholMaxTeam = 1024;      // CTAs that may share one front (bounded by the cooperative grid)
constexpr int kCholBigFront = 192;    // fronts with F above this go to the team kernel
constexpr int kNB = 24;                 // pivot columns per panel of the backward substitution and of the single-CTA front kernel
constexpr int kNBTeam = 24;             // pivot columns per panel of the team front kernel (48 was measured: 6.2 -> 7.1 ms per solve,
                                        // the diagonal block and the row solve grow faster than the barriers shrink)
constexpr int kCholSmemFront = 64;      // fronts up to this many rows are factored in shared memory (single-CTA kernel)
constexpr int kTile = 96;               // trailing-update tile (16 x 16 threads, 6 x 6 outputs each)
template <bool TEAM>
struct FrontSmem {                      // dynamic shared memory of chol_front_kernel<TEAM>, in doubles
  static constexpr int NB = TEAM ? kNBTeam : kNB;
  static constexpr size_t doubles = 2 * NB * (NB + 1) + 2 * NB * (kTile + 2) + NB + (TEAM ? 0 : kCholSmemFront * kCholSmemFront);
  static constexpr size_t bytes = doubles * sizeof(double);
};

struct CholView {
  const CholFront* fronts;
  const int *rows, *rel, *child_idx, *level_fronts, *perm;
  const long long *diag_dst, *off_dst, *rhs_dst;
  const int *diag_ld, *off_ld;
  const unsigned char* off_tr;
  double* arena;
  double* xp;
  int* fail_flag;
  int n, n_off;
};

__global__ void __launch_bounds__(kCholThreads) chol_scatter_kernel(CholView V, const double* __restrict__ Hd, const double* __restrict__ Ho,
                                                                    const double* __restrict__ b, double lambda) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nd = (long long)V.n * 36, no = (long long)V.n_off * 36;
  if (t < nd) {
    const int v = (int)(t / 36), e = (int)(t % 36), a = e / 6, c = e % 6;
    V.arena[V.diag_dst[v] + (long long)c * V.diag_ld[v] + a] = Hd[t] + (a == c ? lambda : 0.0);
  } else if (t < nd + no) {
    const long long u = t - nd;
    const int o = (int)(u / 36), e = (int)(u % 36), a = e / 6, c = e % 6;     // target element (row part a, column part c)
    V.arena[V.off_dst[o] + (long long)c * V.off_ld[o] + a] = V.off_tr[o] ? Ho[(size_t)o * 36 + c * 6 + a] : Ho[(size_t)o * 36 + a * 6 + c];
  } else if (t < nd + no + (long long)V.n * 6) {
    const long long u = t - nd - no;
    const int v = (int)(u / 6), a = (int)(u % 6);
    V.arena[V.rhs_dst[v] + (long long)a * V.diag_ld[v]] = b[u];
  }
}

// Fronts of one level.  A front is worked on by a TEAM of CTAs (team_size 1 for the many small fronts of the lower levels, tens
// of CTAs for the few large fronts near the root); teams take the fronts of the list round-robin.  Inside a team the phases are
// separated by a team barrier: __syncthreads for a single CTA, otherwise an arrive/spin counter in global memory (all CTAs are
// co-resident: cooperative launch).  In team mode every read of the front goes to L2 (__ldcg): L1 is not coherent across SMs.
template <bool TEAM>
__device__ __forceinline__ double ldf(const double* p) { return TEAM ? __ldcg(p) : *p; }

template <bool TEAM>
__device__ __forceinline__ void team_sync(unsigned int* bar, unsigned int& target, int team_size) {
  if (!TEAM) { __syncthreads(); return; }
  __syncthreads();
  if (threadIdx.x == 0) {
    // arrive with release semantics (this CTA's writes, ordered before by the barrier above, become visible with the count) and
    // without waiting for the atomic's return value; poll with acquire loads
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    target += (unsigned)team_size;
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

template <bool TEAM>
__global__ void __launch_bounds__(kCholThreads, TEAM ? 1 : 2) chol_front_kernel(CholView V, const int* __restrict__ list, int n_list, int team_size,
                                                                  unsigned int* __restrict__ bars) {
  const int team = blockIdx.x / team_size, rank = blockIdx.x % team_size, n_teams = gridDim.x / team_size;
  const int tid = rank * kCholThreads + threadIdx.x, nthr = team_size * kCholThreads;
  unsigned int* bar = bars + team;
  unsigned int target = 0;
  // Panel width per kernel variant; all staging lives in dynamic shared memory (FrontSmem<TEAM>).
  constexpr int kNB = FrontSmem<TEAM>::NB;
  extern __shared__ double s_dyn[];
  double (*s_D)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(s_dyn);
  double (*s_S)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(s_dyn + kNB * (kNB + 1));
  double (*s_Li)[kTile + 2] = reinterpret_cast<double (*)[kTile + 2]>(s_dyn + 2 * kNB * (kNB + 1));
  double (*s_Lj)[kTile + 2] = reinterpret_cast<double (*)[kTile + 2]>(s_dyn + 2 * kNB * (kNB + 1) + kNB * (kTile + 2));
  double* s_inv = s_dyn + 2 * kNB * (kNB + 1) + 2 * kNB * (kTile + 2);
  double* s_front = s_inv + kNB;                  // single-CTA launches only: room for a front of kCholSmemFront rows
  for (int fi = team; fi < n_list; fi += n_teams) {
    const CholFront f = V.fronts[list[fi]];
    double* const A_global = V.arena + f.off;
    double* A = A_global;
    const int F = f.F;                            // the right-hand side is row F - 1
    // A small front is worked on in shared memory: one read and one write of it instead of a global-memory round trip in every
    // phase (the many small fronts of the lower levels are pure latency).
    const bool in_smem = !TEAM && F <= kCholSmemFront;
    if (in_smem) {
      for (int t = threadIdx.x; t < F * F; t += kCholThreads) s_front[t] = A_global[t];
      __syncthreads();
      A = s_front;
    }
    // ---- extend-add: U_c (child's trailing block, rows R_c + rhs row) into this front through the relative indices
    for (int ci = f.child_begin; ci < f.child_end; ci++) {
      const CholFront c = V.fronts[V.child_idx[ci]];
      const double* __restrict__ U = V.arena + c.off;
      const int* __restrict__ rel = V.rel + c.rows_off;
      const int rc = c.r, Fc = c.F;
      // tiles (bi >= bj), bi in [0, rc] (rc = rhs row), bj in [0, rc): enumerate the full rectangle and skip the upper part
      const long long total = (long long)(rc + 1) * rc * 36;
      // four elements per thread and step, all loads before the stores: the front is latency-bound otherwise (the compiler must
      // keep a load behind the previous store into the same array)
      constexpr int kEa = 4;
      for (long long t0 = tid; t0 < total; t0 += (long long)nthr * kEa) {
        double* dst[kEa];
        double val[kEa];
#pragma unroll
        for (int u = 0; u < kEa; u++) {
          const long long t = min(t0 + (long long)u * nthr, total - 1);
          const int e = (int)(t % 36), tile = (int)(t / 36);
          const int bi = tile % (rc + 1), bj = tile / (rc + 1);
          const int a = e % 6, b = e / 6;            // a fastest: consecutive threads read consecutive child rows
          const bool live = t0 + (long long)u * nthr < total && bi >= bj && !(bi == rc && a > 0);
          const int src_row = (bi == rc) ? Fc - 1 : 6 * (c.w + bi) + a;
          const int dst_row = (bi == rc) ? F - 1 : 6 * rel[bi] + a;
          double* d = A + (size_t)(6 * rel[bj] + b) * F + dst_row;
          // the loads are unconditional (every address is inside the two fronts) so that all of them are in flight together
          val[u] = ldf<TEAM>(d) + __ldcg(U + (size_t)(6 * (c.w + bj) + b) * Fc + src_row);
          dst[u] = live ? d : nullptr;
        }
#pragma unroll
        for (int u = 0; u < kEa; u++) if (dst[u]) *dst[u] = val[u];
      }
      team_sync<TEAM>(bar, target, team_size);     // children are added one after the other: fixed summation order
    }
    // ---- partial Cholesky of the p = 6 w pivot columns, right-looking in panels of kNB columns.  The right-hand side is simply
    // the last row (F - 1) of the front.  Per panel: (A) the diagonal block is factored in shared memory by warp 0 of EVERY CTA of
    // the team (same arithmetic, same result: no barrier needed before B), (B) the rows below are solved against it, one thread per
    // row, (C) the trailing matrix gets the rank-nb update in 96 x 96 tiles staged through shared memory.
    const int p = 6 * f.w;
    for (int c = 0; c < p; c += kNB) {
      const int nb = min(kNB, p - c);
      // rows of (B) are fetched first so that their latency hides behind (A)
      int i = c + nb + tid;
      double x[kNB];
      if (i < F) {
#pragma unroll
        for (int j = 0; j < kNB; j++) x[j] = ldf<TEAM>(A + (size_t)(c + min(j, nb - 1)) * F + i);
      }
      // (A) diagonal block, all threads, one barrier per column: s_S holds the running Schur complement, column j of the factor is
      // written to s_D while the columns right of j take its rank-1 update (reads column j of s_S, writes columns > j: no conflict)
      for (int t = threadIdx.x; t < kNB * kNB; t += kCholThreads) {
        const int j = t / kNB, i = t % kNB;
        const double v = ldf<TEAM>(A + (size_t)(c + min(j, nb - 1)) * F + c + min(i, nb - 1));
        s_S[i][j] = (i < nb && j <= i) ? v : (i == j ? 1.0 : 0.0);
        s_D[i][j] = 0.0;
      }
      __syncthreads();
      {
        // the lower-triangle elements this thread owns, fixed for the whole panel
        constexpr int kOwn = (kNB * kNB + kCholThreads - 1) / kCholThreads;
        int ei[kOwn], ek[kOwn];
#pragma unroll
        for (int u = 0; u < kOwn; u++) {
          const int t = threadIdx.x + u * kCholThreads;
          ei[u] = t / kNB; ek[u] = t % kNB;
          if (!(t < kNB * kNB && ei[u] < nb && ek[u] <= ei[u])) { ei[u] = -1; ek[u] = kNB; }     // inactive: i = -1 fails every test below
        }
        for (int j = 0; j < nb; j++) {
          bool busy = false;
#pragma unroll
          for (int u = 0; u < kOwn; u++) busy |= ei[u] >= j;
          if (busy) {                                                                                 // rows above j are finished
            double d = s_S[j][j];
            if (!(d > 0.0)) { *V.fail_flag = 1; d = 1.0; }     // every thread that sees it stores the same 1
            const double il = rsqrt(d);                // one reciprocal square root instead of sqrt + divide on the serial path
#pragma unroll
            for (int u = 0; u < kOwn; u++) {
              const int i = ei[u], k = ek[u];
              if (i < j) continue;
              if (k == j) { s_D[i][j] = (i == j) ? d * il : s_S[i][j] * il; if (i == j) s_inv[j] = il; }
              else if (k > j) s_S[i][k] -= (s_S[i][j] * il) * (s_S[k][j] * il);
            }
          }
          __syncthreads();
        }
      }
      // (B) rows below the diagonal block: x L_D^T = a, forward substitution along the row, right-looking: once x_m is final every
      // later entry takes its term.  Each x_j still receives its terms in the order m = 0, 1, ... (the result does not change),
      // but consecutive instructions are independent instead of one 276-long chain of dependent multiply-adds.
      {
        for (; i < F; i += nthr) {
#pragma unroll
          for (int m = 0; m < kNB; m++) {
            if (m < nb) {
              const double xm = x[m] * s_inv[m];
              x[m] = xm;
              A[(size_t)(c + m) * F + i] = xm;
#pragma unroll
              for (int j = m + 1; j < kNB; j++) x[j] -= xm * s_D[j][m];      // rows of s_D past nb are zero
            }
          }
          if (i + nthr < F) {
#pragma unroll
            for (int j = 0; j < kNB; j++) x[j] = ldf<TEAM>(A + (size_t)(c + min(j, nb - 1)) * F + i + nthr);
          }
        }
      }
      team_sync<TEAM>(bar, target, team_size);
      // the factored diagonal block goes back to the front only now: before the barrier other CTAs may still be reading the
      // unfactored block for their own copy of (A); nothing in (C) touches it
      if (rank == 0)
        for (int t = threadIdx.x; t < nb * nb; t += kCholThreads) {
          const int j = t / nb, i = t % nb;
          if (i >= j) A[(size_t)(c + j) * F + c + i] = s_D[i][j];
        }
      // (C) A[i, j] -= sum_q L[i, c+q] L[j, c+q] for c + nb <= j <= i (lower triangle, plus the upper part of diagonal 6x6 blocks)
      const int base = c + nb, rem = F - base;
      const int nt = (rem + kTile - 1) / kTile;
      const int ntiles = nt * (nt + 1) / 2;
      const int ty = threadIdx.x % 16, tx = threadIdx.x / 16;
      for (int tile = rank; tile < ntiles; tile += team_size) {
        int ti = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
        while (ti * (ti + 1) / 2 > tile) ti--;
        while ((ti + 1) * (ti + 2) / 2 <= tile) ti++;
        const int tj = tile - ti * (ti + 1) / 2;
        const int i0 = base + kTile * ti, j0 = base + kTile * tj;
        {
          constexpr int kLd = kNB * kTile / kCholThreads;      // 9 elements of each panel per thread
          double vi[kLd], vj[kLd];
#pragma unroll
          for (int u = 0; u < kLd; u++) {                      // unconditional loads from clamped addresses, all in flight together
            const int t = threadIdx.x + u * kCholThreads, q = min(t / kTile, nb - 1), r = t % kTile;
            vi[u] = ldf<TEAM>(A + (size_t)(c + q) * F + min(i0 + r, F - 1));
            vj[u] = ldf<TEAM>(A + (size_t)(c + q) * F + min(j0 + r, F - 1));
          }
#pragma unroll
          for (int u = 0; u < kLd; u++) {
            const int t = threadIdx.x + u * kCholThreads, q = t / kTile, r = t % kTile;
            const int rp = r + (r >= 48 ? 2 : 0);
            s_Li[q][rp] = (q < nb && i0 + r < F) ? vi[u] : 0.0;
            s_Lj[q][rp] = (q < nb && j0 + r < F) ? vj[u] : 0.0;
          }
        }
        __syncthreads();
        double acc[6][6];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int b2 = 0; b2 < 6; b2++) acc[a][b2] = 0.0;
        const int ro = 6 * ty + (ty >= 8 ? 2 : 0), co = 6 * tx + (tx >= 8 ? 2 : 0);
        for (int q = 0; q < nb; q++) {
          double li[6], lj[6];
#pragma unroll
          for (int a = 0; a < 6; a++) { li[a] = s_Li[q][ro + a]; lj[a] = s_Lj[q][co + a]; }
#pragma unroll
          for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b2 = 0; b2 < 6; b2++) acc[a][b2] += li[a] * lj[b2];
        }
        const int ib = i0 + 6 * ty, jb = j0 + 6 * tx;
        if (ib >= jb) {                       // 6x6 sub-tiles are aligned to the 6x6 blocks of the front: keep lower and diagonal ones
          // all 36 loads first, then the stores (a load cannot be moved above an earlier store into the same array)
#pragma unroll
          for (int b2 = 0; b2 < 6; b2++)
#pragma unroll
            for (int a = 0; a < 6; a++) acc[a][b2] = ldf<TEAM>(A + (size_t)min(jb + b2, F - 2) * F + min(ib + a, F - 1)) - acc[a][b2];
#pragma unroll
          for (int b2 = 0; b2 < 6; b2++)
#pragma unroll
            for (int a = 0; a < 6; a++)
              if (jb + b2 < F - 1 && ib + a < F) A[(size_t)(jb + b2) * F + ib + a] = acc[a][b2];
        }
        __syncthreads();
      }
      team_sync<TEAM>(bar, target, team_size);
    }
    if (in_smem) {
      for (int t = threadIdx.x; t < F * F; t += kCholThreads) A_global[t] = s_front[t];
      __syncthreads();                             // the next front of this CTA reuses the buffer
    }
  }
}

// Backward substitution, one CTA per front of the level (levels top-down): x_piv = L_D^-T (y - B^T x_R), in panels of kNB pivot
// columns from the last one up.  First the ancestors' x_R is folded into the right-hand side (a warp streams a pivot column,
// contiguous in memory, with eight loads in flight); then, per panel, warp 0 solves the 24 x 24 triangle (lane m keeps the
// running sum of its own row) and every earlier pivot takes the panel's contribution (right-looking: a thread streams the 24
// entries of its own column).
__global__ void __launch_bounds__(kCholThreads) chol_backward_kernel(CholView V, int level_begin) {
  extern __shared__ double s_dyn[];                 // xs[F - 1]: y, overwritten by x panel by panel, followed by x_R
  __shared__ double s_D[kNB][kNB + 1], s_dot[kNB];
  const int s = V.level_fronts[level_begin + blockIdx.x];
  const CholFront f = V.fronts[s];
  const double* __restrict__ A = V.arena + f.off;
  const int F = f.F, p = 6 * f.w, nr = 6 * f.r;
  double* xs = s_dyn;
  for (int j = threadIdx.x; j < p; j += kCholThreads) xs[j] = A[(size_t)j * F + (F - 1)];
  for (int i = threadIdx.x; i < nr; i += kCholThreads) xs[p + i] = V.xp[(size_t)V.rows[f.rows_off + i / 6] * 6 + i % 6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kW = kCholThreads / 32;
  __syncthreads();
  // xs[0:p] -= B^T x_R : one warp per pivot column (contiguous in memory), eight loads in flight, fixed shuffle tree
  for (int j = warp; j < p; j += kW) {
    const double* __restrict__ col = A + (size_t)j * F + p;
    double acc = 0.0;
    for (int i0 = lane; i0 < nr; i0 += 32 * 8) {
      // unconditional loads from clamped addresses: with predicated loads the compiler sinks each one next to its use and
      // only one is in flight at a time
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = col[min(i0 + 32 * u, nr - 1)];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int i = i0 + 32 * u; acc += (i < nr) ? v[u] * xs[p + i] : 0.0; }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) xs[j] -= acc;
  }
  const int n_panels = (p + kNB - 1) / kNB;
  for (int pi = n_panels - 1; pi >= 0; pi--) {
    const int c = pi * kNB, nb = min(kNB, p - c);
    __syncthreads();                                // every update of this panel's right-hand side has landed
    for (int t = threadIdx.x; t < kNB * kNB; t += kCholThreads) {
      const int j = t / kNB, i = t % kNB;
      const double v = A[(size_t)(c + min(j, nb - 1)) * F + c + min(i, nb - 1)];
      s_D[i][j] = (i < nb && j <= i) ? v : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // L_D^T x = rhs: x_j = (rhs_j - sum_{i > j} L[i][j] x_i) / L[j][j]; lane m accumulates its own row's sum as the x_i appear
      const int m = threadIdx.x;
      const bool on = m < nb;
      const double rhs = on ? xs[c + m] : 0.0;
      const double inv = on ? 1.0 / s_D[m][m] : 0.0;
      double acc = 0.0, mine = 0.0;
      for (int j = nb - 1; j >= 0; j--) {
        const double xj = __shfl_sync(0xffffffffu, (rhs - acc) * inv, j);
        if (m == j) mine = xj;
        if (m < j) acc += s_D[j][m] * xj;
      }
      if (m < kNB) s_dot[m] = on ? mine : 0.0;       // the panel's solution (zero padded)
      if (on) xs[c + m] = mine;
    }
    __syncthreads();
    // right-looking update of every earlier pivot: xs[i] -= sum_j L[c + j][i] x_j.  Column i of the front is contiguous in j, so a
    // thread streams its 24 entries with all loads in flight.
    for (int i = threadIdx.x; i < c; i += kCholThreads) {
      const double* __restrict__ col = A + (size_t)i * F + c;
      double v[kNB];
#pragma unroll
      for (int j = 0; j < kNB; j++) v[j] = col[min(j, nb - 1)];
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < kNB; j++) acc += v[j] * s_dot[j];
      xs[i] -= acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < p; j += kCholThreads) V.xp[(size_t)f.c0 * 6 + j] = xs[j];
}

// x (original numbering) from the permuted solution, and LM's gain-ratio denominator sum_j x_j (lambda x_j + b_j), one CTA.
__global__ void __launch_bounds__(1024) chol_finish_kernel(CholView V, const double* __restrict__ b, double lambda, double* __restrict__ x,
                                                           double* __restrict__ scale_out, int* __restrict__ ok_out) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (int t = threadIdx.x; t < V.n * 6; t += 1024) {
    const int c = t / 6, a = t % 6;
    const int v = V.perm[c];
    const double xv = V.xp[t];
    x[(size_t)v * 6 + a] = xv;
    acc += xv * (lambda * xv + b[(size_t)v * 6 + a]);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 32; w++) t += s_red[w];
    if (scale_out) *scale_out = t;
    if (ok_out) *ok_out = (*V.fail_flag == 0 && t == t) ? 1 : 0;
  }
}

template <typename T>
static int up(CholDevice& C, T** dst, const std::vector<T>& v, cudaStream_t st) {
  *dst = nullptr;
  CUDA_TRY(cudaMalloc((void**)dst, std::max<size_t>(1, v.size()) * sizeof(T)));
  C.allocs.push_back((void*)*dst);
  if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return LVS_OK;
}

int chol_upload(const CholSymbolic& S, int n_off, CholDevice& C, cudaStream_t st) {
  chol_free(C);
  C.n = S.n; C.n_off = n_off; C.n_fronts = (int)S.fronts.size(); C.n_levels = (int)S.level_ptr.size() - 1; C.max_front = S.max_front;
  C.level_ptr = S.level_ptr;
  C.level_big.assign(std::max(C.n_levels, 0), 0);
  for (const CholFront& f : S.fronts) C.level_big[f.level] = std::max(C.level_big[f.level], f.F);
  int rc;
  if ((rc = up(C, &C.fronts, S.fronts, st)) || (rc = up(C, &C.rows, S.rows, st)) || (rc = up(C, &C.rel, S.rel, st)) ||
      (rc = up(C, &C.child_idx, S.child_idx, st)) || (rc = up(C, &C.level_fronts, S.level_fronts, st)) || (rc = up(C, &C.perm, S.perm, st)) ||
      (rc = up(C, &C.col_front, S.col_front, st)) || (rc = up(C, &C.diag_dst, S.diag_dst, st)) || (rc = up(C, &C.off_dst, S.off_dst, st)) ||
      (rc = up(C, &C.rhs_dst, S.rhs_dst, st)) || (rc = up(C, &C.diag_ld, S.diag_ld, st)) || (rc = up(C, &C.off_ld, S.off_ld, st)) ||
      (rc = up(C, &C.off_tr, S.off_tr, st)))
    return rc;
  {
    std::vector<int> small_list, big_list;
    C.small_ptr.assign(C.n_levels + 1, 0); C.big_ptr.assign(C.n_levels + 1, 0);
    for (int l = 0; l < C.n_levels; l++) {
      for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; k++) {
        const int s = S.level_fronts[k];
        (S.fronts[s].F > kCholBigFront ? big_list : small_list).push_back(s);
      }
      C.small_ptr[l + 1] = (int)small_list.size(); C.big_ptr[l + 1] = (int)big_list.size();
    }
    if ((rc = up(C, &C.small_list, small_list, st)) || (rc = up(C, &C.big_list, big_list, st))) return rc;
    int per_sm = 0, dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaFuncSetAttribute(chol_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrontSmem<true>::bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_front_kernel<true>, kCholThreads, FrontSmem<true>::bytes));
    C.coop_grid = std::max(1, std::min(per_sm, 2) * sms);
    CUDA_TRY(cudaMalloc((void**)&C.bars, (size_t)C.coop_grid * sizeof(unsigned int)));
    C.allocs.push_back((void*)C.bars);
  }
  C.arena_doubles = S.arena;
  cudaError_t e = cudaMalloc((void**)&C.arena, std::max<long long>(1, S.arena) * sizeof(double));
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(LVS_ERR_OOM, "frontal arena allocation failed (%lld MB)", (long long)(S.arena * 8 >> 20)); }
  C.allocs.push_back((void*)C.arena);
  CUDA_TRY(cudaMalloc((void**)&C.xp, std::max<size_t>(1, (size_t)S.n * 6) * sizeof(double)));
  C.allocs.push_back((void*)C.xp);
  CUDA_TRY(cudaMalloc((void**)&C.fail_flag, sizeof(int)));
  C.allocs.push_back((void*)C.fail_flag);
  const size_t smem = (size_t)S.max_front * sizeof(double);
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(chol_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(chol_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrontSmem<false>::bytes));
  CUDA_TRY(cudaStreamSynchronize(st));    // the host vectors of S may go away
  return LVS_OK;
}

void chol_free(CholDevice& C) {
  for (void* p : C.allocs) cudaFree(p);
  C = CholDevice();
}

int chol_solve(CholDevice& C, cudaStream_t st, const double* Hd, const double* Ho, const double* b, double lambda, double* x, double* scale_out,
               int* ok_out, int* launches) {
  CholView V;
  V.fronts = C.fronts; V.rows = C.rows; V.rel = C.rel; V.child_idx = C.child_idx; V.level_fronts = C.level_fronts; V.perm = C.perm;
  V.diag_dst = C.diag_dst; V.off_dst = C.off_dst; V.rhs_dst = C.rhs_dst; V.diag_ld = C.diag_ld; V.off_ld = C.off_ld; V.off_tr = C.off_tr;
  V.arena = C.arena; V.xp = C.xp; V.fail_flag = C.fail_flag; V.n = C.n; V.n_off = C.n_off;
  CUDA_TRY(cudaMemsetAsync(C.arena, 0, (size_t)C.arena_doubles * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(C.fail_flag, 0, sizeof(int), st));
  const long long total = (long long)C.n * 36 + (long long)C.n_off * 36 + (long long)C.n * 6;
  chol_scatter_kernel<<<(unsigned)((total + kCholThreads - 1) / kCholThreads), kCholThreads, 0, st>>>(V, Hd, Ho, b, lambda);
  int nl = 1;
  for (int l = 0; l < C.n_levels; l++) {
    // small fronts: one CTA each; large fronts: teams of CTAs in one cooperative launch
    const int ns = C.small_ptr[l + 1] - C.small_ptr[l], nb = C.big_ptr[l + 1] - C.big_ptr[l];
    if (ns > 0) {
      chol_front_kernel<false><<<ns, kCholThreads, FrontSmem<false>::bytes, st>>>(V, C.small_list + C.small_ptr[l], ns, 1, C.bars);
      nl++;
    }
    if (nb > 0) {
      const int n_teams = std::min(nb, C.coop_grid);
      // no more CTAs per front than the largest front of the level has 96 x 96 tiles in its trailing update (fewer barrier parties)
      const int nt = (C.level_big[l] + 95) / 96, tiles = nt * (nt + 1) / 2;
      int team_size = std::max(1, std::min(std::min(C.coop_grid / n_teams, kCholMaxTeam), tiles));
      int grid = n_teams * team_size;
      const int* list = C.big_list + C.big_ptr[l];
      int n_list = nb;
      if (team_size == 1) chol_front_kernel<false><<<grid, kCholThreads, FrontSmem<false>::bytes, st>>>(V, list, n_list, 1, C.bars);
      else {
        CUDA_TRY(cudaMemsetAsync(C.bars, 0, (size_t)n_teams * sizeof(unsigned int), st));
        void* args[] = {(void*)&V, (void*)&list, (void*)&n_list, (void*)&team_size, (void*)&C.bars};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)chol_front_kernel<true>, dim3(grid), dim3(kCholThreads), args, FrontSmem<true>::bytes, st));
      }
      nl++;
    }
  }
  const size_t smem = (size_t)C.max_front * sizeof(double);
  for (int l = C.n_levels - 1; l >= 0; l--) {
    const int cnt = C.level_ptr[l + 1] - C.level_ptr[l];
    if (cnt <= 0) continue;
    chol_backward_kernel<<<cnt, kCholThreads, smem, st>>>(V, C.level_ptr[l]);
    nl++;
  }
  chol_finish_kernel<<<1, 1024, 0, st>>>(V, b, lambda, x, scale_out, ok_out);
  nl++;
  CUDA_TRY(cudaGetLastError());
  if (launches) *launches += nl;
  return LVS_OK;
}

}  // namespace lvs
