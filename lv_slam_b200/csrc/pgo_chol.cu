// Supernodal multifrontal block Cholesky for the pose-graph normal equations (see pgo_chol.cuh for what it replaces).
#include <algorithm>
#include <cstdlib>
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
// took 2.4 s on the 50 000-vertex sphere, 0.27 s with approximate degrees and lazy heaps, with identical fill on the sphere graphs).
//
// The queue "lowest degree first, smallest index on a tie": one two-level BITMAP over the variable indices per degree that occurs
// (allocated on first use), with exact deletion - a degree update clears one bit and sets another, a pick is a find-first-set over
// ~13 summary words.  The lazy min-heaps it replaces spent 250 of the ordering's 345 ms (50 000 vertices) pushing 2.9 M entries and
// popping 2.3 M stale ones for 50 000 picks; the picks, hence the ordering, are the same (checked by hash).  Above kBitmapMaxN variables
// the bitmaps would cost too much memory and the heaps stay.
constexpr int kBitmapMaxN = 1 << 18;
struct DegreeBitmaps {
  int words = 0, swords = 0;
  std::vector<std::vector<unsigned long long>> bits;      // per degree: [words] + [swords] summary
  std::vector<int> count;
  void init(int n) { words = (n + 63) / 64; swords = (words + 63) / 64; bits.assign((size_t)n + 1, {}); count.assign((size_t)n + 1, 0); }
  void insert(int d, int i) {
    std::vector<unsigned long long>& b = bits[d];
    if (b.empty()) b.assign((size_t)words + swords, 0ull);
    b[i >> 6] |= 1ull << (i & 63);
    b[words + (i >> 12)] |= 1ull << ((i >> 6) & 63);
    count[d]++;
  }
  void erase(int d, int i) {
    std::vector<unsigned long long>& b = bits[d];
    b[i >> 6] &= ~(1ull << (i & 63));
    if (b[i >> 6] == 0) b[words + (i >> 12)] &= ~(1ull << ((i >> 6) & 63));
    count[d]--;
  }
  int first(int d) const {                                  // smallest index in bucket d (count[d] > 0)
    const std::vector<unsigned long long>& b = bits[d];
    for (int s = 0; s < swords; s++)
      if (b[words + s]) { const int w = 64 * s + __builtin_ctzll(b[words + s]); return 64 * w + __builtin_ctzll(b[w]); }
    return -1;
  }
};

static void minimum_degree(int n, const std::vector<std::vector<int>>& adj, std::vector<int>& order, std::vector<std::vector<int>>& pattern) {
  std::vector<std::vector<int>> av(adj), ae(n);
  pattern.assign(n, {});
  std::vector<std::vector<int>>& el = pattern;      // an element's member list IS the pivot's column pattern (kept; `dead` marks absorption)
  std::vector<char> gone(n, 0), dead(n, 0);
  std::vector<int> mark(n, -1), deg(n), w(n, 0), wmark(n, -1);
  // The pivot is the smallest index of the lowest non-empty degree.  Above kBitmapMaxN variables: one min-heap of variable indices per
  // degree with lazy deletion (an entry of bucket d is live while the variable is still there and its degree is still d).
  const bool use_bitmaps = n <= kBitmapMaxN;
  DegreeBitmaps bm;
  if (use_bitmaps) bm.init(n);
  std::vector<std::vector<int>> bucket(use_bitmaps ? 0 : n + 1);
  int mindeg = n;
  auto push = [&](int d, int i) {                           // heaps: lazy (stale entries stay); bitmaps: the caller erased the old entry
    if (use_bitmaps) bm.insert(d, i);
    else {
      std::vector<int>& b = bucket[d];
      b.push_back(i);
      std::push_heap(b.begin(), b.end(), std::greater<int>());
    }
    mindeg = std::min(mindeg, d);
  };
  for (int i = 0; i < n; i++) {
    std::sort(av[i].begin(), av[i].end());
    av[i].erase(std::unique(av[i].begin(), av[i].end()), av[i].end());
    deg[i] = (int)av[i].size();
    push(deg[i], i);
  }
  order.clear(); order.reserve(n);
  int tag = 0, wtag = 0;
  std::vector<int> Lp;
  for (int step = 0; step < n; step++) {
    int p = -1;
    if (use_bitmaps) {
      while (bm.count[mindeg] == 0) mindeg++;
      p = bm.first(mindeg);
      bm.erase(mindeg, p);
    }
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
    }
    std::vector<int>().swap(av[p]);
    std::vector<int>().swap(ae[p]);
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
      for (int x : ae[i]) if (x != p) { if (w[x] == 0 && !dead[x]) dead[x] = 1; d += w[x]; }
      d = std::min<long long>(d, remaining - 1);
      d = std::min<long long>(d, (long long)deg[i] + (long long)Lp.size() - 1);
      const int nd = (int)std::max<long long>(d, 0);
      if (nd != deg[i]) { if (use_bitmaps) bm.erase(deg[i], i); deg[i] = nd; push(nd, i); }
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
  // per column of the final numbering: pattern size and smallest row (its parent in the elimination tree) - all the supernode
  // detection needs; the sorted patterns themselves are built further down for the LAST column of every front only
  std::vector<int> csize(n), par(n, -1);
  for (int c = 0; c < n; c++) {
    const std::vector<int>& P = pat[S.perm[c]];
    csize[c] = (int)P.size();
    int best = n;
    for (int v : P) best = std::min(best, S.iperm[v]);
    if (!P.empty()) par[c] = best;
  }
  // fundamental supernodes: column c joins c-1 when pattern(c-1) = {c} u pattern(c)
  // plus RELAXED amalgamation: the last child of a front (its columns directly precede the front's in the postorder) is merged into it
  // when that pads the child's columns with few explicit zero rows - fewer, fatter fronts and fewer tree levels for a latency-bound
  // factorisation, at the price of some arithmetic on zeros.  relax = largest number of index blocks a child column may be padded by
  // (LVS_CHOL_RELAX overrides; 0 = fundamental supernodes only).
  int relax = 12;
  if (const char* e = getenv("LVS_CHOL_RELAX")) relax = atoi(e);
  S.col_front.assign(n, -1);
  int pad_budget = 0;            // padding already accepted for the columns of the front being grown
  for (int c = 0; c < n; c++) {
    bool join = c > 0 && par[c - 1] == c && csize[c - 1] == csize[c] + 1;
    if (join) {}                            // fundamental: no padding added
    else if (c > 0 && par[c - 1] == c && relax > 0) {
      const int extra = csize[c] + 1 - csize[c - 1];     // rows the child's last column lacks (>= 1 here)
      if (extra + pad_budget <= relax) { join = true; pad_budget += extra; }
    }
    if (!join) pad_budget = 0;
    if (join) { S.fronts.back().w++; }
    else { CholFront f; memset(&f, 0, sizeof f); f.c0 = c; f.w = 1; f.parent = -1; S.fronts.push_back(f); }
    S.col_front[c] = (int)S.fronts.size() - 1;
  }
  const int nf = (int)S.fronts.size();
  std::vector<std::vector<int>> fkids(nf);
  // index sets R_S = pattern of the front's last column, ascending: two transpositions instead of a sort per column (the row lists fill
  // in column order, the fronts' lists then fill in row order)
  {
    size_t total = 0;
    for (int s = 0; s < nf; s++) { CholFront& f = S.fronts[s]; f.r = csize[f.c0 + f.w - 1]; f.rows_off = (int)total; total += f.r; }
    S.rows.resize(total);
    std::vector<int> rptr(n + 1, 0);
    for (int s = 0; s < nf; s++) for (int v : pat[S.perm[S.fronts[s].c0 + S.fronts[s].w - 1]]) rptr[S.iperm[v] + 1]++;
    for (int r = 0; r < n; r++) rptr[r + 1] += rptr[r];
    std::vector<int> rfront(total), rfill(rptr.begin(), rptr.end() - 1);
    for (int s = 0; s < nf; s++) for (int v : pat[S.perm[S.fronts[s].c0 + S.fronts[s].w - 1]]) rfront[rfill[S.iperm[v]]++] = s;
    std::vector<int> ffill(nf);
    for (int s = 0; s < nf; s++) ffill[s] = S.fronts[s].rows_off;
    for (int r = 0; r < n; r++) for (int k = rptr[r]; k < rptr[r + 1]; k++) S.rows[ffill[rfront[k]]++] = r;
  }
  for (int s = 0; s < nf; s++) {
    CholFront& f = S.fronts[s];
    f.F = 6 * (f.w + f.r) + 1;
    f.parent = f.r ? S.col_front[par[f.c0 + f.w - 1]] : -1;
    if (f.parent >= 0) fkids[f.parent].push_back(s);
    f.off = S.arena;
    S.arena += (long long)f.F * f.F;
    S.max_front = std::max(S.max_front, f.F);
    for (int k = 0; k < f.w; k++) S.nnz_l_blocks += 1 + (long long)csize[f.c0 + k];   // the true fill: padding of relaxed supernodes is stored (arena, flops) but not counted
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
  if (getenv("LVS_CHOL_HIST")) {     // diagnostics: fronts per level by size class
    for (int l = 0; l <= max_level; l++) {
      int cnt[6] = {0, 0, 0, 0, 0, 0}, maxp = 0, maxF = 0, maxkids = 0;
      for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; k++) {
        const CholFront& f = S.fronts[S.level_fronts[k]];
        cnt[f.F <= 32 ? 0 : f.F <= 48 ? 1 : f.F <= 64 ? 2 : f.F <= 96 ? 3 : f.F <= 192 ? 4 : 5]++;
        maxp = std::max(maxp, 6 * f.w); maxF = std::max(maxF, f.F); maxkids = std::max(maxkids, f.child_end - f.child_begin);
      }
      fprintf(stderr, "[chol hist] level %2d: F<=32 %4d  <=48 %4d  <=64 %4d  <=96 %4d  <=192 %4d  >192 %4d   max F %4d  max pivots %4d  max children %d\n", l, cnt[0], cnt[1], cnt[2], cnt[3],
              cnt[4], cnt[5], maxF, maxp, maxkids);
    }
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
constexpr int kCholBigFront = 192;      // fronts with F above this go to the team kernel
constexpr int kNB = 24;                 // pivot columns per panel of the backward substitution
constexpr int kCholSmemFront = 64;      // fronts up to this many rows are factored in shared memory (single-CTA kernel)
constexpr int kMB = 16;                 // micro block: factored and inverted by one warp in registers
constexpr int kSlab = 32;               // rows of one row-solve work item: 4 x 2 m8n8 tiles of a 32 x 16 micro column, one per warp

// Partial dense Cholesky of a front, per kernel variant.  Panels are WIDE (96 pivot columns for the team kernel): a panel costs two
// team barriers and one pass over the trailing matrix whatever its width.  All dense arithmetic but the 16 x 16 micro blocks themselves
// runs on the FP64 tensor cores (mma.sync.m8n8k4.f64, DMMA); every operand array in shared memory has a row stride = 4 (mod 16) doubles,
// which makes the fragment loads of all three phases bank-conflict free.
//   (A) the diagonal block is factored in 16 x 16 micro blocks, each by ONE WARP in registers (a lane owns a row, columns travel by
//       shuffle); the same warp builds the micro block's inverse W row by row as the rows of L complete.  The rows of the diagonal block
//       below it are then X = S W^T and the rest of the block takes S -= X X^T, both as DMMA tiles spread over the warps.
//       In a team, RANK 0 ALONE factors the block and does so one panel AHEAD: right after it has updated the next diagonal block (tile 0
//       of the trailing update) it factors it while the other CTAs are still busy with the rest of the update (look-ahead); the factor
//       and the inverses travel through the front / a per-team scratch and are picked up by everyone after the panel's closing barrier;
//   (B) the rows below are solved in slabs of 32 rows: per micro column a DMMA block product with the columns already solved, then
//       the DMMA product with the inverse;
//   (C) the trailing update: TILE x TILE outputs per pass, K = NB, panels staged by cp.async.
template <bool TEAM>
struct FrontCfg {
  static constexpr int NB = TEAM ? 96 : 48;       // pivot columns per panel
  static constexpr int TILE = TEAM ? 96 : 64;     // trailing-update tile (outputs per CTA pass: TILE x TILE)
  static constexpr int LDL = TILE + 4;            // row stride of the staged panels: = 4 (mod 16)
  static constexpr int LDD = NB + 4;              // row stride of the diagonal block
  static constexpr int LDW = kMB + 4;             // row stride of a micro block's inverse
  static constexpr int LDX = kSlab + 4;           // column stride of a slab (column-major)
  static constexpr int NMB = NB / kMB;
  static constexpr int WM = 4, WN = 2;            // warp grid over a tile
  static constexpr int TM = TILE / 8 / WM, TN = TILE / 8 / WN;   // m8n8 tiles per warp
  static constexpr int WSCR = NMB * kMB * LDW + NB;               // per-team scratch: the inverses and the reciprocal diagonal (doubles)
  // dynamic shared memory, in doubles: phases A + B and phase C alias
  static constexpr size_t oD = 0;
  static constexpr size_t oW = oD + (size_t)NB * LDD;
  static constexpr size_t oIl = oW + (size_t)NMB * kMB * LDW;
  static constexpr size_t oX = oIl + NB;
  static constexpr size_t phaseAB = oX + (size_t)NB * LDX;
  static constexpr size_t phaseC = 2 * (size_t)NB * LDL;
  static constexpr size_t phaseH = TEAM ? oX + (size_t)NB * LDL : 0;     // rank 0's look-ahead: the diagonal block next to the panel rows that update it
  static constexpr size_t oFront = (phaseAB > phaseC ? phaseAB : phaseC) > phaseH ? (phaseAB > phaseC ? phaseAB : phaseC) : phaseH;
  static constexpr size_t doubles = oFront + (TEAM ? 0 : (size_t)kCholSmemFront * kCholSmemFront);
  static constexpr size_t bytes = doubles * sizeof(double);
  static_assert(TILE % (8 * WM) == 0 && TILE % (8 * WN) == 0 && LDL % 16 == 4 && LDD % 16 == 4 && LDW % 16 == 4 && LDX % 16 == 4 && NB % kMB == 0, "tile shape");
  static_assert(!TEAM || TILE == NB, "look-ahead: tile 0 of the trailing update is the next diagonal block");
  static_assert(kCholThreads == 256 && kSlab == 32 && kMB == 16, "phase B maps one m8n8 tile of a 32 x 16 micro column to each of the 8 warps");
};

struct CholView {
  const CholFront* fronts;
  const int *rows, *rel, *child_idx, *level_fronts, *perm;
  const long long *diag_dst, *off_dst, *rhs_dst;
  const int *diag_ld, *off_ld;
  const unsigned char* off_tr;
  double* arena;
  double* xp;
  double* wscratch;        // per team: inverses of the current panel's micro blocks + reciprocal diagonal (team kernel)
  int* fail_flag;
  int n, n_off;
  long long* dbg;          // diagnostics (LVS_DEBUG_TIMING): per-phase SM clocks of ranks 0 and 1 of the LAST front of a team launch, null otherwise
};

__global__ void __launch_bounds__(kCholThreads) chol_scatter_kernel(CholView V, const double* __restrict__ Hd, const double* __restrict__ Ho,
                                                                    const double* __restrict__ b, double lambda) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nd = (long long)V.n * 36, no = (long long)V.n_off * 36;
  if (t < nd) {
    const unsigned tu = (unsigned)t;                          // n * 36 < 2^32 (checked at upload): 32-bit divisions by constants
    const int v = (int)(tu / 36u), e = (int)(tu % 36u), a = e / 6, c = e % 6;
    V.arena[V.diag_dst[v] + (long long)c * V.diag_ld[v] + a] = Hd[t] + (a == c ? lambda : 0.0);
  } else if (t < nd + no) {
    const unsigned u = (unsigned)(t - nd);
    const int o = (int)(u / 36u), e = (int)(u % 36u), a = e / 6, c = e % 6;     // target element (row part a, column part c)
    V.arena[V.off_dst[o] + (long long)c * V.off_ld[o] + a] = V.off_tr[o] ? Ho[(size_t)o * 36 + c * 6 + a] : Ho[(size_t)o * 36 + a * 6 + c];
  } else if (t < nd + no + (long long)V.n * 6) {
    const unsigned u = (unsigned)(t - nd - no);
    const int v = (int)(u / 6u), a = (int)(u % 6u);
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

// 8-byte asynchronous copy global -> shared; n_src = 0 writes zeros instead (out-of-range rows / columns of a tile)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src, int n_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(n_src) : "memory");
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 1 / sqrt(d) for the pivot of a column: the hardware approximation (about 28 bits) and one Newton step - half the dependent latency of
// the library routine, on the serial path of every pivot column.  The pivot itself is stored as d * il, the column as S * il.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d * y;                 // Newton: y <- y (1.5 - 0.5 d y^2), twice: 28 -> 56 -> full
  y = fma(y, fma(-h, y, 0.5), y);
  const double h2 = 0.5 * d * y;
  return fma(y, fma(-h2, y, 0.5), y);
}

// The 16 x 16 micro block at S (row stride ld, lower triangle; the strict upper part is ignored), by TWO WARPS: the producer factors it
// in registers - lane l (and l + 16, redundantly: every shuffle source is a lane below 16) owns row l & 15, columns travel by shuffle - and
// stores every finished column to S; the consumer trails one column behind (one mbarrier per column) and builds the inverse W = L^-1 row
// by row: row j of L is complete once column j is stored, and W[j][c] = (delta_jc - sum_{k < j} L[j][k] W[k][c]) / L[j][j] needs nothing
// else.  The producer's column is bound by the shuffle rate of one warp (a 64-bit shuffle issues every ~10 cycles: tools/ubench/chain_lat.cu),
// so the inverse must not share its instruction stream (fused into one warp the block took 6 000 cycles instead of 3 000).
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// producer: writes L over the lower triangle of S and the reciprocals of its diagonal to il_out[16]; a non-positive pivot raises
// *fail_flag and is replaced by 1
__device__ __forceinline__ void micro_factor(double* S, int ld, double* il_out, int* fail_flag, unsigned long long* bars) {
  const int lane = threadIdx.x & 31, row = lane & 15;
  double s[kMB];
#pragma unroll
  for (int k = 0; k < kMB; k++) s[k] = (k <= row) ? S[row * ld + k] : 0.0;
#pragma unroll
  for (int j = 0; j < kMB; j++) {
    double d = __shfl_sync(0xffffffffu, s[j], j);
    if (!(d > 0.0)) { *fail_flag = 1; d = 1.0; }
    const double il = fast_rsqrt(d);
    const double lij = (row == j) ? d * il : s[j] * il;                // rows above j hold 0 there
    if (lane < kMB && row >= j) S[row * ld + j] = lij;
    if (lane == j) il_out[j] = il;
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + j);
#pragma unroll
    for (int k = j + 1; k < kMB; k++) {
      const double lkj = __shfl_sync(0xffffffffu, lij, k);
      s[k] -= lij * lkj;                                               // meaningful for row >= k; the rest is never read
    }
  }
}

// consumer: W to Wout[16][ldw] (upper part zero); lane c (and c + 16, redundantly) owns column c
__device__ __forceinline__ void micro_invert(const double* S, int ld, const double* il, double* Wout, int ldw, unsigned long long* bars, unsigned parity) {
  const int lane = threadIdx.x & 31, col = lane & 15;
  double w[kMB];
#pragma unroll
  for (int j = 0; j < kMB; j++) {
    mbar_wait(bars + j, parity);
    double a0 = (col == j) ? 1.0 : 0.0, a1 = 0.0;                      // two partial sums: half the dependent chain
#pragma unroll
    for (int k = 0; k < j; k += 2) {
      const double2 v = *reinterpret_cast<const double2*>(S + j * ld + k);     // L[j][k], L[j][k + 1]: the same address for every lane
      a0 -= v.x * w[k];
      if (k + 1 < j) a1 -= v.y * w[k + 1];
    }
    w[j] = (a0 + a1) * il[j];                                          // zero for j < column
    if (lane < kMB) Wout[j * ldw + col] = (j >= col) ? w[j] : 0.0;
  }
}

// (A) The nb x nb diagonal block of the panel at column c of the front A (leading dimension F), in three steps that the callers combine:
// stage_diag copies its lower triangle to s_D (rows / columns past nb padded with the identity), factor_core factors it there - leaving L
// in s_D, the inverses of the micro blocks in s_W and the reciprocal diagonal in s_il - and writeback_diag stores the factor to the front.
// Whole CTA; stage_diag does not end with a barrier, the other two do.
template <bool TEAM>
__device__ __forceinline__ void stage_diag(const double* A, int F, int c, int nb, double* s_D) {
  using Cfg = FrontCfg<TEAM>;
  constexpr int NB = Cfg::NB, LDD = Cfg::LDD;
  // A warp copies whole columns (j = warp, warp + 8, ...), its lanes the rows lane, lane + 32, ...: one address per column instead of a
  // division per element.  Every global load is unconditional, from a clamped address, and three columns' worth are issued before the
  // first use (with predicated loads the compiler sinks each one next to its store and only one is in flight at a time).
  constexpr int RPL = (NB + 31) / 32, CU = 3, kW = kCholThreads / 32;
  static_assert((NB / kW) % CU == 0 && NB % kW == 0, "diagonal block staging");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = warp; j0 < NB; j0 += kW * CU) {
    double v[CU][RPL];
#pragma unroll
    for (int cu = 0; cu < CU; cu++) {
      const double* col = A + (size_t)(c + min(j0 + kW * cu, nb - 1)) * F + c;
#pragma unroll
      for (int u = 0; u < RPL; u++) v[cu][u] = ldf<TEAM>(col + min(lane + 32 * u, nb - 1));
    }
#pragma unroll
    for (int cu = 0; cu < CU; cu++) {
      const int j = j0 + kW * cu;
#pragma unroll
      for (int u = 0; u < RPL; u++) {
        const int i = lane + 32 * u;
        if (i < NB) s_D[i * LDD + j] = (i < nb && j <= i) ? v[cu][u] : (i == j ? 1.0 : 0.0);
      }
    }
  }
}

template <bool TEAM>
__device__ __forceinline__ void factor_core(int nb, double* s_D, double* s_W, double* s_il, int* fail_flag, unsigned long long* mbars, unsigned& mb_uses,
                                            long long* tm = nullptr) {
  long long tl = clock64();
#define LVS_TM(k) if (tm && threadIdx.x == 0) { const long long tn = clock64(); tm[k] += tn - tl; tl = tn; }
  using Cfg = FrontCfg<TEAM>;
  constexpr int LDD = Cfg::LDD, LDW = Cfg::LDW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nmb = (nb + kMB - 1) / kMB;
  for (int mb = 0; mb < nmb; mb++) {
    const int m0 = kMB * mb;
    double* W = s_W + mb * kMB * LDW;
    if (warp == 0) micro_factor(s_D + m0 * LDD + m0, LDD, s_il + m0, fail_flag, mbars);
    else if (warp == 1) micro_invert(s_D + m0 * LDD + m0, LDD, s_il + m0, W, LDW, mbars, mb_uses & 1);
    mb_uses++;
    LVS_TM(0)
    __syncthreads();
    LVS_TM(1)
    const int nrt = (kMB * nmb - (m0 + kMB)) / 8;          // 8-row tiles of the diagonal block below this micro block
    if (nrt > 0) {
      // X = S W^T for the rows below, in place: a warp owns whole rows, all fragments are in registers before the first store
      for (int rt = warp; rt < nrt; rt += kCholThreads / 32) {
        const int i0 = m0 + kMB + 8 * rt;
        const double* pa = s_D + (i0 + (lane >> 2)) * LDD + m0 + (lane & 3);
        const double* pw = W + (lane >> 2) * LDW + (lane & 3);
        double fa[4], x[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, y[2][2] = {{0.0, 0.0}, {0.0, 0.0}};      // x / y: two halves of K, independent chains
#pragma unroll
        for (int ks = 0; ks < 4; ks++) fa[ks] = pa[4 * ks];
#pragma unroll
        for (int ks = 0; ks < 2; ks++)
#pragma unroll
          for (int nh = 0; nh < 2; nh++) {
            dmma884(x[nh][0], x[nh][1], fa[ks], pw[8 * nh * LDW + 4 * ks]);
            dmma884(y[nh][0], y[nh][1], fa[ks + 2], pw[8 * nh * LDW + 4 * ks + 8]);
          }
        __syncwarp();
        double* px = s_D + (i0 + (lane >> 2)) * LDD + m0 + 2 * (lane & 3);
#pragma unroll
        for (int nh = 0; nh < 2; nh++) { px[8 * nh] = x[nh][0] + y[nh][0]; px[8 * nh + 1] = x[nh][1] + y[nh][1]; }
      }
      __syncthreads();
      LVS_TM(2)
      // S[i][k] -= sum_q X[i][q] X[k][q] for the rest of the diagonal block: 8 x 8 tiles of the lower part (the strict upper part of a
      // diagonal tile receives values nobody reads)
      // (nrt is even: the triangle of tiles folds into an nrt / 2 x (nrt + 1) rectangle - row a holds tile row a and tile row nrt - 1 - a)
      const int ntl = nrt * (nrt + 1) / 2, nw = nrt + 1;
      for (int t = warp; t < ntl; t += kCholThreads / 32) {
        const int a = t / nw, b = t - a * nw;
        const int ti = (b <= a) ? a : nrt - 1 - a, tj = (b <= a) ? b : b - a - 1;
        const int i0 = m0 + kMB + 8 * ti, k0 = m0 + kMB + 8 * tj;
        const double* pa = s_D + (i0 + (lane >> 2)) * LDD + m0 + (lane & 3);
        const double* pb = s_D + (k0 + (lane >> 2)) * LDD + m0 + (lane & 3);
        double u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
        dmma884(u0, u1, pa[0], pb[0]);
        dmma884(v0, v1, pa[8], pb[8]);
        dmma884(u0, u1, pa[4], pb[4]);
        dmma884(v0, v1, pa[12], pb[12]);
        double* pc = s_D + (i0 + (lane >> 2)) * LDD + k0 + 2 * (lane & 3);
        pc[0] -= u0 + v0; pc[1] -= u1 + v1;
      }
      __syncthreads();
      LVS_TM(3)
    }
  }
#undef LVS_TM
}

// also publishes the inverses and the reciprocal diagonal to the team's scratch (team kernel; wscr = nullptr otherwise)
template <bool TEAM>
__device__ __forceinline__ void writeback_diag(double* A, int F, int c, int nb, const double* s_D, const double* s_W, double* wscr) {
  using Cfg = FrontCfg<TEAM>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < nb; j += kCholThreads / 32) {          // a warp per column: no division by the (run-time) block size per element
    double* col = A + (size_t)(c + j) * F + c;
    for (int i = j + lane; i < nb; i += 32) col[i] = s_D[i * Cfg::LDD + j];
  }
  if (TEAM && wscr)
    for (int t = threadIdx.x; t < Cfg::WSCR; t += kCholThreads) wscr[t] = s_W[t];      // s_il follows s_W in shared memory
  __syncthreads();
}

template <bool TEAM>
__global__ void __launch_bounds__(kCholThreads, TEAM ? 1 : 2) chol_front_kernel(CholView V, const int* __restrict__ list, int n_list, int team_size,
                                                                  unsigned int* __restrict__ bars) {
  using Cfg = FrontCfg<TEAM>;
  constexpr int NB = Cfg::NB, TILE = Cfg::TILE, LDL = Cfg::LDL, LDD = Cfg::LDD, LDW = Cfg::LDW, LDX = Cfg::LDX;
  const int team = blockIdx.x / team_size, rank = blockIdx.x % team_size, n_teams = gridDim.x / team_size;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int* bar = bars + team;
  double* const wscr = TEAM ? V.wscratch + (size_t)team * Cfg::WSCR : nullptr;
  unsigned int target = 0;
  extern __shared__ __align__(16) double s_dyn[];
  double* s_D = s_dyn + Cfg::oD;                  // [NB][LDD] diagonal block of the panel: Schur complement, then its factor
  double* s_W = s_dyn + Cfg::oW;                  // [NMB][16][LDW] inverses of the diagonal micro blocks
  double* s_il = s_dyn + Cfg::oIl;                // [NB] reciprocals of the factor's diagonal
  double* s_X = s_dyn + Cfg::oX;                  // [NB][LDX] one slab of rows below the panel, column-major
  double* s_Li = s_dyn;                           // phase C (aliases A / B): [NB][LDL] panel rows of the tile's row range
  double* s_Lj = s_dyn + (size_t)NB * LDL;        //                          and of its column range
  double* s_front = s_dyn + Cfg::oFront;          // single-CTA launches only: room for a front of kCholSmemFront rows
  __shared__ unsigned long long s_mbar[kMB];      // one per column of a micro block: the factoring warp releases the inverting warp
  unsigned mb_uses = 0;                           // micro blocks factored so far by this CTA: the mbarriers' phase
  if (threadIdx.x < kMB) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(s_mbar + threadIdx.x)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
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
      // The team's warps take the scalar columns of the child's update block round-robin; a column's rows (from its own 6 x 6 diagonal
      // block down, contiguous in the child) run along the lanes and land at 6 rel[i / 6] + i % 6 of the target column, the column's
      // right-hand-side entry goes with lane 0.  (Enumerating the full (rc + 1) x rc x 36 rectangle element by element cost two 64-bit and
      // two run-time 32-bit divisions per element and skipped half of them.)  Four elements per lane and step, all loads before the
      // stores: the compiler must keep a load behind the previous store into the same array.
      {
        constexpr int kEa = 4, kW = kCholThreads / 32;
        const int ncol = 6 * rc, gw = rank * kW + warp, nwarps = team_size * kW;
        for (int jc = gw; jc < ncol; jc += nwarps) {
          const int bj = jc / 6;
          const double* __restrict__ Ucol = U + (size_t)(6 * c.w + jc) * Fc + 6 * c.w;       // row i of the update block at Ucol[i]
          double* const Dcol = A + (size_t)(6 * rel[bj] + (jc - 6 * bj)) * F;
          for (int i0 = 6 * bj + lane; i0 < ncol; i0 += 32 * kEa) {
            double* dst[kEa];
            double val[kEa];
#pragma unroll
            for (int u = 0; u < kEa; u++) {
              const int i = i0 + 32 * u, ic = min(i, ncol - 1), bi = ic / 6;
              double* d = Dcol + 6 * rel[bi] + (ic - 6 * bi);
              // the loads are unconditional (every address is inside the two fronts) so that all of them are in flight together
              val[u] = ldf<TEAM>(d) + __ldcg(Ucol + ic);
              dst[u] = i < ncol ? d : nullptr;
            }
#pragma unroll
            for (int u = 0; u < kEa; u++) if (dst[u]) *dst[u] = val[u];
          }
          if (lane == 0) { double* d = Dcol + (F - 1); *d = ldf<TEAM>(d) + __ldcg(Ucol - 6 * c.w + (Fc - 1)); }
        }
      }
      team_sync<TEAM>(bar, target, team_size);     // children are added one after the other: fixed summation order
    }
    // ---- partial Cholesky of the p = 6 w pivot columns, right-looking in panels of NB columns.  The right-hand side is simply
    // the last row (F - 1) of the front.
    const int p = 6 * f.w;
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tl = clock64();
    const bool dbg = TEAM && V.dbg != nullptr && rank <= 1 && threadIdx.x == 0;
#define LVS_PH(k) if (dbg) { const long long tn = clock64(); tph[k] += tn - tl; tl = tn; }
    if (TEAM) {
      // the first diagonal block: rank 0, everyone else waits (later ones are factored ahead, inside phase C of the panel before)
      if (rank == 0) {
        const int nb0 = min(NB, p);
        stage_diag<TEAM>(A, F, 0, nb0, s_D);
        __syncthreads();
        factor_core<TEAM>(nb0, s_D, s_W, s_il, V.fail_flag, s_mbar, mb_uses);
        writeback_diag<TEAM>(A, F, 0, nb0, s_D, s_W, wscr);
      }
      team_sync<TEAM>(bar, target, team_size);
    }
    for (int c = 0; c < p; c += NB) {
      const int nb = min(NB, p - c);
      const int nmb = (nb + kMB - 1) / kMB;
      LVS_PH(5)
      // (A) single CTA: factor the diagonal block now.  Team: rank 0 still holds the factor, its inverses and the reciprocal diagonal in
      // shared memory; the others fetch them (the factor's lower triangle from the front, the rest from the team's scratch).
      if (!TEAM) {
        stage_diag<TEAM>(A, F, c, nb, s_D);
        __syncthreads();
        factor_core<TEAM>(nb, s_D, s_W, s_il, V.fail_flag, s_mbar, mb_uses);
        writeback_diag<TEAM>(A, F, c, nb, s_D, s_W, nullptr);
      } else if (rank != 0) {
        stage_diag<TEAM>(A, F, c, nb, s_D);
        constexpr int kG = 4;
        for (int t0 = threadIdx.x; t0 < Cfg::WSCR; t0 += kG * kCholThreads) {
          double v[kG];
#pragma unroll
          for (int u = 0; u < kG; u++) v[u] = __ldcg(wscr + min(t0 + u * kCholThreads, Cfg::WSCR - 1));
#pragma unroll
          for (int u = 0; u < kG; u++) if (t0 + u * kCholThreads < Cfg::WSCR) s_W[t0 + u * kCholThreads] = v[u];
        }
        __syncthreads();
      }
      LVS_PH(0)
      // (B) rows below the panel: X L_D^T = A, slab by slab.  In s_X the slab is column-major.  Warp (rt, nh) owns the 8 x 8 tile (rows
      // 8 rt.., columns 8 nh..) of every 32 x 16 micro column: first the block product with the columns already solved (K = m0), then
      // the product with the micro block's inverse (K = 16), both on DMMA.
      {
        const int n_below = F - (c + nb);
        const int n_slabs = (n_below + kSlab - 1) / kSlab;
        const int rt = warp >> 1, nh = warp & 1;
        for (int slab = rank; slab < n_slabs; slab += team_size) {
          const int r0 = c + nb + slab * kSlab, nr = min(kSlab, F - r0);
          {
            constexpr int kG = (NB * kSlab / kCholThreads) % 12 == 0 ? 12 : 6;
            static_assert((NB * kSlab) % (kG * kCholThreads) == 0, "slab staging");
            for (int t0 = threadIdx.x; t0 < NB * kSlab; t0 += kG * kCholThreads) {
              double v[kG];
#pragma unroll
              for (int u = 0; u < kG; u++) {
                const int t = t0 + u * kCholThreads, r = t % kSlab, q = t / kSlab;
                v[u] = ldf<TEAM>(A + (size_t)(c + min(q, nb - 1)) * F + r0 + min(r, nr - 1));
              }
#pragma unroll
              for (int u = 0; u < kG; u++) {
                const int t = t0 + u * kCholThreads, r = t % kSlab, q = t / kSlab;
                s_X[q * LDX + r] = (q < nb && r < nr) ? v[u] : 0.0;
              }
            }
          }
          __syncthreads();
          for (int mb = 0; mb < nmb; mb++) {
            const int m0 = kMB * mb;
            double* pt = s_X + (m0 + 8 * nh + 2 * (lane & 3)) * LDX + 8 * rt + (lane >> 2);      // this lane's two outputs of the tile
            if (m0 > 0) {
              const double* pa = s_X + (lane & 3) * LDX + 8 * rt + (lane >> 2);
              const double* pb = s_D + (m0 + 8 * nh + (lane >> 2)) * LDD + (lane & 3);
              double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;           // two accumulator pairs: half the dependent DMMA chain
              for (int q = 0; q < m0; q += 8) {
                dmma884(a0, a1, pa[q * LDX], pb[q]);
                dmma884(b0, b1, pa[(q + 4) * LDX], pb[q + 4]);
              }
              pt[0] -= a0 + b0; pt[LDX] -= a1 + b1;                     // own entries only
            }
            __syncthreads();
            double x0 = 0.0, x1 = 0.0;
            {
              const double* pa = s_X + (m0 + (lane & 3)) * LDX + 8 * rt + (lane >> 2);
              const double* pw = s_W + mb * kMB * LDW + (8 * nh + (lane >> 2)) * LDW + (lane & 3);
#pragma unroll
              for (int ks = 0; ks < 4; ks++) dmma884(x0, x1, pa[4 * ks * LDX], pw[4 * ks]);
            }
            __syncthreads();
            pt[0] = x0; pt[LDX] = x1;
            __syncthreads();
          }
          for (int t = threadIdx.x; t < nb * kSlab; t += kCholThreads) {
            const int r2 = t % kSlab, q = t / kSlab;
            if (r2 < nr) A[(size_t)(c + q) * F + r0 + r2] = s_X[q * LDX + r2];
          }
          __syncthreads();
        }
      }
      LVS_PH(1)
      team_sync<TEAM>(bar, target, team_size);
      LVS_PH(2)
      // (C) A[i, j] -= sum_q L[i, c+q] L[j, c+q] for c + nb <= j <= i (block-lower part), TILE x TILE outputs per pass on the tensor
      // cores.  Warp (wm, wn) of the 4 x 2 grid owns TM x TN m8n8 tiles; per k-step of 4 it loads TM + TN fragments for TM * TN DMMAs.
      // Team, and another panel follows: rank 0 takes the next diagonal block (tile 0 when that panel is a full one) and factors it while
      // the other CTAs share the rest of the update.
      const int base = c + nb, rem = F - base;
      const int nt = (rem + TILE - 1) / TILE;
      const int ntiles = nt * (nt + 1) / 2;
      const int nbk = (nb + 3) & ~3;
      const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
      const bool has_next = TEAM && base < p;
      const int nbn = has_next ? min(NB, p - base) : 0;
      const bool fused_head = has_next && nbn == NB;          // rank 0 updates the next diagonal block in shared memory and factors it from there
      // tiles of this CTA: everything round-robin, unless a panel follows - then rank 0 has the next diagonal block (nothing else when it
      // updates that block in shared memory, tile 0 otherwise) and the others share tiles 1, 2, ...
      int tile_first = rank, tile_step = team_size;
      if (has_next) {
        tile_first = (rank == 0) ? (fused_head ? ntiles : 0) : rank;
        tile_step = (rank == 0) ? ntiles : team_size - 1;
      }
      if (has_next && fused_head && rank == 0) {
        // the next diagonal block S and the NB panel rows X that update it: S -= X X^T on the lower 8 x 8 tiles, spread over the warps
        double* s_H = s_dyn + Cfg::oX;                        // [nbk][LDL]: X[base + r][c + k] at k * LDL + r  (fits: oX + NB * LDL doubles of the phase C area)
        stage_diag<TEAM>(A, F, base, NB, s_D);
        for (int q = warp; q < nbk; q += kCholThreads / 32) {      // a warp per panel column, its lanes the 96 rows
          const double* src = A + (size_t)(c + min(q, nb - 1)) * F + base;
          double* dst = s_H + q * LDL;
#pragma unroll
          for (int u = 0; u < NB / 32; u++) cp_async8(dst + lane + 32 * u, src + lane + 32 * u, q < nb ? 8 : 0);
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // (NT8 is even: the triangle of tiles folds into an NT8 / 2 x (NT8 + 1) rectangle - row a holds tile row a and tile row NT8 - 1 - a)
        constexpr int NT8 = NB / 8, NTL = NT8 * (NT8 + 1) / 2, NW8 = NT8 + 1;
        static_assert(NT8 % 2 == 0, "folded tile triangle");
        for (int t = warp; t < NTL; t += kCholThreads / 32) {
          const int fa = t / NW8, fb = t - fa * NW8;
          const int ti = (fb <= fa) ? fa : NT8 - 1 - fa, tj = (fb <= fa) ? fb : fb - fa - 1;
          const double* pa = s_H + (lane & 3) * LDL + 8 * ti + (lane >> 2);
          const double* pb = s_H + (lane & 3) * LDL + 8 * tj + (lane >> 2);
          double u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
          for (int k0 = 0; k0 < nbk; k0 += 8) {
            dmma884(u0, u1, pa[k0 * LDL], pb[k0 * LDL]);
            if (k0 + 4 < nbk) dmma884(v0, v1, pa[(k0 + 4) * LDL], pb[(k0 + 4) * LDL]);
          }
          double* pc = s_D + (8 * ti + (lane >> 2)) * LDD + 8 * tj + 2 * (lane & 3);
          pc[0] -= u0 + v0; pc[1] -= u1 + v1;
        }
        __syncthreads();
        factor_core<TEAM>(NB, s_D, s_W, s_il, V.fail_flag, s_mbar, mb_uses);
        writeback_diag<TEAM>(A, F, base, NB, s_D, s_W, wscr);
      }
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        int ti = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);      // the two loops below make it exact
        while (ti * (ti + 1) / 2 > tile) ti--;
        while ((ti + 1) * (ti + 2) / 2 <= tile) ti++;
        const int tj = tile - ti * (ti + 1) / 2;
        const int i0 = base + TILE * ti, j0 = base + TILE * tj;
        // The two panels go to shared memory by cp.async (every copy of the tile in flight at once, no staging registers).  The lines
        // land in L1 as well, which is safe in team mode although L1 is not coherent across SMs: panel columns are final when this
        // phase starts (the barrier after (B)) and nothing else in the kernel reads the arena through L1.  A front held in shared memory
        // (single-CTA kernel) is copied with plain loads.
        if (in_smem) {
          for (int t = threadIdx.x; t < nbk * TILE; t += kCholThreads) {
            const int r = t % TILE, q = t / TILE;
            s_Li[q * LDL + r] = (q < nb && i0 + r < F) ? A[(size_t)(c + q) * F + i0 + r] : 0.0;
            s_Lj[q * LDL + r] = (q < nb && j0 + r < F) ? A[(size_t)(c + q) * F + j0 + r] : 0.0;
          }
        } else {
          // a warp copies whole panel columns (q = warp, warp + 8, ...), its lanes the rows lane, lane + 32, ...: one address per column
          // and panel instead of a division, a clamp and a predicate per element (the element-wise loop issued more instructions than
          // the tile's DMMA loop)
          const int ri = F - 1 - i0, rj = F - 1 - j0;         // last valid row of each panel, relative to the tile
          for (int q = warp; q < nbk; q += kCholThreads / 32) {
            const int qc = min(q, nb - 1);
            const double* srci = A + (size_t)(c + qc) * F + i0;
            const double* srcj = A + (size_t)(c + qc) * F + j0;
            double* di = s_Li + q * LDL;
            double* dj = s_Lj + q * LDL;
            const bool qok = q < nb;
#pragma unroll
            for (int u = 0; u < TILE / 32; u++) {
              const int r = lane + 32 * u;
              cp_async8(di + r, srci + min(r, ri), (qok && r <= ri) ? 8 : 0);
              cp_async8(dj + r, srcj + min(r, rj), (qok && r <= rj) ? 8 : 0);
            }
          }
          asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        double acc[Cfg::TM][Cfg::TN][2];
#pragma unroll
        for (int a = 0; a < Cfg::TM; a++)
#pragma unroll
          for (int b2 = 0; b2 < Cfg::TN; b2++) acc[a][b2][0] = acc[a][b2][1] = 0.0;
        const double* pa = s_Li + (lane & 3) * LDL + 8 * (wm * Cfg::TM) + (lane >> 2);
        const double* pb = s_Lj + (lane & 3) * LDL + 8 * (wn * Cfg::TN) + (lane >> 2);
        for (int k0 = 0; k0 < nbk; k0 += 4) {
          double fa[Cfg::TM], fb[Cfg::TN];
#pragma unroll
          for (int a = 0; a < Cfg::TM; a++) fa[a] = pa[k0 * LDL + 8 * a];
#pragma unroll
          for (int b2 = 0; b2 < Cfg::TN; b2++) fb[b2] = pb[k0 * LDL + 8 * b2];
#pragma unroll
          for (int a = 0; a < Cfg::TM; a++)
#pragma unroll
            for (int b2 = 0; b2 < Cfg::TN; b2++) dmma884(acc[a][b2][0], acc[a][b2][1], fa[a], fb[b2]);
        }
        // a lane holds C[8 a + lane / 4][8 b + 2 (lane % 4) + e]; keep the block-lower part (6 x 6 blocks), all loads before the stores.
        // Row / column block indices and the column addresses are worked out once per lane (TM + 2 TN of them), not per element and pass.
        int rblk[Cfg::TM], cblk[Cfg::TN][2];
        const int gi_0 = i0 + 8 * (wm * Cfg::TM) + (lane >> 2), gj_0 = j0 + 8 * (wn * Cfg::TN) + 2 * (lane & 3);
#pragma unroll
        for (int a = 0; a < Cfg::TM; a++) rblk[a] = (gi_0 + 8 * a < F) ? (gi_0 + 8 * a) / 6 : -1;                       // -1: past the front
#pragma unroll
        for (int b2 = 0; b2 < Cfg::TN; b2++)
#pragma unroll
          for (int e = 0; e < 2; e++) cblk[b2][e] = (gj_0 + 8 * b2 + e < F - 1) ? (gj_0 + 8 * b2 + e) / 6 : 0x7fffffff;     // never <= a row block
        double* const C0 = A + (size_t)gj_0 * F + gi_0;
#pragma unroll
        for (int a = 0; a < Cfg::TM; a++)
#pragma unroll
          for (int b2 = 0; b2 < Cfg::TN; b2++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const bool live = rblk[a] >= cblk[b2][e];
              acc[a][b2][e] = live ? ldf<TEAM>(C0 + (size_t)(8 * b2 + e) * F + 8 * a) - acc[a][b2][e] : 0.0;
            }
#pragma unroll
        for (int a = 0; a < Cfg::TM; a++)
#pragma unroll
          for (int b2 = 0; b2 < Cfg::TN; b2++)
#pragma unroll
            for (int e = 0; e < 2; e++)
              if (rblk[a] >= cblk[b2][e]) C0[(size_t)(8 * b2 + e) * F + 8 * a] = acc[a][b2][e];
        __syncthreads();
      }
      if (has_next && !fused_head && rank == 0) {
        // the (partial) last panel's diagonal block: tile 0 above has brought it up to date
        stage_diag<TEAM>(A, F, base, nbn, s_D);
        __syncthreads();
        factor_core<TEAM>(nbn, s_D, s_W, s_il, V.fail_flag, s_mbar, mb_uses);
        writeback_diag<TEAM>(A, F, base, nbn, s_D, s_W, wscr);
      }
      LVS_PH(3)
      team_sync<TEAM>(bar, target, team_size);
      LVS_PH(4)
    }
    if (dbg) { for (int k = 0; k < 6; k++) V.dbg[8 * rank + k] = tph[k]; V.dbg[8 * rank + 6] = F; V.dbg[8 * rank + 7] = team_size; }
#undef LVS_PH
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
__global__ void __launch_bounds__(kCholThreads) chol_backward_kernel(CholView V, const int* __restrict__ list) {
  extern __shared__ double s_dyn[];                 // xs[F - 1]: y, overwritten by x panel by panel, followed by x_R
  __shared__ double s_D[kNB][kNB + 1], s_dot[kNB];
  const int s = list[blockIdx.x];
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

// Backward substitution of the LARGE fronts of a level by teams of CTAs (cooperative launch, like the forward kernel): one CTA needs
// 4 ms for the 4 932 pivots of the 50 000-vertex graph's root - it streams the whole factor of the front through one SM.  Here every
// CTA of the team holds the front's right-hand side / solution vector in shared memory; the panels are solved LEFT-looking from the
// last one up: the product of the panel's block column with the part of the solution that is already known is split over the CTAs by
// rows (coalesced: a column of the front is contiguous), the partial sums meet in global memory (one team barrier per panel, double
// buffered), and every CTA adds them in the same fixed order and solves the 32 x 32 triangle itself - identical bits everywhere, no
// broadcast needed.
constexpr int kBNB = 96;                // pivot columns per panel: one team barrier each (with 32 the barriers were most of a large front's time)
constexpr int kBSlots = kBNB / 32;      // rows of the panel's triangle per lane of the solving warp
__global__ void __launch_bounds__(kCholThreads, 1) chol_backward_team_kernel(CholView V, const int* __restrict__ list, int n_list, int team_size,
                                                                             unsigned int* __restrict__ bars, double* __restrict__ scratch, int max_chunks) {
  extern __shared__ double s_dyn[];                 // s_T[kBNB][kBNB + 1] (the panel's triangle), then xs[F - 1]
  __shared__ double s_rhs[kBNB];
  double (*s_T)[kBNB + 1] = reinterpret_cast<double (*)[kBNB + 1]>(s_dyn);
  const int team = blockIdx.x / team_size, rank = blockIdx.x % team_size, n_teams = gridDim.x / team_size;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kW = kCholThreads / 32;
  unsigned int* bar = bars + team;
  unsigned int target = 0;
  double* xs = s_dyn + kBNB * (kBNB + 1);
  double* part = scratch + (size_t)team * 2 * kBNB * max_chunks;     // [2 parity][kBNB][max_chunks]
  for (int fi = team; fi < n_list; fi += n_teams) {
    const CholFront f = V.fronts[list[fi]];
    const double* __restrict__ A = V.arena + f.off;
    const int F = f.F, p = 6 * f.w, nr = 6 * f.r;
    for (int j = threadIdx.x; j < p; j += kCholThreads) xs[j] = __ldcg(A + (size_t)j * F + (F - 1));
    for (int i = threadIdx.x; i < nr; i += kCholThreads) xs[p + i] = __ldcg(V.xp + (size_t)V.rows[f.rows_off + i / 6] * 6 + i % 6);
    __syncthreads();
    if (nr > 0) {
      // y - B^T x_R: the team's warps share the pivot columns, the results travel through the front's own slice of xp
      for (int j = rank * kW + warp; j < p; j += team_size * kW) {
        const double* __restrict__ col = A + (size_t)j * F + p;
        double acc = 0.0;
        for (int i0 = lane; i0 < nr; i0 += 32 * 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = __ldcg(col + min(i0 + 32 * u, nr - 1));
#pragma unroll
          for (int u = 0; u < 8; u++) { const int i = i0 + 32 * u; acc += (i < nr) ? v[u] * xs[p + i] : 0.0; }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) V.xp[(size_t)f.c0 * 6 + j] = xs[j] - acc;
      }
      team_sync<true>(bar, target, team_size);
      for (int j = threadIdx.x; j < p; j += kCholThreads) xs[j] = __ldcg(V.xp + (size_t)f.c0 * 6 + j);
      __syncthreads();
    }
    const int n_panels = (p + kBNB - 1) / kBNB;
    for (int pi = n_panels - 1; pi >= 0; pi--) {
      const int c = pi * kBNB, nb = min(kBNB, p - c);
      // (a) partial[q][chunk] = sum over a chunk of 256 later rows i of L[i][c + q] xs[i]: the (column, chunk) items are dealt to the
      // team's warps round-robin, one pass of eight coalesced loads per lane each (a warp that walked a whole column chunk after chunk
      // had one pass of loads in flight at a time: the kernel was bound by that latency, not by the barrier)
      double* mine = part + (size_t)(pi & 1) * kBNB * max_chunks;
      const int later = p - (c + nb), n_chunks = (later + 255) / 256;
      for (int item = rank * kW + warp; item < kBNB * n_chunks; item += team_size * kW) {
        const int q = item % kBNB, ch = item / kBNB;
        double acc = 0.0;
        if (q < nb) {
          const double* __restrict__ col = A + (size_t)(c + q) * F;
          const int i0 = c + nb + 256 * ch + lane;
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = __ldcg(col + min(i0 + 32 * u, p - 1));
#pragma unroll
          for (int u = 0; u < 8; u++) { const int i = i0 + 32 * u; acc += (i < p) ? v[u] * xs[i] : 0.0; }
          for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) mine[q * max_chunks + ch] = acc;
      }
      // the panel's triangle, while the partial sums of the other CTAs arrive
      for (int t = threadIdx.x; t < kBNB * kBNB; t += kCholThreads) {
        const int j = t / kBNB, i = t % kBNB;
        const double v = __ldcg(A + (size_t)(c + min(j, nb - 1)) * F + c + min(i, nb - 1));
        s_T[i][j] = (i < nb && j <= i) ? v : (i == j ? 1.0 : 0.0);
      }
      team_sync<true>(bar, target, team_size);
      // (b) rhs = xs[panel] - sum of the partials in chunk order
      if (threadIdx.x < kBNB) {
        const int q = threadIdx.x;
        const double* src = mine + q * max_chunks;
        double ssum = 0.0;
        for (int r0 = 0; r0 < n_chunks; r0 += 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = __ldcg(src + min(r0 + u, n_chunks - 1));
#pragma unroll
          for (int u = 0; u < 8; u++) ssum += (r0 + u < n_chunks) ? v[u] : 0.0;
        }
        s_rhs[q] = (q < nb) ? xs[c + q] - ssum : 0.0;
      }
      __syncthreads();
      // (c) L_D^T x = rhs by warp 0: lane m keeps rows m, m + 32, m + 64 of the triangle and subtracts every x_j from their running
      // right-hand sides as it appears; the dependent chain per column is multiply -> shuffle -> multiply-add
      if (warp == 0) {
        double run[kBSlots], inv[kBSlots];
#pragma unroll
        for (int sl = 0; sl < kBSlots; sl++) {
          const int r = lane + 32 * sl;
          run[sl] = s_rhs[r];
          inv[sl] = 1.0 / s_T[r][r];                               // 1 on the identity padding
        }
#pragma unroll
        for (int sj = kBSlots - 1; sj >= 0; sj--) {
#pragma unroll 8
          for (int jl = 31; jl >= 0; jl--) {
            const int j = 32 * sj + jl;                            // columns past nb: identity rows, zero right-hand side - harmless
            double t[kBSlots];
#pragma unroll
            for (int sl = 0; sl <= sj; sl++) t[sl] = s_T[j][lane + 32 * sl];          // independent of x_j: issued ahead of the chain
            const double xj = __shfl_sync(0xffffffffu, run[sj] * inv[sj], jl);
            if (lane == jl && j < nb) xs[c + j] = xj;
#pragma unroll
            for (int sl = 0; sl <= sj; sl++)
              if (lane + 32 * sl < j) run[sl] -= t[sl] * xj;
          }
        }
      }
      __syncthreads();
    }
    if (rank == 0)
      for (int j = threadIdx.x; j < p; j += kCholThreads) V.xp[(size_t)f.c0 * 6 + j] = xs[j];
    team_sync<true>(bar, target, team_size);       // xp of this front is complete before the team's next front (a descendant level never shares a launch)
  }
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
  if (((long long)S.n * 42 + (long long)n_off * 36) >= (1ll << 32)) return fail(LVS_ERR_INVALID_ARG, "graph too large for the scatter kernel's 32-bit element index");
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
    CUDA_TRY(cudaFuncSetAttribute(chol_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrontCfg<true>::bytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_front_kernel<true>, kCholThreads, FrontCfg<true>::bytes));
    C.coop_grid = std::max(1, std::min(per_sm, 2) * sms);
    // one set of team barrier counters per cooperative launch of a solve (2 per level at most), zeroed by ONE memset per solve
    CUDA_TRY(cudaMalloc((void**)&C.bars, (size_t)C.coop_grid * (2 * std::max(C.n_levels, 1)) * sizeof(unsigned int)));
    C.allocs.push_back((void*)C.bars);
    CUDA_TRY(cudaMalloc((void**)&C.back_scratch, (size_t)C.coop_grid * 2 * kBNB * ((S.max_front + 255) / 256) * sizeof(double)));      // [team][2 parity][kBNB][chunks of 256 rows]
    C.allocs.push_back((void*)C.back_scratch);
    CUDA_TRY(cudaMalloc((void**)&C.wscratch, (size_t)C.coop_grid * FrontCfg<true>::WSCR * sizeof(double)));
    C.allocs.push_back((void*)C.wscratch);
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&C.side, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&C.ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&C.ev_join, cudaEventDisableTiming));

  C.arena_doubles = S.arena;
  cudaError_t e = cudaMalloc((void**)&C.arena, std::max<long long>(1, S.arena) * sizeof(double));
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(LVS_ERR_OOM, "frontal arena allocation failed (%lld MB)", (long long)(S.arena * 8 >> 20)); }
  C.allocs.push_back((void*)C.arena);
  CUDA_TRY(cudaMalloc((void**)&C.xp, std::max<size_t>(1, (size_t)S.n * 6) * sizeof(double)));
  C.allocs.push_back((void*)C.xp);
  CUDA_TRY(cudaMalloc((void**)&C.fail_flag, sizeof(int)));
  C.allocs.push_back((void*)C.fail_flag);
  if (getenv("LVS_DEBUG_TIMING")) {
    CUDA_TRY(cudaMallocManaged((void**)&C.dbg, 16 * sizeof(long long)));
    memset(C.dbg, 0, 16 * sizeof(long long));
    C.allocs.push_back((void*)C.dbg);
  }
  const size_t smem = (size_t)S.max_front * sizeof(double);
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(chol_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaFuncSetAttribute(chol_backward_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + kBNB * (kBNB + 1) * sizeof(double))));
  CUDA_TRY(cudaFuncSetAttribute(chol_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrontCfg<false>::bytes));
  CUDA_TRY(cudaStreamSynchronize(st));    // the host vectors of S may go away
  return LVS_OK;
}

void chol_free(CholDevice& C) {
  if (C.side) cudaStreamDestroy(C.side);
  if (C.ev_fork) cudaEventDestroy(C.ev_fork);
  if (C.ev_join) cudaEventDestroy(C.ev_join);

  for (void* p : C.allocs) cudaFree(p);
  C = CholDevice();
}

int chol_solve(CholDevice& C, cudaStream_t st, const double* Hd, const double* Ho, const double* b, double lambda, double* x, double* scale_out,
               int* ok_out, int* launches) {
  CholView V;
  V.fronts = C.fronts; V.rows = C.rows; V.rel = C.rel; V.child_idx = C.child_idx; V.level_fronts = C.level_fronts; V.perm = C.perm;
  V.diag_dst = C.diag_dst; V.off_dst = C.off_dst; V.rhs_dst = C.rhs_dst; V.diag_ld = C.diag_ld; V.off_ld = C.off_ld; V.off_tr = C.off_tr;
  V.arena = C.arena; V.xp = C.xp; V.fail_flag = C.fail_flag; V.n = C.n; V.n_off = C.n_off;
  V.dbg = C.dbg; V.wscratch = C.wscratch;
  // (zeroing the arena on the side stream right after the previous solve's last kernel was measured: 2.53 -> 2.46 ms for an isolated solve, but
  // 0.4 ms SLOWER per LM run - there the update / error kernels follow the solve at once and the memset only gets in their way)
  CUDA_TRY(cudaMemsetAsync(C.arena, 0, (size_t)C.arena_doubles * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(C.fail_flag, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(C.bars, 0, (size_t)C.coop_grid * (2 * std::max(C.n_levels, 1)) * sizeof(unsigned int), st));
  unsigned int* bars_next = C.bars;
  const long long total = (long long)C.n * 36 + (long long)C.n_off * 36 + (long long)C.n * 6;
  chol_scatter_kernel<<<(unsigned)((total + kCholThreads - 1) / kCholThreads), kCholThreads, 0, st>>>(V, Hd, Ho, b, lambda);
  int nl = 1;
  for (int l = 0; l < C.n_levels; l++) {
    // small fronts: one CTA each; large fronts: teams of CTAs in one cooperative launch
    // (a level that has both kinds runs them side by side: the fronts of a level are independent, a team launch rarely fills the GPU,
    // and the few small fronts that sit high in the tree would otherwise add their whole latency to the critical path)
    const int ns = C.small_ptr[l + 1] - C.small_ptr[l], nb = C.big_ptr[l + 1] - C.big_ptr[l];
    const bool beside = ns > 0 && nb > 0;
    if (beside) { CUDA_TRY(cudaEventRecord(C.ev_fork, st)); CUDA_TRY(cudaStreamWaitEvent(C.side, C.ev_fork, 0)); }
    if (ns > 0) {
      chol_front_kernel<false><<<ns, kCholThreads, FrontCfg<false>::bytes, beside ? C.side : st>>>(V, C.small_list + C.small_ptr[l], ns, 1, C.bars);
      nl++;
    }
    if (nb > 0) {
      const int n_teams = std::min(nb, C.coop_grid);
      // no more CTAs per front than the largest front of the level has 96 x 96 tiles in its trailing update (fewer barrier parties)
      const int nt = (C.level_big[l] + FrontCfg<true>::TILE - 1) / FrontCfg<true>::TILE, tiles = std::max(nt * (nt + 1) / 2, (C.level_big[l] + kSlab - 1) / kSlab);
      int team_size = std::max(1, std::min(std::min(C.coop_grid / n_teams, kCholMaxTeam), tiles));
      int grid = n_teams * team_size;
      const int* list = C.big_list + C.big_ptr[l];
      int n_list = nb;
      if (team_size == 1) chol_front_kernel<false><<<grid, kCholThreads, FrontCfg<false>::bytes, st>>>(V, list, n_list, 1, C.bars);
      else {
        unsigned int* bars = bars_next; bars_next += C.coop_grid;
        void* args[] = {(void*)&V, (void*)&list, (void*)&n_list, (void*)&team_size, (void*)&bars};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)chol_front_kernel<true>, dim3(grid), dim3(kCholThreads), args, FrontCfg<true>::bytes, st));
      }
      nl++;
    }
    if (beside) { CUDA_TRY(cudaEventRecord(C.ev_join, C.side)); CUDA_TRY(cudaStreamWaitEvent(st, C.ev_join, 0)); }
  }
  const size_t smem = (size_t)C.max_front * sizeof(double);
  for (int l = C.n_levels - 1; l >= 0; l--) {
    const int ns = C.small_ptr[l + 1] - C.small_ptr[l], nbig = C.big_ptr[l + 1] - C.big_ptr[l];
    const bool beside = ns > 0 && nbig > 0;
    if (beside) { CUDA_TRY(cudaEventRecord(C.ev_fork, st)); CUDA_TRY(cudaStreamWaitEvent(C.side, C.ev_fork, 0)); }
    if (nbig > 0) {
      const int n_teams = std::min(nbig, C.coop_grid);
      int team_size = std::max(1, std::min(C.coop_grid / n_teams, (C.level_big[l] + 63) / 64));
      int max_chunks = (C.max_front + 255) / 256;
      const int* list = C.big_list + C.big_ptr[l];
      int n_list = nbig;
      if (team_size == 1) chol_backward_kernel<<<nbig, kCholThreads, smem, st>>>(V, list);
      else {
        int grid = n_teams * team_size;
        unsigned int* bars = bars_next; bars_next += C.coop_grid;
        void* args[] = {(void*)&V, (void*)&list, (void*)&n_list, (void*)&team_size, (void*)&bars, (void*)&C.back_scratch, (void*)&max_chunks};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)chol_backward_team_kernel, dim3(grid), dim3(kCholThreads), args, smem + kBNB * (kBNB + 1) * sizeof(double), st));
      }
      nl++;
    }
    if (ns > 0) {
      chol_backward_kernel<<<ns, kCholThreads, smem, beside ? C.side : st>>>(V, C.small_list + C.small_ptr[l]);
      nl++;
    }
    if (beside) { CUDA_TRY(cudaEventRecord(C.ev_join, C.side)); CUDA_TRY(cudaStreamWaitEvent(st, C.ev_join, 0)); }
  }
  chol_finish_kernel<<<1, 1024, 0, st>>>(V, b, lambda, x, scale_out, ok_out);
  nl++;

  if (C.dbg) {
    cudaStreamSynchronize(st);
    for (int r = 0; r < 2; r++)
      fprintf(stderr, "[chol dbg] last team front F=%lld team=%lld rank %d: diag %.1f us, B %.1f us, barrier1 %.1f us, C (rank 0: next diagonal block) %.1f us, barrier2 %.1f us, other %.1f us\n",
              C.dbg[8 * r + 6], C.dbg[8 * r + 7], r, C.dbg[8 * r + 0] / 1965.0, C.dbg[8 * r + 1] / 1965.0, C.dbg[8 * r + 2] / 1965.0, C.dbg[8 * r + 3] / 1965.0, C.dbg[8 * r + 4] / 1965.0,
              C.dbg[8 * r + 5] / 1965.0);
  }
  CUDA_TRY(cudaGetLastError());
  if (launches) *launches += nl;
  return LVS_OK;
}

}  // namespace lvs
