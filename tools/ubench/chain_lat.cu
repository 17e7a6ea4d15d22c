// Latencies that bound the serial pivot chain of the direct solver's micro block, one warp on one SM:
// dependent DFMA, dependent 64-bit shuffle, the reciprocal square root (hardware seed + Newton steps) and its accuracy,
// and three ways to factor a 16 x 16 block in registers.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_lat chain_lat.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
constexpr int kMB = 16;
__device__ __forceinline__ double rsq_seed(double d) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); return y; }
__device__ __forceinline__ double rsq1(double d) { double y = rsq_seed(d); const double h = 0.5 * d * y; return fma(y, fma(-h, y, 0.5), y); }
__device__ __forceinline__ double rsq2(double d) { double y = rsq1(d); const double h = 0.5 * d * y; return fma(y, fma(-h, y, 0.5), y); }
// seed in single precision (MUFU.RSQ, 23 bits) + two Newton steps in double
__device__ __forceinline__ double rsq_f32(double d) { double y = (double)rsqrtf((float)d); double h = 0.5 * d * y; y = fma(y, fma(-h, y, 0.5), y); h = 0.5 * d * y; return fma(y, fma(-h, y, 0.5), y); }

__global__ void k_lat(long long* out, double* sink, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) x = fma(x, b, a);
  }
  long long t1 = clock64();
  out[0] = (t1 - t0) / 1024;
  double y = x;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 1) & 31);
  }
  t1 = clock64();
  out[1] = (t1 - t0) / 1024;
  double z = fabs(y) + 1.5;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) z = rsq2(z) + 1.5;
  }
  t1 = clock64();
  out[2] = (t1 - t0) / 256;
  double z1 = z;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) z1 = rsq1(z1) + 1.5;
  }
  t1 = clock64();
  out[3] = (t1 - t0) / 256;
  double z2 = z1;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) z2 = rsq_seed(z2) + 1.5;
  }
  t1 = clock64();
  out[4] = (t1 - t0) / 256;
  double z3 = z2;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) z3 = rsq_f32(z3) + 1.5;
  }
  t1 = clock64();
  out[5] = (t1 - t0) / 256;
  // independent DFMA throughput of one warp
  double c[8];
  for (int u = 0; u < 8; u++) c[u] = z3 + u;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) c[u & 7] = fma(c[u & 7], b, a);
  }
  t1 = clock64();
  out[6] = (t1 - t0) * 100 / 1024;
  double sh[8];
  for (int u = 0; u < 8; u++) sh[u] = c[u];
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) sh[u & 7] = __shfl_sync(0xffffffffu, sh[u & 7], (threadIdx.x + 1) & 31);
  }
  t1 = clock64();
  out[7] = (t1 - t0) * 100 / 1024;
  double s = 0;
  for (int u = 0; u < 8; u++) s += sh[u];
  sink[threadIdx.x] = s + x;
}

__global__ void k_acc(double* err) {
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  unsigned long long x = 88172645463325252ull + threadIdx.x * 7919 + blockIdx.x * 104729;
  for (int i = 0; i < 20000; i++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const double d = ldexp((double)(x >> 11) / 9007199254740992.0 + 0.5, (int)(x & 63) - 32);
    const double ref = 1.0 / sqrt(d);
    e0 = fmax(e0, fabs(rsq_seed(d) - ref) / ref); e1 = fmax(e1, fabs(rsq1(d) - ref) / ref); e2 = fmax(e2, fabs(rsq2(d) - ref) / ref); e3 = fmax(e3, fabs(rsq_f32(d) - ref) / ref);
  }
  for (int o = 16; o; o >>= 1) { e0 = fmax(e0, __shfl_xor_sync(~0u, e0, o)); e1 = fmax(e1, __shfl_xor_sync(~0u, e1, o)); e2 = fmax(e2, __shfl_xor_sync(~0u, e2, o)); e3 = fmax(e3, __shfl_xor_sync(~0u, e3, o)); }
  if (threadIdx.x == 0) { err[4 * blockIdx.x] = e0; err[4 * blockIdx.x + 1] = e1; err[4 * blockIdx.x + 2] = e2; err[4 * blockIdx.x + 3] = e3; }
}

// factor-only variants.  0: lane = row, 15 - j shuffles per column (what the solver did before the inverse was fused in); 1: the same with
// one Newton step; 2: the two half-warps share a row (even / odd columns): half the shuffles and multiply-adds per instruction stream;
// 3: column broadcast through shared memory (one STS, LDS.128 pairs) except the next pivot's operand
template <int VAR>
__device__ __forceinline__ void micro(double* S, int ld, double* colbuf) {
  const int lane = threadIdx.x & 31, row = lane & 15;
  if (VAR <= 1) {
    double s[kMB];
#pragma unroll
    for (int k = 0; k < kMB; k++) s[k] = (k <= row) ? S[row * ld + k] : 0.0;
#pragma unroll
    for (int j = 0; j < kMB; j++) {
      const double d = __shfl_sync(0xffffffffu, s[j], j);
      const double il = VAR == 0 ? rsq2(d) : rsq1(d);
      const double lij = (row == j) ? d * il : s[j] * il;
      s[j] = lij;
#pragma unroll
      for (int k = j + 1; k < kMB; k++) s[k] -= lij * __shfl_sync(0xffffffffu, lij, k);
    }
    if (lane < kMB)
#pragma unroll
      for (int k = 0; k < kMB; k++) if (k <= row) S[row * ld + k] = s[k];
  } else if (VAR == 2) {
    const int h = lane >> 4;                       // this lane keeps columns k = 2 kk + h of its row
    double s[kMB / 2];
#pragma unroll
    for (int kk = 0; kk < kMB / 2; kk++) { const int k = 2 * kk + h; s[kk] = (k <= row) ? S[row * ld + k] : 0.0; }
#pragma unroll
    for (int j = 0; j < kMB; j++) {
      const int hj = j & 1, jj = j >> 1;           // column j lives in half hj, slot jj
      const double d = __shfl_sync(0xffffffffu, s[jj], j + 16 * hj);
      const double il = rsq2(d);
      const double mine = __shfl_sync(0xffffffffu, s[jj], row + 16 * hj);      // this row's entry of column j (from the owning half)
      const double lij = (row == j) ? d * il : mine * il;
      if (h == hj) s[jj] = lij;
#pragma unroll
      for (int kk = jj; kk < kMB / 2; kk++) {
        const int k = 2 * kk + h;                  // per-lane column; entries with k <= j are never read again
        const double lkj = __shfl_sync(0xffffffffu, lij, k & 15);
        if (k > j) s[kk] -= lij * lkj;
      }
    }
#pragma unroll
    for (int kk = 0; kk < kMB / 2; kk++) { const int k = 2 * kk + h; if (k <= row) S[row * ld + k] = s[kk]; }
  } else {
    double s[kMB];
#pragma unroll
    for (int k = 0; k < kMB; k++) s[k] = (k <= row) ? S[row * ld + k] : 0.0;
#pragma unroll
    for (int j = 0; j < kMB; j++) {
      const double d = __shfl_sync(0xffffffffu, s[j], j);
      const double il = rsq2(d);
      const double lij = (row == j) ? d * il : s[j] * il;
      s[j] = lij;
      if (lane < kMB) colbuf[j * kMB + row] = lij;
      if (j + 1 < kMB) s[j + 1] -= lij * __shfl_sync(0xffffffffu, lij, j + 1);       // the next pivot's operand stays on the short path
      __syncwarp();
#pragma unroll
      for (int k = j + 2; k < kMB; k += 2) {
        if (k + 1 < kMB && (k & 1) == 0) {
          const double2 v = *reinterpret_cast<const double2*>(colbuf + j * kMB + k);
          s[k] -= lij * v.x; s[k + 1] -= lij * v.y;
        } else {
          s[k] -= lij * colbuf[j * kMB + k];
          if (k + 1 < kMB) s[k + 1] -= lij * colbuf[j * kMB + k + 1];
        }
      }
    }
    if (lane < kMB)
#pragma unroll
      for (int k = 0; k < kMB; k++) if (k <= row) S[row * ld + k] = s[k];
  }
}

template <int VAR>
__global__ void k_micro(const double* A0, double* Lout, long long* cyc, int reps) {
  __shared__ __align__(16) double S[16 * 20], P[16 * 20], colbuf[256];
  for (int t = threadIdx.x; t < 256; t += 32) P[(t / 16) * 20 + t % 16] = A0[t];
  __syncwarp();
  long long tot = 0;
  for (int r = 0; r < reps; r++) {
    for (int t = threadIdx.x; t < 320; t += 32) S[t] = P[t];
    __syncwarp();
    const long long t0 = clock64();
    micro<VAR>(S, 20, colbuf);
    __syncwarp();
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) *cyc = tot / reps;
  for (int t = threadIdx.x; t < 256; t += 32) Lout[t] = (t % 16 <= t / 16) ? S[(t / 16) * 20 + t % 16] : 0.0;
}

int main() {
  long long* dT; double *dS, *dE;
  cudaMalloc(&dT, 64); cudaMalloc(&dS, 256); cudaMalloc(&dE, 4 * 64 * 8);
  k_lat<<<1, 32>>>(dT, dS, 1.0000001, 0.9999999);
  long long t[8]; cudaMemcpy(t, dT, 64, cudaMemcpyDeviceToHost);
  printf("dependent DFMA %lld cycles, dependent 64-bit SHFL %lld, rsqrt seed+2 Newton (+DADD) %lld, seed+1 Newton %lld, seed alone %lld, f32 seed + 2 Newton %lld; independent: DFMA %.2f cycles / warp instruction, 64-bit SHFL %.2f  (%s)\n",
         t[0], t[1], t[2], t[3], t[4], t[5], t[6] / 100.0, t[7] / 100.0, cudaGetErrorString(cudaGetLastError()));
  k_acc<<<64, 32>>>(dE);
  std::vector<double> e(256); cudaMemcpy(e.data(), dE, 256 * 8, cudaMemcpyDeviceToHost);
  double m[4] = {0, 0, 0, 0};
  for (int b = 0; b < 64; b++) for (int k = 0; k < 4; k++) m[k] = std::fmax(m[k], e[4 * b + k]);
  printf("max relative error of 1/sqrt(d): seed %.2e, + 1 Newton %.2e, + 2 Newton %.2e, f32 seed + 2 Newton %.2e\n", m[0], m[1], m[2], m[3]);
  const int n = 16;
  std::vector<double> A(256), B(256);
  unsigned long long x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (double)(x >> 11) / 9007199254740992.0 - 0.5; };
  for (auto& v : B) v = rnd();
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = (i == j) ? 2.0 : 0.0; for (int k = 0; k < n; k++) s += B[i * n + k] * B[j * n + k]; A[i * n + j] = s; }
  double *dA, *dL; cudaMalloc(&dA, 2048); cudaMalloc(&dL, 2048);
  cudaMemcpy(dA, A.data(), 2048, cudaMemcpyHostToDevice);
  for (int var = 0; var < 4; var++) {
    if (var == 0) k_micro<0><<<1, 32>>>(dA, dL, dT, 200); else if (var == 1) k_micro<1><<<1, 32>>>(dA, dL, dT, 200); else if (var == 2) k_micro<2><<<1, 32>>>(dA, dL, dT, 200); else k_micro<3><<<1, 32>>>(dA, dL, dT, 200);
    std::vector<double> L(256); long long cyc;
    cudaMemcpy(L.data(), dL, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dT, 8, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) { double s = 0; for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k]; worst = std::fmax(worst, std::fabs(s - A[i * n + j])); scale = std::fmax(scale, std::fabs(A[i * n + j])); }
    printf("factor-only variant %d: %lld cycles per 16 x 16 block, |L L^T - A| / |A| = %.2e (%s)\n", var, cyc, worst / scale, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
